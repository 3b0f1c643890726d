#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_layer_gpu.py -k "synthetic or vision_only" > gpurun_out/layer2.log 2>&1; tail -3 gpurun_out/layer2.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 300 python bench.py --steps 20 --warmup 5 --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err; tail -c 1500 gpurun_out/bench_c2_lora.json; tail -5 gpurun_out/bench_c2_lora.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_grouped_gemm -s 8 -c 4 -o gpurun_out/prof_gemm_r1 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
