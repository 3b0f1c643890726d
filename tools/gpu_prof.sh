#!/bin/bash
# regression + bench + ncu captures of the three kernel families
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err
python tools/show_bench.py gpurun_out/bench_c2.json gpurun_out/bench_c2_lora.json
B="python bench.py --steps 2 --warmup 3 --no-cpu --graph 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_r1b.csv $B > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_grouped_gemm_pair -s 8 -c 4 -o gpurun_out/prof_gemm_pair_r1 -f $B > gpurun_out/ncu_full_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc -s 2 -c 1 -o gpurun_out/prof_attn_tc_r1 -f $B > gpurun_out/ncu_full_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_rmsnorm -s 4 -c 2 -o gpurun_out/prof_rmsnorm_r1 -f $B > gpurun_out/ncu_full_norm.log 2>&1
ls -la gpurun_out | tail -20
