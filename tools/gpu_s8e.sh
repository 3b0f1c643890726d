#!/bin/bash
# session 8, call E: persistent attention kernel as default -- full regression, smoke, benches (c2, LoRA, vision, 32-layer
# stacks), ncu launch list of the bench command + full capture of the attention kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 python bench.py --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err; tail -3 gpurun_out/bench_c2_lora.err
python tools/show_bench.py gpurun_out/bench_c2_lora.json | head -4
timeout 600 python bench.py --layers 32 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_stack32_c2.json 2> gpurun_out/bench_stack32_c2.err; tail -3 gpurun_out/bench_stack32_c2.err
python tools/show_bench.py gpurun_out/bench_stack32_c2.json | head -9
timeout 600 python bench.py --layers 32 --workload c4s --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_stack32_c4s.json 2> gpurun_out/bench_stack32_c4s.err; tail -3 gpurun_out/bench_stack32_c4s.err
python tools/show_bench.py gpurun_out/bench_stack32_c4s.json | head -9
timeout 600 python bench.py --vision --steps 5 --warmup 3 > gpurun_out/bench_vision63.json 2> gpurun_out/bench_vision63.err; tail -3 gpurun_out/bench_vision63.err
python -c "
import json
d=json.load(open('gpurun_out/bench_vision63.json')); print('vision', d['value'], d['ms_per_step'], d['encoder_tflops_per_gpu'])"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_s8.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc3 -s 6 -c 1 -o gpurun_out/prof_attn_tc3_s8e -f python tools/bench_kernels.py attention:tc3 > gpurun_out/ncu_attn_tc3.log 2>&1
tail -2 gpurun_out/ncu_attn_tc3.log
