#!/bin/bash
# session 5, call A: attention forward with packed-math softmax (tests + micro-bench + ncu), bench with NVML clocks
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py tests/test_vision_gpu.py -k "k4 or attention or k9 or vision" --tb=line > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -5 gpurun_out/k.log
timeout 300 python tools/bench_kernels.py attention
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc -s 6 -c 1 -o gpurun_out/prof_attn_fwd_s5a -f python tools/bench_kernels.py attention > gpurun_out/ncu_attn_fwd.log 2>&1
tail -2 gpurun_out/ncu_attn_fwd.log
ls -la gpurun_out/*.ncu-rep
