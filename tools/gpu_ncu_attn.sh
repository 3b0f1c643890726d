#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc -s 6 -c 1 -o gpurun_out/prof_attn_fwd_r1b -f python tools/bench_kernels.py attention > gpurun_out/ncu_attn_fwd.log 2>&1
tail -3 gpurun_out/ncu_attn_fwd.log
ls -la gpurun_out/*.ncu-rep
