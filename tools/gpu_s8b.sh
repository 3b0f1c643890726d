#!/bin/bash
# session 8, call B: early-S schedule of the two-tile attention kernel: parity + micro-bench; pipelined host prefill test
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k4_attention and early" --tb=short > gpurun_out/k4.log 2>&1; echo "k4 exit=$?"; tail -8 gpurun_out/k4.log
timeout 200 python tools/bench_kernels.py attention 2>&1 | tail -18 | tee gpurun_out/attn_bench.log
timeout 200 python -m pytest -q -m gpu -p no:cacheprovider tests/test_layer_gpu.py -k "pipelined or graphed" --tb=short 2>&1 | tail -5
