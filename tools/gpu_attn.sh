#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k4" > gpurun_out/k4tc.log 2>&1; tail -15 gpurun_out/k4tc.log
timeout 300 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; cat gpurun_out/bench_kernels.log
