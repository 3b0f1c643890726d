#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k3 and pair" -x > gpurun_out/k3pair.log 2>&1; echo "pair exit=$?"; tail -25 gpurun_out/k3pair.log
timeout 240 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k3 and single" -x > gpurun_out/k3single.log 2>&1; echo "single exit=$?"; tail -3 gpurun_out/k3single.log
VEX_GEMM_PAIR=1 timeout 200 python tools/bench_kernels.py gemm > gpurun_out/bench_gemm_pair.log 2>&1; cat gpurun_out/bench_gemm_pair.log
VEX_GEMM_PAIR=0 timeout 200 python tools/bench_kernels.py gemm > gpurun_out/bench_gemm_single.log 2>&1; cat gpurun_out/bench_gemm_single.log
