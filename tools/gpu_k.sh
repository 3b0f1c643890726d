#!/bin/bash
# usage: tools/gpu_k.sh "<pytest -k expression>" [file]
mkdir -p gpurun_out
F=${2:-tests/test_kernels_gpu.py}
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider $F -k "$1" -x > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -60 gpurun_out/k.log
