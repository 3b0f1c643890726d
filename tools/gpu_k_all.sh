#!/bin/bash
# like gpu_k.sh but without -x (see every failing case)
mkdir -p gpurun_out
F=${2:-tests/test_kernels_gpu.py}
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider $F -k "$1" --tb=line > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -40 gpurun_out/k.log
