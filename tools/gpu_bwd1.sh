#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k7 or dgrad" > gpurun_out/bwd1.log 2>&1; echo "exit=$?"; tail -40 gpurun_out/bwd1.log
