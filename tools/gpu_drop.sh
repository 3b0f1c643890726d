#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "dropout or dgrad" --tb=short > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -30 gpurun_out/k.log
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_layer_gpu.py -k "training" --tb=short > gpurun_out/k2.log 2>&1; echo "exit=$?"; tail -30 gpurun_out/k2.log
