#!/bin/bash
# K4: parity tests, then old (git HEAD at build time) vs new timing in alternating processes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_layer_gpu.py tests/test_vision_gpu.py -m gpu -q -x -k "k4 or attention or layer_vs or c4 or c2 or sorted or vision" 2>&1 | tail -4
timeout 400 python tools/k4_ab.py run
