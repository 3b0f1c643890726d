#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_layer_gpu.py -m gpu -q -x -s -k "full_width_r64" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_pytest_grad.log; cat gpurun_out/r2_pytest_grad.log
bash tools/r2_ncu.sh
