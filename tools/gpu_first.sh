#!/bin/bash
# first GPU session: every group under its own timeout so a hung kernel cannot eat the whole call
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 300 python -m pytest -q -m gpu -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -5 gpurun_out/$name.log; }
run k1 tests/test_kernels_gpu.py -k "k1"
run k2k5k6 tests/test_kernels_gpu.py -k "k2 or k5 or k6"
run k4 tests/test_kernels_gpu.py -k "k4"
run k3plain tests/test_kernels_gpu.py -k "k3_plain"
run k3small tests/test_kernels_gpu.py -k "k3_small"
run k3res tests/test_kernels_gpu.py -k "k3_scatter"
run k3lora tests/test_kernels_gpu.py -k "k3_lora"
run k3swiglu tests/test_kernels_gpu.py -k "k3_swiglu"
run k3rope tests/test_kernels_gpu.py -k "k3_rope"
run layer tests/test_layer_gpu.py
cat gpurun_out/summary.txt
