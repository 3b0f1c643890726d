#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k4 or k9" --tb=line > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -8 gpurun_out/k.log
timeout 300 python tools/bench_kernels.py attention
timeout 600 python bench.py --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; python tools/show_bench.py gpurun_out/bench_c2.json
