#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k9" --tb=line > gpurun_out/k.log 2>&1; echo "exit=$?"; tail -12 gpurun_out/k.log
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_layer_gpu.py -k "training" --tb=line > gpurun_out/k2.log 2>&1; echo "exit=$?"; tail -5 gpurun_out/k2.log
timeout 600 python bench.py --train --layers 4 --steps 5 --warmup 3 > gpurun_out/bench_train4.json 2> gpurun_out/bench_train4.err; tail -5 gpurun_out/bench_train4.err; python tools/show_bench.py gpurun_out/bench_train4.json
