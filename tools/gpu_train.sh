#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_layer_gpu.py -k "training or error" --tb=short > gpurun_out/train_test.log 2>&1; tail -30 gpurun_out/train_test.log
timeout 600 python bench.py --train --layers 4 --steps 5 --warmup 3 > gpurun_out/bench_train4.json 2> gpurun_out/bench_train4.err; tail -5 gpurun_out/bench_train4.err; cat gpurun_out/bench_train4.json | cut -c1-300
