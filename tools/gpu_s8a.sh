#!/bin/bash
# session 8, call A: the two-tile ping-pong attention kernel (default stays tc1 until this is green): parity,
# micro-bench of every implementation, layer tests + bench with VEX_ATTN_IMPL=tc2 when parity passes, ncu capture
mkdir -p gpurun_out
timeout 400 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k4_attention" --tb=short > gpurun_out/k4.log 2>&1; K4=$?; echo "k4 exit=$K4"; tail -25 gpurun_out/k4.log
timeout 300 python tools/bench_kernels.py attention 2>&1 | tail -14 | tee gpurun_out/attn_bench.log
if [ $K4 -eq 0 ]; then export VEX_ATTN_IMPL=tc2; echo "running layer tests + bench with tc2"; fi
timeout 600 python -m pytest tests/test_layer_gpu.py tests/test_vision_gpu.py -q -m gpu -p no:cacheprovider --tb=line > gpurun_out/pytest_layer.log 2>&1; echo "layer exit=$?"; tail -8 gpurun_out/pytest_layer.log
timeout 400 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; python tools/show_bench.py gpurun_out/bench_c2.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc2 -s 6 -c 1 -o gpurun_out/prof_attn_tc2_s8a -f python tools/bench_kernels.py attention:tc2 > gpurun_out/ncu_attn_tc2.log 2>&1
tail -2 gpurun_out/ncu_attn_tc2.log
ls -la gpurun_out/*.ncu-rep
