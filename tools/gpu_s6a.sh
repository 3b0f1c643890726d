#!/bin/bash
# session 6, call A: two-tile ping-pong attention kernel (P in TMEM / shared memory) -- parity, micro-bench, ncu; then the
# full GPU regression and the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_kernels_gpu.py -k "k4_attention" --tb=short -x > gpurun_out/k4.log 2>&1; echo "k4 exit=$?"; tail -15 gpurun_out/k4.log
timeout 300 python tools/bench_kernels.py attention 2>&1 | tail -14
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --tb=line > gpurun_out/pytest_gpu.log 2>&1; echo "full exit=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err; python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc2 -s 6 -c 1 -o gpurun_out/prof_attn_tc2_s6a -f python tools/bench_kernels.py attention:tc2 > gpurun_out/ncu_attn_tc2.log 2>&1
tail -2 gpurun_out/ncu_attn_tc2.log
ls -la gpurun_out/*.ncu-rep
