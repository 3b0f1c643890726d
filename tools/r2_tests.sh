#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode.json 2> gpurun_out/r2_bench_decode.err; tail -5 gpurun_out/r2_bench_decode.err
VEX_DECODE_GEMM=k3 timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode_k3.json 2> gpurun_out/r2_bench_decode_k3.err; tail -5 gpurun_out/r2_bench_decode_k3.err
python tools/show_bench.py gpurun_out/r2_bench_decode.json gpurun_out/r2_bench_decode_k3.json
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; tail -3 gpurun_out/r2_bench_c2_b.err
python tools/show_bench.py gpurun_out/r2_bench_c2_b.json
bash tools/r2_sanitizer.sh
