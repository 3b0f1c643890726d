#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_b.json 2> gpurun_out/r2_bench_c2_b.err; tail -3 gpurun_out/r2_bench_c2_b.err
timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode.json 2> gpurun_out/r2_bench_decode.err; tail -5 gpurun_out/r2_bench_decode.err
python tools/show_bench.py gpurun_out/r2_bench_c2_b.json gpurun_out/r2_bench_decode.json
