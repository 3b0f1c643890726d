#!/bin/bash
# round 2, first GPU pass: full GPU test suite, then the default (c3, 32-layer, strong-scaling) bench at N = 1, the
# single-layer c2 bench, and the reference arm.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; tail -3 gpurun_out/r2_bench_c3_n1.err
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -3 gpurun_out/r2_bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -3 gpurun_out/r2_bench_ref.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n1.json gpurun_out/r2_bench_c2.json gpurun_out/r2_bench_ref.json
timeout 600 python bench.py --train --workload c2 --layers 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_train4.json 2> gpurun_out/r2_bench_train4.err; tail -3 gpurun_out/r2_bench_train4.err
cut -c1-900 gpurun_out/r2_bench_train4.json
