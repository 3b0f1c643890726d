#!/bin/bash
# full GPU regression + bench (what the driver runs at round end)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1500 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
timeout 600 python bench.py --graph 0 --no-cpu > gpurun_out/bench_c2_eager.json 2> gpurun_out/bench_c2_eager.err; tail -5 gpurun_out/bench_c2_eager.err
timeout 600 python bench.py --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err; tail -5 gpurun_out/bench_c2_lora.err
timeout 600 python bench.py --workload c4s --no-cpu > gpurun_out/bench_c4s.json 2> gpurun_out/bench_c4s.err; tail -5 gpurun_out/bench_c4s.err
timeout 900 python bench.py --layers 32 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c3_32l.json 2> gpurun_out/bench_c3_32l.err; tail -5 gpurun_out/bench_c3_32l.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err
python tools/show_bench.py gpurun_out/bench_c2.json gpurun_out/bench_c2_eager.json gpurun_out/bench_c2_lora.json gpurun_out/bench_c4s.json gpurun_out/bench_c3_32l.json gpurun_out/bench_ref.json
