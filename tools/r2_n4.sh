#!/bin/bash
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n4.json 2> gpurun_out/r2_bench_c3_n4.err; tail -2 gpurun_out/r2_bench_c3_n4.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n4.json | grep -v "^      "
