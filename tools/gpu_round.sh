#!/bin/bash
# baseline numbers for the wider configs: 32-layer stack (c3 per-GPU shard at N=8), c4 shard, LoRA, training step
mkdir -p gpurun_out
timeout 600 python bench.py --layers 32 --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_stack32.json 2> gpurun_out/bench_stack32.err; tail -3 gpurun_out/bench_stack32.err
timeout 600 python bench.py --workload c4s --layers 32 --no-cpu --steps 5 --warmup 3 > gpurun_out/bench_c4s_stack32.json 2> gpurun_out/bench_c4s_stack32.err; tail -3 gpurun_out/bench_c4s_stack32.err
timeout 600 python bench.py --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err; tail -3 gpurun_out/bench_c2_lora.err
timeout 600 python bench.py --train --layers 4 --steps 5 --warmup 3 > gpurun_out/bench_train4.json 2> gpurun_out/bench_train4.err; tail -5 gpurun_out/bench_train4.err
python tools/show_bench.py gpurun_out/bench_stack32.json gpurun_out/bench_c4s_stack32.json gpurun_out/bench_c2_lora.json
cut -c1-1200 gpurun_out/bench_train4.json
