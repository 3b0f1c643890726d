#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python tools/show_bench.py gpurun_out/bench_n$N.json | head -12
python -c "
import json;d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]);print(d.get('host_numa'))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --train --gpus $N --layers 32 --steps 3 --warmup 3 > gpurun_out/bench_train32_n$N.json 2> gpurun_out/bench_train32_n$N.err; tail -5 gpurun_out/bench_train32_n$N.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_train32_n$N.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('value','ms_per_step','allreduce_ms','allreduce_bytes','trainable_params','step_tflops_per_gpu','step_frac_of_bf16_peak')})"
