#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --train --gpus $N --layers 32 --steps 3 --warmup 3 > gpurun_out/bench_train32_n$N.json 2> gpurun_out/bench_train32_n$N.err; tail -5 gpurun_out/bench_train32_n$N.err
python tools/show_bench.py gpurun_out/bench_train32_n$N.json; cut -c1-200 gpurun_out/bench_train32_n$N.json; python -c "
import json;d=json.loads(open('gpurun_out/bench_train32_n$N.json').read().strip().splitlines()[-1]);print({k:d[k] for k in ('allreduce_ms','allreduce_bytes','trainable_params','step_tflops_per_gpu','step_frac_of_bf16_peak')})"
timeout 600 python -m pytest -q -p no:cacheprovider tests/test_sharding_gloo.py 2>&1 | tail -3
