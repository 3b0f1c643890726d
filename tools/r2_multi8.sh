#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; tail -2 gpurun_out/r2_bench_c3_n$N.err
if [ "$N" = "8" ]; then
timeout 600 $RUN bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err; tail -2 gpurun_out/r2_bench_c4_n$N.err
timeout 600 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 6 --warmup 3 > gpurun_out/r2_bench_train32_n$N.json 2> gpurun_out/r2_bench_train32_n$N.err; tail -2 gpurun_out/r2_bench_train32_n$N.err
timeout 300 $RUN bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; tail -2 gpurun_out/r2_bench_ref_n$N.err
fi
python tools/show_bench.py gpurun_out/r2_bench_c3_n$N.json gpurun_out/r2_bench_c4_n$N.json 2>/dev/null | grep -v "^      "
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_train32_n$N.json"))
    print({k: d.get(k) for k in ("value","ms_per_step","ms_per_step_without_collectives","allreduce_exposed_ms","allreduce_tail_ms","step_frac_of_bf16_peak")}, d.get("clocks"))
except Exception as e: print("train ERR", e)
PY
wc -l gpurun_out/r2_bench_ref_n$N.json 2>/dev/null
