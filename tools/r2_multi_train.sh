#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for tag in a b; do
for g in 0 8 32; do
timeout 900 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 6 --warmup 3 --ar-group $g > gpurun_out/r2_train32_g${g}_${tag}_n$N.json 2> gpurun_out/r2_train32_g${g}_${tag}_n$N.err; tail -1 gpurun_out/r2_train32_g${g}_${tag}_n$N.err
done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_train32_g*_n$N.json")):
    try:
        d=json.load(open(f))
        print(f, {k: round(d.get(k) or 0, 2) for k in ("ms_per_step","ms_per_step_without_collectives","allreduce_exposed_ms","allreduce_tail_ms")}, d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
