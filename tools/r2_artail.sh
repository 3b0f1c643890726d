#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for tail in 0 2 0 2; do
timeout 600 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 6 --warmup 3 --ar-tail $tail > gpurun_out/r2_artail${tail}_n$N.json 2> gpurun_out/r2_artail${tail}_n$N.err; tail -1 gpurun_out/r2_artail${tail}_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_artail${tail}_n$N.json").read().strip().splitlines()[-1])
print("tail $tail", {k: round(d.get(k) or 0, 2) for k in ("ms_per_step","ms_per_step_without_collectives","allreduce_exposed_ms","allreduce_tail_ms")}, d["clocks"]["sm_mhz"])
PY
done
