#!/bin/bash
# end-of-round 8-GPU record on the final code: c3 (default), c4, the training step in its default (auto-keep) mode
N=${1:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n${N}_final.json 2> gpurun_out/r2_bench_c3_n${N}_final.err; tail -2 gpurun_out/r2_bench_c3_n${N}_final.err
timeout 600 $RUN bench.py --gpus $N --workload c4 --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_n${N}_final.json 2> gpurun_out/r2_bench_c4_n${N}_final.err; tail -2 gpurun_out/r2_bench_c4_n${N}_final.err
timeout 600 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 6 --warmup 3 > gpurun_out/r2_bench_train32_auto_n$N.json 2> gpurun_out/r2_bench_train32_auto_n$N.err; tail -2 gpurun_out/r2_bench_train32_auto_n$N.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n${N}_final.json gpurun_out/r2_bench_c4_n${N}_final.json 2>/dev/null | grep -v "^      "
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_train32_auto_n$N.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value","ms_per_step","ms_per_step_without_collectives","allreduce_exposed_ms","allreduce_tail_ms","step_frac_of_bf16_peak","peak_mem_gb")}, d["config"].get("recomputed_layer_fraction"), d.get("clocks"))
except Exception as e: print("train ERR", e)
PY
