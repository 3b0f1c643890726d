#!/bin/bash
# usage: tools/r2_multi.sh N   -- strong-scaling c3 bench + LoRA training step with the overlapped all-reduce on N GPUs
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $RUN bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err; tail -2 gpurun_out/r2_bench_c3_n$N.err
timeout 900 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 3 --warmup 2 > gpurun_out/r2_bench_train32_n$N.json 2> gpurun_out/r2_bench_train32_n$N.err; tail -2 gpurun_out/r2_bench_train32_n$N.err
timeout 900 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 3 --warmup 2 --ar-group 8 > gpurun_out/r2_bench_train32_g8_n$N.json 2> gpurun_out/r2_bench_train32_g8_n$N.err; tail -2 gpurun_out/r2_bench_train32_g8_n$N.err
timeout 900 $RUN bench.py --gpus $N --train --workload c2 --layers 32 --steps 3 --warmup 2 --allreduce post > gpurun_out/r2_bench_train32_post_n$N.json 2> gpurun_out/r2_bench_train32_post_n$N.err; tail -2 gpurun_out/r2_bench_train32_post_n$N.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n$N.json
python - <<PY
import json
for f in ("gpurun_out/r2_bench_train32_n$N.json", "gpurun_out/r2_bench_train32_g8_n$N.json", "gpurun_out/r2_bench_train32_post_n$N.json"):
    try:
        d=json.load(open(f))
        print(f, {k: d.get(k) for k in ("value","ms_per_step","ms_per_step_without_collectives","allreduce_exposed_ms","allreduce_tail_ms","allreduce_bytes","allreduce_busbw_gbs_over_exposed_time","step_frac_of_bf16_peak","peak_mem_gb")}, d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
