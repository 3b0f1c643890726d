#!/bin/bash
# programmatic dependent launch in the decode step: parity, then the decode bench with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "k12 or k4d or decode or static_cache or tuple_cache or k2 or rmsnorm or layer_vs or sorted" 2>&1 | tail -5
for pdl in 0 1 0 1; do
  VEX_PDL=$pdl timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode_pdl$pdl.json 2> gpurun_out/r2_bench_decode_pdl$pdl.err || tail -5 gpurun_out/r2_bench_decode_pdl$pdl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_decode_pdl$pdl.json").read().strip().splitlines()[-1])
print("VEX_PDL=$pdl ms_per_step %.4f" % d["ms_per_step"], "value", round(d["value"],1), d.get("roofline"))
PY
done
