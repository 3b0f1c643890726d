#!/bin/bash
# decode attention: single-pass (default) vs the exact-rounding two-pass kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "k4d or decode or static_cache or tuple_cache" 2>&1 | tail -4
for impl in 2pass 1pass 2pass 1pass; do
  VEX_K4D_IMPL=$impl timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode_$impl.json 2> gpurun_out/r2_bench_decode_$impl.err || tail -5 gpurun_out/r2_bench_decode_$impl.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_decode_$impl.json").read().strip().splitlines()[-1])
k=d.get("kernels") or {}
print("$impl ms_per_step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {n: round(v["ms"],4) for n,v in k.items() if "atten" in n})
PY
done
