#!/bin/bash
# decode step after the single-resident-wave changes (K12 4-warp / 5-per-SM forms, K4d split rule): parity + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "k12 or k4d or decode or static_cache or tuple_cache" 2>&1 | tail -3
for f in a b; do
  timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode_$f.json 2> gpurun_out/r2_bench_decode_$f.err || tail -5 gpurun_out/r2_bench_decode_$f.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_decode_$f.json").read().strip().splitlines()[-1])
print("run $f ms_per_step %.4f" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "e2e", round(d["e2e"]["value"],1))
PY
done
