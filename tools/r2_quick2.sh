#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 900 python bench.py --train --workload c2 --layers 32 --steps 4 --warmup 2 > gpurun_out/r2_bench_train32_n1.json 2> gpurun_out/r2_bench_train32_n1.err; tail -2 gpurun_out/r2_bench_train32_n1.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_train32_n1.json")); k=d["kernels"]
print({x: d.get(x) for x in ("value","ms_per_step","step_frac_of_bf16_peak","step_frac_of_bf16_sustained")}, d["clocks"]["sm_mhz"])
for n,v in sorted(k.items(), key=lambda kv:-kv[1]["ms"])[:12]: print("  %-26s %8.3f ms x%d" % (n, v["ms"], v["calls_per_step"]))
PY
