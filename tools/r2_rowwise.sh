#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "k7 or k5 or k6 or silu or rmsnorm or gather or dropout or training or residual" 2>&1 | tail -3
timeout 900 python bench.py --train --workload c2 --layers 32 --steps 4 --warmup 2 > gpurun_out/r2_bench_train32_auto_n1.json 2> gpurun_out/r2_bench_train32_auto_n1.err || tail -3 gpurun_out/r2_bench_train32_auto_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_train32_auto_n1.json").read().strip().splitlines()[-1]); k=d["kernels"]
print("ms/step %.1f" % d["ms_per_step"], "tok/s %.0f" % d["value"], "clk", d["clocks"]["sm_mhz"])
for n in ("lora_wgrad","silu_mul_backward","rmsnorm_backward","silu_mul","gather_rows","rmsnorm"):
    if n in k: print("  %-22s %7.3f ms x%d  %.1f us/call" % (n, k[n]["ms"], k[n]["calls_per_step"], 1e3*k[n]["ms"]/k[n]["calls_per_step"]))
PY
