#!/bin/bash
# memory-adaptive activation keeping: training parity, then the 32-layer step in auto / forced-partial / checkpoint mode
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layer_gpu.py -m gpu -q -k "training" 2>&1 | tail -3
run() {  # name, extra env, extra args
  env $2 timeout 900 python bench.py --train --workload c2 --layers 32 --steps 4 --warmup 2 $3 > gpurun_out/r2_bench_train32_$1.json 2> gpurun_out/r2_bench_train32_$1.err || tail -3 gpurun_out/r2_bench_train32_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_train32_$1.json").read().strip().splitlines()[-1])
c=d["config"]
print("$1", "ms/step %.1f" % d["ms_per_step"], "tok/s %.0f" % d["value"], "recomputed fraction", c.get("recomputed_layer_fraction"), "peak mem %.1f GB" % d["peak_mem_gb"], "frac burst %.3f" % d["step_frac_of_bf16_peak"], "clk", d["clocks"]["sm_mhz"])
PY
}
run auto "X=1" ""
run partial "VEX_TRAIN_KEEP_RESERVE_GB=140" ""
run ckpt "X=1" "--recompute 1"
