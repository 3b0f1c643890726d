#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "k9" 2>&1 | tail -15 > gpurun_out/r2_k9_test.log; echo "rc=$?"; cat gpurun_out/r2_k9_test.log
timeout 300 python -m pytest tests/test_layer_gpu.py -m gpu -q -x -k "training" 2>&1 | tail -8
for impl in grid persistent; do
  VEX_K9_IMPL=$impl timeout 300 python bench.py --train --workload c2 --layers 4 --steps 6 --warmup 3 > gpurun_out/r2_k9_train4_$impl.json 2> gpurun_out/r2_k9_train4_$impl.err; tail -2 gpurun_out/r2_k9_train4_$impl.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_k9_train4_$impl.json")); k=d["kernels"]
print("$impl", "ms/step %.2f" % d["ms_per_step"], "attention_backward %.3f ms x%d" % (k["attention_backward"]["ms"], k["attention_backward"]["calls_per_step"]))
PY
done
