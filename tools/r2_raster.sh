#!/bin/bash
# raster-group sweep at the c3 N=1 shape (64 samples = 95 040 rows per GEMM): per-kernel times of one layer
mkdir -p gpurun_out
for rg in 8 12 16 24 32 40 64; do
  VEX_GEMM_RASTER=$rg timeout 300 python bench.py --workload c3 --layers 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_raster_$rg.json 2> gpurun_out/r2_raster_$rg.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_raster_$rg.json"))
k=d["kernels"]
print("raster $rg: step %.3f ms clk %s | swiglu %.3f rope %.3f residual(2) %.3f attn %.3f" % (d["ms_per_step"], d["clocks"]["sm_mhz"], k["gemm_swiglu"]["ms"], k["gemm_rope"]["ms"], k["gemm_residual"]["ms"], k["attention"]["ms"]))
PY
done
