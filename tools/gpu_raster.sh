#!/bin/bash
# raster-group A/B on one box: GEMM micro-bench with the minimal group (16) and the L2-sized default, parity, bench
for r in 16 default 16 default; do
  if [ $r = default ]; then unset VEX_GEMM_RASTER; else export VEX_GEMM_RASTER=$r; fi
  echo "== raster $r"; timeout 100 python tools/bench_kernels.py gemm 2>&1 | grep kernel | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' ', d['kernel'], round(d['ms'],4), round(d['tflops'],1))"
done
unset VEX_GEMM_RASTER
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_layer_gpu.py -q -m gpu -p no:cacheprovider --tb=line -x 2>&1 | tail -2
for r in 16 default; do
  if [ $r = default ]; then unset VEX_GEMM_RASTER; else export VEX_GEMM_RASTER=$r; fi
  timeout 300 python bench.py --layers 32 --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('stack32 raster $r:', round(d['ms_per_step'],2), 'ms', round(d['value']), 'tok/s', d['clocks'].get('sm_mhz'))"
done
