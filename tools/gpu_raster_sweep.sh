for r in 24 40 64 80 40 64; do
  export VEX_GEMM_RASTER=$r
  echo "== raster $r"; timeout 60 python tools/bench_kernels.py gemm 2>&1 | grep kernel | python -c "
import sys,json
print('  '+'  '.join(f\"{json.loads(l)['kernel'][5:]} {json.loads(l)['ms']:.4f}\" for l in sys.stdin))"
done
