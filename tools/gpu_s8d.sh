#!/bin/bash
# session 8, call D: warp-uniform elected MMA issue in K3 / K8 / K9 -- full regression, c2 bench, 4-layer training bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 python bench.py --train --layers 4 --steps 5 --warmup 3 > gpurun_out/bench_train4.json 2> gpurun_out/bench_train4.err; tail -5 gpurun_out/bench_train4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train4.json'))
print('train4', d['value'], d['ms_per_step'])
for k,v in sorted(d['kernels'].items(), key=lambda kv:-kv[1]['ms']): print('  ', k, round(v['ms'],3), v['calls_per_step'])
PY
timeout 200 python tools/bench_kernels.py gemm 2>&1 | tail -12
