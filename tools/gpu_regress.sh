#!/bin/bash
# full GPU regression + default bench + vision bench (after a K3 epilogue change)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 python bench.py --lora 64 --no-cpu > gpurun_out/bench_c2_lora.json 2> gpurun_out/bench_c2_lora.err; tail -3 gpurun_out/bench_c2_lora.err
python tools/show_bench.py gpurun_out/bench_c2_lora.json
timeout 600 python bench.py --vision --steps 5 --warmup 3 > gpurun_out/bench_vision63.json 2> gpurun_out/bench_vision63.err; tail -5 gpurun_out/bench_vision63.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_vision63.json'))
print(d['value'], d['ms_per_step'], d['encoder_tflops_per_gpu'], d['roofline'])
for k,v in d['kernels'].items(): print(k, v)
PY
