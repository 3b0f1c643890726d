#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vision_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_vision.log 2>&1; tail -5 gpurun_out/pytest_vision.log
timeout 600 python -m pytest tests/test_layer_gpu.py -q -m gpu -p no:cacheprovider -k training > gpurun_out/pytest_train.log 2>&1; tail -5 gpurun_out/pytest_train.log
timeout 600 python bench.py --vision --steps 5 --warmup 3 > gpurun_out/bench_vision63_v2.json 2> gpurun_out/bench_vision63.err; tail -5 gpurun_out/bench_vision63.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_vision63_v2.json'))
print(d['value'], d['ms_per_step'], d['encoder_tflops_per_gpu'], d['roofline'])
for k,v in d['kernels'].items(): print(k, v)
PY
for rc in 1 0; do
timeout 600 python bench.py --train --layers 4 --steps 5 --warmup 3 --recompute $rc > gpurun_out/bench_train4_rc$rc.json 2> gpurun_out/bench_train4.err; tail -3 gpurun_out/bench_train4.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_train4_rc$rc.json'))
print('recompute', $rc, d['value'], d['ms_per_step'], d['step_frac_of_bf16_peak'], d.get('peak_mem_gb'))
PY
done
