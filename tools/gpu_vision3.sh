#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vision_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_vision.log 2>&1; tail -3 gpurun_out/pytest_vision.log
timeout 600 python bench.py --vision --steps 5 --warmup 3 > gpurun_out/bench_vision63_v3.json 2> gpurun_out/bench_vision63.err; tail -5 gpurun_out/bench_vision63.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_vision63_v3.json'))
print(d['value'], d['ms_per_step'], d['encoder_tflops_per_gpu'], d['roofline'])
for k,v in d['kernels'].items(): print(k, v)
PY
B="python bench.py --vision --vision-layers 2 --steps 1 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_vision.csv $B > gpurun_out/ncu_vision_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_grouped_gemm_pair -s 3 -c 1 -o gpurun_out/prof_vision_fc1 -f $B > gpurun_out/ncu_full_fc1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k11_layernorm -s 2 -c 1 -o gpurun_out/prof_vision_ln -f $B > gpurun_out/ncu_full_ln.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k4_attention_tc -s 1 -c 1 -o gpurun_out/prof_vision_attn -f $B > gpurun_out/ncu_full_attn.log 2>&1
ls -la gpurun_out | tail -8
