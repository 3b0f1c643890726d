#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_layer_gpu.py -m gpu -q -k "k8 or wgrad or training or lora" 2>&1 | tail -5
bash tools/r2_k8_ncu.sh
timeout 300 python bench.py --train --workload c2 --layers 1 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('roofline'))"
