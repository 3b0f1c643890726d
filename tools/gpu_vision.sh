#!/bin/bash
# vision encoder: new parity tests first, then the full regression, smoke, default bench, vision bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vision_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_vision.log 2>&1; tail -25 gpurun_out/pytest_vision.log
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider --deselect tests/test_vision_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -3 gpurun_out/bench_c2.err
python tools/show_bench.py gpurun_out/bench_c2.json
timeout 600 python bench.py --vision --steps 5 --warmup 3 > gpurun_out/bench_vision63.json 2> gpurun_out/bench_vision63.err; tail -5 gpurun_out/bench_vision63.err
cut -c1-3000 gpurun_out/bench_vision63.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_vision.csv python bench.py --vision --vision-layers 2 --steps 1 --warmup 3 > gpurun_out/ncu_vision_list.log 2>&1
tail -3 gpurun_out/ncu_vision_list.log
