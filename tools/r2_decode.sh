#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "k12 or k4d or decode or static_cache or tuple_cache or full_width_r64 or inference_mode" 2>&1 | tail -25 > gpurun_out/r2_pytest_c.log; cat gpurun_out/r2_pytest_c.log
timeout 600 python bench.py --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode.json 2> gpurun_out/r2_bench_decode.err; tail -5 gpurun_out/r2_bench_decode.err
python tools/show_bench.py gpurun_out/r2_bench_decode.json
timeout 900 ncu --clock-control none --set full --import-source on -k regex:'k12_decode_gemm|k4_attention_decode|k2_rmsnorm' --launch-skip 14 -c 7 \
  -o gpurun_out/r2_decode -f python bench.py --decode --workload c2 --layers 2 --steps 4 --warmup 1 --graph 0 > gpurun_out/r2_ncu_decode.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_decode.ncu-rep > gpurun_out/r2_decode.md
grep -n "launch [0-9]\|duration\|DRAM read (\|DRAM throughput" gpurun_out/r2_decode.md
