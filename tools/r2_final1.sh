#!/bin/bash
# round 2 single-GPU record: tests, every bench line, ncu evidence
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
B="timeout 900 python bench.py"
$B --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; tail -2 gpurun_out/r2_bench_c3_n1.err
$B --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -2 gpurun_out/r2_bench_c2.err
$B --workload c2 --lora 64 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_lora64.json 2> gpurun_out/r2_bench_c2_lora64.err; tail -2 gpurun_out/r2_bench_c2_lora64.err
$B --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; tail -2 gpurun_out/r2_bench_c4_n1.err
$B --workload c2 --layers 32 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_stack32_c2.json 2> gpurun_out/r2_bench_stack32_c2.err; tail -2 gpurun_out/r2_bench_stack32_c2.err
$B --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode.json 2> gpurun_out/r2_bench_decode.err; tail -2 gpurun_out/r2_bench_decode.err
$B --train --workload c2 --layers 32 --steps 4 --warmup 2 > gpurun_out/r2_bench_train32_n1.json 2> gpurun_out/r2_bench_train32_n1.err; tail -2 gpurun_out/r2_bench_train32_n1.err
$B --train --workload c2 --layers 32 --steps 4 --warmup 2 --recompute 0 > gpurun_out/r2_bench_train32_keep_n1.json 2> gpurun_out/r2_bench_train32_keep_n1.err; tail -2 gpurun_out/r2_bench_train32_keep_n1.err
$B --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -2 gpurun_out/r2_bench_ref.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n1.json gpurun_out/r2_bench_c2.json gpurun_out/r2_bench_c2_lora64.json gpurun_out/r2_bench_c4_n1.json gpurun_out/r2_bench_stack32_c2.json gpurun_out/r2_bench_decode.json gpurun_out/r2_bench_ref.json
python - <<PY
import json
for f in ("gpurun_out/r2_bench_train32_n1.json", "gpurun_out/r2_bench_train32_keep_n1.json"):
    try:
        d=json.load(open(f)); print(f, {k: d.get(k) for k in ("value","ms_per_step","step_frac_of_bf16_peak","step_frac_of_bf16_sustained","peak_mem_gb")}, d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
bash tools/r2_ncu.sh
