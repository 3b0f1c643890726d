#!/bin/bash
# end-of-round single-GPU record: smoke, every bench line that changed in the second half of round 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
B="timeout 900 python bench.py"
$B --steps 10 --warmup 3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; tail -2 gpurun_out/r2_bench_c3_n1.err
$B --workload c2 --steps 20 --warmup 5 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; tail -2 gpurun_out/r2_bench_c2.err
$B --workload c2 --lora 64 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_c2_lora64.json 2> gpurun_out/r2_bench_c2_lora64.err; tail -2 gpurun_out/r2_bench_c2_lora64.err
$B --workload c4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; tail -2 gpurun_out/r2_bench_c4_n1.err
$B --workload c2 --layers 32 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_stack32_c2.json 2> gpurun_out/r2_bench_stack32_c2.err; tail -2 gpurun_out/r2_bench_stack32_c2.err
$B --decode --workload c2 --layers 32 --steps 30 --warmup 5 > gpurun_out/r2_bench_decode.json 2> gpurun_out/r2_bench_decode.err; tail -2 gpurun_out/r2_bench_decode.err
$B --impl reference --steps 5 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -2 gpurun_out/r2_bench_ref.err
python tools/show_bench.py gpurun_out/r2_bench_c3_n1.json gpurun_out/r2_bench_c2.json gpurun_out/r2_bench_c2_lora64.json gpurun_out/r2_bench_c4_n1.json gpurun_out/r2_bench_stack32_c2.json gpurun_out/r2_bench_decode.json gpurun_out/r2_bench_ref.json
