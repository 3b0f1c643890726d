#!/bin/bash
# round 2 ncu evidence: launch list of the default bench command + --set full captures of the kernels that changed
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (1) launch list of the default command (c3: 32 layers, 64 samples on one GPU); numbers printed under ncu are not bench values
# VEX_PROFILER_RANGE=1 brackets the timed region with cudaProfilerStart/Stop, so the list holds exactly its launches (the
# kernel nodes of the replayed CUDA graph), not the weight initialisation
VEX_PROFILER_RANGE=1 timeout 1500 $NCU --profile-from-start off --metrics gpu__time_duration.sum -c 4000 --csv \
  --log-file gpurun_out/r2_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_c3.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_c3.csv
# (2) the four grouped GEMMs + norm + attention of ONE layer at the c3 N=1 shape (64 samples): DRAM traffic per launch
timeout 900 $NCU --set full --import-source on -k regex:'k3_grouped_gemm_pair|k2_rmsnorm|k4_attention_tc3' --launch-skip 14 -c 7 \
  -o gpurun_out/r2_c3_layer -f python bench.py --workload c3 --layers 1 --steps 1 --warmup 1 --no-cpu --graph 0 > gpurun_out/r2_ncu_c3_layer.log 2>&1
echo "c3 layer rc=$?"
# (3) decode step: K12 (4 modes) + cluster split-KV attention + norm
timeout 900 $NCU --set full --import-source on -k regex:'k12_decode_gemm|k4_attention_decode' --launch-skip 10 -c 5 \
  -o gpurun_out/r2_decode -f python bench.py --decode --workload c2 --layers 2 --steps 4 --warmup 1 --graph 0 > gpurun_out/r2_ncu_decode.log 2>&1
echo "decode rc=$?"
for f in r2_c3_layer r2_decode; do
  python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.md 2>/dev/null; head -c 600 gpurun_out/$f.md
done
ls -la gpurun_out/*.ncu-rep
