#!/bin/bash
# end-of-round ncu evidence for the kernels changed late in round 2: K4 (c2 shape, merged tail waves) and K8 (coalesced dB)
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:'k4_attention_tc3' --launch-skip 3 -c 1 \
  -o gpurun_out/r2_k4_final -f python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --graph 0 > gpurun_out/r2_ncu_k4_final.log 2>&1
echo "k4 rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:'k8_lora_wgrad' --launch-skip 30 -c 4 \
  -o gpurun_out/r2_k8_final -f python bench.py --train --workload c2 --layers 1 --steps 2 --warmup 2 > gpurun_out/r2_ncu_k8_final.log 2>&1
echo "k8 rc=$?"
for f in r2_k4_final r2_k8_final; do
  python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.md 2>/dev/null; grep -n "launch [0-9]\|duration\|tensor pipe\|DRAM throughput\|DRAM read (" gpurun_out/$f.md | head -24
done
