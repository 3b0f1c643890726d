#!/bin/bash
mkdir -p gpurun_out
for impl in grid persistent; do
VEX_K9_IMPL=$impl timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.avg.per_second \
  -k regex:'k9_attn_bwd' --launch-skip 4 -c 4 --csv --log-file gpurun_out/r2_k9_ncu_$impl.csv \
  python bench.py --train --workload c2 --layers 1 --steps 2 --warmup 2 > gpurun_out/r2_k9_ncu_$impl.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_k9_ncu_$impl.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; c={n:i for i,n in enumerate(h)}
per={}
for r in rows[hi+1:]:
    if len(r)!=len(h): continue
    per.setdefault((r[c["ID"]], r[c["Kernel Name"]].split("(")[0]),{})[r[c["Metric Name"]]]=r[c["Metric Value"]]+" "+r[c["Metric Unit"]]
for k,m in per.items(): print("$impl", k[1], m)
PY
done
