#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size \
  -k regex:'k8_lora_wgrad|k3_grouped_gemm<64' --launch-skip 30 -c 25 --csv --log-file gpurun_out/r2_k8_ncu.csv \
  python bench.py --train --workload c2 --layers 1 --steps 2 --warmup 2 > gpurun_out/r2_k8_ncu.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_k8_ncu.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; c={n:i for i,n in enumerate(h)}
per={}
for r in rows[hi+1:]:
    if len(r)!=len(h): continue
    per.setdefault((r[c["ID"]], r[c["Kernel Name"]].split("(")[0][:40], r[c["Grid Size"]]),{})[r[c["Metric Name"]]]=r[c["Metric Value"]]+r[c["Metric Unit"]]
for k,m in per.items():
    print(k[1], k[2], m.get("gpu__time_duration.sum"), "rd", m.get("dram__bytes_read.sum"), "wr", m.get("dram__bytes_write.sum"), "dram%", m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), "hit", m.get("lts__t_sector_hit_rate.pct"), "warps", m.get("sm__warps_active.avg.pct_of_peak_sustained_active"))
PY
