#!/bin/bash
# DRAM bytes of the four grouped GEMMs of one c3 layer at N = 1 (64 samples) as a function of the raster group height
mkdir -p gpurun_out
for rg in 4 8 12 16 24 40; do
  VEX_GEMM_RASTER=$rg timeout 600 ncu --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
    -k regex:k3_grouped_gemm_pair --launch-skip 8 -c 4 --csv --log-file gpurun_out/r2_raster_traffic_$rg.csv \
    python bench.py --workload c3 --layers 1 --steps 1 --warmup 1 --no-cpu --graph 0 > gpurun_out/r2_raster_traffic_$rg.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r2_raster_traffic_$rg.csv")))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]; c={n:i for i,n in enumerate(h)}
per={}
for r in rows[hi+1:]:
    if len(r)!=len(h): continue
    per.setdefault(r[c["ID"]],{})[r[c["Metric Name"]]]=(float(r[c["Metric Value"]].replace(",","")), r[c["Metric Unit"]])
def b(v,u): return v*{"Gbyte":1e9,"Mbyte":1e6,"Kbyte":1e3,"byte":1}.get(u,1)
def t(v,u): return v*{"ms":1e-3,"us":1e-6,"ns":1e-9,"s":1}.get(u,1)
out=[]
for k,m in per.items():
    out.append("%.2f GB / %.2f ms / hit %.0f%%" % ((b(*m["dram__bytes_read.sum"])+b(*m["dram__bytes_write.sum"]))/1e9, t(*m["gpu__time_duration.sum"])*1e3, m["lts__t_sector_hit_rate.pct"][0]))
print("raster $rg:", " | ".join(out))
PY
done
