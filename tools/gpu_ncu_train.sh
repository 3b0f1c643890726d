#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --train --layers 1 --steps 1 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k9|k8|k7" -c 60 --csv --log-file gpurun_out/launches_train.csv $B > gpurun_out/ncu_train_list.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_train.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); 
for r in rows[1:40]: print(r[ki][:60], r[vi])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k9_attn_bwd -s 2 -c 2 -o gpurun_out/prof_attn_bwd_r1 -f $B > gpurun_out/ncu_full_attn_bwd.log 2>&1
ls -la gpurun_out | tail -5
