#!/bin/bash
mkdir -p gpurun_out
BBMPC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_ncu.log 2>&1
tail -2 gpurun_out/r2l_ncu.log | cut -c1-300
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches.csv')) if len(r)>5]
hdr=rows[0]; 
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); iu=hdr.index('Metric Unit')
for r in rows[1:61]:
    print(r[ik][:60].ljust(60), r[iv], r[iu])
PY
