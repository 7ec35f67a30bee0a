#!/bin/bash
mkdir -p gpurun_out
BBMPC_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r2_launches_c5.csv python bench.py --workload C5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_c5.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:70]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[iv].replace(',',''))
for k,(n,t) in agg.items(): print(k.ljust(70), n, round(t/n/1e3,1), 'us avg')
PY
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool memcheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck.log
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool racecheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2_sanitizer_racecheck.log
BBMPC_TC_PIPE=0 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool memcheck python tools/debug/san_act.py > gpurun_out/r2_sanitizer_memcheck_tc.log 2>&1; echo "memcheck(tc) rc=$?"; tail -3 gpurun_out/r2_sanitizer_memcheck_tc.log
