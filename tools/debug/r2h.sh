#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 200 $B 2>> gpurun_out/r2h_err.log | ext "pipe_P10000" | tee -a gpurun_out/r2h_ab.log
BBMPC_TC_X=32 timeout 200 $B 2>> gpurun_out/r2h_err.log | ext "pipe_whole" | tee -a gpurun_out/r2h_ab.log
timeout 200 $B --population 1250 2>> gpurun_out/r2h_err.log | ext "pipe_P1250" | tee -a gpurun_out/r2h_ab.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_pipe -s 3 -c 1 -f -o gpurun_out/prof_pipe2_p10000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2h_ncu.log 2>&1
tail -n 3 gpurun_out/r2h_err.log
