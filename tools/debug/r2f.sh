#!/bin/bash
mkdir -p gpurun_out
BBMPC_TC_TRACE=gpurun_out/r2f_trace.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/r2f_err.log
BBMPC_PIPE_AUNITS=7 BBMPC_TC_TRACE=gpurun_out/r2f_trace_au7.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/r2f_err.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
BBMPC_PIPE_AUNITS=7 timeout 200 $B 2>> gpurun_out/r2f_err.log | ext "pipe_AU7" | tee -a gpurun_out/r2f_ab.log
BBMPC_PIPE_AUNITS=7 timeout 200 $B --population 1250 2>> gpurun_out/r2f_err.log | ext "pipe_AU7_P1250" | tee -a gpurun_out/r2f_ab.log
tail -n 3 gpurun_out/r2f_err.log
