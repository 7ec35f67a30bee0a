#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -3
timeout 200 $B 2>> gpurun_out/r2g_err.log | ext "pipe_P10000" | tee -a gpurun_out/r2g_ab.log
for ST in 0 350 1000; do BBMPC_PIPE_STAGGER=$ST timeout 200 $B 2>> gpurun_out/r2g_err.log | ext "pipe_stagger$ST" | tee -a gpurun_out/r2g_ab.log; done
BBMPC_PIPE_AUNITS=7 timeout 200 $B 2>> gpurun_out/r2g_err.log | ext "pipe_AU7" | tee -a gpurun_out/r2g_ab.log
BBMPC_PIPE_MT=2 timeout 200 $B 2>> gpurun_out/r2g_err.log | ext "pipe_MT2" | tee -a gpurun_out/r2g_ab.log
timeout 200 $B --population 1250 2>> gpurun_out/r2g_err.log | ext "pipe_P1250" | tee -a gpurun_out/r2g_ab.log
BBMPC_PIPE_STAGGER=0 timeout 200 $B --population 1250 2>> gpurun_out/r2g_err.log | ext "pipe_P1250_stagger0" | tee -a gpurun_out/r2g_ab.log
tail -n 3 gpurun_out/r2g_err.log
