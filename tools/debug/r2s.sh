#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['roofline']['kernel'], 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 200 $B 2>> gpurun_out/r2s_err.log | ext "base(stagger700)"
for S in 0 300 1200; do BBMPC_PIPE_STAGGER=$S timeout 200 $B 2>> gpurun_out/r2s_err.log | ext "stagger$S"; done
timeout 200 $B 2>> gpurun_out/r2s_err.log | ext "base again"
tail -n 3 gpurun_out/r2s_err.log
