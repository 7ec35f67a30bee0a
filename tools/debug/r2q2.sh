#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 4 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['roofline']['kernel'], 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4))"; }
timeout 200 $B 2>> gpurun_out/r2q2_err.log | ext "base"
timeout 200 $B --precision bf16 2>> gpurun_out/r2q2_err.log | ext "bf16x1"
