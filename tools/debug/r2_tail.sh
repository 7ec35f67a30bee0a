#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3))"; }
timeout 200 $B 2>> gpurun_out/r2tail_err.log | ext "C4"
timeout 200 $B --workload C2 2>> gpurun_out/r2tail_err.log | ext "C2"
timeout 200 $B --workload C3 2>> gpurun_out/r2tail_err.log | ext "C3"
BBMPC_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:step_mlp -c 4 --csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | grep step_mlp | awk -F, '{print $5, $NF}' | head -4
tail -3 gpurun_out/r2tail_err.log
