#!/bin/bash
# A/B timing of rollout variants selected by environment knobs.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'steps/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
$B | ext nsplit
BBMPC_NO_NSPLIT=1 $B | ext nosplit
BBMPC_NO_EARLY_L0=1 $B | ext nsplit_noearly
BBMPC_NO_NSPLIT=1 BBMPC_NO_EARLY_L0=1 $B | ext nosplit_noearly
