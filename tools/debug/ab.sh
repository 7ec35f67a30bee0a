#!/bin/bash
# A/B timing of rollout variants selected by environment knobs.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'steps/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
$B | ext groups
BBMPC_NO_GROUPS=1 $B | ext nogroups
$B --population 1250 | ext groups_P1250
BBMPC_NO_GROUPS=1 $B --population 1250 | ext nogroups_P1250
$B --population 5000 | ext groups_P5000
BBMPC_NO_GROUPS=1 $B --population 5000 | ext nogroups_P5000
