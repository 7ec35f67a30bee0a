#!/bin/bash
# A/B timing of rollout variants selected by environment knobs.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
$B | ext base
$B --population 1250 | ext base_P1250
$B --population 5000 | ext base_P5000
BBMPC_NO_GROUPS=1 $B | ext nogroups
