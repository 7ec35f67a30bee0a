#!/bin/bash
# A/B timing of rollout variants selected by environment knobs.
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3), 'share', round(d['roofline']['kernel_share_of_step'],3))"; }
$B | ext C4
$B --population 1250 | ext C4_P1250
