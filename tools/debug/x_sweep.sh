#!/bin/bash
# Timing experiments on the tensor-core rollout (BBMPC_TC_X knobs produce garbage results on purpose).
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('$1', 'steps/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
$B | ext base
BBMPC_TC_X=1 $B | ext noload
BBMPC_TC_X=2 $B | ext noact
BBMPC_TC_X=3 $B | ext noload+noact
$B --precision bf16 | ext bf16
BBMPC_TC_X=1 $B --precision bf16 | ext bf16+noload
BBMPC_TC_X=3 $B --precision bf16 | ext bf16+noload+noact
