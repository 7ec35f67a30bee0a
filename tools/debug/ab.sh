#!/bin/bash
# A/B timing of rollout variants selected by environment knobs.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
$B | ext stagger250
BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_s0.so $B | ext stagger0
BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_s500.so $B | ext stagger500
$B --population 1250 | ext stagger250_P1250
BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_s0.so $B --population 1250 | ext stagger0_P1250
BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_s500.so $B --population 1250 | ext stagger500_P1250
