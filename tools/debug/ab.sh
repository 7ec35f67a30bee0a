#!/bin/bash
# A/B timing of rollout variants selected by environment knobs / alternative builds (BBMPC_LIB).
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],3))"; }
for V in "" 100 200; do
  L=""; [ -n "$V" ] && L=$PWD/blackbox_mpc_b200/libbbmpc_s$V.so
  BBMPC_LIB=$L $B | ext "stagger${V:-150}"
  BBMPC_LIB=$L $B --population 1250 | ext "stagger${V:-150}_P1250"
done
for V in 100 200; do BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_s$V.so timeout 200 python -m pytest tests/test_gpu_rollout.py -x -q 2>&1 | tail -1; done
