#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
for V in "" iv1 iv2; do
  L=""; [ -n "$V" ] && L=$PWD/blackbox_mpc_b200/libbbmpc_$V.so
  BBMPC_LIB=$L timeout 200 $B 2>> gpurun_out/r2i_err.log | ext "variant${V:-0}_P10000" | tee -a gpurun_out/r2i_ab.log
  BBMPC_LIB=$L timeout 200 $B --population 1250 2>> gpurun_out/r2i_err.log | ext "variant${V:-0}_P1250" | tee -a gpurun_out/r2i_ab.log
done
tail -n 3 gpurun_out/r2i_err.log
