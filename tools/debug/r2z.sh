#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), d['roofline']['kernel'], 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 200 $B 2>> gpurun_out/r2z_err.log | ext "base"
timeout 200 $B --workload C5 2>> gpurun_out/r2z_err.log | ext "base C5"
for V in "$@"; do
  export BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_$V.so
  timeout 300 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -2
  timeout 200 $B 2>> gpurun_out/r2z_err.log | ext "$V"
  timeout 200 $B --population 5000 2>> gpurun_out/r2z_err.log | ext "$V P5000"
  timeout 200 $B --workload C5 2>> gpurun_out/r2z_err.log | ext "$V C5"
done
tail -n 3 gpurun_out/r2z_err.log
