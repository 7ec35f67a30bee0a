#!/bin/bash
# First hardware run of the pipelined rollout kernel: parity tests, then A/B timing against the single-tile kernel.
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -25 > gpurun_out/r2a_pipe_tests.log
cat gpurun_out/r2a_pipe_tests.log
for P in 10000 5000 2500 1250; do
  timeout 200 $B --population $P 2> gpurun_out/r2a_err_pipe_$P.log | ext "pipe_P$P" | tee -a gpurun_out/r2a_ab.log
  BBMPC_TC_PIPE=0 timeout 200 $B --population $P 2> gpurun_out/r2a_err_old_$P.log | ext "old_P$P" | tee -a gpurun_out/r2a_ab.log
done
for MT in 1 2; do
  BBMPC_PIPE_MT=$MT timeout 200 $B 2>> gpurun_out/r2a_err.log | ext "pipe_MT$MT" | tee -a gpurun_out/r2a_ab.log
done
timeout 200 $B --workload C3 2>> gpurun_out/r2a_err.log | ext "pipe_C3" | tee -a gpurun_out/r2a_ab.log
BBMPC_TC_PIPE=0 timeout 200 $B --workload C3 2>> gpurun_out/r2a_err.log | ext "old_C3" | tee -a gpurun_out/r2a_ab.log
BBMPC_TC_TRACE=gpurun_out/r2a_trace.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/r2a_err.log
BBMPC_TC_TRACE=gpurun_out/r2a_trace_p1250.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --population 1250 > /dev/null 2>> gpurun_out/r2a_err.log
tail -5 gpurun_out/r2a_err*.log | tail -40
