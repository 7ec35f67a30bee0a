#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],4))"; }
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -5
for P in 10000 5000 1250; do
  timeout 200 $B --population $P 2>> gpurun_out/r2d_err.log | ext "pipe_P$P" | tee -a gpurun_out/r2d_ab.log
done
timeout 200 $B --workload C3 2>> gpurun_out/r2d_err.log | ext "pipe_C3" | tee -a gpurun_out/r2d_ab.log
BBMPC_TC_TRACE=gpurun_out/r2d_trace.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/r2d_err.log
BBMPC_TC_TRACE=gpurun_out/r2d_trace_p1250.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --population 1250 > /dev/null 2>> gpurun_out/r2d_err.log
tail -n 5 gpurun_out/r2d_err.log
