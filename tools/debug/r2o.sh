#!/bin/bash
mkdir -p gpurun_out
BBMPC_NO_GRAPH=1 BBMPC_TC_TRACE=gpurun_out/r2o_trace.txt timeout 200 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/r2o_err.log
tail -n 3 gpurun_out/r2o_err.log
