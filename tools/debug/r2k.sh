#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r2k_tests.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4), d['roofline']['kernel'], 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])"; }
timeout 300 $B 2>> gpurun_out/r2k_err.log | ext "P10000" | tee -a gpurun_out/r2k_ab.log
timeout 300 $B --population 1250 2>> gpurun_out/r2k_err.log | ext "P1250" | tee -a gpurun_out/r2k_ab.log
tail -n 5 gpurun_out/r2k_err.log
