#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optimizers.py tests/test_gpu_golden.py tests/test_gpu_cmaes.py tests/test_gpu_engine.py -x -q 2>&1 | tail -3
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline"
ext() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$1', 'steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms_avg'],4))"; }
timeout 200 $B 2>> gpurun_out/r2topk_err.log | ext "C4"
timeout 200 $B --workload C2 2>> gpurun_out/r2topk_err.log | ext "C2"
timeout 200 $B --workload C5 --steps 20 2>> gpurun_out/r2topk_err.log | ext "C5"
tail -3 gpurun_out/r2topk_err.log
