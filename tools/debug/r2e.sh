#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_pipe -s 3 -c 1 -f -o gpurun_out/prof_pipe_p1250 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --population 1250 > gpurun_out/r2e_ncu1.log 2>&1
tail -3 gpurun_out/r2e_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_pipe -s 3 -c 1 -f -o gpurun_out/prof_pipe_p10000 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_ncu2.log 2>&1
tail -3 gpurun_out/r2e_ncu2.log
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_user_reward.py tests/test_gpu_golden.py tests/test_gpu_optimizers.py -q 2>&1 | tail -40 > gpurun_out/r2e_tests.log
cat gpurun_out/r2e_tests.log
