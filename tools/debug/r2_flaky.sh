#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -1; done
BBMPC_TC_PIPE=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_golden.py tests/test_gpu_user_reward.py -x -q 2>&1 | tail -1
BBMPC_TC_PIPE=1 BBMPC_NO_GRAPH=1 timeout 500 compute-sanitizer --tool synccheck python tools/debug/san_act.py 2>&1 | tail -2
