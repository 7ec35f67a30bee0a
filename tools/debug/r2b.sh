#!/bin/bash
# debugging run of the pipelined kernel: watchdog records, sanitizer, reduced cases
mkdir -p gpurun_out
cat > /tmp/pipe_case.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers
from blackbox_mpc_b200.utils import workloads
name, P, H = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
w = workloads.make(name, population_size=P, planning_horizon=H, bias_scale=0.1)
policy = workloads.build_policy(w, precision="bf16x3")
ev = policy._trajectory_evaluator
actions = helpers.random_actions(w, P, seed=21)
got = ev(torch.from_numpy(w.state), actions, 0).cpu().numpy()
os.environ["BBMPC_TC_PIPE"] = "0"
old = ev(torch.from_numpy(w.state), actions, 0).cpu().numpy()
print(name, P, H, "finite", np.isfinite(got).all(), "equal", np.array_equal(got, old), "maxdiff", float(np.abs(got - old).max()))
PY
for c in "C3 128 1" "C3 128 2" "C4 128 1" "C4 128 3" "C3 300 30" "C4 700 30" "C4 3000 4" "C4 10000 30"; do
  echo "== $c"; BBMPC_DEBUG=1 timeout 120 python /tmp/pipe_case.py $c 2>&1 | grep -v Warning | tail -12
done > gpurun_out/r2b_cases.log 2>&1
cat gpurun_out/r2b_cases.log
echo "== sanitizer memcheck C3 128 2"
timeout 300 compute-sanitizer --tool memcheck python /tmp/pipe_case.py C3 128 2 2>&1 | tail -30 > gpurun_out/r2b_memcheck.log
cat gpurun_out/r2b_memcheck.log
