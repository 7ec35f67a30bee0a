#!/bin/bash
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
echo "== memcheck C3 128 1"
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/pipe_case.py C3 128 1 2>&1 | grep -v "Host Frame" | head -60 > gpurun_out/r2c_memcheck.log
cat gpurun_out/r2c_memcheck.log
for c in "C3 128 1" "C3 300 30" "C4 700 30" "C4 10000 30"; do
  echo "== nosmr $c"; BBMPC_LIB=$PWD/blackbox_mpc_b200/libbbmpc_nosmr.so BBMPC_DEBUG=1 timeout 120 python /tmp/pipe_case.py $c 2>&1 | grep -v Warning | tail -4
done 2>&1 | tee gpurun_out/r2c_nosmr.log
