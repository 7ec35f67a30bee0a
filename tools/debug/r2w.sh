#!/bin/bash
mkdir -p gpurun_out
for P in 5000 10000; do
BBMPC_TC_PIPE=1 BBMPC_DEBUG=1 PP=$P timeout 120 python - > gpurun_out/r2w_dbg_$P.log 2>&1 <<'PY'
import torch, numpy as np, sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import helpers
from blackbox_mpc_b200.utils import workloads
P = int(os.environ["PP"])
w = workloads.make("C4", population_size=P, planning_horizon=4, bias_scale=0.1)
p = workloads.build_policy(w)
ev = p._trajectory_evaluator
a = helpers.random_actions(w, P, seed=1)
r = ev(torch.from_numpy(w.state), a, 0)
torch.cuda.synchronize()
print("ok", ev.engine().last_rollout_kernel, r[:3].cpu().numpy().ravel())
PY
echo "P=$P"; tail -12 gpurun_out/r2w_dbg_$P.log | cut -c1-300
done
