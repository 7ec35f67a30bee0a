"""One small C4 act() for compute-sanitizer (member-parallel pipelined rollout when BBMPC_TC_PIPE=1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from blackbox_mpc_b200.utils import workloads
w = workloads.make("C4", population_size=600, planning_horizon=6, bias_scale=0.1)
w.max_iterations = 2
p = workloads.build_policy(w)
for t in range(2):
    a, n, r = p.act(w.state[0], t)
print("rollout kernel:", p._optimizer._engine.last_rollout_kernel, "action:", a)
