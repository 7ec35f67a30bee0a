"""Prints return-error statistics of the tensor-core rollout against the float64 oracle (debug)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from blackbox_mpc_b200.utils import workloads

name, P, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3]
w = workloads.make(name, population_size=P, bias_scale=0.1)
pol = workloads.build_policy(w, precision=prec)
ev = pol._trajectory_evaluator
actions = helpers.random_actions(w, P, seed=1)
state = torch.from_numpy(w.state)
got = ev(state, actions, 0).cpu().numpy().ravel()
ref = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy().ravel()
d = np.abs(got - ref)
tol = 2e-2 + 2e-4 * np.abs(ref)
print(f"{name} P={P} {prec} lib={os.environ.get('BBMPC_LIB','default')[-16:]}: max {d.max():.4f} p99 {np.percentile(d,99):.4f} median {np.median(d):.5f} "
      f"n_bad {(d>tol).sum()} |ref| median {np.median(np.abs(ref)):.1f}")
