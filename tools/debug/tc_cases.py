"""Debug harness: runs the tensor-core rollout on a few (members, H, P) shapes against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers
from blackbox_mpc_b200.utils import workloads

def run(nm, H, P, prec="bf16x3"):
    w = workloads.make("C4", population_size=P, planning_horizon=H, bias_scale=0.1)
    w.weights, w.biases = w.weights[:nm], w.biases[:nm]
    pol = workloads.build_policy(w, precision=prec)
    ev = pol._trajectory_evaluator
    actions = helpers.random_actions(w, P, seed=1)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0)
    torch.cuda.synchronize()
    ref = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy()
    err = np.abs(got.cpu().numpy() - ref).max()
    print(f"members={nm} H={H} P={P} {prec}: max abs err {err:.3e}", flush=True)

if __name__ == "__main__":
    nm, H, P = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    run(nm, H, P, sys.argv[4] if len(sys.argv) > 4 else "bf16x3")
