"""Generates tests/golden/*.npz from the CPU restatement (oracle/, float64 "truth" twin) with every
random draw injected from numpy.random.default_rng(seed).  PARITY UNPINNED by the reference (it
has no tests and TF 2.0 cannot be installed here): these vectors pin the restatement against
regressions and give the CUDA path fixed inputs/outputs that travel to the GPU box.

    python tests/golden/make_golden.py          # rewrites the fixtures (small: < 300 KB in total)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import build as ob  # noqa: E402
from blackbox_mpc_b200.utils import workloads  # noqa: E402

F64 = torch.float64


def uniform_actions(w, P, rng):
    lb, ub = w.lb.astype(np.float64), w.ub.astype(np.float64)
    return (lb + (ub - lb) * rng.random((P, w.num_agents, w.planning_horizon, w.dU))).astype(np.float32)


def std_truncnorm(shape, rng):
    z = rng.standard_normal(shape)
    bad = np.abs(z) > 2
    while bad.any():
        z[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(z) > 2
    return z


def rollout_case(name, P, A, H, seed):
    rng = np.random.default_rng(seed)
    w = workloads.make(name, population_size=P, planning_horizon=H, num_agents=A, bias_scale=0.1)
    actions = uniform_actions(w, P, rng)
    ev = ob.evaluator(w, F64)
    state = torch.from_numpy(w.state).double()
    returns = ev(state, torch.from_numpy(actions).double(), 0).numpy()
    nxt = ev.predict_next_state(state, torch.from_numpy(actions[0, :, 0]).double())
    rew = ev.evaluate_next_reward(state, nxt, torch.from_numpy(actions[0, :, 0]).double())
    return dict(actions=actions, returns=returns, next_state=nxt.numpy(), reward=rew.numpy())


def rollout_full_case(name, P, seed):
    """BASELINE-size rollout: the actions come from the seeded torch generator the GPU tests use
    (tests/helpers.random_actions), so only the float64 returns and a checksum of the actions are stored."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    w = workloads.make(name, population_size=P, bias_scale=0.1)
    actions = helpers.random_actions(w, P, seed=seed)
    ev = ob.evaluator(w, F64)
    returns = ev(torch.from_numpy(w.state).double(), actions.double(), 0).numpy()
    return dict(returns=returns, actions_checksum=np.array([actions.double().sum().item(), actions.double().abs().sum().item()]),
                seed=np.array([seed]))


class RecordingDraws:
    """numpy-rng draws with TF sampler semantics; records the *standard* variates so a test can
    re-inject them."""

    def __init__(self, rng):
        self.rng, self.rec = rng, {}

    def _keep(self, tag, v):
        v32 = v.numpy().astype(np.float32)   # fixtures hold fp32 samples: round BEFORE use so that
        self.rec.setdefault(tag, []).append(v32)  # re-injecting the stored draws reproduces the run
        return torch.from_numpy(v32).to(v.dtype)

    # The STANDARD variate behind every affine draw is rounded to fp32 first and recorded as "std.<tag>": the GPU
    # tests inject exactly these into the CUDA samplers (bbmpc_opt_set_draw_injection), tests/test_gpu_golden.py.
    def _std(self, tag, v):
        v32 = np.asarray(v).astype(np.float32)
        self.rec.setdefault("std." + tag, []).append(v32)
        return torch.from_numpy(v32.astype(np.float64))

    def truncated_normal(self, shape, mean, std, tag=""):
        z = self._std(tag, std_truncnorm(tuple(shape), self.rng))
        return self._keep(tag, mean + std * z.to(mean.dtype))

    def uniform(self, shape, lo, hi, tag=""):
        u = np.minimum(self.rng.random(tuple(shape)), 1.0 - 2.0 ** -24)   # stays < 1 after the fp32 rounding
        return self._keep(tag, lo + (hi - lo) * self._std(tag, u).to(lo.dtype))

    def normal(self, shape, tag=""):
        return self._keep(tag, torch.from_numpy(np.asarray(self.rng.standard_normal(tuple(shape)))).to(F64))

    def rademacher(self, shape, tag=""):
        return self._keep(tag, torch.from_numpy(self.rng.integers(0, 2, tuple(shape)) * 2.0 - 1.0).to(F64))


def optimizer_case(wname, opt_name, P, A, H, seed, **extra):
    rng = np.random.default_rng(seed)
    w = workloads.make(wname, population_size=P, planning_horizon=H, num_agents=A, bias_scale=0.1)
    if opt_name == "CEM":
        w.optimizer_args = dict(num_elite=16, alpha=0.25)
    opt = ob.optimizer(w, opt_name, dtype=F64, **extra)
    draws = RecordingDraws(rng)
    state = torch.from_numpy(w.state).double()
    out = {}
    for call in range(2):  # two act() calls: exercises the warm start / no-warm-start behaviour
        a, n, r = opt(state, call, False, draws)
        out[f"action{call}"], out[f"next{call}"], out[f"reward{call}"] = a.numpy(), n.numpy(), r.numpy()
    for tag, lst in draws.rec.items():
        out["draws." + tag] = np.stack(lst)
    return out


def main():
    cases = {
        "rollout_c1": rollout_case("C1", 64, 1, 30, 1),
        "rollout_c2": rollout_case("C2", 48, 2, 30, 2),
        "rollout_c3": rollout_case("C3", 32, 1, 30, 3),
        "rollout_c4": rollout_case("C4", 24, 1, 30, 4),
        "rollout_c4_full": rollout_full_case("C4", 10000, 31),
        "rollout_c3_full": rollout_full_case("C3", 5000, 32),
        "opt_cem_c2": optimizer_case("C2", "CEM", 64, 2, 12, 5, max_iterations=3),
        "opt_pi2_c2": optimizer_case("C2", "PI2", 64, 2, 12, 6, max_iterations=3),
        "opt_rs_c1": optimizer_case("C1", "RandomSearch", 64, 2, 12, 7),
        "opt_spsa_c2": optimizer_case("C2", "SPSA", 32, 1, 12, 8, max_iterations=3),
        "opt_cmaes_c2": optimizer_case("C2", "CMA-ES", 48, 1, 8, 9, max_iterations=3, num_elite=12),
    }
    for name, arrs in cases.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
        print(name, {k: v.shape for k, v in arrs.items()})


if __name__ == "__main__":
    main()
