"""Shared test helpers: build oracle objects (oracle/ is the checker) from a synthetic Workload."""
import numpy as np
import torch

import oracle
from blackbox_mpc_b200.utils import workloads


from oracle import build as _ob

oracle_spaces, oracle_evaluator, oracle_optimizer, ORACLE_OPT = _ob.spaces, _ob.evaluator, _ob.optimizer, _ob.OPTIMIZERS


def random_actions(w, P, seed=0):
    g = torch.Generator().manual_seed(seed)
    lb, ub = torch.from_numpy(w.lb), torch.from_numpy(w.ub)
    return lb + (ub - lb) * torch.rand(P, w.num_agents, w.planning_horizon, w.dU, generator=g)


def compare_returns(got, ref, atol, rtol, jump=10.0, max_jump_frac=0.0):
    """|got-ref| <= atol + rtol|ref| row-wise; rows that differ by a multiple of `jump` (a reward
    threshold that flipped under rounding noise, tutorials/mujoco/cost_func.py:9-17) are tolerated up
    to max_jump_frac of the rows.  Returns (n_bad, n_jump)."""
    got, ref = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    d = np.abs(got - ref)
    tol = atol + rtol * np.abs(ref)
    ok = d <= tol
    k = np.round((got - ref) / jump)
    jumped = (~ok) & (k != 0) & (np.abs(got - ref - k * jump) <= tol)
    n_bad = int((~ok & ~jumped).sum())
    n_jump = int(jumped.sum())
    assert n_bad == 0, f"{n_bad} rows out of tolerance (max err {d[~ok & ~jumped].max() if n_bad else 0})"
    assert n_jump <= max_jump_frac * got.size, f"{n_jump} threshold flips out of {got.size} rows"
    return n_bad, n_jump
