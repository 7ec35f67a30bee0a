"""Shared test helpers: build oracle objects (oracle/ is the checker) from a synthetic Workload."""
import numpy as np
import torch

import oracle
from blackbox_mpc_b200.utils import workloads


def oracle_spaces(w):
    return oracle.Space(w.lb, w.ub), oracle.Space(-np.ones(w.dS, np.float32), np.ones(w.dS, np.float32))


def oracle_evaluator(w, dtype=torch.float32):
    if w.dynamics == "pendulum_true":
        handler = oracle.Handler(oracle.PendulumTrueModel(), true_model=True, dtype=dtype)
    else:
        members = [oracle.MLP([torch.from_numpy(x) for x in ws], [torch.from_numpy(x) for x in bs], w.activations)
                   for ws, bs in zip(w.weights, w.biases)]
        fn = members[0] if len(members) == 1 else oracle.Ensemble(members)
        handler = oracle.Handler(fn, true_model=False, is_normalized=True, stats=w.stats, dtype=dtype)
    reward = oracle.pendulum_reward_function if w.reward == "pendulum" else oracle.halfcheetah_reward_function
    return oracle.Evaluator(reward, handler)


ORACLE_OPT = {"CEM": oracle.CEM, "PI2": oracle.PI2, "RandomSearch": oracle.RandomSearch, "PSO": oracle.PSO,
              "SPSA": oracle.SPSA, "CMA-ES": oracle.CMAES}


def oracle_optimizer(w, name=None, dtype=torch.float32, **extra):
    name = name or w.optimizer_name
    a_sp, o_sp = oracle_spaces(w)
    args = dict(w.optimizer_args) if name == w.optimizer_name else {}
    args.update(planning_horizon=w.planning_horizon, population_size=w.population_size, num_agents=w.num_agents)
    if name != "RandomSearch":
        args["max_iterations"] = w.max_iterations or 5
    args.update(extra)
    opt = ORACLE_OPT[name](a_sp, o_sp, dtype=dtype, **args)
    opt.set_trajectory_evaluator(oracle_evaluator(w, dtype))
    return opt


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
