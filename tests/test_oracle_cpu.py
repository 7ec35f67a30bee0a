"""CPU-only: the oracle (CPU restatement of the reference hot path) against closed forms, against
its own committed golden vectors (tests/golden/, PARITY UNPINNED by the reference — see
oracle/__init__.py), and the sharded protocol restatement against the unsharded one."""
import glob
import math
import os

import numpy as np
import pytest
import torch

import helpers
import oracle
from oracle import sharded
from blackbox_mpc_b200.utils import workloads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F64 = torch.float64


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


# ------------------------------------------------------------------ closed forms
def test_pendulum_step_matches_gym_closed_form():
    """utils/pendulum.py:78-92 against the textbook update written independently in numpy."""
    rng = np.random.default_rng(0)
    th, om, u = rng.uniform(-np.pi, np.pi, 200), rng.uniform(-8, 8, 200), rng.uniform(-2, 2, 200)
    x = torch.tensor(np.stack([np.cos(th), np.sin(th), om, u], 1), dtype=F64)
    dev = oracle.PendulumTrueModel()(x).numpy()
    g, m, l, dt = 10.0, 1.0, 1.0, 0.05
    th_n = np.arctan2(np.sin(th), np.cos(th))
    om2 = om + (-3 * g / (2 * l) * np.sin(th_n + np.pi) + 3.0 / (m * l ** 2) * u) * dt
    th2 = th_n + om2 * dt
    om2c = np.clip(om2, -8, 8)   # clipped AFTER the angle update (reference behaviour)
    want = np.stack([np.cos(th2), np.sin(th2), om2c], 1) - x[:, :3].numpy()
    np.testing.assert_allclose(dev, want, atol=1e-12)


def test_pendulum_reward_argument_order_quirk():
    """deterministic.py:65-66 passes (s, a, s') into a function declared (s, s', a)."""
    w = workloads.make("C1", population_size=4)
    ev = helpers.oracle_evaluator(w, F64)
    s = torch.tensor([[math.cos(0.3), math.sin(0.3), 0.5]], dtype=F64)
    a = torch.tensor([[1.7]], dtype=F64)
    s2 = ev.predict_next_state(s, a)
    got = ev.evaluate_next_reward(s, s2, a).item()
    want = -(0.3 ** 2 + 0.1 * 0.25) - 0.001 * float((s2 ** 2).sum())   # control cost from NEXT STATE
    assert abs(got - want) < 1e-12
    assert abs(got - (-(0.3 ** 2 + 0.1 * 0.25) - 0.001 * 1.7 ** 2)) > 1e-4


def test_halfcheetah_reward():
    s = torch.zeros(3, 20, dtype=F64); s2 = torch.zeros(3, 20, dtype=F64)
    s[0, 5], s[1, 6], s[2, 7] = 0.2, -0.1, 0.0
    s2[:, 17] = torch.tensor([0.05, 0.0, -0.01])
    r = oracle.halfcheetah_reward_function(s, torch.ones(3, 6, dtype=F64), s2).numpy()
    np.testing.assert_allclose(r, [-10 - 10 - 10 + 5.0, -10 + 0.0, -10 - 10 - 1.0])


def test_evaluator_nan_guard_and_row_order():
    """NaN (not inf) -> -1e6 (deterministic.py:75-77); row p*A+a starts from current_states[a] (:57)."""
    w = workloads.make("C2", population_size=3, num_agents=2, bias_scale=0.1)
    ev = helpers.oracle_evaluator(w, F64)
    acts = torch.zeros(3, 2, w.planning_horizon, 1, dtype=F64)
    acts[1, 0, 4, 0] = float("nan")
    out = ev(torch.from_numpy(w.state).double(), acts, 0)
    assert out[1, 0].item() == -1e6 and torch.isfinite(out).all()
    assert out[0, 0] == out[2, 0] and out[0, 1] == out[2, 1] and out[0, 0] != out[0, 1]


def _quadratic_evaluator(target):
    class Quad:
        def __call__(self, states, seqs, t=0):
            return -((seqs - target) ** 2).sum(dim=(2, 3))

        def predict_next_state(self, s, a):
            return s

        def evaluate_next_reward(self, s, s2, a):
            return torch.zeros(s.shape[0], dtype=s.dtype)
    return Quad()


@pytest.mark.parametrize("name", ["CEM", "PI2", "CMA-ES", "SPSA"])
def test_optimizers_climb_a_concave_quadratic(name):
    a_sp, o_sp = oracle.Space([-1.0, -1.0], [1.0, 1.0]), oracle.Space([-1.0] * 3, [1.0] * 3)
    target = torch.tensor([0.4, -0.3], dtype=F64)
    kw = dict(planning_horizon=4, population_size=400, num_agents=1, dtype=F64)
    cls = helpers.ORACLE_OPT[name]
    extra = {"CEM": dict(num_elite=40, max_iterations=8), "PI2": dict(max_iterations=8, lamda=0.3),
             "CMA-ES": dict(num_elite=40, max_iterations=8), "SPSA": dict(max_iterations=60, a_par=0.05)}[name]
    opt = cls(a_sp, o_sp, **kw, **extra)
    opt.set_trajectory_evaluator(_quadratic_evaluator(target))
    state = torch.zeros(1, 3, dtype=F64)
    action, _, _ = opt(state, 0, False, oracle.TorchDraws(0, F64))
    assert torch.linalg.vector_norm(action[0] - target) < {"CEM": 0.1, "CMA-ES": 0.15, "PI2": 0.25, "SPSA": 0.3}[name], action  # start: |target| = 0.5


def test_pi2_weights_sum_to_one_and_cem_ddof0():
    w = workloads.make("C2", population_size=50, num_agents=2, planning_horizon=6, bias_scale=0.1)
    o = helpers.oracle_optimizer(w, "PI2", dtype=F64, max_iterations=2)
    o(torch.from_numpy(w.state).double(), 0, False, oracle.TorchDraws(1, F64))
    np.testing.assert_allclose(o.trace[-1]["omega"].sum(dim=1).numpy(), 1.0, atol=1e-12)
    w.optimizer_args = dict(num_elite=5, alpha=0.0)
    c = helpers.oracle_optimizer(w, "CEM", dtype=F64, max_iterations=1)
    c(torch.from_numpy(w.state).double(), 0, False, oracle.TorchDraws(2, F64))
    tr = c.trace[0]
    el = tr["samples"].permute(1, 0, 2, 3)[0][tr["elite_idx"][0]]
    np.testing.assert_allclose(tr["variance"][0].numpy(), el.numpy().var(axis=0, ddof=0), atol=1e-14)
    # no warm start: the persistent mean is untouched by _optimize (cem.py:133-134)
    np.testing.assert_array_equal(c._previous_solution.numpy(), c._midpoint().numpy())


def test_cmaes_constants_match_hansen():
    """cma_es.py:61-92 against the formulas of Hansen's tutorial (arXiv:1604.00772), numpy."""
    a_sp, o_sp = oracle.Space([-1.0] * 6, [1.0] * 6), oracle.Space([-1.0] * 20, [1.0] * 20)
    o = oracle.CMAES(a_sp, o_sp, planning_horizon=50, population_size=500, num_elite=50, num_agents=1, dtype=F64)
    n, mu = 300, 50
    wts = np.log(mu + 0.5) - np.log(np.arange(1, mu + 1))
    wts /= wts.sum()
    mu_eff = 1.0 / (wts ** 2).sum()
    assert abs(o._mu_eff.item() - mu_eff) < 1e-9
    assert abs(o._c_sigma.item() - (mu_eff + 2) / (n + mu_eff + 5)) < 1e-12
    assert abs(o._cc.item() - (4 + mu_eff / n) / (n + 4 + 2 * mu_eff / n)) < 1e-12
    assert abs(o._c1.item() - 2 / ((n + 1.3) ** 2 + mu_eff)) < 1e-12
    # reference quirk (cma_es.py:118-126): sqrt(n * (1 - 1/(4n) + 1/(21 n^2))), the factor INSIDE the
    # root, not Hansen's sqrt(n) * (1 - 1/(4n) + 1/(21 n^2)).  Parity target = the reference's form.
    assert abs(o._expectation_of_normal.item() - math.sqrt(n * (1 - 1 / (4 * n) + 1 / (21 * n * n)))) < 1e-9
    assert abs(o._expectation_of_normal.item() - math.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))) > 1e-3
    assert o._weights.shape == (500, 1) and (o._weights[50:] == 0).all()


def test_truncated_normal_semantics():
    d = oracle.TorchDraws(0, F64)
    x = d.truncated_normal([20000, 3], torch.tensor([0.0, 1.0, -1.0], dtype=F64), torch.tensor([1.0, 0.5, 2.0], dtype=F64))
    z = (x - torch.tensor([0.0, 1.0, -1.0])) / torch.tensor([1.0, 0.5, 2.0])
    assert z.abs().max() <= 2.0
    assert abs(z.std().item() - 0.8796) < 0.01   # std of N(0,1) truncated to +-2


# ------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("case,wname,P,A", [("rollout_c1", "C1", 64, 1), ("rollout_c2", "C2", 48, 2),
                                            ("rollout_c3", "C3", 32, 1), ("rollout_c4", "C4", 24, 1)])
def test_oracle_reproduces_golden_rollouts(case, wname, P, A):
    g = load(case)
    w = workloads.make(wname, population_size=P, planning_horizon=30, num_agents=A, bias_scale=0.1)
    ev = helpers.oracle_evaluator(w, F64)
    got = ev(torch.from_numpy(w.state).double(), torch.from_numpy(g["actions"]).double(), 0).numpy()
    np.testing.assert_allclose(got, g["returns"], rtol=1e-12, atol=1e-12)
    # the fp32 twin (the reference's arithmetic type) stays within fp32 noise of the fp64 truth
    ev32 = helpers.oracle_evaluator(w, torch.float32)
    got32 = ev32(torch.from_numpy(w.state), torch.from_numpy(g["actions"]), 0).numpy()
    helpers.compare_returns(got32, g["returns"], atol=5e-3, rtol=1e-5, max_jump_frac=0.05)


GOLDEN_OPT = [("opt_cem_c2", "C2", "CEM", 64, 2, 12, dict(max_iterations=3), "cem.samples"),
              ("opt_pi2_c2", "C2", "PI2", 64, 2, 12, dict(max_iterations=3), "pi2.samples"),
              ("opt_rs_c1", "C1", "RandomSearch", 64, 2, 12, {}, "rs.samples"),
              ("opt_spsa_c2", "C2", "SPSA", 32, 1, 12, dict(max_iterations=3), "spsa.delta"),
              ("opt_cmaes_c2", "C2", "CMA-ES", 48, 1, 8, dict(max_iterations=3, num_elite=12), "cmaes.z")]


@pytest.mark.parametrize("case,wname,opt_name,P,A,H,extra,tag", GOLDEN_OPT)
def test_oracle_reproduces_golden_optimizer_runs(case, wname, opt_name, P, A, H, extra, tag):
    g = load(case)
    w = workloads.make(wname, population_size=P, planning_horizon=H, num_agents=A, bias_scale=0.1)
    if opt_name == "CEM":
        w.optimizer_args = dict(num_elite=16, alpha=0.25)
    opt = helpers.oracle_optimizer(w, opt_name, dtype=F64, **extra)
    draws = oracle.InjectedDraws({tag: list(g["draws." + tag])}, dtype=F64)
    state = torch.from_numpy(w.state).double()
    for call in range(2):
        a, n, r = opt(state, call, False, draws)
        np.testing.assert_allclose(a.numpy(), g[f"action{call}"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(n.numpy(), g[f"next{call}"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(r.numpy(), g[f"reward{call}"], rtol=1e-9, atol=1e-10)


def test_golden_fixture_inventory():
    assert len(glob.glob(os.path.join(GOLDEN, "*.npz"))) == 11   # 4 rollouts + 2 full-size return vectors + 5 optimizer runs
    assert os.path.exists(os.path.join(GOLDEN, "make_golden.py"))


# ------------------------------------------------------------------ sharded protocol (single process)
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_cem_merge_equals_unsharded(world):
    from blackbox_mpc_b200.sharding import shard_range
    rng = np.random.default_rng(3)
    P, A, H, U, E = 203, 2, 5, 3, 16
    samples = torch.from_numpy(rng.standard_normal((P, A, H, U)))
    returns = torch.from_numpy(np.round(rng.standard_normal((P, A)), 1))   # rounded: forces ties
    mean, var = torch.zeros(A, H * U, dtype=F64), torch.ones(A, H * U, dtype=F64)
    full = sharded.cem_merge(sharded.cem_partial(samples, returns, 0, E)[None], E, mean, var, 0.25)
    parts = []
    for r in range(world):
        p0, p1 = shard_range(P, r, world)
        parts.append(sharded.cem_partial(samples[p0:p1], returns[p0:p1], p0, E))
    got = sharded.cem_merge(torch.stack(parts), E, mean, var, 0.25)
    assert torch.equal(got[0], full[0]) and torch.equal(got[1], full[1])
    # and equals optimizers/cem.py:98-125 restated directly
    idx = torch.sort(-returns.t(), dim=-1, stable=True).indices[:, :E]
    el = torch.stack([samples.permute(1, 0, 2, 3)[a][idx[a]] for a in range(A)]).reshape(A, E, -1)
    np.testing.assert_allclose(full[0].numpy(), 0.25 * 0 + 0.75 * el.mean(1).numpy(), atol=1e-15)


def test_sharded_pi2_merge_equals_unsharded():
    from blackbox_mpc_b200.sharding import shard_range
    rng = np.random.default_rng(4)
    P, A, H, U = 101, 2, 4, 2
    samples = torch.from_numpy(rng.standard_normal((P, A, H, U)))
    rewards = torch.from_numpy(rng.standard_normal((P, A)) * 3)
    costs = (-rewards).t()
    prob = torch.exp(-(costs - costs.min(dim=1).values[:, None]))
    omega = prob / prob.sum(dim=1)[:, None]
    want = (samples.permute(1, 0, 2, 3) * omega[:, :, None, None]).sum(dim=1).reshape(A, -1)   # pi2.py:79-87
    parts = [sharded.pi2_partial(samples[a:b], rewards[a:b], 1.0) for a, b in (shard_range(P, r, 4) for r in range(4))]
    np.testing.assert_allclose(sharded.pi2_merge(torch.stack(parts), 1.0).numpy(), want.numpy(), rtol=1e-12, atol=1e-14)
