"""GPU parity: bbmpc_rollout / predict_next_state / reward (through the Python plugin classes, i.e.
through the C ABI) against the oracle on identical seeded inputs."""
import numpy as np
import pytest
import torch

import helpers
from blackbox_mpc_b200.utils import workloads

pytestmark = pytest.mark.gpu

# fp32 path: same arithmetic as the reference up to summation order.  bf16x3: operands carry 16
# mantissa bits, products ~2^-17 relative; measured max |error| on H=30 returns of magnitude 10^2..10^3
# is 2e-4..6e-4 (tools/debug/err_stats.py).  The bound is kept within ~10x of that: a systematic operand
# error (e.g. a truncating instead of rounding hi/lo split: 4e-2) must fail.
TOL = {"fp32": dict(atol=2e-3, rtol=2e-5), "bf16x3": dict(atol=3e-3, rtol=3e-5)}
STEP_TOL = {"fp32": 2e-5, "bf16x3": 1e-4}


def _policy_and_eval(w, precision):
    policy = workloads.build_policy(w, precision=precision)
    return policy, policy._trajectory_evaluator


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name,P,A", [("C2", 300, 1), ("C3", 200, 1), ("C4", 131, 1), ("C4", 64, 3), ("C2", 129, 2)])
def test_rollout_matches_oracle(cuda_device, name, P, A, precision):
    w = workloads.make(name, population_size=P, num_agents=A, bias_scale=0.1)
    _, ev = _policy_and_eval(w, precision)
    assert ev.engine().effective_precision == precision
    actions = helpers.random_actions(w, P, seed=1)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0).cpu().numpy()
    ref = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy()
    assert got.shape == (P, A)
    helpers.compare_returns(got, ref, max_jump_frac=0.02, **TOL[precision])


def test_rollout_pendulum_true_model(cuda_device):
    w = workloads.make("C1", population_size=500, num_agents=2)
    _, ev = _policy_and_eval(w, "auto")
    actions = helpers.random_actions(w, 500, seed=2)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0).cpu().numpy()
    ref32 = helpers.oracle_evaluator(w, torch.float32)(state, actions, 0).numpy()
    ref64 = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy()
    np.testing.assert_allclose(got, ref64, rtol=2e-4, atol=2e-3)
    np.testing.assert_allclose(got, ref32, rtol=2e-4, atol=2e-3)


@pytest.mark.parametrize("name", ["C1", "C2", "C4"])
def test_single_step_predict_and_reward(cuda_device, name):
    w = workloads.make(name, num_agents=37, bias_scale=0.1)
    _, ev = _policy_and_eval(w, "fp32")
    g = torch.Generator().manual_seed(3)
    s = torch.from_numpy(w.state)
    a = torch.from_numpy(w.lb) + torch.from_numpy(w.ub - w.lb) * torch.rand(37, w.dU, generator=g)
    o = helpers.oracle_evaluator(w, torch.float64)
    nxt = ev.predict_next_state(s, a).cpu().numpy()
    ref_nxt = o.predict_next_state(s.double(), a.double()).numpy()
    np.testing.assert_allclose(nxt, ref_nxt, rtol=2e-5, atol=2e-5)
    rew = ev.evaluate_next_reward(s, torch.from_numpy(ref_nxt.astype(np.float32)), a).cpu().numpy()
    ref_rew = o.evaluate_next_reward(s.double(), torch.from_numpy(ref_nxt.astype(np.float32)).double(), a.double()).numpy()
    helpers.compare_returns(rew, ref_rew, atol=2e-3, rtol=2e-5, max_jump_frac=0.1)


def test_dynamics_function_callable(cuda_device):
    """DeterministicMLP.__call__(x, train) and PendulumTrueModel.__call__ are the reference's
    dynamics_function plug point (deterministic_mlp.py:27-51, pendulum.py:58-92)."""
    from blackbox_mpc_b200.dynamics_functions.deterministic_mlp import DeterministicMLP
    from blackbox_mpc_b200.utils.pendulum import PendulumTrueModel
    import oracle
    m = DeterministicMLP([26, 200, 200, 200, 20], ["tanh", "tanh", "tanh", None], seed=5)
    x = torch.randn(77, 26, generator=torch.Generator().manual_seed(4))
    got = m(x, False).cpu()
    ref = oracle.MLP([t.cpu().double() for t in m.weights], [t.cpu().double() for t in m.biases], ["tanh"] * 3 + [None])(x.double())
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)
    xp = torch.randn(50, 4, generator=torch.Generator().manual_seed(6))
    np.testing.assert_allclose(PendulumTrueModel()(xp, False).cpu().numpy(), oracle.PendulumTrueModel()(xp.double()).numpy(),
                               rtol=1e-5, atol=1e-5)


def test_nan_guard_and_empty(cuda_device):
    """NaN (not inf) -> -1e6 (deterministic.py:75-77); P = 0 is a no-op."""
    w = workloads.make("C2", population_size=40, bias_scale=0.1)
    _, ev = _policy_and_eval(w, "fp32")
    actions = helpers.random_actions(w, 40, seed=7)
    actions[3, 0, 5, 0] = float("nan")
    got = ev(torch.from_numpy(w.state), actions, 0).cpu().numpy()
    assert got[3, 0] == np.float32(-1e6) and np.isfinite(np.delete(got[:, 0], 3)).all()
    empty = ev(torch.from_numpy(w.state), actions[:0], 0)
    assert tuple(empty.shape) == (0, 1)


def test_full_size_properties(cuda_device):
    """BASELINE C4 at full size: properties that do not need the oracle — permutation equivariance
    of the population axis, tile-boundary independence, and fp32 vs bf16x3 agreement."""
    w = workloads.make("C4", bias_scale=0.1)
    _, ev3 = _policy_and_eval(w, "bf16x3")
    actions = helpers.random_actions(w, w.population_size, seed=8)
    state = torch.from_numpy(w.state)
    r = ev3(state, actions, 0).cpu()
    perm = torch.randperm(w.population_size, generator=torch.Generator().manual_seed(9))
    r_perm = ev3(state, actions[perm], 0).cpu()
    assert torch.equal(r[perm], r_perm)
    r_head = ev3(state, actions[:1000], 0).cpu()
    assert torch.equal(r[:1000], r_head)
    _, ev32 = _policy_and_eval(w, "fp32")
    r32 = ev32(state, actions, 0).cpu()
    helpers.compare_returns(r.numpy(), r32.numpy(), max_jump_frac=0.02, **TOL["bf16x3"])


@pytest.mark.parametrize("P,A", [(700, 1), (130, 3), (20000, 1)])
def test_member_parallel_matches_single_cta(cuda_device, monkeypatch, P, A):
    """Ensembles run member-parallel (n_members CTAs share a tile and exchange their raw outputs every
    horizon step); BBMPC_NO_GROUPS=1 forces the one-CTA-per-tile pipeline.  Both contract the same
    products; they differ only in where the member sum is formed (fp32 adds vs the TMEM accumulator).
    Covers ragged last tiles, several agents and more tiles than groups (several rounds per group)."""
    w = workloads.make("C4", population_size=P, num_agents=A, bias_scale=0.1)
    _, ev = _policy_and_eval(w, "bf16x3")
    actions = helpers.random_actions(w, P, seed=11)
    state = torch.from_numpy(w.state)
    r_group = ev(state, actions, 0).cpu().numpy()
    r_again = ev(state, actions, 0).cpu().numpy()
    assert np.array_equal(r_group, r_again)                 # deterministic exchange
    monkeypatch.setenv("BBMPC_NO_GROUPS", "1")
    r_single = ev(state, actions, 0).cpu().numpy()
    monkeypatch.delenv("BBMPC_NO_GROUPS")
    assert np.isfinite(r_group).all()
    helpers.compare_returns(r_group, r_single, max_jump_frac=0.02, **TOL["bf16x3"])


def _custom_workload(layers, acts, n_members, dS, dU, P, A, H, seed):
    """A halfcheetah-style (dS >= 18) or pendulum-style workload with arbitrary MLP shape."""
    base = workloads.make("C4" if dS >= 18 else "C2", population_size=P, planning_horizon=H, num_agents=A, seed=seed, bias_scale=0.1)
    members = [workloads.glorot_mlp(layers, seed + 10 * m, 0.1) for m in range(n_members)]
    base.layers, base.activations = layers, acts
    base.weights, base.biases = [m[0] for m in members], [m[1] for m in members]
    base.dS, base.dU = dS, dU
    base.lb, base.ub = -np.ones(dU, np.float32), np.ones(dU, np.float32)
    base.stats = workloads._stats(dS, dU, seed)
    g = torch.Generator().manual_seed(77 + seed)
    base.state = (torch.randn(A, dS, generator=g) * 0.5).numpy().astype(np.float32)
    if dS < 18:
        base.state[:, :2] /= np.linalg.norm(base.state[:, :2], axis=1, keepdims=True)
    return base


@pytest.mark.parametrize("layers,acts,n_members,dS,dU,P,A", [
    ([4, 16, 3], ["tanh", None], 2, 3, 1, 200, 1),                          # tiny ensemble, pendulum-style reward
    ([26, 100, 100, 20], ["tanh", "relu", None], 3, 20, 6, 300, 2),        # mixed activations, 3 members, 2 agents
    ([26, 200, 20], ["sigmoid", None], 1, 20, 6, 150, 1),                   # one hidden layer, sigmoid
    ([42, 64, 64, 64, 64, 30], ["tanh"] * 4 + [None], 8, 30, 12, 257, 1),  # 8 members, dS=30, dU=12, 5 layers
    ([32, 200, 200, 20], ["tanh", "tanh", None], 2, 20, 12, 129, 1),       # widest layers the TMEM budget admits with a 16-slot action block
    ([26, 20], [None], 1, 20, 6, 100, 1),                                   # linear single layer
])
def test_rollout_shapes_against_oracle(cuda_device, layers, acts, n_members, dS, dU, P, A):
    """Tensor-core path over model shapes the BASELINE configs do not touch: member counts (member-parallel
    groups of 2..8 CTAs), layer counts, widths with ragged tail chunks, activation mixes, state/action widths
    that select the other kernel instantiations."""
    H = 12
    w = _custom_workload(layers, acts, n_members, dS, dU, P, A, H, seed=3)
    _, ev = _policy_and_eval(w, "bf16x3")
    assert ev.engine().effective_precision == "bf16x3"
    actions = helpers.random_actions(w, P, seed=5)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0).cpu().numpy()
    ref = helpers.oracle_evaluator(w, torch.float64)(state.double(), actions.double(), 0).numpy()
    helpers.compare_returns(got, ref, max_jump_frac=0.05, **TOL["bf16x3"])


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_unnormalised_mlp_rollout(cuda_device, precision):
    """is_normalized=False with a learned MLP: process_input is the raw concat [s, a], process_output is s + y
    (system_dynamics_handler.py:125-126,160-161 of the reference).  Weights scaled down so that the raw recurrence stays bounded."""
    import oracle as ref
    from blackbox_mpc_b200.dynamics_functions.deterministic_mlp import DeterministicMLP
    from blackbox_mpc_b200.dynamics_handlers.system_dynamics_handler import SystemDynamicsHandler
    from blackbox_mpc_b200.spaces import Box
    from blackbox_mpc_b200.trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
    from blackbox_mpc_b200.utils import halfcheetah
    P = 300
    w = workloads.make("C3", population_size=P, planning_horizon=12, bias_scale=0.05)
    ws = [x * np.float32(0.3) for x in w.weights[0]]
    bs = [x * np.float32(0.3) for x in w.biases[0]]
    mlp = DeterministicMLP(w.layers, w.activations)
    mlp.set_weights(ws, bs)
    act_space, obs_space = Box(w.lb, w.ub), Box(-np.ones(w.dS, np.float32), np.ones(w.dS, np.float32))
    handler = SystemDynamicsHandler(act_space, obs_space, dynamics_function=mlp, true_model=False, is_normalized=False, precision=precision)
    ev = DeterministicTrajectoryEvaluator(reward_function=halfcheetah.reward_function, system_dynamics_handler=handler)
    assert ev.engine().effective_precision == precision
    actions = helpers.random_actions(w, P, seed=13)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0).cpu().numpy()
    o_handler = ref.Handler(ref.MLP([torch.from_numpy(x) for x in ws], [torch.from_numpy(x) for x in bs], w.activations),
                            true_model=False, is_normalized=False, dtype=torch.float64)
    o_ev = ref.Evaluator(ref.halfcheetah_reward_function, o_handler)
    want = o_ev(state.double(), actions.double(), 0).numpy()
    helpers.compare_returns(got, want, max_jump_frac=0.02, **TOL[precision])
    nxt = ev.predict_next_state(state, actions[0, :, 0]).cpu().numpy()
    np.testing.assert_allclose(nxt, o_ev.predict_next_state(state.double(), actions[0, :, 0].double()).numpy(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_infinite_returns_pass_through(cuda_device, precision):
    """deterministic.py:75-77 replaces NaN returns only: +inf and -inf reach the optimizer unchanged."""
    from blackbox_mpc_b200.trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
    from blackbox_mpc_b200.utils import rewards
    src = """
    __device__ float reward(const float* s, const float* a, const float* s2) {
      if (a[0] > 0.9f) return __int_as_float(0x7f800000);        // +inf
      if (a[0] < -0.9f) return -__int_as_float(0x7f800000);      // -inf
      return s2[0] - s[0];
    }"""
    P = 400
    w = workloads.make("C3", population_size=P, planning_horizon=6, bias_scale=0.1)
    policy = workloads.build_policy(w, precision=precision)
    ev = DeterministicTrajectoryEvaluator(reward_function=rewards.cuda_reward(src),
                                          system_dynamics_handler=policy._trajectory_evaluator._system_dynamics_handler)
    actions = helpers.random_actions(w, P, seed=14) * 0.5            # |a| <= 0.5: finite by default
    actions[3, 0, 2, 0] = 0.95                                       # +inf once
    actions[5, 0, 4, 0] = -0.95                                      # -inf once
    actions[7, 0, 1, 0], actions[7, 0, 3, 0] = 0.95, -0.95           # +inf - inf = NaN -> -1e6
    got = ev(torch.from_numpy(w.state), actions, 0).cpu().numpy()[:, 0]
    assert got[3] == np.inf and got[5] == -np.inf and got[7] == np.float32(-1e6)
    assert np.isfinite(np.delete(got, [3, 5, 7])).all()
