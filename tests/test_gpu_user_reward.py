"""User-supplied reward functions as CUDA source (bbmpc_reward_set_nvrtc, SURVEY 8b plug point (2)): the reference accepts
an arbitrary reward_function(current_state, actions, next_state) (policies/mpc_policy.py:42-44) and its headline tutorial's
reward is user code (tutorials/mujoco/cost_func.py:5-22)."""
import numpy as np
import pytest
import torch

import helpers
import oracle
from blackbox_mpc_b200 import _lib
from blackbox_mpc_b200.utils import halfcheetah, rewards, workloads

pytestmark = pytest.mark.gpu


def _policy(w, reward, precision):
    """workloads.build_policy with the reward function swapped."""
    policy = workloads.build_policy(w, precision=precision)
    from blackbox_mpc_b200.trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
    ev = DeterministicTrajectoryEvaluator(reward_function=reward, system_dynamics_handler=policy._trajectory_evaluator._system_dynamics_handler)
    policy._optimizer.set_trajectory_evaluator(ev)
    policy._trajectory_evaluator = ev
    return policy, ev


@pytest.mark.parametrize("name,P,precision", [("C4", 700, "bf16x3"), ("C4", 10000, "bf16x3"), ("C3", 300, "fp32"), ("C3", 5000, "bf16x3")])
def test_halfcheetah_reward_from_source_equals_builtin(cuda_device, name, P, precision):
    """cost_func.py:5-22 supplied as CUDA source must reproduce the built-in device reward bit for bit: rollout returns
    (tensor-core and fp32 paths), evaluate_next_reward, and a whole act()."""
    w = workloads.make(name, population_size=P, bias_scale=0.1)
    actions = helpers.random_actions(w, P, seed=41)
    state = torch.from_numpy(w.state)
    p_builtin = workloads.build_policy(w, precision=precision)
    want = p_builtin._trajectory_evaluator(state, actions, 0).cpu().numpy()
    p_user, ev = _policy(w, rewards.cuda_reward(rewards.HALFCHEETAH_SOURCE), precision)
    got = ev(state, actions, 0).cpu().numpy()
    assert np.array_equal(got, want), f"max |d| = {np.abs(got - want).max()}"
    a0 = actions[0, :, 0]
    nxt = ev.predict_next_state(state, a0)
    r_user = ev.evaluate_next_reward(state, nxt, a0).cpu().numpy()
    r_builtin = p_builtin._trajectory_evaluator.evaluate_next_reward(state, nxt, a0).cpu().numpy()
    assert np.array_equal(r_user, r_builtin)
    if P <= 1000:
        act_u = p_user.act(w.state[0], 0)
        act_b = p_builtin.act(w.state[0], 0)
        for u, b in zip(act_u, act_b):
            assert np.array_equal(np.asarray(u), np.asarray(b))


def test_custom_reward_matches_oracle(cuda_device):
    """A reward that is NOT built in (quadratic tracking cost with an action penalty and a NaN trap) against the float64
    oracle evaluator driven by the same function written in torch."""
    src = """
    __device__ float reward(const float* s, const float* a, const float* s2) {
      float c = 0.0f;
      for (int i = 0; i < BBMPC_DS; ++i) { const float d = s2[i] - 0.25f * s[i]; c = c + d * d; }
      float u = 0.0f;
      for (int i = 0; i < BBMPC_DU; ++i) u = u + a[i] * a[i];
      return -c - 0.1f * u + 3.0f * s2[17];
    }"""

    def torch_reward(s, a, s2):
        return -((s2 - 0.25 * s) ** 2).sum(-1) - 0.1 * (a ** 2).sum(-1) + 3.0 * s2[:, 17]

    P = 600
    w = workloads.make("C4", population_size=P, bias_scale=0.1)
    _, ev = _policy(w, rewards.cuda_reward(src, "tracking"), "bf16x3")
    actions = helpers.random_actions(w, P, seed=42)
    actions[7, 0, 3, 1] = float("nan")          # NaN return -> -1e6 (deterministic.py:75-77)
    state = torch.from_numpy(w.state)
    got = ev(state, actions, 0).cpu().numpy()
    oev = helpers.oracle_evaluator(w, torch.float64)
    oev._reward_function = torch_reward
    ref = oev(state.double(), actions.double(), 0).numpy()
    assert got[7, 0] == np.float32(-1e6) and ref[7, 0] == -1e6
    np.testing.assert_allclose(got, ref, rtol=3e-5, atol=3e-3)
    # stand-alone call of the reward object (B rows)
    s = torch.randn(33, 20); a = torch.rand(33, 6); s2 = torch.randn(33, 20)
    r = rewards.cuda_reward(src)(s, a, s2).cpu().numpy()
    np.testing.assert_allclose(r, torch_reward(s.double(), a.double(), s2.double()).numpy(), rtol=1e-5, atol=1e-4)


def test_compile_error_is_reported(cuda_device):
    w = workloads.make("C2", population_size=64, bias_scale=0.1)
    _, ev = _policy(w, rewards.cuda_reward("__device__ float reward(const float* s, const float* a, const float* s2) { return undefined_symbol; }"), "fp32")
    with pytest.raises(_lib.BBMPCError) as ei:
        ev(torch.from_numpy(w.state), helpers.random_actions(w, 64, seed=1), 0)
    assert "undefined_symbol" in str(ei.value)
    with pytest.raises(TypeError):
        from blackbox_mpc_b200.trajectory_evaluators.deterministic import DeterministicTrajectoryEvaluator
        DeterministicTrajectoryEvaluator(reward_function=lambda s, a, s2: 0.0, system_dynamics_handler=ev._system_dynamics_handler)
