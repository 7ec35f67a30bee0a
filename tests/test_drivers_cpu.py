"""CPU tests of the host-side drivers: episode collection with a model-free policy on the
dependency-free pendulum environment (policy.act on the GPU path is covered by tests/test_gpu_training.py)."""
import numpy as np

from blackbox_mpc_b200.environment_utils import PendulumVecEnv
from blackbox_mpc_b200.policies.random_policy import RandomPolicy
from blackbox_mpc_b200.utils.rollouts import perform_rollouts


class _Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, value, step):
        self.rows.append((tag, value, step))


def test_pendulum_env_matches_gym_closed_form():
    env = PendulumVecEnv(num_of_agents=3, seed=1)
    obs = env.reset()
    assert obs.shape == (3, 3) and np.allclose(obs[:, 0] ** 2 + obs[:, 1] ** 2, 1, atol=1e-6) and (np.abs(obs[:, 2]) <= 1).all()
    th, thdot = np.arctan2(obs[:, 1], obs[:, 0]).astype(np.float64), obs[:, 2].astype(np.float64)
    u = np.array([[2.5], [-0.3], [0.0]])                      # first action is clipped to +2
    nxt, rew, done, info = env.step(u)
    uc = np.clip(u[:, 0], -2, 2)
    thdot2 = np.clip(thdot + (-15.0 * np.sin(th + np.pi) + 3.0 * uc) * 0.05, -8, 8)
    th2 = th + (thdot + (-15.0 * np.sin(th + np.pi) + 3.0 * uc) * 0.05) * 0.05
    np.testing.assert_allclose(nxt, np.stack([np.cos(th2), np.sin(th2), thdot2], 1), atol=1e-5)
    np.testing.assert_allclose(rew, -(th ** 2 + 0.1 * thdot ** 2 + 0.001 * uc ** 2), atol=1e-5)
    assert not done.any() and len(info) == 3


def test_perform_rollouts_shapes_and_logging():
    env = PendulumVecEnv(num_of_agents=4, seed=0)
    policy = RandomPolicy(number_of_agents=4, env_action_space=env.action_space, seed=0)
    writer = _Writer()
    obs, acts, rews = perform_rollouts(env, number_of_rollouts=3, task_horizon=25, policy=policy, tf_writer=writer)
    assert len(obs) == len(acts) == len(rews) == 3
    assert obs[0].shape == (26, 4, 3) and acts[0].shape == (25, 4, 1) and rews[0].shape == (25, 4)
    assert (acts[0] >= -2).all() and (acts[0] <= 2).all() and acts[0].std() > 0.5
    assert writer.rows == []                                   # a RandomPolicy logs no reward scalars (rollouts.py:103-108)
    # consecutive observations follow the environment dynamics
    th = np.arctan2(obs[1][:-1, :, 1], obs[1][:-1, :, 0])
    thdot = obs[1][:-1, :, 2] + (-15.0 * np.sin(th + np.pi) + 3.0 * acts[1][:, :, 0]) * 0.05
    np.testing.assert_allclose(obs[1][1:, :, 2], np.clip(thdot, -8, 8), atol=1e-4)
