"""PendulumVecEnv — n independent Pendulum-v0 instances stepped in numpy: a stand-in for
`EnvironmentWrapper.make_standard_gym_env("Pendulum-v0", num_of_agents=n)` of the reference's tutorials
(e.g. tutorials/true_model_mpc/tutorial_one.py:13-14) in an image without gym.  Dynamics, reward, action
clipping and reset distribution follow gym's Pendulum-v0 (g = 10, m = l = 1, dt = 0.05, |u| <= 2, |thdot| <= 8)."""
import numpy as np

from ..spaces import Box


class PendulumVecEnv:
    def __init__(self, num_of_agents=1, seed=0):
        self.num_of_agents = int(num_of_agents)
        self.action_space = Box(np.array([-2.0], np.float32), np.array([2.0], np.float32))
        self.observation_space = Box(np.array([-1.0, -1.0, -8.0], np.float32), np.array([1.0, 1.0, 8.0], np.float32))
        self._rng = np.random.default_rng(seed)
        self._th = np.zeros(self.num_of_agents)
        self._thdot = np.zeros(self.num_of_agents)

    def seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def _obs(self):
        return np.stack([np.cos(self._th), np.sin(self._th), self._thdot], axis=1).astype(np.float32)

    def reset(self):
        self._th = self._rng.uniform(-np.pi, np.pi, self.num_of_agents)
        self._thdot = self._rng.uniform(-1.0, 1.0, self.num_of_agents)
        return self._obs()

    def step(self, actions):
        u = np.clip(np.asarray(actions, np.float64).reshape(self.num_of_agents, -1)[:, 0], -2.0, 2.0)
        wrapped = ((self._th + np.pi) % (2 * np.pi)) - np.pi
        rewards = -(wrapped ** 2 + 0.1 * self._thdot ** 2 + 0.001 * u ** 2)
        self._thdot = self._thdot + (-3 * 10.0 / 2 * np.sin(self._th + np.pi) + 3.0 * u) * 0.05
        self._th = self._th + self._thdot * 0.05
        self._thdot = np.clip(self._thdot, -8.0, 8.0)
        return self._obs(), rewards.astype(np.float32), np.zeros(self.num_of_agents, bool), [{} for _ in range(self.num_of_agents)]

    def close(self):
        return
