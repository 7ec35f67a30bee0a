"""RandomPolicy (blackbox_mpc/policies/random_policy.py:5-56): uniform actions for data collection.
The reference passes (minval=high, maxval=low) to tf.random.uniform (:20-23,:45-47), which still draws
high + u (low - high), u in [0,1): the same uniform distribution over the box, reproduced here with numpy."""
import numpy as np

from .model_free_base_policy import ModelFreeBasePolicy


class RandomPolicy(ModelFreeBasePolicy):
    def __init__(self, number_of_agents, env_action_space, seed=None):
        self._num_of_agents = int(number_of_agents)
        self._high = np.asarray(env_action_space.high, np.float32)
        self._low = np.asarray(env_action_space.low, np.float32)
        self._rng = np.random.default_rng(seed)

    def act(self, observations, t, exploration_noise=False):
        u = self._rng.random((self._num_of_agents, *self._high.shape), dtype=np.float32)
        return self._high + u * (self._low - self._high)

    def reset(self):
        return
