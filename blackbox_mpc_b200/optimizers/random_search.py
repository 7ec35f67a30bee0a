"""RandomSearchOptimizer (blackbox_mpc/optimizers/random_search.py:6-54): one shot of uniform
samples in the action box, argmax (first index on ties), first action of the best sequence."""
from .. import _lib
from .optimizer_base import OptimizerBase


class RandomSearchOptimizer(OptimizerBase):
    KIND = _lib.OPT_RANDOM_SEARCH

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, population_size=1024,
                 num_agents=5):
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=None, num_agents=num_agents,
                         env_action_space=env_action_space, env_observation_space=env_observation_space)
        self._population_size = int(population_size)

    def reset(self):
        return
