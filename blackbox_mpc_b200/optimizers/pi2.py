"""PI2Optimizer (blackbox_mpc/optimizers/pi2.py:8-105): truncated-normal sampling with a fixed
variance, clip + squared-excess penalty, soft-min weights exp(-(cost-min)/lamda), weighted mean;
warm start by shifting the solution one step left (pi2.py:92-93)."""
from .. import _lib
from .optimizer_base import OptimizerBase


class PI2Optimizer(OptimizerBase):
    KIND = _lib.OPT_PI2

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, lamda=1.0):
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=max_iterations,
                         num_agents=num_agents, env_action_space=env_action_space,
                         env_observation_space=env_observation_space)
        self._population_size, self._lamda = int(population_size), float(lamda)

    def _config(self):
        return dict(lamda=self._lamda)
