"""CEMOptimizer (blackbox_mpc/optimizers/cem.py:6-149): constrained-variance truncated-normal
sampling, top-k elites, mean / ddof-0 variance refit, alpha smoothing.  No warm start between
act() calls (cem.py:133-134) and `epsilon` stored but unused (cem.py:53) — reference behaviour."""
from .. import _lib
from .optimizer_base import OptimizerBase


class CEMOptimizer(OptimizerBase):
    KIND = _lib.OPT_CEM

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_elite=50, num_agents=5, epsilon=0.001, alpha=0.25):
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=max_iterations,
                         num_agents=num_agents, env_action_space=env_action_space,
                         env_observation_space=env_observation_space)
        self._population_size, self._num_elite = int(population_size), int(num_elite)
        self._epsilon, self._alpha = float(epsilon), float(alpha)

    def _config(self):
        return dict(num_elite=self._num_elite, alpha=self._alpha, epsilon=self._epsilon)
