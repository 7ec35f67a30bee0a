"""SPSAOptimizer (blackbox_mpc/optimizers/spsa.py:6-127): Rademacher perturbations, one evaluator
call on 2P rows (:94-96), gradient estimate mean_P[(r+ - r-)/(2 c_k delta)], clipped ascent step,
shift-left warm start (:114-115).  a_k = a/(t+1+iters/10)^alpha, c_k = c/(t+1)^gamma (:56,69-70)."""
from .. import _lib
from .optimizer_base import OptimizerBase


class SPSAOptimizer(OptimizerBase):
    KIND = _lib.OPT_SPSA

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, alpha=0.602, gamma=0.101, a_par=0.01, noise_parameter=0.3):
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=max_iterations,
                         num_agents=num_agents, env_action_space=env_action_space,
                         env_observation_space=env_observation_space)
        self._population_size = int(population_size)
        self._alpha, self._gamma, self._a_par, self._noise_parameter = float(alpha), float(gamma), float(a_par), float(noise_parameter)

    def _config(self):
        return dict(alpha=self._alpha, gamma=self._gamma, a_par=self._a_par, noise_parameter=self._noise_parameter)
