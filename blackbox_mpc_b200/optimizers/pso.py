"""PSOOptimizer (blackbox_mpc/optimizers/pso.py:6-160).  Reference quirks kept: r1/r2 are ONE
N(0,1) scalar each per iteration (:108-109); all swarm state starts at zero until reset() or the
re-seeding tail of the first _optimize (:50-68, :116-138); the re-seed variance uses the un-shifted
global best while the mean is shifted (:116-127)."""
from .. import _lib
from .optimizer_base import OptimizerBase


class PSOOptimizer(OptimizerBase):
    KIND = _lib.OPT_PSO

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, c1=0.3, c2=0.5, w=0.2, initial_velocity_fraction=0.01):
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=max_iterations,
                         num_agents=num_agents, env_action_space=env_action_space,
                         env_observation_space=env_observation_space)
        self._population_size = int(population_size)
        self._c1, self._c2, self._w = float(c1), float(c2), float(w)
        self._initial_velocity_fraction = float(initial_velocity_fraction)

    def _config(self):
        return dict(c1=self._c1, c2=self._c2, w=self._w, initial_velocity_fraction=self._initial_velocity_fraction)
