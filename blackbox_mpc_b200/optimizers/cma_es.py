"""CMAESOptimizer (blackbox_mpc/optimizers/cma_es.py:6-227).  Runs entirely inside libbbmpc
(csrc/cmaes.cu): fused Philox + `z @ (B D)` sampling GEMM, rollout, exact global top-E, evolution
paths / per-coordinate step size / rank-one + rank-mu covariance update on the E elite rows (the
reference materialises a [P,N,N] tensor there), symmetric eigendecomposition (cuSOLVER syevd).
State (`m`, `sigma`, `C`, `B`, `D`, `p_sigma`, `p_C`) persists across act() calls; reset() restores
`m` and `sigma` only, as the reference does.  `get_tensor("D")` is the diagonal of D."""
from .. import _lib
from .optimizer_base import OptimizerBase


class CMAESOptimizer(OptimizerBase):
    KIND = _lib.OPT_CMAES

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_elite=50, num_agents=5, alpha_cov=2.0, h_sigma=1.0):   # reference order (cma_es.py:7-9)
        super().__init__(name=None, planning_horizon=planning_horizon, max_iterations=max_iterations,
                         num_agents=num_agents, env_action_space=env_action_space,
                         env_observation_space=env_observation_space)
        self._population_size, self._num_elite = int(population_size), int(num_elite)
        self._h_sigma, self._alpha_cov = float(h_sigma), float(alpha_cov)

    def _config(self):
        return dict(num_elite=self._num_elite, h_sigma=self._h_sigma, alpha_cov=self._alpha_cov)
