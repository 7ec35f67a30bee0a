from .optimizer_base import OptimizerBase
from .cem import CEMOptimizer
from .pi2 import PI2Optimizer
from .random_search import RandomSearchOptimizer
from .pso import PSOOptimizer
from .spsa import SPSAOptimizer
from .cma_es import CMAESOptimizer

BY_NAME = {"CEM": CEMOptimizer, "CMA-ES": CMAESOptimizer, "PI2": PI2Optimizer, "PSO": PSOOptimizer,
           "SPSA": SPSAOptimizer, "RandomSearch": RandomSearchOptimizer}
