"""Builds oracle objects from a synthetic Workload (blackbox_mpc_b200.utils.workloads: pure numpy
data).  TEST INFRASTRUCTURE — see oracle/__init__.py: used by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only."""
import numpy as np
import torch

from . import reference_port as ref


def spaces(w):
    return ref.Space(w.lb, w.ub), ref.Space(-np.ones(w.dS, np.float32), np.ones(w.dS, np.float32))


def evaluator(w, dtype=torch.float32):
    if w.dynamics == "pendulum_true":
        handler = ref.Handler(ref.PendulumTrueModel(), true_model=True, dtype=dtype)
    else:
        members = [ref.MLP([torch.from_numpy(x) for x in ws], [torch.from_numpy(x) for x in bs], w.activations)
                   for ws, bs in zip(w.weights, w.biases)]
        fn = members[0] if len(members) == 1 else ref.Ensemble(members)
        handler = ref.Handler(fn, true_model=False, is_normalized=True, stats=w.stats, dtype=dtype)
    reward = ref.pendulum_reward_function if w.reward == "pendulum" else ref.halfcheetah_reward_function
    return ref.Evaluator(reward, handler)


OPTIMIZERS = {"CEM": ref.CEM, "PI2": ref.PI2, "RandomSearch": ref.RandomSearch, "PSO": ref.PSO,
              "SPSA": ref.SPSA, "CMA-ES": ref.CMAES}


def optimizer(w, name=None, dtype=torch.float32, **extra):
    name = name or w.optimizer_name
    a_sp, o_sp = spaces(w)
    args = dict(w.optimizer_args) if name == w.optimizer_name else {}
    args.update(planning_horizon=w.planning_horizon, population_size=w.population_size, num_agents=w.num_agents)
    if name != "RandomSearch":
        args["max_iterations"] = w.max_iterations or 5
    args.update(extra)
    opt = OPTIMIZERS[name](a_sp, o_sp, dtype=dtype, **args)
    opt.set_trajectory_evaluator(evaluator(w, dtype))
    return opt
