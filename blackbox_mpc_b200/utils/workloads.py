"""Synthetic workloads C1..C5 of BASELINE.json / SURVEY §8(d), generated on the CPU with fixed
seeds (torch.Generator().manual_seed) so every rank — and the CPU oracle — sees identical values.
Pure data: numpy arrays only, no model/engine objects, importable by bench.py, tests and the
oracle legs alike."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch


@dataclass
class Workload:
    name: str
    optimizer_name: str
    dS: int
    dU: int
    lb: np.ndarray
    ub: np.ndarray
    population_size: int
    planning_horizon: int
    max_iterations: Optional[int]
    num_agents: int
    reward: str                       # "pendulum" | "halfcheetah"
    dynamics: str                     # "pendulum_true" | "mlp"
    layers: Optional[List[int]] = None
    activations: Optional[List[Optional[str]]] = None
    weights: List[List[np.ndarray]] = field(default_factory=list)   # [member][layer] [in,out]
    biases: List[List[np.ndarray]] = field(default_factory=list)
    stats: Optional[List[np.ndarray]] = None   # mean_s, std_s, mean_a, std_a, mean_t, std_t
    state: Optional[np.ndarray] = None         # [A, dS]
    optimizer_args: Dict = field(default_factory=dict)

    @property
    def n_members(self) -> int:
        return len(self.weights)

    def flops_per_row_step(self) -> int:
        if self.dynamics != "mlp":
            return 0
        return 2 * sum(a * b for a, b in zip(self.layers[:-1], self.layers[1:])) * self.n_members

    def flops_per_iteration(self) -> int:
        mult = 2 if self.optimizer_name == "SPSA" else 1
        return self.flops_per_row_step() * self.population_size * mult * self.num_agents * self.planning_horizon


def glorot_mlp(layers, seed, bias_scale=0.0):
    g = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    for fi, fo in zip(layers[:-1], layers[1:]):
        lim = math.sqrt(6.0 / (fi + fo))
        ws.append(((torch.rand(fi, fo, generator=g) * 2 - 1) * lim).numpy().astype(np.float32))
        b = torch.randn(fo, generator=g) * bias_scale if bias_scale else torch.zeros(fo)
        bs.append(b.numpy().astype(np.float32))
    return ws, bs


def _stats(dS, dU, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    n = lambda k, s: (torch.randn(k, generator=g) * s).numpy().astype(np.float32)          # noqa: E731
    u = lambda k, a, b: (torch.rand(k, generator=g) * (b - a) + a).numpy().astype(np.float32)  # noqa: E731
    return [n(dS, 0.1), u(dS, 0.5, 1.5), n(dU, 0.1), u(dU, 0.5, 1.5), n(dS, 0.1), u(dS, 0.02, 0.08)]


def _pendulum_state(A, seed):
    g = torch.Generator().manual_seed(2000 + seed)
    th = (torch.rand(A, generator=g) * 2 - 1) * math.pi
    om = torch.rand(A, generator=g) * 2 - 1
    return torch.stack([torch.cos(th), torch.sin(th), om], 1).numpy().astype(np.float32)


def _cheetah_state(A, seed):
    g = torch.Generator().manual_seed(3000 + seed)
    return torch.randn(A, 20, generator=g).numpy().astype(np.float32)


def make(name: str, population_size: Optional[int] = None, planning_horizon: Optional[int] = None,
         num_agents: int = 1, seed: int = 0, bias_scale: float = 0.0) -> Workload:
    """name in {"C1".."C5"}; population/horizon overridable for reduced-size parity cases."""
    name = name.upper()
    pend_lb, pend_ub = np.array([-2.0], np.float32), np.array([2.0], np.float32)
    ch_lb, ch_ub = -np.ones(6, np.float32), np.ones(6, np.float32)
    tanh3 = ["tanh", "tanh", "tanh", None]
    if name == "C1":
        w = Workload("C1", "RandomSearch", 3, 1, pend_lb, pend_ub, population_size or 500, planning_horizon or 30,
                     None, num_agents, "pendulum", "pendulum_true", state=_pendulum_state(num_agents, seed))
    elif name == "C2":
        layers = [4, 64, 64, 3]
        ws, bs = glorot_mlp(layers, seed, bias_scale)
        w = Workload("C2", "CEM", 3, 1, pend_lb, pend_ub, population_size or 2000, planning_horizon or 30, 5, num_agents,
                     "pendulum", "mlp", layers, ["tanh", "tanh", None], [ws], [bs], _stats(3, 1, seed),
                     _pendulum_state(num_agents, seed), dict(num_elite=50, alpha=0.25))
    elif name in ("C3", "C4", "C5"):
        layers = [26, 200, 200, 200, 20]
        n_members = 5 if name == "C4" else 1
        members = [glorot_mlp(layers, seed + m, bias_scale) for m in range(n_members)]
        opt = {"C3": "PI2", "C4": "CEM", "C5": "CMA-ES"}[name]
        P = {"C3": 5000, "C4": 10000, "C5": 50000}[name]
        H = {"C3": 30, "C4": 30, "C5": 50}[name]
        args = {"C3": dict(lamda=1.0), "C4": dict(num_elite=50, alpha=0.25),
                "C5": dict(num_elite=50, alpha_cov=2.0, h_sigma=1.0)}[name]
        w = Workload(name, opt, 20, 6, ch_lb, ch_ub, population_size or P, planning_horizon or H, 5, num_agents,
                     "halfcheetah", "mlp", layers, tanh3, [m[0] for m in members], [m[1] for m in members],
                     _stats(20, 6, seed), _cheetah_state(num_agents, seed), args)
    else:
        raise ValueError(f"unknown workload {name}")
    return w


def build_policy(w: Workload, precision: str = "auto", seed: int = 0, optimizer_name: Optional[str] = None):
    """MPCPolicy over the B200 engine for a workload (the product path)."""
    from ..dynamics_functions.deterministic_mlp import DeterministicMLP, EnsembleMLP
    from ..dynamics_handlers.system_dynamics_handler import SystemDynamicsHandler
    from ..policies.mpc_policy import MPCPolicy
    from ..spaces import Box
    from . import halfcheetah, pendulum
    act_space, obs_space = Box(w.lb, w.ub), Box(-np.ones(w.dS, np.float32) * np.inf, np.ones(w.dS, np.float32) * np.inf)
    reward = pendulum.pendulum_reward_function if w.reward == "pendulum" else halfcheetah.reward_function
    if w.dynamics == "pendulum_true":
        handler = SystemDynamicsHandler(act_space, obs_space, dynamics_function=pendulum.PendulumTrueModel(),
                                        true_model=True, seed=seed, precision=precision)
    else:
        members = []
        for ws, bs in zip(w.weights, w.biases):
            m = DeterministicMLP(w.layers, w.activations)
            m.set_weights(ws, bs)
            members.append(m)
        fn = members[0] if len(members) == 1 else EnsembleMLP(members)
        handler = SystemDynamicsHandler(act_space, obs_space, dynamics_function=fn, true_model=False,
                                        is_normalized=True, seed=seed, precision=precision)
        handler.set_normalization(*w.stats)
    name = optimizer_name or w.optimizer_name
    args = dict(w.optimizer_args) if name == w.optimizer_name else {}
    args.update(planning_horizon=w.planning_horizon, population_size=w.population_size)
    if name != "RandomSearch":
        args["max_iterations"] = w.max_iterations or 5
    policy = MPCPolicy(reward_function=reward, env_action_space=act_space, env_observation_space=obs_space,
                       dynamics_handler=handler, optimizer_name=name, num_agents=w.num_agents, **args)
    return policy
