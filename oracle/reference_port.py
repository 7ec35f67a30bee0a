"""CPU restatement (torch-CPU, dtype-parametric) of the reference hot path.

TEST INFRASTRUCTURE — see oracle/__init__.py.  PARITY UNPINNED (no reference tests/fixtures; TF
not installable).  All citations are relative to /root/reference/blackbox_mpc/ unless a path
starts with tutorials/.

Random draws are *injected*: every sampler call goes through a `draws` object so that a test can
feed the exact samples the CUDA path produced (the reference never seeds tf.random, so there is
no reference bit-stream to match; TF semantics of each sampler are restated in `TorchDraws`).

`dtype=torch.float64` gives the "truth" twin, `torch.float32` the expected-rounding twin.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

__all__ = [
    "Space", "MLP", "Ensemble", "PendulumTrueModel", "pendulum_reward_function",
    "halfcheetah_reward_function", "Handler", "Evaluator", "TorchDraws", "InjectedDraws",
    "OptimizerBase", "CEM", "PI2", "RandomSearch", "PSO", "SPSA", "CMAES", "policy_act",
    "ACTIVATIONS",
]


class Space:
    """Stand-in for a gym Box: the reference reads only .shape[0], .high, .low
    (optimizers/optimizer_base.py:31-36)."""

    def __init__(self, low, high):
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape


ACTIVATIONS: Dict[Optional[str], Optional[Callable]] = {
    None: None, "linear": None, "tanh": torch.tanh, "relu": torch.relu, "sigmoid": torch.sigmoid,
}


# --------------------------------------------------------------------------- L0: models + rewards
class MLP:
    """dynamics_functions/deterministic_mlp.py:20-24 (Dense chain), :49-51 (__call__).
    Keras Dense = act(x @ W + b), W [in, out]."""

    def __init__(self, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                 activations: Sequence[Optional[str]]):
        self.weights, self.biases, self.activations = list(weights), list(biases), list(activations)

    def to(self, dtype):
        return MLP([w.to(dtype) for w in self.weights], [b.to(dtype) for b in self.biases],
                   self.activations)

    def __call__(self, x, train=False):
        for w, b, a in zip(self.weights, self.biases, self.activations):
            x = x @ w + b
            fn = ACTIVATIONS[a]
            if fn is not None:
                x = fn(x)
        return x


class Ensemble:
    """Not a reference type: a plain reference-API dynamics_function that averages
    the raw outputs of n DeterministicMLP members, summed in member order then divided by n."""

    def __init__(self, members: Sequence[MLP]):
        self.members = list(members)

    def to(self, dtype):
        return Ensemble([m.to(dtype) for m in self.members])

    def __call__(self, x, train=False):
        acc = self.members[0](x)
        for m in self.members[1:]:
            acc = acc + m(x)
        return acc / len(self.members)


class PendulumTrueModel:
    """utils/pendulum.py:50-56 (constants), :78-92 (__call__): returns the deviation new - s."""

    def __init__(self):
        self.g, self.max_speed, self.m, self.l, self.dt = 10.0, 8.0, 1.0, 1.0, 0.05

    def to(self, dtype):
        return self

    def __call__(self, x, train=False):
        dt_ = x.dtype
        c = lambda v: torch.tensor(v, dtype=dt_)  # noqa: E731  (tf.constant(..., float32))
        u, thdot, th_cos, th_sin = x[:, 3], x[:, 2], x[:, 0], x[:, 1]
        theta = torch.atan2(th_sin, th_cos)
        pi = c(float(np.pi))
        # pendulum.py:83-85: (-3*g/(2*l) * sin(theta+pi) + 3/(m*l**2) * u) * dt
        newthdot = thdot + (-c(3.0) * c(self.g) / (c(2.0) * c(self.l)) * torch.sin(theta + pi)
                            + c(3.0) / (c(self.m) * c(self.l) ** c(2.0)) * u) * c(self.dt)
        newth = theta + newthdot * c(self.dt)
        newthdot = torch.clamp(newthdot, -self.max_speed, self.max_speed)  # clipped AFTER newth
        new_state = torch.stack([torch.cos(newth), torch.sin(newth), newthdot], dim=1)
        return new_state - x[:, :3]


def _pendulum_angle_normalize(x):
    """utils/pendulum.py:5-7; `%` on tf floats is floormod (divisor's sign) = torch.remainder."""
    return torch.remainder(x + np.pi, 2 * np.pi) - np.pi


def pendulum_reward_function(current_state, next_state, actions):
    """utils/pendulum.py:10-35.  NOTE the declared order (current, next_state, actions): the
    evaluator passes (current, actions, next) positionally (deterministic.py:65-66), so the
    parameter named `actions` here receives next_state.  Restated verbatim so the quirk falls
    out of the call, not of this function."""
    ang = _pendulum_angle_normalize(torch.atan2(current_state[:, 1], current_state[:, 0]))
    return -(ang ** 2 + 0.1 * current_state[:, 2] ** 2) - 0.001 * torch.sum(actions * actions, dim=1)


def halfcheetah_reward_function(current_state, actions, next_state):
    """tutorials/mujoco/cost_func.py:5-22."""
    rewards = torch.zeros(current_state.shape[0], dtype=current_state.dtype)
    rewards = torch.where(current_state[:, 5] >= 0.2, rewards + (-10), rewards)
    rewards = torch.where(current_state[:, 6] >= 0.0, rewards + (-10), rewards)
    rewards = torch.where(current_state[:, 7] >= 0.0, rewards + (-10), rewards)
    rewards = rewards + ((next_state[:, 17] - current_state[:, 17]) / 0.01)
    rewards = rewards - (0.0 * torch.sum(actions * actions, dim=1))
    return rewards


# --------------------------------------------------------------------------- L1: handler (inference half)
class Handler:
    """dynamics_handlers/system_dynamics_handler.py:97-126 (process_input), :128-161
    (process_output); utils/transforms.py:34 (inverse transform = delta + state)."""

    def __init__(self, dynamics_function, true_model=False, is_normalized=True, stats=None,
                 dtype=torch.float32):
        self._dynamics_function = dynamics_function.to(dtype)
        self._is_true_model, self._is_normalized = true_model, is_normalized
        self.dtype = dtype
        if stats is not None:
            (self._mean_states, self._std_states, self._mean_actions, self._std_actions,
             self._mean_targets, self._std_targets) = [torch.as_tensor(s).to(dtype) for s in stats]

    def process_input(self, states, actions):
        if self._is_true_model or not self._is_normalized:
            return torch.cat([states, actions], dim=-1)
        new_states = (states - self._mean_states) / (self._std_states + 1e-7)
        new_actions = (actions - self._mean_actions) / (self._std_actions + 1e-7)
        return torch.cat([new_states, new_actions], dim=-1)

    def process_output(self, input_states, raw_output):
        if self._is_true_model or not self._is_normalized:
            deviation = raw_output
        else:
            deviation = self._mean_targets + raw_output * (self._std_targets + 1e-7)
        return deviation + input_states


# --------------------------------------------------------------------------- L2: evaluator
class Evaluator:
    """trajectory_evaluators/deterministic.py:26-77 (__call__), :79-103, :105-127."""

    def __init__(self, reward_function, system_dynamics_handler: Handler):
        self._reward_function = reward_function
        self._system_dynamics_handler = system_dynamics_handler

    def __call__(self, current_states, action_sequences, time_step=0):
        nopt, n_agents, horizon, dim_u = action_sequences.shape
        rewards = torch.zeros(nopt * n_agents, dtype=current_states.dtype)
        seq = action_sequences.reshape(-1, horizon, dim_u).permute(1, 0, 2)  # [H, P*A, dU]
        state = current_states.repeat(nopt, 1)  # tf.tile -> row p*A+a starts at states[a]
        for t in range(horizon):
            actions = seq[t]
            next_state = self.predict_next_state(state, actions)
            rewards = rewards + self._reward_function(state, actions, next_state)
            state = next_state
        rewards = rewards.reshape(nopt, n_agents)
        return torch.where(torch.isnan(rewards), torch.full_like(rewards, -1e6), rewards)

    def predict_next_state(self, current_states, current_actions):
        h = self._system_dynamics_handler
        x = h.process_input(current_states, current_actions)
        raw = h._dynamics_function(x, train=False)
        return h.process_output(current_states, raw)

    def evaluate_next_reward(self, current_states, next_states, current_actions):
        return self._reward_function(current_states, current_actions, next_states)


# --------------------------------------------------------------------------- samplers
class TorchDraws:
    """TF sampler semantics [TF] on a seeded torch generator.
    truncated_normal: N(0,1) redrawn until |z| <= 2, then mean + std*z (mean/std broadcast over
    the leading population axis); uniform: lo + (hi-lo)*U[0,1); normal: N(0,1); rademacher:
    randint(0,2)*2-1 (optimizers/spsa.py:73-75)."""

    def __init__(self, seed=0, dtype=torch.float32):
        self.gen = torch.Generator().manual_seed(seed)
        self.dtype = dtype
        self.log: List = []

    def _std_truncnorm(self, shape):
        z = torch.randn(shape, generator=self.gen, dtype=torch.float64)
        bad = z.abs() > 2
        while bad.any():
            z[bad] = torch.randn(int(bad.sum()), generator=self.gen, dtype=torch.float64)
            bad = z.abs() > 2
        return z.to(self.dtype)

    def truncated_normal(self, shape, mean, std, tag=""):
        out = mean + std * self._std_truncnorm(tuple(shape))
        self.log.append((tag, out))
        return out

    def uniform(self, shape, lo, hi, tag=""):
        out = lo + (hi - lo) * torch.rand(tuple(shape), generator=self.gen, dtype=torch.float64).to(self.dtype)
        self.log.append((tag, out))
        return out

    def normal(self, shape, tag=""):
        out = torch.randn(tuple(shape), generator=self.gen, dtype=torch.float64).to(self.dtype)
        self.log.append((tag, out))
        return out

    def rademacher(self, shape, tag=""):
        out = (torch.randint(0, 2, tuple(shape), generator=self.gen) * 2 - 1).to(self.dtype)
        self.log.append((tag, out))
        return out


class InjectedDraws:
    """Feeds pre-recorded final samples (e.g. dumped from the CUDA path) in call order per tag."""

    def __init__(self, recorded: Dict[str, List[torch.Tensor]], dtype=torch.float32):
        self.rec = {k: list(v) for k, v in recorded.items()}
        self.dtype = dtype

    def _pop(self, tag, shape):
        out = torch.as_tensor(self.rec[tag].pop(0)).to(self.dtype)
        assert tuple(out.shape) == tuple(shape), (tag, out.shape, shape)
        return out

    def truncated_normal(self, shape, mean, std, tag=""):
        return self._pop(tag, shape)

    def uniform(self, shape, lo, hi, tag=""):
        return self._pop(tag, shape)

    def normal(self, shape, tag=""):
        return self._pop(tag, shape)

    def rademacher(self, shape, tag=""):
        return self._pop(tag, shape)


def _topk_desc(values, k):
    """tf.nn.top_k / argsort(DESCENDING) / argmax tie rule [TF]: lowest index first."""
    order = torch.sort(-values, dim=-1, stable=True).indices
    return order[..., :k]


def _shift_left(x):
    """concat([x[:,1:], x[:,-1:]], 1)  (pi2.py:92-93, spsa.py:114-115, pso.py:123-125)."""
    return torch.cat([x[:, 1:], x[:, -1:]], dim=1)


# --------------------------------------------------------------------------- L3: optimizers
class OptimizerBase:
    """optimizers/optimizer_base.py:6-50 (state), :55-95 (__call__)."""

    def __init__(self, planning_horizon, max_iterations, num_agents, env_action_space,
                 env_observation_space, dtype=torch.float32):
        self.dtype = dtype
        self._planning_horizon = int(planning_horizon)
        self._dim_U = int(env_action_space.shape[0])
        self._dim_S = int(env_observation_space.shape[0])
        f32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float32))  # noqa: E731
        self._ub = f32(env_action_space.high).to(dtype)
        self._lb = f32(env_action_space.low).to(dtype)
        self._ub_h = self._ub[None, :].repeat(self._planning_horizon, 1)
        self._lb_h = self._lb[None, :].repeat(self._planning_horizon, 1)
        self._num_agents = int(num_agents)
        self._max_iterations = max_iterations
        self._trajectory_evaluator: Optional[Evaluator] = None
        self._exploration_variance = ((self._lb - self._ub) ** 2 / 16) * 0.05
        self._exploration_mean = (self._ub + self._lb) / 2

    def set_trajectory_evaluator(self, ev):
        self._trajectory_evaluator = ev

    def _midpoint(self):
        return ((self._lb + self._ub) / 2)[None, None, :].repeat(self._num_agents, self._planning_horizon, 1)

    def _init_variance(self):
        return ((self._lb - self._ub) ** 2 / 16)[None, None, :].repeat(self._num_agents, self._planning_horizon, 1)

    def _clip_h(self, x):
        return torch.minimum(torch.maximum(x, self._lb_h), self._ub_h)

    def _penalty(self, raw, feasible, pop):
        """tf.norm(reshape(raw-feasible,[P,A,-1]),axis=2)**2 — (sqrt(sum sq))**2 [TF]."""
        d = (raw - feasible).reshape(pop, self._num_agents, -1)
        return torch.sqrt(torch.sum(d * d, dim=2)) ** 2

    def _optimize(self, current_state, time_step, draws):
        raise Exception("__call__ function is not implemented yet")

    def __call__(self, current_state, time_step, add_exploration_noise, draws):
        action = self._optimize(current_state, time_step, draws)
        if add_exploration_noise:
            noise = draws.truncated_normal([self._num_agents, self._dim_U], self._exploration_mean,
                                           torch.sqrt(self._exploration_variance), tag="explore")
            action = torch.minimum(torch.maximum(action + noise, self._lb), self._ub)
        next_state = self._trajectory_evaluator.predict_next_state(current_state, action)
        reward = self._trajectory_evaluator.evaluate_next_reward(current_state, next_state, action)
        return action, next_state, reward


class CEM(OptimizerBase):
    """optimizers/cem.py:7-72 (init), :74-136 (_optimize), :138-149 (reset)."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_elite=50, num_agents=5, epsilon=0.001, alpha=0.25,
                 dtype=torch.float32):
        super().__init__(planning_horizon, max_iterations, num_agents, env_action_space,
                         env_observation_space, dtype)
        self._population_size, self._num_elite = population_size, num_elite
        self._epsilon, self._alpha = epsilon, alpha  # epsilon stored, never used (cem.py:53)
        self._previous_solution = self._midpoint()
        self._solution_variance = self._init_variance()
        self.trace: List[dict] = []

    def _optimize(self, current_state, time_step, draws):
        mean, variance = self._previous_solution, self._solution_variance
        alpha = torch.tensor(self._alpha, dtype=self.dtype)
        one = torch.tensor(1.0, dtype=self.dtype)
        two = torch.tensor(2.0, dtype=self.dtype)
        self.trace = []
        for _ in range(self._max_iterations):
            cvar = torch.minimum(torch.minimum(((mean - self._lb_h) / two) ** 2,
                                               ((self._ub_h - mean) / two) ** 2), variance)
            samples = draws.truncated_normal(
                [self._population_size, self._num_agents, self._planning_horizon, self._dim_U],
                mean, torch.sqrt(cvar), tag="cem.samples")
            rewards = self._trajectory_evaluator(current_state, samples, time_step).t()  # [A,P]
            idx = _topk_desc(rewards, self._num_elite)  # [A,E]
            per_agent = samples.permute(1, 0, 2, 3)  # [A,P,H,dU]
            elites = torch.stack([per_agent[a][idx[a]] for a in range(self._num_agents)], 0)
            new_mean = elites.mean(dim=1)
            new_var = ((elites - new_mean[:, None]) ** 2).mean(dim=1)  # ddof = 0
            mean = alpha * mean + (one - alpha) * new_mean
            variance = alpha * variance + (one - alpha) * new_var
            self.trace.append(dict(samples=samples, rewards=rewards, elite_idx=idx, mean=mean,
                                   variance=variance))
        # cem.py:133-134: the assign is commented out -> no warm start between act() calls.
        return mean[:, 0]

    def reset(self):
        self._previous_solution = self._midpoint()  # variance NOT reset (cem.py:138-149)


class PI2(OptimizerBase):
    """optimizers/pi2.py:9-56 (init), :58-96 (_optimize), :98-105 (reset)."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, lamda=1.0, dtype=torch.float32):
        super().__init__(planning_horizon, max_iterations, num_agents, env_action_space,
                         env_observation_space, dtype)
        self._population_size, self._lamda = population_size, lamda
        self._previous_solution = self._midpoint()
        self._solution_variance = self._init_variance()
        self.trace: List[dict] = []

    def _optimize(self, current_state, time_step, draws):
        mean = self._previous_solution
        lam = torch.tensor(self._lamda, dtype=self.dtype)
        self.trace = []
        for _ in range(self._max_iterations):
            samples = draws.truncated_normal(
                [self._population_size, self._num_agents, self._planning_horizon, self._dim_U],
                mean, torch.sqrt(self._solution_variance), tag="pi2.samples")
            feasible = self._clip_h(samples)
            penalty = self._penalty(samples, feasible, self._population_size)
            samples = feasible
            rewards = self._trajectory_evaluator(current_state, samples, time_step) - penalty
            costs = (-rewards).t()  # [A,P]
            beta = costs.min(dim=1).values
            prob = torch.exp(-(1 / lam) * (costs - beta[:, None]))
            eta = prob.sum(dim=1)
            omega = (1 / eta)[:, None] * prob
            per_agent = samples.permute(1, 0, 2, 3)
            mean = (per_agent * omega[:, :, None, None]).sum(dim=1)
            self.trace.append(dict(samples=samples, rewards=rewards, omega=omega, mean=mean))
        self._previous_solution = _shift_left(mean)
        return mean[:, 0]

    def reset(self):
        self._previous_solution = self._midpoint()


class RandomSearch(OptimizerBase):
    """optimizers/random_search.py:7-36 (init), :38-48 (_optimize)."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50,
                 population_size=1024, num_agents=5, dtype=torch.float32):
        super().__init__(planning_horizon, None, num_agents, env_action_space, env_observation_space, dtype)
        self._population_size = population_size
        self.trace: List[dict] = []

    def _optimize(self, current_state, time_step, draws):
        samples = draws.uniform(
            [self._population_size, self._num_agents, self._planning_horizon, self._dim_U],
            self._lb_h, self._ub_h, tag="rs.samples")
        rewards = self._trajectory_evaluator(current_state, samples, time_step)  # [P,A]
        best = _topk_desc(rewards.t(), 1)[:, 0]  # argmax over P, first index on ties [TF]
        per_agent = samples.permute(1, 0, 2, 3)
        self.trace = [dict(samples=samples, rewards=rewards, best=best)]
        return torch.stack([per_agent[a, best[a], 0] for a in range(self._num_agents)], 0)

    def reset(self):
        return


class PSO(OptimizerBase):
    """optimizers/pso.py:7-68 (init: every Variable starts at ZERO), :70-141 (_optimize incl. the
    re-seed tail), :143-160 (reset)."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, c1=0.3, c2=0.5, w=0.2,
                 initial_velocity_fraction=0.01, dtype=torch.float32):
        super().__init__(planning_horizon, max_iterations, num_agents, env_action_space,
                         env_observation_space, dtype)
        P, A, H, U = population_size, self._num_agents, self._planning_horizon, self._dim_U
        self._population_size = P
        z = lambda *s: torch.zeros(*s, dtype=dtype)  # noqa: E731
        self._x, self._v, self._pbest_x = z(P, A, H, U), z(P, A, H, U), z(P, A, H, U)
        self._pbest_r, self._gbest_x, self._gbest_r = z(P, A), z(A, H, U), z(A)
        self._solution_variance = self._init_variance()
        self._c1, self._c2, self._w, self._v0 = c1, c2, w, initial_velocity_fraction
        self._solution = z(A, U)
        self.trace: List[dict] = []

    def _c(self, v):
        return torch.tensor(v, dtype=self.dtype)

    def _optimize(self, current_state, time_step, draws):
        P, A = self._population_size, self._num_agents
        self.trace = []
        for _ in range(self._max_iterations):
            feasible = self._clip_h(self._x)
            penalty = self._penalty(self._x, feasible, P)
            self._x = feasible
            rewards = self._trajectory_evaluator(current_state, self._x, time_step) - penalty
            cond = self._pbest_r < rewards
            self._pbest_x = torch.where(cond[:, :, None, None], self._x, self._pbest_x)
            self._pbest_r = torch.where(cond, rewards, self._pbest_r)
            best = _topk_desc(self._pbest_r.t(), 1)[:, 0]  # argmax over axis 0 (pso.py:97)
            per_agent = self._pbest_x.permute(1, 0, 2, 3)
            self._gbest_x = torch.stack([per_agent[a, best[a]] for a in range(A)], 0)
            # pso.py:99,102-103: index a*P+p applied to a row-major [P,A] flatten (quirk (c)).
            flat_idx = best + torch.arange(A) * P
            self._gbest_r = self._pbest_r.reshape(-1)[flat_idx]
            r1 = draws.normal([], tag="pso.r1")
            r2 = draws.normal([], tag="pso.r2")
            self._v = (self._v * self._c(self._w)) + (self._pbest_x - self._x) * self._c(self._c1) * r1 \
                + (self._gbest_x - self._x) * self._c(self._c2) * r2
            self._x = self._x + self._v
            self.trace.append(dict(rewards=rewards, best=best, gbest_x=self._gbest_x, x=self._x, v=self._v))
        self._solution = self._gbest_x[:, 0, :]
        two = self._c(2.0)
        cvar = torch.minimum(torch.minimum(((self._gbest_x - self._lb_h) / two) ** 2,
                                           ((self._ub_h - self._gbest_x) / two) ** 2),
                             self._solution_variance)  # un-shifted gbest (quirk (d))
        shape = [P, A, self._planning_horizon, self._dim_U]
        pos = draws.truncated_normal(shape, _shift_left(self._gbest_x), torch.sqrt(cvar), tag="pso.reseed_x")
        v0 = self._c(self._v0) * (self._ub_h - self._lb_h)
        vel = draws.uniform(shape, -v0, v0, tag="pso.reseed_v")
        self._x, self._v, self._pbest_x = pos, vel, pos
        self._pbest_r = torch.full((P, A), -float("inf"), dtype=self.dtype)
        self._gbest_r = torch.full((A,), -float("inf"), dtype=self.dtype)
        return self._solution

    def reset(self, draws):
        P, A = self._population_size, self._num_agents
        shape = [P, A, self._planning_horizon, self._dim_U]
        pos = draws.uniform(shape, self._lb_h, self._ub_h, tag="pso.reset_x")
        v0 = self._c(self._v0) * (self._ub_h - self._lb_h)
        vel = draws.uniform(shape, -v0, v0, tag="pso.reset_v")
        self._x, self._v, self._pbest_x = pos, vel, pos
        self._pbest_r = torch.full((P, A), -float("inf"), dtype=self.dtype)
        self._gbest_r = torch.full((A,), -float("inf"), dtype=self.dtype)


class SPSA(OptimizerBase):
    """optimizers/spsa.py:7-59 (init), :61-117 (_optimize), :119-127 (reset)."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_agents=5, alpha=0.602, gamma=0.101, a_par=0.01,
                 noise_parameter=0.3, dtype=torch.float32):
        super().__init__(planning_horizon, max_iterations, num_agents, env_action_space,
                         env_observation_space, dtype)
        self._population_size = population_size
        c = lambda v: torch.tensor(v, dtype=dtype)  # noqa: E731
        self._alpha, self._gamma, self._a_par, self._noise = c(alpha), c(gamma), c(a_par), c(noise_parameter)
        self._big_a = c(float(max_iterations)) / c(10.0)
        self._current_parameters = self._midpoint()
        self.trace: List[dict] = []

    def _optimize(self, current_state, time_step, draws):
        P = self._population_size
        sol = self._current_parameters
        self.trace = []
        for t in range(self._max_iterations):
            tf_ = torch.tensor(float(t), dtype=self.dtype)
            ak = self._a_par / (tf_ + 1 + self._big_a) ** self._alpha
            ck = self._noise / (tf_ + 1) ** self._gamma
            delta = draws.rademacher([P, self._num_agents, self._planning_horizon, self._dim_U], tag="spsa.delta")
            plus, minus = sol + ck * delta, sol - ck * delta
            plus_f, minus_f = self._clip_h(plus), self._clip_h(minus)
            pen_p, pen_m = self._penalty(plus, plus_f, P), self._penalty(minus, minus_f, P)
            full = self._trajectory_evaluator(current_state, torch.cat([plus_f, minus_f], 0), time_step)
            r_p, r_m = full[:P] - pen_p, full[P:] - pen_m
            ghat = ((r_p - r_m)[:, :, None, None] / (2.0 * ck * delta)).mean(dim=0)
            sol = self._clip_h(sol + ak * ghat)
            self.trace.append(dict(r_plus=r_p, r_minus=r_m, ghat=ghat, sol=sol, ak=ak, ck=ck))
        self._current_parameters = _shift_left(sol)
        return sol[:, 0]

    def reset(self):
        self._current_parameters = self._midpoint()


class CMAES(OptimizerBase):
    """optimizers/cma_es.py:7-127 (init + constants), :129-213 (_optimize), :215-227 (reset).
    sigma is a per-coordinate VECTOR (:97); y = z @ (B@D) (:140); rewards are summed over agents
    (:158); svd(C) -> (s, U): D = diag(sqrt(s)), B = U (:195-198).  `eig_fn(C) -> (s desc, U)`
    is injectable because eigenvector signs / degenerate subspaces are solver-specific."""

    def __init__(self, env_action_space, env_observation_space, planning_horizon=50, max_iterations=5,
                 population_size=500, num_elite=50, h_sigma=1.0, alpha_cov=2.0, num_agents=5,
                 dtype=torch.float32, eig_fn=None):
        super().__init__(planning_horizon, max_iterations, num_agents, env_action_space,
                         env_observation_space, dtype)
        self._population_size, self._num_elite = population_size, num_elite
        n = self._num_agents * self._planning_horizon * self._dim_U
        self._n = n
        nf = torch.tensor(float(n), dtype=dtype)
        e = torch.tensor(float(num_elite), dtype=dtype)
        w = torch.log(e + 0.5) - torch.log(torch.arange(1, num_elite + 1, dtype=dtype))
        w = torch.cat([w, torch.zeros(population_size - num_elite, dtype=dtype)])
        self._weights = (w / w.sum())[:, None]
        self._mu_eff = self._weights.sum() ** 2 / (self._weights ** 2).sum()
        self._c_sigma = (self._mu_eff + 2) / (nf + self._mu_eff + 5)
        self._d_sigma = 1 + 2 * torch.clamp(torch.sqrt((self._mu_eff - 1) / (nf + 1)) - 1, min=0) + self._c_sigma
        self._cc = (4 + self._mu_eff / nf) / (nf + 4 + 2 * self._mu_eff / nf)
        self._alpha_cov, self._h_sigma = alpha_cov, h_sigma
        self._c1 = alpha_cov / ((nf + 1.3) ** 2 + self._mu_eff)
        c_mu2 = alpha_cov * (self._mu_eff - 2 + 1 / self._mu_eff) / ((nf + 2) ** 2 + alpha_cov * self._mu_eff / 2)
        self._c_mu = torch.minimum(1 - self._c1, c_mu2)
        self._m = self._midpoint().reshape(-1)
        self._sigma = torch.sqrt(self._init_variance().reshape(-1))
        self._C = torch.eye(n, dtype=dtype)
        self._p_sigma, self._p_C = torch.zeros(n, dtype=dtype), torch.zeros(n, dtype=dtype)
        self._B, self._D = torch.eye(n, dtype=dtype), torch.eye(n, dtype=dtype)
        self._expectation_of_normal = torch.sqrt(nf * (1 - 1 / (4 * nf) + 1 / (21 * nf ** 2)))
        self._eig_fn = eig_fn or self._svd
        self.trace: List[dict] = []

    @staticmethod
    def _svd(C):
        U, s, _ = torch.linalg.svd(C)
        return s, U

    def _optimize(self, current_state, time_step, draws):
        P, n = self._population_size, self._n
        self.trace = []
        for _ in range(self._max_iterations):
            z = draws.normal([P, n], tag="cmaes.z")
            y = z @ (self._B @ self._D)
            samples = (self._m + self._sigma * y).reshape(P, self._num_agents, self._planning_horizon, self._dim_U)
            feasible = self._clip_h(samples)
            penalty = self._penalty(samples, feasible, P)
            samples = feasible
            rewards = self._trajectory_evaluator(current_state, samples, time_step) - penalty
            rewards = rewards.sum(dim=1)
            order = _topk_desc(rewards, P)
            x_sorted = samples[order]
            x_diff = x_sorted.reshape(P, n) - self._m
            x_mean = (x_diff * self._weights).sum(dim=0)
            m = self._m + x_mean
            y_mean = x_mean / self._sigma
            D_inv = torch.diag(1.0 / torch.diagonal(self._D))
            C_inv_half = (self._B @ D_inv) @ self._B.t()
            p_sigma = (1 - self._c_sigma) * self._p_sigma + \
                torch.sqrt(self._c_sigma * (2 - self._c_sigma) * self._mu_eff) * (C_inv_half @ y_mean)
            sigma = self._sigma * torch.exp((self._c_sigma / self._d_sigma) *
                                            (torch.linalg.vector_norm(p_sigma) / self._expectation_of_normal - 1))
            p_C = (1 - self._cc) * self._p_C + self._h_sigma * torch.sqrt(self._cc * (2 - self._cc) * self._mu_eff) * y_mean
            E = self._num_elite  # weights are zero beyond E: the reference's [P,N,N] map_fn reduces to this
            yu = x_diff[:E] / self._sigma
            y_s = (yu * self._weights[:E]).t() @ yu
            C = (1 - self._c1 - self._c_mu) * self._C + self._c1 * torch.outer(p_C, p_C) + self._c_mu * y_s
            C_upper = torch.triu(C)
            C = C_upper + (C_upper - torch.diag(torch.diagonal(C_upper))).t()
            s, U = self._eig_fn(C)
            self._p_C, self._p_sigma, self._C, self._sigma = p_C, p_sigma, C, sigma
            self._B, self._D, self._m = U, torch.diag(torch.sqrt(s)), m
            self.trace.append(dict(z=z, samples=samples, rewards=rewards, order=order[:E], m=m, sigma=sigma,
                                   p_sigma=p_sigma, p_C=p_C, C=C, s=s, B=U))
        return self._m.reshape(self._num_agents, self._planning_horizon, self._dim_U)[:, 0]

    def reset(self):
        self._m = self._midpoint().reshape(-1)
        self._sigma = torch.sqrt(self._init_variance().reshape(-1))  # C,B,D,p_* NOT reset (:215-227)


# --------------------------------------------------------------------------- L4: policy marshalling
def policy_act(optimizer: OptimizerBase, observations: np.ndarray, t: int, draws, exploration_noise=False):
    """policies/mpc_policy.py:149-172: 1-D obs tiled to [A,dS]; outputs un-batched again."""
    obs = np.array(observations)
    batched = np.tile(obs[None, :], (optimizer._num_agents, 1)) if obs.ndim == 1 else obs
    state = torch.as_tensor(batched).to(optimizer.dtype)
    action, next_state, reward = optimizer(state, t, exploration_noise, draws)
    action, next_state = action.numpy(), next_state.numpy()
    if obs.ndim == 1:
        action, next_state, reward = action[0], next_state[0], reward[0]
    return action, next_state, reward
