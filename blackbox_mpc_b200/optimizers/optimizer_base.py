"""OptimizerBase — host side of the six sampling optimizers.

Mirrors blackbox_mpc/optimizers/optimizer_base.py:5-115: constructor arguments, bounds handling
(:31-42), exploration statistics (:46-50), `__call__` = _optimize -> optional exploration noise +
clip -> predict_next_state -> evaluate_next_reward (:55-95), reset, set_trajectory_evaluator.

Built-in subclasses set KIND and run entirely inside libbbmpc (bbmpc_opt_*): sampling (Philox,
in-kernel), rollout, refit.  With a sharded population (`shard(rank, world)`), every iteration is
bbmpc_opt_iter_local -> one small all_gather over torch.distributed -> bbmpc_opt_iter_merge.
A user subclass that overrides `_optimize` in Python (KIND = None) still composes with the fused
evaluator through this base class."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from .. import _lib
from ..sharding import all_gather_partials


class OptimizerBase:
    KIND: Optional[int] = None

    def __init__(self, name, planning_horizon, max_iterations, num_agents, env_action_space, env_observation_space):
        self.name = name
        self._planning_horizon = int(planning_horizon)
        self._env_action_space, self._env_observation_space = env_action_space, env_observation_space
        self._dim_U = int(env_action_space.shape[0])
        self._dim_S = int(env_observation_space.shape[0])
        self._action_upper_bound = np.asarray(env_action_space.high, dtype=np.float32).reshape(-1)
        self._action_lower_bound = np.asarray(env_action_space.low, dtype=np.float32).reshape(-1)
        self._action_upper_bound_horizon = np.tile(self._action_upper_bound[None, :], (self._planning_horizon, 1))
        self._action_lower_bound_horizon = np.tile(self._action_lower_bound[None, :], (self._planning_horizon, 1))
        self._num_agents = int(num_agents)
        self._max_iterations = max_iterations
        self._trajectory_evaluator = None
        self._exploration_variance = (np.square(self._action_lower_bound - self._action_upper_bound) / 16) * 0.05
        self._exploration_mean = (self._action_upper_bound + self._action_lower_bound) / 2
        self._handle = None
        self._engine = None
        self._rank, self._world, self._group = 0, 1, None
        self._gather_buf = None
        self._partial = None
        self._p2p = False

    # -- hyper-parameters of the C handle; subclasses extend --------------------------------------
    def _config(self) -> dict:
        return {}

    def set_trajectory_evaluator(self, trajectory_evaluator):
        self._trajectory_evaluator = trajectory_evaluator
        self._destroy_handle()

    # -- C handle ---------------------------------------------------------------------------------
    def _destroy_handle(self):
        if self._handle is not None and self._engine is not None and self._engine.handle:
            self._engine.lib.bbmpc_opt_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._destroy_handle()
        except Exception:
            pass

    def _ensure_handle(self):
        if self.KIND is None:
            raise Exception("__call__ function is not implemented yet")
        ev = self._trajectory_evaluator
        if ev is None:
            raise RuntimeError("set_trajectory_evaluator() must be called before the optimizer is used")
        if not hasattr(ev, "engine"):
            raise TypeError("built-in optimizers need a DeterministicTrajectoryEvaluator (fused path)")
        e = ev.engine()  # stages model + reward if they changed
        if self._handle is None or self._engine is not e:
            self._engine = e
            cfg = _lib.OptConfig()
            cfg.kind = self.KIND
            cfg.population_size = int(self._population_size)
            cfg.num_agents, cfg.planning_horizon = self._num_agents, self._planning_horizon
            cfg.max_iterations = int(self._max_iterations) if self._max_iterations is not None else 1
            cfg.dS, cfg.dU = self._dim_S, self._dim_U
            self._lb_c = (C.c_float * self._dim_U)(*self._action_lower_bound.tolist())
            self._ub_c = (C.c_float * self._dim_U)(*self._action_upper_bound.tolist())
            cfg.lb_host, cfg.ub_host = self._lb_c, self._ub_c
            for k, v in self._config().items():
                setattr(cfg, k, v)
            h = C.c_void_p()
            e.check(e.lib.bbmpc_opt_create(e.handle, C.byref(cfg), C.byref(h)))
            self._handle = h
            if self._world > 1:
                e.check(e.lib.bbmpc_opt_set_shard(h, self._rank, self._world))
            self._alloc_io()
            self._p2p = self._connect_p2p() if self._world > 1 else False
        return self._engine

    def _alloc_io(self):
        dev = self._engine.device
        A = self._num_agents
        self._d_state = torch.empty(A, self._dim_S, dtype=torch.float32, device=dev)
        self._d_action = torch.empty(A, self._dim_U, dtype=torch.float32, device=dev)
        self._d_next = torch.empty(A, self._dim_S, dtype=torch.float32, device=dev)
        self._d_reward = torch.empty(A, dtype=torch.float32, device=dev)
        n = self._engine.lib.bbmpc_opt_partial_floats(self._handle)
        self._partial = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        self._gather_buf = torch.empty(self._world, max(n, 1), dtype=torch.float32, device=dev)

    def _connect_p2p(self) -> bool:
        """Peer-memory exchange (bbmpc_opt_p2p_*): every rank exports its exchange buffer as a CUDA IPC
        handle, the handles are all-gathered once, and from then on a sharded act() is ONE call into
        libbbmpc — partial messages are pulled over NVLink inside the merge-side kernel instead of an NCCL
        all_gather per iteration.  Needs all ranks on one node; BBMPC_P2P=0 keeps the all_gather path."""
        import os
        import torch.distributed as dist
        if os.environ.get("BBMPC_P2P", "1") == "0" or not (dist.is_available() and dist.is_initialized()):
            return False
        if dist.get_backend(self._group) != "nccl":
            return False
        if int(os.environ.get("LOCAL_WORLD_SIZE", self._world)) != self._world:
            return False                      # ranks on several nodes: IPC handles do not travel
        e, lib = self._engine, self._engine.lib
        ok = 1
        try:
            handle = (C.c_ubyte * 64)()
            e.check(lib.bbmpc_opt_p2p_export(self._handle, handle, None))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=e.device)
        except Exception:
            ok, mine = 0, torch.zeros(64, dtype=torch.uint8, device=e.device)
        gathered = torch.empty(self._world * 64, dtype=torch.uint8, device=e.device)
        dist.all_gather_into_tensor(gathered, mine, group=self._group)
        if ok:
            try:
                buf = gathered.cpu().numpy().tobytes()
                e.check(lib.bbmpc_opt_p2p_connect(self._handle, buf, None))
            except Exception:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=e.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._group)   # all ranks or none
        return bool(flag.item())

    # -- population sharding (one process per GPU) ------------------------------------------------
    def shard(self, rank: int, world: int, group=None):
        """Evaluate global population rows [rank*P/world, (rank+1)*P/world) on this GPU; `group`
        is the torch.distributed process group used for the per-iteration all_gather."""
        self._rank, self._world, self._group = int(rank), int(world), group
        self._destroy_handle()

    # -- reference API ----------------------------------------------------------------------------
    def _optimize(self, current_state, time_step):
        if self.KIND is None:
            raise Exception("__call__ function is not implemented yet")
        return self.__call__(current_state, time_step, False)[0]

    def __call__(self, current_state, time_step, add_exploration_noise):
        if self.KIND is None:
            return self._python_call(current_state, time_step, add_exploration_noise)
        e = self._ensure_handle()
        lib, h, st = e.lib, self._handle, e.stream()
        state = torch.as_tensor(current_state, dtype=torch.float32).to(e.device).contiguous()
        if tuple(state.shape) != (self._num_agents, self._dim_S):
            raise ValueError(f"current_state must be [{self._num_agents}, {self._dim_S}]")
        action, nxt, rew = (torch.empty_like(self._d_action), torch.empty_like(self._d_next), torch.empty_like(self._d_reward))
        noise = 1 if bool(add_exploration_noise) else 0
        if self._world == 1 or self._p2p:
            e.check(lib.bbmpc_opt_call(h, _lib.ptr(state), int(time_step), noise, _lib.ptr(action), _lib.ptr(nxt), _lib.ptr(rew), st))
        else:
            e.check(lib.bbmpc_opt_begin(h, _lib.ptr(state), int(time_step), st))
            for it in range(lib.bbmpc_opt_num_iterations(h)):
                e.check(lib.bbmpc_opt_iter_local(h, it, _lib.ptr(self._partial), st))
                all_gather_partials(self._partial, self._gather_buf, group=self._group)
                e.check(lib.bbmpc_opt_iter_merge(h, it, _lib.ptr(self._gather_buf), self._world, st))
            e.check(lib.bbmpc_opt_finish(h, noise, _lib.ptr(action), _lib.ptr(nxt), _lib.ptr(rew), st))
        return action, nxt, rew

    def call_host(self, observations: np.ndarray, time_step: int, add_exploration_noise: bool):
        """MPCPolicy.act's fast path: numpy [A,dS] in, numpy out, one synchronisation."""
        e = self._ensure_handle()
        if self._world != 1 and not self._p2p:
            a, n, r = self.__call__(torch.from_numpy(np.ascontiguousarray(observations, dtype=np.float32)), time_step, add_exploration_noise)
            return a.cpu().numpy(), n.cpu().numpy(), r.cpu().numpy()
        obs = np.ascontiguousarray(observations, dtype=np.float32)
        A = self._num_agents
        action = np.empty((A, self._dim_U), dtype=np.float32)
        nxt = np.empty((A, self._dim_S), dtype=np.float32)
        rew = np.empty((A,), dtype=np.float32)
        e.check(e.lib.bbmpc_opt_call_host(self._handle, obs.ctypes.data, int(time_step), 1 if add_exploration_noise else 0,
                                          action.ctypes.data, nxt.ctypes.data, rew.ctypes.data, e.stream()))
        return action, nxt, rew

    def _python_call(self, current_state, time_step, add_exploration_noise):
        """optimizer_base.py:80-95 for user subclasses whose `_optimize` is Python code."""
        ev = self._trajectory_evaluator
        action = self._optimize(current_state, time_step)
        if add_exploration_noise:
            dev = action.device
            mean = torch.as_tensor(self._exploration_mean, device=dev)
            std = torch.sqrt(torch.as_tensor(self._exploration_variance, device=dev))
            z = torch.empty(self._num_agents, self._dim_U, device=dev)
            torch.nn.init.trunc_normal_(z, 0.0, 1.0, -2.0, 2.0)
            action = torch.minimum(torch.maximum(action + (mean + std * z), torch.as_tensor(self._action_lower_bound, device=dev)),
                                   torch.as_tensor(self._action_upper_bound, device=dev))
        next_state = ev.predict_next_state(current_state, action)
        return action, next_state, ev.evaluate_next_reward(current_state, next_state, action)

    def reset(self):
        if self.KIND is None:
            raise Exception("reset function is not implemented yet")
        e = self._ensure_handle()
        e.check(e.lib.bbmpc_opt_reset(self._handle, e.stream()))

    # -- inspection (tests) -----------------------------------------------------------------------
    def get_tensor(self, name: str) -> torch.Tensor:
        e = self._ensure_handle()
        n = e.check(e.lib.bbmpc_opt_get_tensor(self._handle, name.encode(), None, 0, e.stream()))
        out = torch.empty(n, dtype=torch.float32, device=e.device)
        e.check(e.lib.bbmpc_opt_get_tensor(self._handle, name.encode(), _lib.ptr(out), n, e.stream()))
        return out

    def enable_sample_trace(self, n_iters: Optional[int] = None) -> torch.Tensor:
        e = self._ensure_handle()
        per_iter = self.get_tensor("samples").numel()
        n_iters = e.lib.bbmpc_opt_num_iterations(self._handle) if n_iters is None else n_iters
        self._trace = torch.zeros(n_iters, per_iter, dtype=torch.float32, device=e.device)
        e.check(e.lib.bbmpc_opt_set_sample_trace(self._handle, _lib.ptr(self._trace), self._trace.numel()))
        return self._trace

    def set_draw_injection(self, std_draws: Optional[torch.Tensor]) -> None:
        """Test hook: feed the samplers STANDARD variates (one [P, A, H*dU] block per optimizer iteration) instead of
        Philox (bbmpc_opt_set_draw_injection); None switches back.  The tensor is kept alive by the optimizer."""
        e = self._ensure_handle()
        if std_draws is None:
            self._inject = None
            e.check(e.lib.bbmpc_opt_set_draw_injection(self._handle, None, 0))
            return
        self._inject = std_draws.to(device=e.device, dtype=torch.float32).contiguous()
        e.check(e.lib.bbmpc_opt_set_draw_injection(self._handle, _lib.ptr(self._inject), self._inject.numel()))
