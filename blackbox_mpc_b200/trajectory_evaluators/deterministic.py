"""DeterministicTrajectoryEvaluator — the rollout hot path.

Mirrors blackbox_mpc/trajectory_evaluators/deterministic.py:6-127.  `__call__` is one launch of the
fused sm_100a rollout kernel (bbmpc_rollout): process_input -> dynamics -> process_output ->
reward, H times, with trajectory state on chip; NaN returns -> -1e6 (:75-77).  `time_step` is
accepted and ignored, as in the reference."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from .evaluator_base import EvaluatorBase


def reward_id_of(reward_function) -> int:
    rid = getattr(reward_function, "bbmpc_reward_id", None)
    if rid is None and isinstance(getattr(reward_function, "cuda_source", None), str):
        rid = _lib.REWARD_USER     # any callable that carries its CUDA source (utils/rewards.py)
    if rid is None:
        raise TypeError(
            "reward_function must be a built-in device reward (utils.pendulum.pendulum_reward_function, "
            "utils.halfcheetah.reward_function, ...) or carry CUDA source (utils.rewards.cuda_reward(...), or any callable "
            "with a `cuda_source` attribute): a plain Python callable cannot run inside the sm_100a rollout kernel")
    return int(rid)


class DeterministicTrajectoryEvaluator(EvaluatorBase):
    def __init__(self, reward_function, system_dynamics_handler):
        super().__init__(reward_function=reward_function, system_dynamics_handler=system_dynamics_handler, name=None)
        self._reward_id = reward_id_of(reward_function)

    # -- engine plumbing ----------------------------------------------------------------------
    def engine(self):
        e = self._system_dynamics_handler.ensure_staged()
        if self._reward_id == _lib.REWARD_USER:   # compiled once per engine and source; a no-op afterwards
            e.check(e.lib.bbmpc_reward_set_nvrtc(e.handle, self._reward_function.cuda_source.encode()))
        else:
            e.check(e.lib.bbmpc_reward_set_builtin(e.handle, self._reward_id))
        return e

    def _dev(self, x):
        e = self._system_dynamics_handler.engine
        if not torch.is_tensor(x):
            x = torch.as_tensor(np.asarray(x, dtype=np.float32))
        return x.to(device=e.device, dtype=torch.float32).contiguous()

    # -- reference API ------------------------------------------------------------------------
    def __call__(self, current_states, action_sequences, time_step=0):
        e = self.engine()
        s, a = self._dev(current_states), self._dev(action_sequences)
        if a.dim() != 4 or s.dim() != 2 or a.shape[1] != s.shape[0]:
            raise ValueError("expected current_states [A,dS] and action_sequences [P,A,H,dU]")
        P, A, H, _ = a.shape
        out = torch.empty(P, A, dtype=torch.float32, device=e.device)
        e.check(e.lib.bbmpc_rollout(e.handle, _lib.ptr(s), _lib.ptr(a), _lib.ptr(out), P, A, H, e.stream()))
        return out

    def predict_next_state(self, current_states, current_actions):
        e = self.engine()
        s, a = self._dev(current_states), self._dev(current_actions)
        out = torch.empty_like(s)
        e.check(e.lib.bbmpc_predict_next_state(e.handle, _lib.ptr(s), _lib.ptr(a), _lib.ptr(out), s.shape[0], e.stream()))
        return out

    def evaluate_next_reward(self, current_states, next_states, current_actions):
        """reward_function(current, actions, next) — note the argument order (:105-127)."""
        e = self.engine()
        s, s2, a = self._dev(current_states), self._dev(next_states), self._dev(current_actions)
        out = torch.empty(s.shape[0], dtype=torch.float32, device=e.device)
        e.check(e.lib.bbmpc_reward(e.handle, _lib.ptr(s), _lib.ptr(a), _lib.ptr(s2), _lib.ptr(out), s.shape[0], e.stream()))
        return out
