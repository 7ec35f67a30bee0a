"""Pendulum true model and reward as built-in device functions.

Mirrors blackbox_mpc/utils/pendulum.py: PendulumTrueModel (:37-92) and pendulum_reward_function
(:10-35).  The objects are plugin markers: the arithmetic lives in csrc/device_fns.cuh and is
fused into the rollout kernel.

Argument-order note (reference behaviour, kept): the evaluator calls
reward_function(current_state, actions, next_state) (deterministic.py:65-66) while
pendulum_reward_function is declared (current_state, next_state, actions), so its control-cost
term is computed from next_state.  `pendulum_reward_function` reproduces exactly that;
`pendulum_reward_function_gym_order` is the reward with arguments as its docstring intends."""
from .. import _lib


class _BuiltinReward:
    def __init__(self, name, reward_id, doc):
        self.__name__, self.bbmpc_reward_id, self.__doc__ = name, reward_id, doc

    def __call__(self, current_state, actions, next_state):
        """Stand-alone evaluation (B rows) through bbmpc_reward."""
        import ctypes as C
        import torch
        from ..engine import Engine
        s = torch.as_tensor(current_state, dtype=torch.float32).cuda().contiguous()
        a = torch.as_tensor(actions, dtype=torch.float32).cuda().contiguous()
        s2 = torch.as_tensor(next_state, dtype=torch.float32).cuda().contiguous()
        e = Engine(s.device.index)
        e.check(e.lib.bbmpc_model_set_norm(e.handle, s.shape[1], a.shape[1], None, None, None, None, None, None, e.stream()))
        e.check(e.lib.bbmpc_reward_set_builtin(e.handle, self.bbmpc_reward_id))
        out = torch.empty(s.shape[0], dtype=torch.float32, device=s.device)
        e.check(e.lib.bbmpc_reward(e.handle, _lib.ptr(s), _lib.ptr(a), _lib.ptr(s2), _lib.ptr(out), s.shape[0], e.stream()))
        torch.cuda.synchronize(s.device)
        e.close()
        return out

    def __repr__(self):
        return f"<built-in device reward {self.__name__}>"


pendulum_reward_function = _BuiltinReward(
    "pendulum_reward_function", _lib.REWARD_PENDULUM,
    "-(wrap(atan2(s1,s0))^2 + 0.1 s2^2) - 0.001*sum(third^2), third = the evaluator's 3rd positional (next_state)")
pendulum_reward_function_gym_order = _BuiltinReward(
    "pendulum_reward_function_gym_order", _lib.REWARD_PENDULUM_GYM,
    "-(wrap(atan2(s1,s0))^2 + 0.1 s2^2) - 0.001*sum(actions^2)")


class PendulumTrueModel:
    """x = [cos th, sin th, thdot, u] -> deviation of the next state (pendulum.py:58-92);
    g=10, m=l=1, dt=0.05, |thdot| <= 8 (clipped after the angle update), u not clipped."""
    bbmpc_dynamics_id = _lib.DYN_PENDULUM
    dim_S, dim_U = 3, 1
    version = 0

    def __init__(self, name=None):
        self.name = name
        self._engine = None

    def __call__(self, x, train=False):
        import torch
        from ..engine import Engine
        if self._engine is None:
            self._engine = Engine()
            e = self._engine
            e.check(e.lib.bbmpc_model_set_builtin(e.handle, self.bbmpc_dynamics_id, 3, 1))
        e = self._engine
        x = torch.as_tensor(x, dtype=torch.float32, device=e.device).contiguous()
        out = torch.empty(x.shape[0], 3, dtype=torch.float32, device=e.device)
        e.check(e.lib.bbmpc_dynamics_forward(e.handle, _lib.ptr(x), _lib.ptr(out), x.shape[0], e.stream()))
        return out
