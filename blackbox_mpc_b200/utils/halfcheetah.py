"""HalfCheetah reward (tutorials/mujoco/cost_func.py:5-22) as a built-in device function:
-10*[s5 >= 0.2] - 10*[s6 >= 0] - 10*[s7 >= 0] + (s'_17 - s_17)/0.01 - 0.0*sum(a^2),
state layout qpos[1:] | qvel | torso-COM (tutorials/mujoco/env_modified.py:22-27), dS >= 18."""
from .. import _lib
from .pendulum import _BuiltinReward

reward_function = _BuiltinReward("halfcheetah_reward_function", _lib.REWARD_HALFCHEETAH, __doc__)
