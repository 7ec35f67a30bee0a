"""User-supplied reward functions for the fused rollout.

The reference's plug point is an arbitrary Python callable `reward_function(current_state, actions, next_state)`
(policies/mpc_policy.py:42-44; invoked at trajectory_evaluators/deterministic.py:65-66,126-127).  A Python callable cannot
run inside the sm_100a kernels; a CUDA one can.  `cuda_reward(source)` wraps CUDA source that defines

    __device__ float reward(const float* s, const float* a, const float* s2)   // dS, dU, dS floats; BBMPC_DS / BBMPC_DU macros

into an object every evaluator / MPCPolicy of this package accepts as `reward_function`: it is compiled once per engine with
NVRTC (bbmpc_reward_set_nvrtc, --fmad=false so that + - * / are the IEEE fp32 operations of the reference's graph) and applied
by a JIT-compiled kernel to the states the rollout visits.  Example — the HalfCheetah reward of tutorials/mujoco/cost_func.py:5-22:

    reward = cuda_reward('''
    __device__ float reward(const float* s, const float* a, const float* s2) {
      float r = 0.0f;
      if (s[5] >= 0.2f) r += -10.0f;
      if (s[6] >= 0.0f) r += -10.0f;
      if (s[7] >= 0.0f) r += -10.0f;
      r = r + (s2[17] - s[17]) / 0.01f;
      float ss = 0.0f;
      for (int i = 0; i < BBMPC_DU; ++i) ss = ss + a[i] * a[i];
      return r - 0.0f * ss;
    }''')
    policy = MPCPolicy(reward_function=reward, ...)
"""
from __future__ import annotations

from .. import _lib

HALFCHEETAH_SOURCE = """
// tutorials/mujoco/cost_func.py:5-22
__device__ float reward(const float* s, const float* a, const float* s2) {
  float r = 0.0f;
  if (s[5] >= 0.2f) r += -10.0f;
  if (s[6] >= 0.0f) r += -10.0f;
  if (s[7] >= 0.0f) r += -10.0f;
  r = r + (s2[17] - s[17]) / 0.01f;
  float ss = 0.0f;
  for (int i = 0; i < BBMPC_DU; ++i) ss = ss + a[i] * a[i];
  return r - 0.0f * ss;
}
"""


class CudaReward:
    """A reward_function given as CUDA source (see module docstring)."""

    bbmpc_reward_id = _lib.REWARD_USER

    def __init__(self, source: str, name: str = "user_cuda_reward"):
        if "reward" not in source:
            raise ValueError("the source must define `__device__ float reward(const float* s, const float* a, const float* s2)`")
        self.cuda_source = source
        self.__name__ = name

    def __call__(self, current_state, actions, next_state):
        """Stand-alone evaluation of B rows through bbmpc_reward (compiles for the given dS / dU)."""
        import torch
        from ..engine import Engine
        s = torch.as_tensor(current_state, dtype=torch.float32).cuda().contiguous()
        a = torch.as_tensor(actions, dtype=torch.float32).cuda().contiguous()
        s2 = torch.as_tensor(next_state, dtype=torch.float32).cuda().contiguous()
        e = Engine(s.device.index)
        try:
            e.check(e.lib.bbmpc_model_set_norm(e.handle, s.shape[1], a.shape[1], None, None, None, None, None, None, e.stream()))
            e.check(e.lib.bbmpc_reward_set_nvrtc(e.handle, self.cuda_source.encode()))
            out = torch.empty(s.shape[0], dtype=torch.float32, device=s.device)
            e.check(e.lib.bbmpc_reward(e.handle, _lib.ptr(s), _lib.ptr(a), _lib.ptr(s2), _lib.ptr(out), s.shape[0], e.stream()))
            torch.cuda.synchronize(s.device)
        finally:
            e.close()
        return out

    def __repr__(self):
        return f"<CUDA-source reward {self.__name__}>"


def cuda_reward(source: str, name: str = "user_cuda_reward") -> CudaReward:
    return CudaReward(source, name)
