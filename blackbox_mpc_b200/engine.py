"""Engine: owner of one bbmpc_ctx (one model + one reward function on one GPU)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

from . import _lib


def default_device_index() -> int:
    """One process per GPU: LOCAL_RANK picks the device under torchrun, else the current device."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("blackbox_mpc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if "LOCAL_RANK" in os.environ:
        return int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
    return torch.cuda.current_device()


class Engine:
    """Thin RAII wrapper of bbmpc_ctx_create / bbmpc_ctx_destroy."""

    def __init__(self, device: Optional[int] = None, seed: int = 0, precision: str = "auto"):
        import torch
        self.lib = _lib.load()
        self.device_index = default_device_index() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        rc = self.lib.bbmpc_ctx_create(self.device_index, C.c_uint64(seed & (2**64 - 1)), C.byref(h))
        if rc < 0:
            raise _lib.BBMPCError(rc, (self.lib.bbmpc_last_error(None) or b"").decode())
        self.handle = h
        self.seed = seed
        self.set_precision(precision)

    def check(self, rc: int) -> int:
        return _lib.check(self.handle, rc)

    def set_precision(self, precision: str) -> None:
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS)}")
        self.check(self.lib.bbmpc_set_precision(self.handle, _lib.PRECISIONS[precision]))
        self.precision = precision

    @property
    def effective_precision(self) -> str:
        code = self.lib.bbmpc_get_effective_precision(self.handle)
        return {v: k for k, v in _lib.PRECISIONS.items()}.get(code, "?")

    @property
    def launch_count(self) -> int:
        return int(self.lib.bbmpc_launch_count(self.handle))

    @property
    def last_rollout_kernel(self) -> str:
        return (self.lib.bbmpc_last_rollout_kernel(self.handle) or b"").decode()

    def profile_enable(self, on: bool = True) -> None:
        self.check(self.lib.bbmpc_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self):
        """(summed rollout-kernel device time in ms, number of rollout launches) since the last read."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        self.check(self.lib.bbmpc_profile_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stream(self) -> int:
        import torch
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.bbmpc_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
