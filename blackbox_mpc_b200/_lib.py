"""ctypes binding of libbbmpc.so (include/bbmpc.h).  There is no CPU fallback: a missing or
unloadable library raises at import of the first class that needs it."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BBMPC_LIB") or os.path.join(_HERE, "libbbmpc.so")   # BBMPC_LIB: A/B builds (tools/debug)

# mirrors of include/bbmpc.h
OK, EINVAL, ECUDA, ESTATE, ENOMEM = 0, -1, -2, -3, -4
DYN_MLP, DYN_PENDULUM = 0, 1
REWARD_PENDULUM, REWARD_HALFCHEETAH, REWARD_PENDULUM_GYM, REWARD_USER = 1, 2, 3, 4
ACT_NONE, ACT_TANH, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3
PREC_AUTO, PREC_FP32, PREC_BF16X3, PREC_BF16 = 0, 1, 2, 3
OPT_CEM, OPT_PI2, OPT_RANDOM_SEARCH, OPT_PSO, OPT_SPSA, OPT_CMAES = 1, 2, 3, 4, 5, 6
PRECISIONS = {"auto": PREC_AUTO, "fp32": PREC_FP32, "bf16x3": PREC_BF16X3, "bf16": PREC_BF16}

c_float_p = C.POINTER(C.c_float)
MIN_VERSION = 200


class OptConfig(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("population_size", C.c_int), ("num_agents", C.c_int),
        ("planning_horizon", C.c_int), ("max_iterations", C.c_int), ("dS", C.c_int), ("dU", C.c_int),
        ("lb_host", c_float_p), ("ub_host", c_float_p), ("num_elite", C.c_int),
        ("alpha", C.c_float), ("epsilon", C.c_float), ("lamda", C.c_float),
        ("c1", C.c_float), ("c2", C.c_float), ("w", C.c_float), ("initial_velocity_fraction", C.c_float),
        ("gamma", C.c_float), ("a_par", C.c_float), ("noise_parameter", C.c_float),
        ("h_sigma", C.c_float), ("alpha_cov", C.c_float),
    ]


# symbol -> (restype, argtypes); the CPU test-suite checks every symbol of bbmpc.h is exported.
_VP, _I, _I64, _U64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
SIGNATURES = {
    "bbmpc_version": (_I, []),
    "bbmpc_abi_config_size": (_I, []),
    "bbmpc_last_error": (C.c_char_p, [_VP]),
    "bbmpc_ctx_create": (_I, [_I, _U64, C.POINTER(_VP)]),
    "bbmpc_ctx_destroy": (None, [_VP]),
    "bbmpc_set_precision": (_I, [_VP, _I]),
    "bbmpc_get_effective_precision": (_I, [_VP]),
    "bbmpc_launch_count": (_U64, [_VP]),
    "bbmpc_last_rollout_kernel": (C.c_char_p, [_VP]),
    "bbmpc_profile_enable": (_I, [_VP, _I]),
    "bbmpc_profile_read": (_I, [_VP, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "bbmpc_model_set_mlp": (_I, [_VP, _I, _I, C.POINTER(_I), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_I), _VP]),
    "bbmpc_model_set_norm": (_I, [_VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "bbmpc_model_set_builtin": (_I, [_VP, _I, _I, _I]),
    "bbmpc_reward_set_builtin": (_I, [_VP, _I]),
    "bbmpc_reward_set_nvrtc": (_I, [_VP, C.c_char_p]),
    "bbmpc_rollout": (_I, [_VP, _VP, _VP, _VP, _I, _I, _I, _VP]),
    "bbmpc_predict_next_state": (_I, [_VP, _VP, _VP, _VP, _I, _VP]),
    "bbmpc_reward": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _VP]),
    "bbmpc_dynamics_forward": (_I, [_VP, _VP, _VP, _I, _VP]),
    "bbmpc_opt_create": (_I, [_VP, C.POINTER(OptConfig), C.POINTER(_VP)]),
    "bbmpc_opt_destroy": (None, [_VP]),
    "bbmpc_opt_reset": (_I, [_VP, _VP]),
    "bbmpc_opt_set_shard": (_I, [_VP, _I, _I]),
    "bbmpc_opt_call": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "bbmpc_opt_call_host": (_I, [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP]),
    "bbmpc_opt_num_iterations": (_I, [_VP]),
    "bbmpc_opt_partial_floats": (_I, [_VP]),
    "bbmpc_opt_begin": (_I, [_VP, _VP, _I, _VP]),
    "bbmpc_opt_iter_local": (_I, [_VP, _I, _VP, _VP]),
    "bbmpc_opt_iter_merge": (_I, [_VP, _I, _VP, _I, _VP]),
    "bbmpc_opt_finish": (_I, [_VP, _I, _VP, _VP, _VP, _VP]),
    "bbmpc_opt_p2p_export": (_I, [_VP, _VP, C.POINTER(_VP)]),
    "bbmpc_opt_p2p_connect": (_I, [_VP, _VP, C.POINTER(_VP)]),
    "bbmpc_opt_get_tensor": (_I64, [_VP, C.c_char_p, _VP, _I64, _VP]),
    "bbmpc_opt_set_sample_trace": (_I, [_VP, _VP, _I64]),
    "bbmpc_opt_set_draw_injection": (_I, [_VP, _VP, _I64]),
    "bbmpc_philox4x32_host": (None, [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

_lib: Optional[C.CDLL] = None


class BBMPCError(RuntimeError):
    """A negative BBMPC_E* return code from libbbmpc.so."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libbbmpc error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Loads the in-tree library, building it first when it is missing (nvcc cross-compiles without a GPU).  A library
    that does not match this binding (older version, different bbmpc_opt_config layout, missing symbol) is rejected
    here instead of corrupting configurations later: rebuild with `python -m blackbox_mpc_b200._build`."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from ._build import build_library
        build_library()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.bbmpc_version() < MIN_VERSION or lib.bbmpc_abi_config_size() != C.sizeof(OptConfig):
        raise ImportError(f"{LIB_PATH} is stale (version {lib.bbmpc_version()}, bbmpc_opt_config of "
                          f"{lib.bbmpc_abi_config_size()} bytes; this binding needs >= {MIN_VERSION} and "
                          f"{C.sizeof(OptConfig)} bytes): run `python -m blackbox_mpc_b200._build --force`")
    _lib = lib
    return lib


def check(ctx_handle, rc: int) -> int:
    if rc < 0:
        msg = load().bbmpc_last_error(ctx_handle)
        raise BBMPCError(rc, msg.decode() if msg else "")
    return rc


def ptr(t) -> Optional[int]:
    """Device pointer of a CUDA fp32 contiguous torch tensor (None passes NULL)."""
    if t is None:
        return None
    import torch
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous CUDA fp32 tensor"
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
