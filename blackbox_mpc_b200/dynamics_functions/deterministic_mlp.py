"""DeterministicMLP — Dense-chain dynamics function (s_t, a_t) -> raw model output.

Mirrors blackbox_mpc/dynamics_functions/deterministic_mlp.py:5-51 (constructor signature, Keras
Dense semantics y = act(x @ W + b), W [in, out], glorot-uniform kernel / zero bias).  Weights are
torch CUDA tensors; the forward pass itself runs inside libbbmpc (fused into the rollout kernel on
the hot path, bbmpc_dynamics_forward when called directly).  Training hooks (get_loss, :53-95)
are out of scope."""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from .. import _lib

_ACT_IDS = {None: _lib.ACT_NONE, "none": _lib.ACT_NONE, "linear": _lib.ACT_NONE, "tanh": _lib.ACT_TANH,
            "relu": _lib.ACT_RELU, "sigmoid": _lib.ACT_SIGMOID}


def activation_id(act: Union[None, str, Callable]) -> int:
    """Accepts names or callables (torch.tanh, tf.math.tanh, ...): the tutorials pass
    `tf.math.tanh` objects (tutorials/mujoco/tutorial_two.py:28-31); they are matched by name."""
    if act is None or isinstance(act, str):
        key = act.lower() if isinstance(act, str) else None
    else:
        key = getattr(act, "__name__", str(act)).lower()
    if key not in _ACT_IDS:
        raise ValueError(f"unsupported activation {act!r}; supported: tanh, relu, sigmoid, None")
    return _ACT_IDS[key]


class DeterministicMLP:
    def __init__(self, layers: Sequence[int], activation_functions: Sequence, loss_fn=None, name=None,
                 device: Optional[torch.device] = None, seed: Optional[int] = None):
        if len(activation_functions) != len(layers) - 1:
            raise ValueError("need one activation per Dense layer")
        from ..engine import default_device_index
        self.device = torch.device("cuda", default_device_index()) if device is None else torch.device(device)
        self.layer_sizes = [int(v) for v in layers]
        self.activation_ids = [activation_id(a) for a in activation_functions]
        self.loss_fn, self.name = loss_fn, name
        gen = torch.Generator().manual_seed(0 if seed is None else int(seed))
        self.weights: List[torch.Tensor] = []
        self.biases: List[torch.Tensor] = []
        for fan_in, fan_out in zip(self.layer_sizes[:-1], self.layer_sizes[1:]):
            limit = math.sqrt(6.0 / (fan_in + fan_out))  # Keras glorot_uniform [TF]
            w = (torch.rand(fan_in, fan_out, generator=gen, dtype=torch.float32) * 2 - 1) * limit
            self.weights.append(w.to(self.device).contiguous())
            self.biases.append(torch.zeros(fan_out, dtype=torch.float32, device=self.device))
        self._version = 0
        self._engine = None

    # -- weight management ------------------------------------------------------------------
    @property
    def version(self) -> int:
        return self._version

    def mark_dirty(self) -> None:
        """Call after mutating .weights/.biases in place: consumers re-stage on the next use."""
        self._version += 1

    def set_weights(self, weights, biases) -> None:
        for i, (w, b) in enumerate(zip(weights, biases)):
            w = torch.as_tensor(np.asarray(w) if not torch.is_tensor(w) else w, dtype=torch.float32)
            b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b, dtype=torch.float32)
            if tuple(w.shape) != tuple(self.weights[i].shape) or tuple(b.shape) != tuple(self.biases[i].shape):
                raise ValueError(f"layer {i}: shape mismatch")
            self.weights[i] = w.to(self.device).contiguous()
            self.biases[i] = b.to(self.device).contiguous()
        self.mark_dirty()

    def members(self) -> List["DeterministicMLP"]:
        return [self]

    # -- dynamics_function(x, train) --------------------------------------------------------
    def __call__(self, x: torch.Tensor, train=False) -> torch.Tensor:
        from ..dynamics_handlers.system_dynamics_handler import stage_model
        from ..engine import Engine
        if self._engine is None:
            self._engine = Engine(self.device.index)
            self._staged = -1
        if self._staged != self.version:
            stage_model(self._engine, self)
            self._staged = self.version
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device).contiguous()
        out = torch.empty(x.shape[0], self.layer_sizes[-1], dtype=torch.float32, device=self.device)
        e = self._engine
        e.check(e.lib.bbmpc_dynamics_forward(e.handle, _lib.ptr(x), _lib.ptr(out), x.shape[0], e.stream()))
        return out


class EnsembleMLP:
    """n DeterministicMLP members of identical shape; dynamics_function(x) = mean of the members'
    raw outputs (summed in member order, then / n).  Not a reference type: it is what a user of
    the reference would write as a plain `dynamics_function` averaging several DeterministicMLPs
    (SURVEY §8d), made a class so the kernel can fuse it."""

    def __init__(self, members: Sequence[DeterministicMLP]):
        members = list(members)
        if not members:
            raise ValueError("empty ensemble")
        ref = members[0]
        for m in members[1:]:
            if m.layer_sizes != ref.layer_sizes or m.activation_ids != ref.activation_ids:
                raise ValueError("ensemble members must share layer sizes and activations")
        self._members = members
        self.device, self.layer_sizes, self.activation_ids = ref.device, ref.layer_sizes, ref.activation_ids
        self._engine = None

    @property
    def version(self) -> int:
        return sum(m.version for m in self._members)

    def members(self) -> List[DeterministicMLP]:
        return self._members

    __call__ = DeterministicMLP.__call__
