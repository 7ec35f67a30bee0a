"""Minimal stand-in for gym.spaces.Box.  The reference reads spaces only through .shape[0], .high
and .low (blackbox_mpc/optimizers/optimizer_base.py:31-36), so any object with those attributes —
including a real gym Box — is accepted everywhere a space is expected."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        if shape is not None:
            low, high = np.broadcast_to(low, shape).copy(), np.broadcast_to(high, shape).copy()
        if low.shape != high.shape or low.ndim != 1:
            raise ValueError("Box expects 1-D low/high of equal shape")
        self.low, self.high, self.shape, self.dtype = low, high, low.shape, dtype

    def __repr__(self):
        return f"Box({self.low.tolist()}, {self.high.tolist()})"
