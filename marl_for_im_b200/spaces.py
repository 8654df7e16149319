"""Minimal stand-in for ``gym.spaces.Box`` (gym 0.18 is what the reference pins; it is not a
dependency here).  Carries exactly what the reference's callers read: low / high / dtype / shape
(e.g. CC_inv_management.py:133-137) plus ``sample`` and ``contains`` for convenience."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float64, shape=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(int(s) for s in shape)          # the reference leaks numpy int8 here (quirk 14)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).astype(self.dtype)
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).astype(self.dtype)

    def sample(self, rng=None):
        rng = rng or np.random.default_rng()
        high = np.where(np.isfinite(self.high.astype(np.float64)), self.high, 1e6)
        return rng.uniform(self.low, high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
