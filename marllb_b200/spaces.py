"""Observation / action spaces.

The reference builds gym spaces (env.py:156-184) and refuses to run without
gym (env.py:108-109).  gym is optional here: when it is importable its classes
are used, otherwise these minimal equivalents provide what the reference's
callers and tests touch (.shape/.low/.high/.nvec/.sample()/.contains()).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gym.spaces import Box, MultiDiscrete  # type: ignore
    HAVE_GYM = True
except Exception:  # gym absent (this image) or broken
    HAVE_GYM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class MultiDiscrete:
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            self.shape = self.nvec.shape
            self.dtype = np.dtype(np.int64)
            self._rng = np.random.RandomState()

        def seed(self, seed=None):
            self._rng = np.random.RandomState(seed)
            return [seed]

        def sample(self):
            return (self._rng.random_sample(self.nvec.shape) * self.nvec).astype(np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

        def __repr__(self):
            return f"MultiDiscrete({self.nvec.tolist()})"
