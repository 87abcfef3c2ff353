"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference's simulation-mode env refuses to construct without gym
(reference: simulation-mode/problem-03-rl-environment/src/env.py:20-26,108-109).
gym is not installed in this image, so the oracle harness injects this stub to
import the UNMODIFIED reference.  It provides exactly what env.py touches:
`gym.Env`, `spaces.Box(low, high, shape, dtype)` (.low/.high/.shape/.sample())
and `spaces.MultiDiscrete(nvec)` (.nvec/.sample()).
"""
import numpy as np


class Env:
    metadata = {}


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.RandomState()

    def sample(self):
        hi = np.where(np.isinf(self.high), 1e6, self.high)
        return self._rng.uniform(self.low, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class _MultiDiscrete:
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(np.int64)
        self._rng = np.random.RandomState()

    def sample(self):
        return (self._rng.random_sample(self.nvec.shape) * self.nvec).astype(np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class _Spaces:
    Box = _Box
    MultiDiscrete = _MultiDiscrete


spaces = _Spaces()
