"""Reservoir sampling on the GPU with the reference's Python API.

API mirror of simulation-mode/problem-01-reservoir-sampling/src/reservoir.py
(`ReservoirSampler` :17-233, `MultiMetricReservoir` :236-316) and the per-server
state assembly of src/features.py:232-303, computed by the batched kernels of
csrc/mlb_ops.cu (`mlb_reservoir_add`, `mlb_reservoir_features`).
`BatchedReservoirs` is the native batched form; `ReservoirSampler` is a batch
of one.  Timestamps are stored on the device as float32 seconds (src/vpp/lb/shm.h:23-25)
RELATIVE to a float64 host-side epoch (the first timestamp seen): the reference keeps float64
`time.time()` values (reservoir.py:42,62,140), whose differences -- all the decay weights depend
on -- would be lost in float32 at epoch scale (128 s resolution near 1.8e9).
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class BatchedReservoirs:
    """R independent reservoirs of capacity K; reservoir r replays RandomState(seeds[r])."""

    def __init__(self, num: int, capacity: int = 128, seeds=None, table_len: int = 65536, device: int = 0):
        if not 1 <= capacity <= 128:
            raise ValueError("capacity must be in [1, 128]")
        if not torch.cuda.is_available():
            raise RuntimeError("marllb_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = _lib.load()
        self.num, self.capacity, self.table_len = num, capacity, table_len
        self.kp = (capacity + 31) // 32 * 32
        self.device = torch.device("cuda", device)
        seeds = np.arange(num) if seeds is None else np.asarray(seeds)
        uniq, inv = np.unique(seeds.astype(np.uint32), return_inverse=True)
        tab = np.empty((len(uniq), table_len), np.uint32)
        for i, s in enumerate(uniq):
            check(self._L.mlb_mt19937_fill(int(s), tab[i].ctypes.data_as(C.c_void_p), table_len))
        self._table = torch.as_tensor(tab.view(np.int32)).to(self.device)
        self._seed_row = torch.as_tensor(inv.astype(np.int32)).to(self.device)
        self.values = torch.zeros((num, self.kp), dtype=torch.float32, device=self.device)
        self.timestamps = torch.zeros((num, self.kp), dtype=torch.float32, device=self.device)
        self.count = torch.zeros(num, dtype=torch.int32, device=self.device)
        self.cursor = torch.zeros(num, dtype=torch.int32, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.epoch = None          # float64 host-side time base; device timestamps are (t - epoch) in float32

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _rel_time(self, t, set_epoch=False):
        """float64 timestamps (array-like or tensor) -> float32 device tensor of seconds since `epoch`."""
        if isinstance(t, torch.Tensor):
            t64 = t.detach().to(dtype=torch.float64)
        else:
            t64 = torch.as_tensor(np.asarray(t, np.float64))
        if self.epoch is None:
            if not set_epoch or t64.numel() == 0:
                return t64.to(torch.float32).to(self.device)
            # simulated time (seconds since reset, float32-exact in the env kernels) is kept as is; wall-clock
            # time (time.time() ~ 1.8e9, float32 resolution 128 s) is rebased on the first timestamp seen
            t0 = float(t64.reshape(-1)[0].item())
            self.epoch = t0 if abs(t0) >= 2.0 ** 20 else 0.0
        return (t64 - self.epoch).to(torch.float32).to(self.device)

    def add(self, values, timestamps, n_add=None) -> torch.Tensor:
        """values/timestamps: (R, M) samples per reservoir, fed in order; n_add: (R,) counts
        (default M each).  Returns the (R, M) uint8 accepted flags."""
        v = torch.as_tensor(values, dtype=torch.float32).to(self.device).reshape(self.num, -1).contiguous()
        t = self._rel_time(timestamps, set_epoch=True).reshape(self.num, -1).contiguous()
        M = v.shape[1]
        n = (torch.full((self.num,), M, dtype=torch.int32, device=self.device) if n_add is None
             else torch.as_tensor(n_add, dtype=torch.int32).to(self.device).contiguous())
        acc = torch.zeros((self.num, M), dtype=torch.uint8, device=self.device)
        check(self._L.mlb_reservoir_add(_p(self.values), _p(self.timestamps), _p(self.count), _p(self.cursor),
                                        _p(self._table), _p(self._seed_row), self.table_len, self.num,
                                        self.capacity, _p(v), _p(t), _p(n), M, _p(acc), _p(self._status),
                                        self._stream()))
        return acc

    def features(self, decay: float = 0.9, now=0.0) -> torch.Tensor:
        """(R, 5) float32: mean, p90, std, mean_decay, p90_decay (reservoir.py:105-163)."""
        nw = self._rel_time(now).expand(self.num).contiguous()
        out = torch.empty((self.num, 5), dtype=torch.float32, device=self.device)
        check(self._L.mlb_reservoir_features(_p(self.values), _p(self.timestamps), _p(self.count), self.num,
                                             self.capacity, float(decay), _p(nw), _p(out), self._stream()))
        return out

    def check_status(self):
        if int(self._status.item()) != 0:
            raise _lib.MlbError(_lib.ERNG, "replayed MT19937 stream exhausted: raise table_len")

    def reset(self):
        """reservoir.py:220-225: clears samples and counters; like the reference the random
        stream is NOT rewound."""
        self.values.zero_()
        self.timestamps.zero_()
        self.count.zero_()
        self.epoch = None


class ReservoirSampler:
    """reservoir.py:17-233 on the GPU (a batch of one)."""

    def __init__(self, capacity: int = 128, seed: Optional[int] = None, table_len: int = 65536):
        self.capacity = capacity
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        self._b = BatchedReservoirs(1, capacity, seeds=[seed], table_len=table_len)

    @property
    def count(self) -> int:
        return int(self._b.count.item())

    @property
    def values(self) -> np.ndarray:
        return self._b.values[0, :self.capacity].cpu().numpy()

    @property
    def timestamps(self) -> np.ndarray:
        ts = self._b.timestamps[0, :self.capacity].cpu().numpy().astype(np.float64)
        if self._b.epoch is not None:                       # zero-filled (unused) slots stay 0 like the reference's
            ts[:self.get_size()] += self._b.epoch
        return ts

    def add(self, value: float, timestamp: Optional[float] = None) -> bool:
        if timestamp is None:
            timestamp = time.time()
        acc = self._b.add([[value]], [[timestamp]])
        self._b.check_status()
        return bool(acc.item())

    def add_many(self, values, timestamps) -> np.ndarray:
        acc = self._b.add(np.asarray(values, np.float32)[None], np.asarray(timestamps, np.float64)[None])
        self._b.check_status()
        return acc[0].cpu().numpy().astype(bool)

    def get_size(self) -> int:
        return min(self.count, self.capacity)

    def is_full(self) -> bool:
        return self.count >= self.capacity

    def get_samples(self) -> Tuple[np.ndarray, np.ndarray]:
        n = self.get_size()
        return self.values[:n].copy(), self.timestamps[:n].copy()

    def get_features(self, decay_factor: float = 0.9, current_time: Optional[float] = None) -> Dict[str, float]:
        v = self.get_feature_vector(decay_factor, current_time)
        return {k: float(x) for k, x in zip(('mean', 'p90', 'std', 'mean_decay', 'p90_decay'), v)}

    def get_feature_vector(self, decay_factor: float = 0.9, current_time: Optional[float] = None) -> np.ndarray:
        if current_time is None:
            current_time = time.time()
        return self._b.features(decay_factor, float(current_time))[0].cpu().numpy()

    def reset(self):
        self._b.reset()

    def __len__(self) -> int:
        return self.get_size()

    def __repr__(self) -> str:
        return f"ReservoirSampler(capacity={self.capacity}, count={self.count}, size={self.get_size()})"


class MultiMetricReservoir:
    """reservoir.py:236-316: one reservoir per metric, all seeded alike (:261-265)."""

    def __init__(self, metrics: List[str] = None, capacity: int = 128, seed: Optional[int] = None):
        self.metrics = metrics if metrics is not None else ['fct', 'flow_duration']
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        self.reservoirs = {m: ReservoirSampler(capacity=capacity, seed=seed) for m in self.metrics}

    def add(self, metric: str, value: float, timestamp: Optional[float] = None):
        if metric not in self.reservoirs:
            raise ValueError(f"Unknown metric: {metric}")
        return self.reservoirs[metric].add(value, timestamp)

    def get_all_features(self, decay_factor: float = 0.9, current_time: Optional[float] = None):
        return {m: r.get_features(decay_factor, current_time) for m, r in self.reservoirs.items()}

    def get_feature_vector(self, decay_factor: float = 0.9, current_time: Optional[float] = None) -> np.ndarray:
        return np.concatenate([self.reservoirs[m].get_feature_vector(decay_factor, current_time)
                               for m in self.metrics])

    def reset(self):
        for r in self.reservoirs.values():
            r.reset()

    def __repr__(self) -> str:
        return f"MultiMetricReservoir(metrics={self.metrics})"


class PerServerFeatures:
    """features.py:232-303: (S, 11) state = [n_flow_on | 10 reservoir features]."""

    def __init__(self, num_servers: int):
        self.num_servers = num_servers
        self.n_flow_on = np.zeros(num_servers, dtype=np.int32)

    def update_flow_count(self, server_id: int, count: int):
        self.n_flow_on[server_id] = count

    def get_state_vector(self, reservoir_features, active_servers=None) -> np.ndarray:
        state = np.zeros((self.num_servers, 11), dtype=np.float32)
        for i in range(self.num_servers):
            state[i, 0] = self.n_flow_on[i]
            if i < len(reservoir_features):
                state[i, 1:] = reservoir_features[i]
        if active_servers is not None:
            mask = np.zeros(self.num_servers, dtype=bool)
            mask[active_servers] = True
            state[~mask] = 0
        return state
