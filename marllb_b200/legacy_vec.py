"""E reference envs in their actual simulation mode, stepped by one kernel launch.

What `LoadBalanceEnv.step()` does in the reference's simulation mode is draw a random observation from the env's
own `np.random.RandomState(seed)` and score it (problem-03-rl-environment/src/env.py:215-286, 425-448).
`VecLegacyEnv` is that, for E independent envs: `mlb_legacy_step` replays every env's MT19937 stream on the device
(one warp per env) and evaluates the reward metric in the same launch.  Env e reproduces
`LoadBalanceEnv(num_servers=S, seed=seeds[e], ...)` bit for bit: observations exactly, rewards to float64 rounding.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check


class VecLegacyEnv:
    def __init__(self, num_envs: int, num_servers: int = 4, seeds=0, reward_metric: str = "jain",
                 reward_field="flow_duration_avg_decay", max_steps: int = 10000, device: int = 0):
        if reward_metric not in _lib.METRICS:
            raise ValueError(f"Unsupported metric: {reward_metric}. Supported: {list(_lib.METRICS.keys())}")
        if isinstance(reward_field, str):
            if reward_field not in _lib.FEATURE_NAMES:
                raise ValueError(f"Unknown reward_field: {reward_field}")
            reward_field = _lib.FEATURE_NAMES.index(reward_field)
        if not torch.cuda.is_available():
            raise RuntimeError("marllb_b200 needs a CUDA device (there is no CPU fallback)")
        self._L = _lib.load()
        self.num_envs, self.num_servers, self.max_steps = num_envs, num_servers, max_steps
        self._metric, self._field = _lib.METRICS[reward_metric], int(reward_field)
        self.device = torch.device("cuda", device)
        s = np.asarray(seeds, dtype=np.uint32)
        self.seeds = (np.arange(num_envs, dtype=np.uint32) + s) if s.ndim == 0 else np.ascontiguousarray(s)
        if self.seeds.shape != (num_envs,):
            raise ValueError("seeds: one integer (env e gets seeds + e) or one per env")
        E, S = num_envs, num_servers
        self._mt = torch.empty((E, 625), dtype=torch.int32, device=self.device)       # numpy's state + position
        self.obs = torch.empty((E, S, 11), dtype=torch.float32, device=self.device)
        self.reward = torch.zeros((E,), dtype=torch.float64, device=self.device)
        self.done = torch.zeros((E,), dtype=torch.bool, device=self.device)
        self._status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.current_step = 0
        self.seed(self.seeds)

    @staticmethod
    def _st():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def seed(self, seeds):
        """np.random.RandomState(seed) per env (env.py:127,327-332)."""
        s = torch.as_tensor(np.ascontiguousarray(np.asarray(seeds, np.uint32)).view(np.int32)).to(self.device)
        check(self._L.mlb_legacy_seed(C.c_void_p(self._mt.data_ptr()), C.c_void_p(s.data_ptr()), self.num_envs, self._st()))
        torch.cuda.current_stream().synchronize()
        return list(np.asarray(seeds).tolist())

    def _launch(self, with_reward):
        check(self._L.mlb_legacy_step(C.c_void_p(self._mt.data_ptr()), self.num_envs, self.num_servers, self._metric,
                                      self._field, C.c_void_p(self.obs.data_ptr()),
                                      C.c_void_p(self.reward.data_ptr()) if with_reward else None,
                                      C.c_void_p(self._status.data_ptr()), self._st()))

    def reset(self):
        """env.py:186-213: counters to zero and a first observation from the (continuing) stream."""
        self.current_step = 0
        self._launch(False)
        return self.obs

    def step(self, action=None):
        """env.py:215-286 in simulation mode: the action does not influence the next state.
        Returns (obs (E,S,11) f32, reward (E,) f64, done (E,) bool) device tensors."""
        self.current_step += 1
        self._launch(True)
        self.done.fill_(self.current_step >= self.max_steps)                       # env.py:267
        return self.obs, self.reward, self.done

    def check_status(self):
        if int(self._status.item()) != 0:
            raise RuntimeError("legacy stream replay reported a corrupt MT19937 state")
