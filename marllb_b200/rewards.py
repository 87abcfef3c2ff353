"""Fairness / makespan rewards evaluated on the GPU.

API mirror of simulation-mode/problem-03-rl-environment/src/rewards.py: the
nine metric functions (rewards.py:21-287) and `RewardFunction` (:290-388) with
the same names, arguments, return conventions and errors.  Arithmetic runs in
float64 in the `reward_metric_kernel` (csrc/mlb_ops.cu) through
`mlb_reward_metric`; inside the fused env step the same device code computes
the reward without leaving the kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Union

import numpy as np
import torch

from . import _lib
from ._lib import check


def reward_metric_batch(metric: str, values: torch.Tensor, n: torch.Tensor) -> torch.Tensor:
    """values: (B, stride) float64 CUDA tensor, n: (B,) int32 valid counts -> (B,) float64."""
    if metric not in _lib.METRICS:
        raise ValueError(f"Unsupported metric: {metric}. Supported: {list(_lib.METRICS.keys())}")
    L = _lib.load()
    values = values.to(dtype=torch.float64).contiguous()
    n = n.to(dtype=torch.int32, device=values.device).contiguous()
    out = torch.empty(values.shape[0], dtype=torch.float64, device=values.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    check(L.mlb_reward_metric(_lib.METRICS[metric], C.c_void_p(values.data_ptr()),
                              C.c_void_p(n.data_ptr()), values.shape[0], values.shape[1] if values.dim() > 1 else 0,
                              C.c_void_p(out.data_ptr()), st))
    return out


def _one(metric: str, values) -> float:
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    if v.size == 0:
        return 1.0 if metric in ("jain", "fair_jain") else 0.0          # rewards.py:49-50,91-92
    t = torch.as_tensor(v).to("cuda").reshape(1, -1)
    n = torch.tensor([v.size], dtype=torch.int32, device="cuda")
    return float(reward_metric_batch(metric, t, n).item())


def jain_fairness(values: Union[List[float], np.ndarray], epsilon: float = 1e-10) -> float:
    """Jain's index (rewards.py:21-67)."""
    return _one("jain", values)


def variance_fairness(values) -> float:
    """-variance (rewards.py:70-94)."""
    return _one("variance", values)


def std_fairness(values) -> float:
    """-std (rewards.py:97-114)."""
    return _one("std", values)


def coefficient_of_variation(values, epsilon: float = 1e-10) -> float:
    """-std/mean (rewards.py:117-144)."""
    return _one("cv", values)


def max_min_fairness(values) -> float:
    """-max, the makespan reward (rewards.py:147-171)."""
    return _one("max", values)


def min_max_fairness(values) -> float:
    """min (rewards.py:174-191)."""
    return _one("min", values)


def product_fairness(values, epsilon: float = 1e-10) -> float:
    """sum log(x+eps) (rewards.py:194-225)."""
    return _one("product", values)


def range_fairness(values) -> float:
    """-(max-min) (rewards.py:228-246)."""
    return _one("range", values)


def gini_coefficient(values) -> float:
    """-Gini (rewards.py:249-287)."""
    return _one("gini", values)


# ---- the original testbed's reward table (src/lb/env.py:73-161), same names ----------------------
def calcul_fair_jain(values) -> float:
    """(sum x)^2 / (n sum x^2), 1.0 when sum x == 0, not clipped (src/lb/env.py:73-85)."""
    return _one("fair_jain", values)


def calcul_fair_product(values) -> float:
    """prod(x / (max x + 1e-6)) (src/lb/env.py:87-96)."""
    return _one("fair_product", values)


def calcul_fair_variance(values) -> float:
    """-var (src/lb/env.py:98-105)."""
    return _one("var", values)


def calcul_fair_variance_exp(values, k=10000) -> float:
    """exp(-k var), k = 10000 (src/lb/env.py:108-115)."""
    if k != 10000:
        raise ValueError("only the reference's k = 10000 is built into the kernel")
    return _one("var_exp", values)


def calcul_fair_variance_log(values) -> float:
    """-log(var) (src/lb/env.py:118-125)."""
    return _one("var_log", values)


def calcul_max(values) -> float:
    """-max = negative makespan (src/lb/env.py:127-132)."""
    return _one("max", values)


def calcul_max_log(values) -> float:
    """-log(max) (src/lb/env.py:135-139)."""
    return _one("max_log", values)


def calcul_max_exp(values, k=10000) -> float:
    """exp(-k max), k = 10000 (src/lb/env.py:142-149)."""
    if k != 10000:
        raise ValueError("only the reference's k = 10000 is built into the kernel")
    return _one("max_exp", values)


fair_fn = {                                                # src/lb/env.py:152-161
    'jain': calcul_fair_jain, 'product': calcul_fair_product, 'var': calcul_fair_variance,
    'var_exp': calcul_fair_variance_exp, 'var_log': calcul_fair_variance_log,
    'max': calcul_max, 'max_exp': calcul_max_exp, 'max_log': calcul_max_log,
}


def calcul_fair(values, type):                             # src/lb/env.py:158-161
    return fair_fn[type](values)


class RewardFunction:
    """Configurable reward (rewards.py:290-388): same attributes, `compute(obs_dict)` contract."""

    SUPPORTED_METRICS = {
        'jain': jain_fairness, 'variance': variance_fairness, 'std': std_fairness,
        'cv': coefficient_of_variation, 'max': max_min_fairness, 'min': min_max_fairness,
        'product': product_fairness, 'range': range_fairness, 'gini': gini_coefficient,
        # beyond rewards.py:297-307: the original testbed's table (src/lb/env.py:152-161) under distinct names
        'fair_jain': calcul_fair_jain, 'fair_product': calcul_fair_product, 'var': calcul_fair_variance,
        'var_exp': calcul_fair_variance_exp, 'var_log': calcul_fair_variance_log,
        'max_exp': calcul_max_exp, 'max_log': calcul_max_log,
    }

    def __init__(self, metric: str = 'jain', reward_field: str = 'flow_duration_avg_decay'):
        if metric not in self.SUPPORTED_METRICS:
            raise ValueError(f"Unsupported metric: {metric}. "
                             f"Supported: {list(self.SUPPORTED_METRICS.keys())}")
        self.metric = metric
        self.reward_field = reward_field
        self._compute_func = self.SUPPORTED_METRICS[metric]

    def compute(self, observations: dict) -> float:
        active_servers = observations.get('active_servers', [])
        server_stats = observations.get('server_stats', {})
        if not active_servers:
            return 0.0                                   # rewards.py:364-365
        values = [server_stats[s][self.reward_field] for s in active_servers
                  if s in server_stats and self.reward_field in server_stats[s]]
        if not values:
            return 0.0                                   # rewards.py:375-376
        return self._compute_func(values)

    def __call__(self, observations: dict) -> float:
        return self.compute(observations)

    def __repr__(self):
        return f"RewardFunction(metric='{self.metric}', reward_field='{self.reward_field}')"


def create_jain_reward(field: str = 'flow_duration_avg_decay') -> RewardFunction:
    return RewardFunction(metric='jain', reward_field=field)


def create_variance_reward(field: str = 'flow_duration_avg_decay') -> RewardFunction:
    return RewardFunction(metric='variance', reward_field=field)


def create_max_reward(field: str = 'flow_duration_avg_decay') -> RewardFunction:
    return RewardFunction(metric='max', reward_field=field)
