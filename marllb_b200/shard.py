"""Env sharding over the GPUs of one box.

Env instances never interact (reference env.py:145-154 keeps all state per object), so the
data-parallel strategy is: whole envs to ranks, contiguous blocks, **no data-path collective**.
The only cross-rank traffic of the env path is the timing reduction in bench.py.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of envs owned by `rank`: returns (first_env, num_envs).

    Blocks differ by at most one env; `first_env` is what VecLoadBalanceEnv takes as `env_id_base`
    so that synthetic arrival streams are keyed by the GLOBAL env id."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total_envs, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def owner_of(env_id: int, total_envs: int, world: int) -> int:
    """Rank that owns global env `env_id` under shard_range()."""
    base, rem = divmod(total_envs, world)
    cut = rem * (base + 1)
    if env_id < cut:
        return env_id // (base + 1)
    return rem + (env_id - cut) // max(base, 1)


def bind_host_to_gpu(device: int = 0):
    """Pin the calling process to the CPU cores (hence the NUMA node) closest to `device`, as NVML reports them.
    Pinned host buffers allocated afterwards land on that node, so the per-step device->host copy of the
    observations does not cross the socket interconnect -- what limits the end-to-end step once several ranks of
    one box copy at the same time.  Returns the previous affinity set (restore with os.sched_setaffinity(0, prev))
    or None when NVML or the affinity call is unavailable."""
    import os
    try:
        import pynvml
        prev = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device
        if vis:
            ids = vis.split(",")
            if device < len(ids) and ids[device].strip().isdigit():
                phys = int(ids[device])
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return prev
    except Exception:
        return None
