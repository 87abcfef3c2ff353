"""Arrival inputs: trace files and host-side Poisson generation.

Trace format (reference data/trace/poisson_for_loop/*.csv, replayed by
src/client/replay_fork_io.py:95-120): TSV with header `time<TAB>query`, column 0
the arrival time in seconds, the query `/dummy.php/?n=<work>` with <work> the
PHP loop count.  (The reference's own loader reads it with the wrong separator
and never uses it: training_pipeline.py:98-139, SURVEY App. C #10.)
"""
from __future__ import annotations

import numpy as np


def load_trace(path: str, horizon: float = None, work_scale: float = 1e-6):
    """Return {'time': float32[n], 'work': float32[n]}; work = n_loops * work_scale."""
    t, w = [], []
    with open(path) as f:
        header = f.readline()
        if not header.lower().startswith("time"):
            f.seek(0)
        for line in f:
            line = line.rstrip("\n")
            if not line:
                continue
            a, b = line.split("\t")[:2]
            ta = float(a)
            if horizon is not None and ta >= horizon:
                break
            t.append(ta)
            w.append(float(b.rsplit("n=", 1)[1]) if "n=" in b else float(b))
    return {"time": np.asarray(t, np.float64).astype(np.float32),
            "work": (np.asarray(w, np.float64) * work_scale).astype(np.float32)}


def split_round_robin(trace: dict, num_agents: int):
    """One trace shared by A agents: row r -> agent r mod A (replay_fork_io.py:112)."""
    return [{k: v[i::num_agents] for k, v in trace.items()} for i in range(num_agents)]


def poisson_trace(rate: float, duration: float, mean_work: float = 1.0, rng=None, servers: int = 0):
    """Host generator with the reference's semantics (training_pipeline.py:141-155):
    cumsum of exponential(1/rate) gaps, int(rate*duration) draws, kept while < duration."""
    rng = rng if rng is not None else np.random
    n = int(rate * duration)
    t = np.cumsum(rng.exponential(1.0 / rate, n))
    t = t[t < duration]
    out = {"time": t.astype(np.float32), "work": rng.exponential(mean_work, len(t)).astype(np.float32)}
    if servers:
        out["bucket"] = rng.randint(0, servers, len(t)).astype(np.int32)
        out["u"] = rng.random_sample(len(t)).astype(np.float32)
    return out
