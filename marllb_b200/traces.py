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


class TraceStream:
    """A long trace cut into chunks of `steps_per_chunk` env steps.

    `source`: one trace (a path, or a dict of 'time' / 'work' [/ 'bucket', 'u'] arrays) that every env
    replays, or a list with one source per env; each env's trace is split round-robin over its agents
    (row r -> agent r mod A, replay_fork_io.py:112).  Files are read lazily, line by line, as the
    chunks are requested, so an hour-long trace never has to be in memory, let alone in HBM.

    Chunk c holds, for every (env, agent) stream, exactly the arrivals the kernel will consume in steps
    [c*steps_per_chunk, (c+1)*steps_per_chunk): those with float32 time < float32(k_end) * float32(dt),
    the same float32 comparison the event kernel makes, so a chunk swap never strands a flow.
    """

    def __init__(self, source, num_envs: int, num_agents: int = 1, dt: float = 0.25, steps_per_chunk: int = 64,
                 work_scale: float = 1e-6):
        self.sources = list(source) if isinstance(source, (list, tuple)) else [source]
        if len(self.sources) not in (1, num_envs):
            raise ValueError("one trace for all envs, or one per env")
        self.num_envs, self.num_agents = num_envs, num_agents
        self.dt, self.steps_per_chunk, self.work_scale = np.float32(dt), steps_per_chunk, work_scale
        self.restart()

    # -- per-source row readers ------------------------------------------------------------------
    def _rows(self, src):
        """Generator of (time f32, work f32, bucket or None, u or None) in file order."""
        if isinstance(src, dict):
            b, u = src.get("bucket"), src.get("u")
            for i in range(len(src["time"])):
                yield (np.float32(src["time"][i]), np.float32(src["work"][i]),
                       None if b is None else int(b[i]), None if u is None else np.float32(u[i]))
            return
        with open(src) as f:
            first = f.readline()
            if not first.lower().startswith("time"):
                f.seek(0)
            for line in f:
                line = line.rstrip("\n")
                if not line:
                    continue
                a, q = line.split("\t")[:2]
                w = float(q.rsplit("n=", 1)[1]) if "n=" in q else float(q)
                yield np.float32(float(a)), np.float32(w * self.work_scale), None, None

    def restart(self):
        self._gens = [self._rows(s) for s in self.sources]
        self._pending = [None] * len(self.sources)     # first row of the next chunk, already read
        self._row_no = [0] * len(self.sources)
        self._chunk = 0
        self._done = [False] * len(self.sources)

    @property
    def exhausted(self):
        return all(self._done) and all(p is None for p in self._pending)

    def next_chunk(self):
        """-> (k_end, CSR dict: time, work, offsets [, bucket, u]) for the next steps_per_chunk steps."""
        k_end = (self._chunk + 1) * self.steps_per_chunk
        t_end = np.float32(np.float32(k_end) * self.dt)              # event kernel: t1 = __fmul_rn((float)step, dt)
        A = self.num_agents
        per_src = []
        for s, gen in enumerate(self._gens):
            cols = [([], [], [], []) for _ in range(A)]
            while True:
                row = self._pending[s]
                if row is None:
                    row = next(gen, None)
                    if row is None:
                        self._done[s] = True
                        break
                if not (row[0] < t_end):
                    self._pending[s] = row
                    break
                self._pending[s] = None
                c = cols[self._row_no[s] % A]
                self._row_no[s] += 1
                c[0].append(row[0]); c[1].append(row[1]); c[2].append(row[2]); c[3].append(row[3])
            per_src.append(cols)
        self._chunk += 1
        streams = [per_src[e if len(per_src) > 1 else 0][a] for e in range(self.num_envs) for a in range(A)]
        off = np.zeros(len(streams) + 1, np.int64)
        off[1:] = np.cumsum([len(c[0]) for c in streams])
        cat = lambda i, dt_: np.concatenate([np.asarray(c[i], dt_) for c in streams]) if off[-1] else np.zeros(0, dt_)
        out = {"time": cat(0, np.float32), "work": cat(1, np.float32), "offsets": off}
        if off[-1] and streams[int(np.argmax(np.diff(off)))][2][0] is not None:
            out["bucket"] = cat(2, np.int32)
            if streams[int(np.argmax(np.diff(off)))][3][0] is not None:
                out["u"] = cat(3, np.float32)
        return k_end, out
