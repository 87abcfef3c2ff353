"""ctypes binding of oracle/_build/libflow_oracle.so (TEST INFRASTRUCTURE ONLY).

May be imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  marllb_b200/ never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libflow_oracle.so")

METRICS = ["jain", "variance", "std", "cv", "max", "min", "product", "range", "gini",
           "fair_jain", "fair_product", "var_exp", "var_log", "max_exp", "max_log"]   # src/lb/env.py:152-161
POLICIES = ["sed", "lsq", "alias", "sed2", "lsq2"]
FIELDS = ['n_flow_on', 'fct_mean', 'fct_p90', 'fct_std', 'fct_mean_decay', 'fct_p90_decay',
          'flow_duration_mean', 'flow_duration_p90', 'flow_duration_std',
          'flow_duration_mean_decay', 'flow_duration_avg_decay']   # env.py:377-381


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "flow_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "_build/libflow_oracle.so"], check=True)
    return _SO


class _Cfg(C.Structure):
    _fields_ = [("num_agents", C.c_int32), ("servers_per_agent", C.c_int32),
                ("reservoir_k", C.c_int32), ("queue_cap", C.c_int32),
                ("policy", C.c_int32), ("action_kind", C.c_int32),
                ("n_discrete", C.c_int32), ("discrete_weights", C.c_float * 8),
                ("min_weight", C.c_float), ("max_weight", C.c_float),
                ("dt", C.c_float), ("decay", C.c_double),
                ("reward_metric", C.c_int32), ("reward_field", C.c_int32),
                ("max_steps", C.c_int32), ("seed_base", C.c_uint32),
                ("rng_mode", C.c_int32), ("env_id", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i32, i64, u32, f64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_double, C.c_float
    L.ora_mt_fill.argtypes = [u32, vp, i64]
    L.ora_philox4x32_10.argtypes = [vp, vp, vp]
    L.ora_reservoir_set_philox.argtypes = [vp, u32, u32, u32]
    L.ora_philox_word.restype = u32
    L.ora_philox_word.argtypes = [u32, u32, u32, u32]
    L.ora_reservoir_create.restype = vp
    L.ora_reservoir_create.argtypes = [C.c_int, u32]
    L.ora_reservoir_destroy.argtypes = [vp]
    L.ora_reservoir_reset.argtypes = [vp]
    L.ora_reservoir_add.restype = C.c_int
    L.ora_reservoir_add.argtypes = [vp, f32, f64]
    L.ora_reservoir_last_slot.restype = C.c_int
    L.ora_reservoir_last_slot.argtypes = [vp]
    L.ora_reservoir_features.argtypes = [vp, f64, f64, vp]
    L.ora_reservoir_get.argtypes = [vp, vp, vp, vp]
    L.ora_features.argtypes = [vp, vp, C.c_int, f64, f64, vp]
    L.ora_reward_metric.restype = f64
    L.ora_reward_metric.argtypes = [C.c_int, vp, C.c_int]
    L.ora_reward_from_obs.restype = f64
    L.ora_reward_from_obs.argtypes = [C.c_int, C.c_int, vp, C.c_int]
    L.ora_alias_build.argtypes = [vp, C.c_int, vp, vp]
    L.ora_legacy_obs.argtypes = [vp, C.c_int, vp]
    L.ora_mt_seed.argtypes = [vp, u32]
    L.ora_env_create.restype = vp
    L.ora_env_create.argtypes = [C.POINTER(_Cfg), vp]
    L.ora_env_destroy.argtypes = [vp]
    L.ora_env_set_arrivals.argtypes = [vp, C.c_int, vp, vp, vp, vp, i64]
    L.ora_env_reset.argtypes = [vp]
    L.ora_env_step.restype = i64
    L.ora_env_step.argtypes = [vp, vp, vp, vp, vp, vp, i64]
    L.ora_env_dump.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ora_env_step_batch.restype = i64
    L.ora_env_step_batch.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int]
    L.ora_max_threads.restype = C.c_int
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def mt_fill(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, np.uint32)
    lib().ora_mt_fill(seed, _p(out), n)
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c, k, out = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.empty(4, np.uint32)
    lib().ora_philox4x32_10(_p(c), _p(k), _p(out))
    return out


def philox_word(c: int, server: int, env: int, key0: int) -> int:
    return int(lib().ora_philox_word(c, server, env, key0))


class Reservoir:
    """C restatement of reference ReservoirSampler (reservoir.py:17-233)."""

    def __init__(self, capacity=128, seed=0):
        self.capacity = capacity
        self._h = lib().ora_reservoir_create(capacity, seed)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_reservoir_destroy(self._h)
            self._h = None

    def add(self, value, timestamp):
        return bool(lib().ora_reservoir_add(self._h, float(np.float32(value)), float(timestamp)))

    def set_philox(self, key0, server, env):
        lib().ora_reservoir_set_philox(self._h, key0, server, env)

    def last_slot(self):
        return lib().ora_reservoir_last_slot(self._h)

    def reset(self):
        lib().ora_reservoir_reset(self._h)

    def features(self, decay=0.9, now=0.0):
        out = np.empty(5, np.float64)
        lib().ora_reservoir_features(self._h, decay, now, _p(out))
        return out

    def state(self):
        v = np.empty(self.capacity, np.float32)
        t = np.empty(self.capacity, np.float64)
        c = np.zeros(1, np.uint64)
        lib().ora_reservoir_get(self._h, _p(v), _p(t), _p(c))
        return v, t, int(c[0])


def features(values, ts, decay=0.9, now=0.0):
    values = np.ascontiguousarray(values, np.float32)
    ts = np.ascontiguousarray(ts, np.float64)
    out = np.empty(5, np.float64)
    lib().ora_features(_p(values), _p(ts), len(values), decay, now, _p(out))
    return out


def reward_metric(metric: str, values) -> float:
    v = np.ascontiguousarray(values, np.float64)
    return lib().ora_reward_metric(METRICS.index(metric), _p(v), len(v))


def reward_from_obs(metric: str, field, obs) -> float:
    obs = np.ascontiguousarray(obs, np.float32)
    f = FIELDS.index(field) if isinstance(field, str) else int(field)
    return lib().ora_reward_from_obs(METRICS.index(metric), f, _p(obs), obs.shape[0])


def alias_build(p):
    p = np.ascontiguousarray(p, np.float64)
    prob = np.empty(len(p), np.float64)
    alias = np.empty(len(p), np.int32)
    lib().ora_alias_build(_p(p), len(p), _p(prob), _p(alias))
    return prob, alias


class LegacyObs:
    """env.py:425-448 random observation stream for RandomState(seed)."""

    def __init__(self, seed):
        self._st = C.create_string_buffer(4 * 624 + 8)
        lib().ora_mt_seed(C.cast(self._st, C.c_void_p), seed)

    def next(self, S):
        obs = np.empty((S, 11), np.float32)
        lib().ora_legacy_obs(C.cast(self._st, C.c_void_p), S, _p(obs))
        return obs


class FlowEnv:
    """C restatement of the composed flow-level env (one env)."""

    def __init__(self, num_agents, servers_per_agent, speeds, arrivals,
                 reservoir_k=128, queue_cap=160, dt=0.25, decay=0.9, policy="sed",
                 action_type="discrete", discrete_weights=None, min_weight=0.1,
                 max_weight=10.0, reward_metric="jain",
                 reward_field="flow_duration_avg_decay", max_steps=10000, seed_base=0,
                 rng_mode="replay", env_id=0):
        cfg = _Cfg()
        cfg.num_agents, cfg.servers_per_agent = num_agents, servers_per_agent
        cfg.reservoir_k, cfg.queue_cap = reservoir_k, queue_cap
        cfg.policy = POLICIES.index(policy)
        cfg.action_kind = 0 if action_type == "discrete" else 1
        dw = list(discrete_weights or [1.0, 1.5, 2.0])
        cfg.n_discrete = len(dw)
        for i, x in enumerate(dw):
            cfg.discrete_weights[i] = x
        cfg.min_weight, cfg.max_weight = min_weight, max_weight
        cfg.dt, cfg.decay = dt, decay
        cfg.reward_metric = METRICS.index(reward_metric)
        cfg.reward_field = FIELDS.index(reward_field) if isinstance(reward_field, str) else int(reward_field)
        cfg.max_steps, cfg.seed_base = max_steps, seed_base
        cfg.rng_mode, cfg.env_id = {"replay": 0, "philox": 1}[rng_mode], env_id
        self.cfg = cfg
        self.A, self.Sa = num_agents, servers_per_agent
        self.S, self.K = num_agents * servers_per_agent, reservoir_k
        self.discrete = cfg.action_kind == 0
        sp = np.ascontiguousarray(speeds, np.float32).reshape(self.S)
        self._h = lib().ora_env_create(C.byref(cfg), _p(sp))
        self._keep = []
        max_flows = 0
        for i, a in enumerate(arrivals):
            t = np.ascontiguousarray(a["time"], np.float32)
            w = np.ascontiguousarray(a["work"], np.float32)
            b = np.ascontiguousarray(a["bucket"], np.int32) if "bucket" in a else None
            u = np.ascontiguousarray(a["u"], np.float32) if "u" in a else None
            self._keep += [t, w, b, u]
            lib().ora_env_set_arrivals(self._h, i, _p(t), _p(w), _p(b), _p(u), len(t))
            max_flows += len(t)
        self._assign = np.empty(max(max_flows, 1), np.int32)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ora_env_destroy(self._h)
            self._h = None

    def reset(self):
        lib().ora_env_reset(self._h)
        return np.zeros((self.S, 11), np.float32)

    def step(self, action):
        a = np.ascontiguousarray(action, np.int32 if self.discrete else np.float32)
        obs = np.empty((self.S, 11), np.float32)
        rew = np.zeros(1, np.float64)
        done = np.zeros(1, np.uint8)
        n = lib().ora_env_step(self._h, _p(a), _p(obs), _p(rew), _p(done),
                               _p(self._assign), len(self._assign))
        return obs, float(rew[0]), bool(done[0]), {"assign": self._assign[:n].copy()}

    def dump(self):
        S, K = self.S, self.K
        n_on = np.empty(S, np.int32)
        vals = np.empty((S, 2, K), np.float32)
        ts = np.empty((S, 2, K), np.float32)
        cnt = np.empty((S, 2), np.int64)
        drp = np.empty(S, np.int64)
        lib().ora_env_dump(self._h, _p(n_on), _p(vals), _p(ts), _p(cnt), _p(drp))
        return {"n_flow_on": n_on, "res_values": vals, "res_ts": ts, "res_count": cnt, "dropped": drp}


def step_batch(envs, actions, nthreads=0):
    """Step a list of FlowEnv with pthreads; actions: (n,S) int32/float32."""
    n = len(envs)
    S = envs[0].S
    arr = (C.c_void_p * n)(*[e._h for e in envs])
    a = np.ascontiguousarray(actions)
    obs = np.empty((n, S, 11), np.float32)
    rew = np.empty(n, np.float64)
    done = np.empty(n, np.uint8)
    flows = lib().ora_env_step_batch(arr, n, _p(a), a.strides[0], _p(obs), _p(rew), _p(done), nthreads)
    return obs, rew, done, flows


def max_threads():
    return lib().ora_max_threads()
