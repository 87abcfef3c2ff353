"""ctypes binding of libmarllb_b200.so -- the C ABI declared in include/marllb_b200.h.

There is no CPU fallback: if the library is missing this module raises, and
every compute entry point needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

# ---- constants mirrored from include/marllb_b200.h
ABI_VERSION = 1
OK, EINVAL, ECUDA, ENOMEM, ESTATE, ERNG, EACTION = 0, -1, -2, -3, -4, -5, -6
HOST, DEVICE = 0, 1
POLICIES = {"sed": 0, "lsq": 1, "alias": 2, "sed2": 3, "lsq2": 4}    # node.c:393-460
ACTION_DISCRETE_I32, ACTION_CONTINUOUS_F32, ACTION_DISCRETE_U8 = 0, 1, 2
RNG_MODES = {"replay": 0, "philox": 1}
METRICS = {"jain": 0, "variance": 1, "std": 2, "cv": 3, "max": 4, "min": 5,
           "product": 6, "range": 7, "gini": 8,           # rewards.py:297-307
           # the original testbed's fair_fn table, src/lb/env.py:152-161
           "fair_jain": 9, "fair_product": 10, "var": 1, "var_exp": 11, "var_log": 12,
           "max_exp": 13, "max_log": 14}
FEATURE_NAMES = ['n_flow_on', 'fct_mean', 'fct_p90', 'fct_std', 'fct_mean_decay', 'fct_p90_decay',
                 'flow_duration_mean', 'flow_duration_p90', 'flow_duration_std',
                 'flow_duration_mean_decay', 'flow_duration_avg_decay']   # env.py:377-381
(F_N_FLOW_ON, F_RES_VALUES, F_RES_TS, F_RES_COUNT, F_RES_CURSOR, F_DROPPED, F_LAST_FIN,
 F_HEAD, F_STEP, F_OBS, F_ARR_CURSOR) = range(11)
PTR_OBS, PTR_REWARD, PTR_DONE, PTR_ASSIGN = 100, 101, 102, 103

EXPORTS = [
    "mlb_abi_version", "mlb_config_default", "mlb_create", "mlb_destroy", "mlb_last_error",
    "mlb_set_speeds", "mlb_load_arrivals", "mlb_gen_poisson", "mlb_get_arrivals", "mlb_reset",
    "mlb_step", "mlb_get_assignments", "mlb_device_ptr", "mlb_get_state", "mlb_status",
    "mlb_launch_count", "mlb_profile_begin", "mlb_profile_end", "mlb_profile_pair_ms", "mlb_mt19937_fill", "mlb_reservoir_add", "mlb_reservoir_features",
    "mlb_reward_metric", "mlb_legacy_seed", "mlb_legacy_obs", "mlb_normalize_obs",
    "mlb_stage_arrivals", "mlb_commit_arrivals", "mlb_legacy_step", "mlb_gen_poisson_window", "mlb_step_changed",
]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("num_envs", C.c_int32),
                ("num_agents", C.c_int32), ("servers_per_agent", C.c_int32),
                ("reservoir_k", C.c_int32), ("queue_cap", C.c_int32), ("policy", C.c_int32),
                ("action_kind", C.c_int32), ("n_discrete", C.c_int32),
                ("discrete_weights", C.c_float * 8), ("min_weight", C.c_float),
                ("max_weight", C.c_float), ("dt", C.c_float), ("decay", C.c_double),
                ("reward_metric", C.c_int32), ("reward_field", C.c_int32),
                ("max_steps", C.c_int32), ("rng_seed_base", C.c_uint32),
                ("rng_table_len", C.c_int32), ("feature_cache", C.c_int32),
                ("record_assign", C.c_int32), ("env_id_base", C.c_int32),
                ("rng_mode", C.c_int32), ("reserved", C.c_int32 * 5)]


class MlbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"marllb_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Load (building first if stale) the CUDA library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MARLLB_B200_LIB") or _build.LIB_PATH   # override: A/B builds of the same library
    if path == _build.LIB_PATH:
        # content-hash staleness check (see _build._source_hash); builds under a file lock
        path = _build.build(force=os.environ.get("MARLLB_B200_REBUILD") == "1")
    L = C.CDLL(path)
    vp, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "mlb_abi_version": (C.c_int, []),
        "mlb_config_default": (C.c_int, [C.POINTER(Config)]),
        "mlb_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
        "mlb_destroy": (C.c_int, [vp]),
        "mlb_last_error": (C.c_char_p, [vp]),
        "mlb_set_speeds": (C.c_int, [vp, vp, i64, C.c_int, vp]),
        "mlb_load_arrivals": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, vp]),
        "mlb_gen_poisson": (C.c_int, [vp, f64, f64, f64, u64, vp]),
        "mlb_gen_poisson_window": (C.c_int, [vp, f64, f64, f64, f64, u64, C.c_uint32, vp]),
        "mlb_get_arrivals": (C.c_int, [vp, i32, i32, vp, vp, vp, vp, i64, C.POINTER(i64)]),
        "mlb_reset": (C.c_int, [vp, vp, vp]),
        "mlb_step": (C.c_int, [vp, vp, C.c_int, vp, vp, vp, C.c_int, vp]),
        "mlb_step_changed": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, C.POINTER(i64), vp]),
        "mlb_get_assignments": (C.c_int, [vp, vp, i64, C.c_int, vp]),
        "mlb_device_ptr": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)]),
        "mlb_get_state": (C.c_int, [vp, C.c_int, vp, C.c_size_t, C.c_int]),
        "mlb_status": (C.c_int, [vp, vp]),
        "mlb_launch_count": (i64, [vp]),
        "mlb_profile_begin": (C.c_int, [vp, i32]),
        "mlb_profile_end": (C.c_int, [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(i32)]),
        "mlb_profile_pair_ms": (f64, [vp]),
        "mlb_mt19937_fill": (C.c_int, [C.c_uint32, vp, i64]),
        "mlb_reservoir_add": (C.c_int, [vp, vp, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp, vp, vp]),
        "mlb_reservoir_features": (C.c_int, [vp, vp, vp, i32, i32, f64, vp, vp, vp]),
        "mlb_reward_metric": (C.c_int, [C.c_int, vp, vp, i32, i32, vp, vp]),
        "mlb_legacy_seed": (C.c_int, [vp, vp, i32, vp]),
        "mlb_legacy_obs": (C.c_int, [vp, i32, i32, vp, vp]),
        "mlb_normalize_obs": (C.c_int, [vp, vp, vp, i64, vp, i64, vp]),
        "mlb_stage_arrivals": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "mlb_commit_arrivals": (C.c_int, [vp, vp]),
        "mlb_legacy_step": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    if L.mlb_abi_version() != ABI_VERSION:
        raise RuntimeError("libmarllb_b200.so ABI version mismatch; rebuild with marllb_b200._build.build(force=True)")
    _lib = L
    return L


def check(rc: int, handle=None):
    if rc != OK:
        msg = load().mlb_last_error(handle)
        raise MlbError(rc, (msg or b"").decode() or f"status {rc}")
    return rc
