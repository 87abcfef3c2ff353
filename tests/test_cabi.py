"""CPU: the C-ABI shared library builds, loads and exports every symbol the header declares;
argument validation and error reporting work without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from marllb_b200 import _build, _lib


def header_functions(name="marllb_b200.h"):
    src = open(os.path.join(ROOT, "include", name)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mlb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/marllb_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == [n for n in names if n in _lib.EXPORTS]
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_policy_header_symbols_are_exported_and_bound():
    from marllb_b200.policy import ops
    L = C.CDLL(_build.build())
    names = header_functions("marllb_b200_policy.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/marllb_b200_policy.h but not exported"
    assert set(names) == set(ops.POLICY_EXPORTS), set(names) ^ set(ops.POLICY_EXPORTS)
    ops._L()    # binds argtypes for every entry point


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _build.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_config_default_matches_reference_defaults():
    L = _lib.load()
    cfg = _lib.Config()
    assert L.mlb_config_default(C.byref(cfg)) == 0
    assert cfg.abi_version == L.mlb_abi_version() == _lib.ABI_VERSION
    assert cfg.servers_per_agent == 4 and cfg.reservoir_k == 128          # env.py:73, reservoir.py:31
    assert list(cfg.discrete_weights)[:3] == [1.0, 1.5, 2.0] and cfg.n_discrete == 3   # env.py:69
    assert cfg.min_weight == pytest.approx(0.1) and cfg.max_weight == 10.0 # env.py:76-77
    assert cfg.dt == 0.25 and cfg.max_steps == 10000 and cfg.decay == 0.9  # env.py:80-81
    assert cfg.reward_metric == _lib.METRICS["jain"] and cfg.reward_field == 10
    assert C.sizeof(_lib.Config) == 152


def test_create_rejects_bad_config_with_message():
    L = _lib.load()
    cfg = _lib.Config()
    L.mlb_config_default(C.byref(cfg))
    h = C.c_void_p()
    for field, val, frag in (("reservoir_k", 129, "reservoir_k"), ("servers_per_agent", 0, "servers_per_agent"),
                             ("reward_metric", 99, "Unsupported metric"), ("action_kind", 7, "Unknown action_type"),
                             ("num_agents", 33, "num_agents"), ("abi_version", 0, "abi_version")):
        bad = _lib.Config.from_buffer_copy(cfg)
        setattr(bad, field, val)
        rc = L.mlb_create(C.byref(bad), C.byref(h))
        assert rc == _lib.EINVAL and not h.value
        assert frag in L.mlb_last_error(None).decode()


def test_no_cpu_fallback_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from marllb_b200 import LoadBalanceEnv, VecLoadBalanceEnv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        VecLoadBalanceEnv(4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LoadBalanceEnv()
    # straight through the C ABI: creation must fail loudly, not degrade
    L = _lib.load()
    cfg = _lib.Config()
    L.mlb_config_default(C.byref(cfg))
    h = C.c_void_p()
    rc = L.mlb_create(C.byref(cfg), C.byref(h))
    assert rc in (_lib.ECUDA, _lib.ENOMEM) and not h.value
    assert L.mlb_last_error(None)


def test_python_api_errors_match_reference():
    from marllb_b200 import VecLoadBalanceEnv
    from marllb_b200.rewards import RewardFunction
    with pytest.raises(ValueError, match="Unknown action_type"):          # env.py:184
        VecLoadBalanceEnv(1, action_type="weird")
    with pytest.raises(ValueError, match="Unsupported metric"):           # rewards.py:321-323
        VecLoadBalanceEnv(1, reward_metric="nope")
    with pytest.raises(ValueError, match="Unsupported metric"):
        RewardFunction(metric="nope")
    rf = RewardFunction("jain", "fct_mean")
    assert rf.compute({"active_servers": [], "server_stats": {}}) == 0.0  # rewards.py:364-365
    assert rf.compute({"active_servers": [0], "server_stats": {0: {"other": 1.0}}}) == 0.0


def test_host_mt19937_matches_numpy():
    L = _lib.load()
    for seed in (0, 42, 63, 2**32 - 1):
        out = np.empty(2000, np.uint32)
        assert L.mlb_mt19937_fill(seed, out.ctypes.data_as(C.c_void_p), out.size) == 0
        ref = np.random.RandomState(seed).randint(0, 2**32, size=out.size, dtype=np.uint32)
        assert np.array_equal(out, ref)


def test_spaces_and_traces_host_logic(tmp_path):
    from marllb_b200.spaces import Box, MultiDiscrete
    from marllb_b200.traces import load_trace, poisson_trace, split_round_robin
    b = Box(low=0.1, high=10.0, shape=(4,), dtype=np.float32)
    assert b.shape == (4,) and np.all(b.low == np.float32(0.1)) and np.all(b.high == 10.0)
    assert b.contains(b.sample())
    m = MultiDiscrete([3] * 4)
    assert np.all(m.nvec == 3) and len(m.nvec) == 4 and m.contains(m.sample())
    p = tmp_path / "t.csv"
    p.write_text("time\tquery\n0.5\t/dummy.php/?n=2000000\n0.75\t/dummy.php/?n=500000\n3.0\t/dummy.php/?n=1\n")
    tr = load_trace(str(p), horizon=2.0)
    assert tr["time"].tolist() == [0.5, 0.75] and tr["work"].tolist() == [2.0, 0.5]
    parts = split_round_robin(load_trace(str(p)), 2)
    assert parts[0]["time"].tolist() == [0.5, 3.0] and parts[1]["time"].tolist() == [0.75]
    pt = poisson_trace(100.0, 2.0, rng=np.random.RandomState(0), servers=4)
    assert np.all(np.diff(pt["time"]) >= 0) and pt["time"].max() < 2.0 and set(pt) == {"time", "work", "bucket", "u"}


def test_bind_host_to_gpu_is_a_noop_without_a_gpu():
    """shard.bind_host_to_gpu never raises: without NVML / a device it returns None and leaves the affinity alone."""
    import os
    from marllb_b200.shard import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    prev = bind_host_to_gpu(0)
    assert prev is None or prev == before
    if prev is not None:
        os.sched_setaffinity(0, prev)
    assert os.sched_getaffinity(0) == before
