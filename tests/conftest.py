import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def flow_case(name):
    """Return (cfg dict, arrivals list, golden dict) of a tests/golden/flow_*.npz fixture."""
    g = load_golden(name)
    cfg = json.loads(str(g["cfg"]))
    A = cfg["A"]
    arrivals = []
    for i in range(A):
        a = {"time": g[f"arr_time_{i}"], "work": g[f"arr_work_{i}"]}
        if f"arr_bucket_{i}" in g:
            a["bucket"], a["u"] = g[f"arr_bucket_{i}"], g[f"arr_u_{i}"]
        arrivals.append(a)
    return cfg, arrivals, g


FLOW_CASES = ["flow_c1_trace", "flow_sed_a2s3", "flow_lsq_s5", "flow_alias_a2s4", "flow_cont_s4",
              "flow_drops_q4", "flow_k8_s4", "flow_s40_var", "flow_a4s16_gini", "flow_sed2_a2s5", "flow_lsq2_s36"]


def env_kwargs(cfg):
    kw = {k: v for k, v in cfg.items() if k not in ("A", "Sa", "steps")}
    return kw


# tolerances (stated once, used by every parity test)
OBS_RTOL = 1e-5      # float32 observations: relative, north_star "1e-5 fp32"
OBS_ATOL = 1e-7
REWARD_RTOL = 1e-9   # float64 rewards computed from bit-identical inputs
