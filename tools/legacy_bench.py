"""Throughput of the reference's actual simulation-mode step (env.py:215-286: random observation from the env's
RandomState stream + reward), E envs per launch through VecLegacyEnv / mlb_legacy_step, with the C oracle's
single-thread restatement of the same step timed beside it.  Not the driver's bench (that is bench.py, the
flow-level step); this documents SURVEY 8a rows a2 / a4 / a6 at speed.
    python tools/legacy_bench.py [--envs 131072] [--servers 64] [--steps 50]
Prints one JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from marllb_b200 import VecLegacyEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--servers", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    E, S = a.envs, a.servers
    env = VecLegacyEnv(E, num_servers=S, seeds=1000)
    env.reset()
    for _ in range(5):
        env.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        env.step()
    e1.record()
    torch.cuda.synchronize()
    env.check_status()
    ms = e0.elapsed_time(e1) / a.steps
    # end to end: pinned device->host copy of obs + reward every step
    h_obs = torch.empty((E, S, 11), dtype=torch.float32, pin_memory=True)
    h_rew = torch.empty((E,), dtype=torch.float64, pin_memory=True)
    t0 = time.perf_counter()
    n2 = max(3, a.steps // 10)
    for _ in range(n2):
        o, r, _ = env.step()
        h_obs.copy_(o, non_blocking=True)
        h_rew.copy_(r, non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / n2 * 1e3
    # algorithmic bytes per env-step: generator state read + written once (2 x 2500 B), obs row writes, reward
    bytes_step = 2 * 625 * 4 + S * 44 + 8
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # CPU: the C oracle's restatement of the same step, one thread
    import flow_oracle as fo
    g = fo.LegacyObs(1000)
    t0 = time.perf_counter()
    n = 0
    while time.perf_counter() - t0 < 3.0:
        for _ in range(200):
            o = g.next(S)
            fo.reward_metric("jain", o[:, 10].astype(np.float64))
        n += 200
    cpu = n / (time.perf_counter() - t0)
    print(json.dumps({"workload": f"legacy simulation-mode step: {E} envs x {S} servers", "ms_per_step": ms,
                      "env_steps_per_s": E / ms * 1e3, "e2e_env_steps_per_s": E / e2e_ms * 1e3,
                      "roofline": {"bound": "hbm", "algorithmic_bytes_per_env_step": bytes_step,
                                   "achieved": bytes_step * E / ms / 1e6, "peak": peak, "unit": "GB/s",
                                   "frac": bytes_step * E / ms / 1e6 / peak},
                      "cpu_baseline": {"value": cpu, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                       "sample": "oracle/flow_oracle.c ora_legacy_obs + jain reward, 3 s"}}))


if __name__ == "__main__":
    main()
