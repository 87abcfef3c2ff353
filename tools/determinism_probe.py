"""Run the bench's c5 env twice in one process and compare per-step, per-env observation sums:
    python tools/determinism_probe.py [--envs N] [--steps K] [--runs R]
Prints the first step at which two runs differ and which envs / columns differ there.  The env step has no
cross-env communication and replays fixed random streams, so any difference is a race or an uninitialised read."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marllb_b200 import VecLoadBalanceEnv  # noqa: E402


def run(a, keep_step=None, keep_env=None, feature_cache=True):
    E, S = a.envs, a.servers
    total = a.steps + 16
    A = a.agents
    env = VecLoadBalanceEnv(E, num_servers=S, num_agents=A, max_steps=total + 1, feature_cache=feature_cache,
                            rng_mode=a.rng_mode)
    env.set_speeds(np.where(np.arange(S * A) % 2 == 0, 1.0, 2.0).astype(np.float32))
    rate = a.rate or 128.0 * S / 64
    env.gen_poisson(rate, 0.8 * 1.5 * S / rate, total * 0.25 + 1.0, seed=1234)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randint(0, 3, (E, S * A), device="cuda", dtype=torch.uint8, generator=g) for _ in range(8)]
    sums = torch.empty((a.steps, E), dtype=torch.float32, device="cuda")
    kept = None
    for k in range(a.steps):
        obs, _, _ = env.step(acts[k % 8])
        sums[k] = obs.view(E, -1).sum(1)
        if keep_step is not None and k == keep_step:
            kept = obs[keep_env].clone()
    torch.cuda.synchronize()
    env.check_status()
    env.close()
    del env
    torch.cuda.empty_cache()
    return sums, kept


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=131072)
    ap.add_argument("--servers", type=int, default=64, help="servers per agent")
    ap.add_argument("--agents", type=int, default=1)
    ap.add_argument("--rate", type=float, default=0.0, help="flows/s/agent (default 2 per server)")
    ap.add_argument("--rng-mode", default="replay", choices=["replay", "philox"])
    ap.add_argument("--steps", type=int, default=2148)
    ap.add_argument("--runs", type=int, default=3)
    ap.add_argument("--vs-resort", action="store_true",
                    help="compare the incremental statistics (feature_cache=True) with a full re-sort of every reservoir every "
                         "step (feature_cache=False): per-(step, env) observation sums within 2e-6 relative")
    a = ap.parse_args()
    ref, _ = run(a)
    if a.vs_resort:
        cur, _ = run(a, feature_cache=False)
        rel = ((cur - ref).abs() / ref.abs().clamp_min(1e-6))
        bad = rel > 2e-6
        print(f"incremental vs re-sort over {a.steps} steps x {a.envs} envs: max relative difference of an env's observation sum "
              f"{float(rel.max()):.3e}; {int(bad.sum())} (step, env) pairs above 2e-6", flush=True)
        if int(bad.sum()):
            st = bad.any(1).nonzero().flatten()
            print("  first steps:", st[:8].tolist(), "envs at the first:", bad[int(st[0])].nonzero().flatten()[:8].tolist())
        return
    for r in range(1, a.runs):
        cur, _ = run(a)
        diff = (cur != ref)
        nd = int(diff.sum())
        print(f"run {r}: {nd} differing (step, env) sums", flush=True)
        if nd:
            steps = diff.any(1).nonzero().flatten()
            first = int(steps[0])
            envs = diff[first].nonzero().flatten().tolist()
            print(f"  first differing step {first}, envs {envs[:16]} ({len(envs)} envs); last step: {int(diff[-1].sum())} envs differ")
            per_env_first = {}
            de = diff.any(0).nonzero().flatten().tolist()
            for e in de[:12]:
                s0 = int(diff[:, e].nonzero().flatten()[0])
                per_env_first[e] = (s0, float(ref[s0, e]), float(cur[s0, e]))
            print("  env -> (first step, ref sum, this sum):", per_env_first)
            print(f"  envs that ever differ: {len(de)}")


if __name__ == "__main__":
    main()
