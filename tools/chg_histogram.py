"""Histogram of (n_old, changed slots) per reservoir in one step of the bench's c5 workload: which entries each
statistics kernel sees.  python tools/chg_histogram.py [--at 300]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marllb_b200 import VecLoadBalanceEnv
ap = argparse.ArgumentParser(); ap.add_argument("--at", type=int, default=300); ap.add_argument("--envs", type=int, default=32768)
a = ap.parse_args()
E, S = a.envs, 64
env = VecLoadBalanceEnv(E, num_servers=S, max_steps=a.at + 20)
env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
env.gen_poisson(128.0, 0.8 * 1.5 * S / 128.0, (a.at + 20) * 0.25, seed=1234)
env.reset()
g = torch.Generator(device="cuda").manual_seed(0)
for k in range(a.at):
    env.step(torch.randint(0, 3, (E, S), device="cuda", dtype=torch.uint8, generator=g))
v0, t0 = env.state_view("res_values").clone(), env.state_view("res_ts").clone()
c0 = env.state_view("res_count").clone()          # [E][2][S]
env.step(torch.randint(0, 3, (E, S), device="cuda", dtype=torch.uint8, generator=g))
v1, t1 = env.state_view("res_values"), env.state_view("res_ts")
nchg = ((v1 != v0) | (t1 != t0)).sum(-1)           # [E][S][2]   (a slot rewritten with identical bytes is not counted)
nold = c0.clamp(max=128).permute(0, 2, 1)          # -> [E][S][2]
for m, name in ((0, "fct"), (1, "flow_duration")):
    n_, c_ = nold[..., m].flatten().cpu().numpy(), nchg[..., m].flatten().cpu().numpy()
    tot = len(n_)
    print(f"{name}: touched {np.mean(c_ > 0):.3f} of reservoirs; full at step start {np.mean(n_ == 128):.3f}")
    for label, sel in (("full, 1 slot   (pair<.,0>)", (n_ == 128) & (c_ == 1)), ("full, 2 slots", (n_ == 128) & (c_ == 2)),
                       ("full, 3 slots", (n_ == 128) & (c_ == 3)), ("full, >3 slots (re-sort)", (n_ == 128) & (c_ > 3)),
                       ("filling, 1 append", (n_ < 128) & (n_ > 0) & (c_ == 1)), ("filling, 2 appends", (n_ < 128) & (n_ > 0) & (c_ == 2)),
                       ("filling, 3 appends", (n_ < 128) & (n_ > 0) & (c_ == 3)), ("filling, >3 / first sample (re-sort)", (n_ < 128) & ((c_ > 3) | ((n_ == 0) & (c_ > 0))))):
        print(f"    {label:40s} {sel.sum() / tot:.4f}")
