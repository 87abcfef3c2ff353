"""One full reference-length episode (max_steps = 10000, env.py:81) of 16384 envs x 64 servers with 128-slot
reservoirs, random policy: step time per 1000-step segment (it falls as Algorithm R's acceptance K/count falls),
flow conservation, sticky status, done flags.    python tools/full_episode.py [--envs N] [--steps T]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marllb_b200 import VecLoadBalanceEnv  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--servers", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10000)
    a = ap.parse_args()
    E, S, T = a.envs, a.servers, a.steps
    env = VecLoadBalanceEnv(E, num_servers=S, max_steps=T, action_dtype="uint8")
    env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0).astype(np.float32))
    rate = 128.0 * S / 64
    env.gen_poisson(rate, 0.8 * 1.5 * S / rate, T * 0.25 + 1.0, seed=2024)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = [torch.randint(0, 3, (E, S), device="cuda", dtype=torch.uint8, generator=g) for _ in range(16)]
    seg = max(T // 10, 1)
    print(f"{E} envs x {S} servers, {T} steps, Poisson {rate:.0f} flows/s/agent, rho 0.8, SED, random policy")
    k = 0
    while k < T:
        n = min(seg, T - k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            obs, rew, done = env.step(acts[(k + i) % 16])
        e1.record()
        torch.cuda.synchronize()
        k += n
        ms = e0.elapsed_time(e1) / n
        print(f"steps {k - n + 1:6d}-{k:6d}: {ms:.3f} ms/step = {E / ms / 1e3:7.2f} M agent-steps/s, "
              f"mean reward {float(rew.mean()):.6f}, flows in system {int(obs[..., 0].sum())}")
    env.check_status()
    assert bool(done.all()), "every env must be done at max_steps"
    arrived = int(env.get_state("arr_cursor").astype(np.int64).sum())
    n_on = int(env.get_state("n_flow_on").astype(np.int64).sum())
    dropped = int(env.get_state("dropped").astype(np.int64).sum())
    cnt = env.get_state("res_count").astype(np.int64)            # [E][2][S]
    completed = int(cnt[:, 0].sum())
    assert arrived == completed + n_on + dropped, (arrived, completed, n_on, dropped)
    print(f"flow conservation: arrived {arrived} = completed {completed} + in system {n_on} + dropped {dropped}")
    print(f"flow_duration samples {int(cnt[:, 1].sum())}; status OK; all {E} envs done at step {T}")


if __name__ == "__main__":
    main()
