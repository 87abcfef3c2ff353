"""Which reward / config gives the training driver a learnable signal?  Env-only baselines (random weights, equal
weights, weights proportional to speed) for several reward fields, then a short SAC run through TrainingPipeline."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from marllb_b200 import VecLoadBalanceEnv
from marllb_b200.training_pipeline import TrainingPipeline, MAX_EPISODE_STEPS

S, E, T = 8, 64, 200
speeds = np.array([2, 2, 2, 2, 1, 1, 1, 1], np.float32)
rate = 24.0
mean_work = 0.8 * speeds.sum() / rate
for K in (128, 16):
    for field in ("flow_duration_avg_decay", "fct_mean", "fct_mean_decay", "flow_duration_mean"):
        env = VecLoadBalanceEnv(E, num_servers=S, action_type="continuous", max_steps=T, reward_field=field, reservoir_capacity=K)
        env.set_speeds(speeds)
        out = {}
        for name in ("random", "equal", "prop", "inv"):
            env.gen_poisson(rate, mean_work, T * 0.25 + 1, seed=3); env.reset()
            g = torch.Generator(device="cuda"); g.manual_seed(0)
            tot = 0.0
            for t in range(T):
                if name == "random": a = torch.rand((E, S), generator=g, device="cuda") * 2 - 1     # SAC's raw action range; env clips to [0.1, 10]
                elif name == "equal": a = torch.full((E, S), 0.5, device="cuda")
                elif name == "prop": a = torch.as_tensor(speeds / 2.0).cuda().expand(E, S).contiguous()
                else: a = torch.as_tensor(1.0 / speeds).cuda().expand(E, S).contiguous()
                _, r, _ = env.step(a); tot += float(r[..., None].mean()) if t >= 50 else 0.0
            out[name] = tot / (T - 50)
        print(f"K={K:3d} {field:26s}", {k: round(v, 4) for k, v in out.items()}, flush=True)
        env.close()

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
upd = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = dict(server_speeds=list(speeds), rates=[rate], seed=1, updates_per_round=upd, batch_size=256)
torch.manual_seed(0); np.random.seed(0)
import random; random.seed(0)
tp = TrainingPipeline('sac-gru', num_servers=S, num_agents=1, trace_dir='/nonexistent', checkpoint_dir='/tmp/lp_ck2',
                      config=cfg, num_envs=16, verbose=False)
ev = lambda: float(tp._run_round(0, explore=False, learn=False)[0].mean()) / MAX_EPISODE_STEPS
print("SAC greedy before:", ev(), flush=True)
t0 = time.time()
for r in range(rounds):
    rets, loss = tp._run_round(r * 16)
    print(f"round {r+1}: train {rets.mean()/MAX_EPISODE_STEPS:.4f} q1loss {loss} greedy {ev():.4f} alpha {float(tp.agent.alpha.item()):.3f} t={time.time()-t0:.0f}s", flush=True)
