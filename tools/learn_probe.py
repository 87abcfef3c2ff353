"""Does the offline training driver learn?  8 servers, one fast (speed 4) and seven slow (speed 1): boosting the fast
server's weight evens out the flow durations (higher Jain reward).  Prints the mean per-step reward of (a) a uniformly
random server choice, (b) always boosting each fixed server, (c) the greedy QMIX policy before / after training."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from marllb_b200.training_pipeline import TrainingPipeline, MAX_EPISODE_STEPS
from marllb_b200.policy import ops

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
upd = int(sys.argv[2]) if len(sys.argv) > 2 else 8
speeds = [4.0] + [1.0] * 7
cfg = dict(server_speeds=speeds, rates=[24], seed=1, updates_per_round=upd, batch_size=32, max_seq_len=50,
           learning_rate=float(sys.argv[3]) if len(sys.argv) > 3 else 0.0005)
torch.manual_seed(0); np.random.seed(0)
tp = TrainingPipeline('qmix', num_servers=8, num_agents=1, trace_dir='/nonexistent', checkpoint_dir='/tmp/lp_ck',
                      config=cfg, num_envs=32, verbose=False)
E = tp.num_envs

def fixed(server):
    obs = tp._load_round(); ret = 0.0
    a = torch.zeros((E, 8), dtype=torch.uint8, device='cuda')
    if server >= 0: a[:, server] = 2
    for t in range(MAX_EPISODE_STEPS):
        if server == -2:
            a.zero_(); idx = torch.randint(0, 8, (E,), device='cuda'); a[torch.arange(E), idx] = 2
        _, r, _ = tp.env.step(a); ret += float(r.mean())
    return ret / MAX_EPISODE_STEPS

print("random choice      :", np.mean([fixed(-2) for _ in range(3)]))
print("no boost           :", fixed(-1))
for s in range(8): print(f"always boost {s}     :", fixed(s))
ev = lambda: float(np.mean([tp._run_round(0, explore=False, learn=False)[0].mean() for _ in range(2)])) / MAX_EPISODE_STEPS
print("greedy before      :", ev())
t0 = time.time()
for r in range(rounds):
    rets, loss = tp._run_round(r * E)
    if r % 5 == 4: print(f"round {r+1}: train reward/step {rets.mean()/MAX_EPISODE_STEPS:.4f} loss {loss} greedy {ev():.4f}  t={time.time()-t0:.0f}s", flush=True)
print("greedy after       :", ev())
