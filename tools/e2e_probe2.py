import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from marllb_b200 import VecLoadBalanceEnv
E, S = 131072, 64
env = VecLoadBalanceEnv(E, num_servers=S, max_steps=10**9, action_dtype="uint8")
env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0)); env.gen_poisson(128.0, 0.6, 90.0, seed=1); env.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
pool = [torch.randint(0, 3, (E, S), generator=g, device="cuda", dtype=torch.uint8) for _ in range(8)]
for k in range(311): env.step(pool[k % 8])
torch.cuda.synchronize()
h_act = [p.cpu().numpy() for p in pool[:2]]
env.step_host(h_act[0])
for k in range(10):
    t0 = time.perf_counter(); env.step_host(h_act[k % 2]); print(f"step_host {k}: {(time.perf_counter()-t0)*1e3:.2f} ms")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); env.step(pool[0]); ev1.record(); torch.cuda.synchronize(); print("device step ms", ev0.elapsed_time(ev1))
