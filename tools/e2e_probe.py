import sys, time, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
from marllb_b200 import VecLoadBalanceEnv, _lib
E, S = 131072, 64
env = VecLoadBalanceEnv(E, num_servers=S, max_steps=10**9, action_dtype="uint8")
env.set_speeds(np.where(np.arange(S) % 2 == 0, 1.0, 2.0)); env.gen_poisson(128.0, 0.6, 40.0, seed=1); env.reset()
act = np.random.randint(0, 3, (E, S)).astype(np.uint8)
for _ in range(20): env.step(act)
torch.cuda.synchronize()
def T(f, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
env.step_host(act)
print("step_host total ms", T(lambda: env.step_host(act)))
print("numpy->pinned copy ms", T(lambda: env._h_action.numpy().__setitem__(Ellipsis, act)))
da = torch.as_tensor(act).cuda()
print("device step ms", T(lambda: env.step(da)))
L = env._L; st = env._stream()
vp = lambda t: C.c_void_p(t.data_ptr())
print("host action only ms", T(lambda: L.mlb_step(env._h, vp(env._h_action), 0, None, None, None, 0, st)))
print("host action + obs D2H ms", T(lambda: L.mlb_step(env._h, vp(env._h_action), 0, vp(env._h_obs), None, None, 0, st)))
print("obs D2H via torch ms", T(lambda: env._h_obs.copy_(env.obs, non_blocking=True)))
