"""Update-path timings at the BASELINE config sizes (C3 QMIX update, C4 SAC update), CUDA events."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from marllb_b200.policy import QMIXAgent, SAC_GRU_Agent, ops

def timeit(f, n=10, warm=3):
    for _ in range(warm): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = ops.LAUNCHES; e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (ops.LAUNCHES - l0) / n

torch.manual_seed(0)
A, Sa, S = 2, 32, 64
agent = QMIXAgent(num_agents=A, state_dim=4 * S + 10, obs_dim=Sa * 11, action_dim=Sa, hidden_dim=64, gru_dim=64,
                  mixing_embed_dim=32, hypernet_embed_dim=64, batch_size=32, max_seq_len=50)
rng = np.random.RandomState(0)
B, T = 32, 50
batch = {'observations': rng.randn(B, T, A, Sa * 11), 'actions': rng.randint(0, Sa, (B, T, A, 1)).astype(np.float64),
         'rewards': rng.rand(B, T, A), 'states': rng.randn(B, T, 4 * S + 10), 'dones': np.zeros((B, T)),
         'seq_lengths': np.full(B, T, np.int32)}
ms, l = timeit(lambda: agent.update(batch=batch))
print("QMIX update  B=32 T=50 A=2 obs=352: %.2f ms, %d kernel launches" % (ms, l), flush=True)
agent.graph_updates = True
ms, l = timeit(lambda: agent.update(batch=batch))
print("QMIX update  (CUDA graph)          : %.2f ms" % ms, flush=True)

S = 256
sac = SAC_GRU_Agent(state_dim=S * 11, action_dim=S, hidden_dim=256, gru_dim=128, batch_size=256)
Bq = 256
b = (torch.randn(Bq, S * 11, device="cuda"), torch.rand(Bq, S, device="cuda") * 2 - 1, torch.rand(Bq, 1, device="cuda"),
     torch.randn(Bq, S * 11, device="cuda"), torch.zeros(Bq, 1, device="cuda"), torch.zeros(1, Bq, 128, device="cuda"))
ms, l = timeit(lambda: sac.update_parameters(1, batch=b))
print("SAC update   batch=256 state=2816 action=256: %.2f ms, %d kernel launches" % (ms, l), flush=True)
x = torch.randn(1024, S * 11, device="cuda")
ms, l = timeit(lambda: sac.select_action_batch(x))
print("SAC actor    1024 envs: %.3f ms, %d kernel launches" % (ms, l), flush=True)

# the same SAC update replayed from a CUDA graph: device time without Python launch overhead
upd = lambda: sac.update_parameters(1, batch=b, sync_stats=False)
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    upd(); upd()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
l0 = ops.LAUNCHES
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    upd()
nl = ops.LAUNCHES - l0
ms, _ = timeit(g.replay, n=30)
print("SAC update   (CUDA graph)          : %.3f ms, %d launches through ops in the graph" % (ms, nl), flush=True)
