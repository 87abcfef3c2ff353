"""One SAC update at C4 sizes (batch 256, state 2816, action 256), eager, for `ncu --metrics gpu__time_duration.sum`:
3 warm-up updates, then the profiled one between cudaProfilerStart/Stop."""
import sys
sys.path.insert(0, '.')
import torch
from marllb_b200.policy import SAC_GRU_Agent
S = 256
torch.manual_seed(0)
sac = SAC_GRU_Agent(state_dim=S * 11, action_dim=S, hidden_dim=256, gru_dim=128, batch_size=256)
B = 256
b = (torch.randn(B, S * 11, device="cuda"), torch.rand(B, S, device="cuda") * 2 - 1, torch.rand(B, 1, device="cuda"),
     torch.randn(B, S * 11, device="cuda"), torch.zeros(B, 1, device="cuda"), torch.zeros(1, B, 128, device="cuda"))
for _ in range(3):
    sac.update_parameters(1, batch=b, sync_stats=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sac.update_parameters(1, batch=b, sync_stats=False)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
