import sys
sys.path.insert(0, '.')
import torch
from marllb_b200.policy import ops
def t(M, N, K, n=20):
    x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    f = lambda: ops.linear_tc(x, W, b, out=out)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for M in (128 * 148, 128 * 148 * 2, 32768):
    for N in (128, 192, 64, 32):
        print("M", M, "N", N, " ".join("K%d: %.1fus" % (K, t(M, N, K)) for K in (32, 128, 352, 1408, 2816)), flush=True)
