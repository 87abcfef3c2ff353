import sys
sys.path.insert(0, '.')
import torch
from marllb_b200.policy import ops
import os
def t(M, N, K, reps=50):
    x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); b = None if os.environ.get("NOBIAS") else torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    ops.linear_tc(x, W, b, out=out); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.linear_tc(x, W, b, out=out)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(reps): ops.linear_tc(x, W, b, out=out)
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for cfg in [(128, 32, 32), (128, 192, 32), (128, 192, 352), (16384, 32, 32), (16384, 192, 32), (16384, 192, 64), (16384, 192, 352), (16384, 64, 128), (32768, 192, 352), (18944, 128, 2816)]:
    print(cfg, "%.1f us per launch (graph replay)" % t(*cfg), flush=True)
