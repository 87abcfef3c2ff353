import sys, time
sys.path.insert(0, '.')
import torch
from marllb_b200.policy import ops
torch.manual_seed(0)
for M, N, K in ((128, 16, 32), (128, 192, 352), (1000, 192, 352), (256, 384, 2816)):
    x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.3; b = torch.randn(N, device="cuda")
    y = ops.linear_tc(x, W, b)
    torch.cuda.synchronize()
    ref = x.double() @ W.double().T + b.double()
    err = (y.double() - ref).abs().max().item()
    print(M, N, K, "max abs err", err, "ref scale", ref.abs().max().item(), flush=True)
x = torch.randn(32768, 352, device="cuda"); W = torch.randn(192, 352, device="cuda"); b = torch.randn(192, device="cuda")
out = torch.empty(32768, 192, device="cuda")
def ffma():
    ops.check(ops._L().mlb_gemm(ops._p(x), 0, 352, 1, ops._p(W), 0, 1, 352, ops._p(out), 0, 192, ops._p(b), 0, 32768, 192, 352, 1, 0.0, 0, ops._st()))
for name, f in (("tc", lambda: ops.linear_tc(x, W, b, out=out)), ("ffma", ffma)):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, "32768x192x352: %.1f us, %.1f TFLOP/s" % (ms * 1e3, 2 * 32768 * 192 * 352 / ms / 1e9), flush=True)
