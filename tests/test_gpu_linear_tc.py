"""GPU: the tcgen05 (3xTF32) nn.Linear forward against a float64 torch reference and against the
FFMA GEMM, over tile tails, strided rows, multi-tile N, bias / activation variants.
Tolerance: 1e-5 relative (north_star) with the absolute floor that an fp32 dot product of length K
of unit-normal data needs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 16, 32), (128, 192, 352), (1000, 192, 352), (4096, 128, 64), (513, 32, 128), (300, 5, 100),
          (256, 384, 2816), (2048, 384, 2944), (129, 256, 36), (2000, 200, 20), (640, 520, 96)]


def _ref(x, W, b, act):
    y = x.double() @ W.double().T
    if b is not None:
        y = y + b.double()
    if act == 1:
        y = y.clamp_min(0)
    elif act == 2:
        y = y.abs()
    return y


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_linear_tc_matches_float64(M, N, K):
    from marllb_b200.policy import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.3
    b = torch.randn(N, device="cuda", generator=g)
    for bias, act in ((b, ops.ACT_NONE), (b, ops.ACT_RELU), (None, ops.ACT_ABS)):
        y = ops.linear_tc(x, W, bias, act)
        ref = _ref(x, W, bias, act)
        np.testing.assert_allclose(y.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=2e-7 * K + 1e-6)
    # and it is at least as accurate as the FFMA kernel on the same data
    G = torch.empty((M, N), device="cuda")
    from marllb_b200.policy.ops import _L, _p, _st, check
    # (two identical batches: batched calls stay on the FFMA tile kernel, single ones may be routed to the tensor-core or
    # the skinny kernels by mlb_gemm itself)
    G2 = torch.empty((2, M, N), device="cuda")
    check(_L().mlb_gemm(_p(x), 0, K, 1, _p(W), 0, 1, K, _p(G2), M * N, N, _p(b), 0, M, N, K, 2, 0.0, 0, _st()))
    G = G2[0]
    ref = _ref(x, W, b, 0)
    e_tc = (ops.linear_tc(x, W, b).double() - ref).abs().max().item()
    e_ff = (G.double() - ref).abs().max().item()
    assert e_tc <= 4 * e_ff + 1e-6, (e_tc, e_ff)


def test_linear_tc_strided_rows_and_output_view():
    from marllb_b200.policy import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randn(1500, 2, 352, device="cuda", generator=g)       # [E, A, obs]: one agent's rows are strided
    W = torch.randn(192, 352, device="cuda", generator=g) * 0.1
    b = torch.randn(192, device="cuda", generator=g)
    out = torch.zeros(1500, 2, 192, device="cuda")
    for a in range(2):
        ops.linear_tc(obs[:, a], W, b, out=out[:, a])
        ref = _ref(obs[:, a], W, b, 0)
        np.testing.assert_allclose(out[:, a].cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-4)


def test_linear_dispatch_uses_tensor_cores_for_large_m():
    from marllb_b200.policy import ops
    x = torch.randn(ops.TC_MIN_M, 64, device="cuda")
    W = torch.randn(128, 64, device="cuda")
    y_big = ops.linear(x, W)                 # tensor-core path
    y_small = ops.linear(x[:64].contiguous(), W)   # FFMA path
    np.testing.assert_allclose(y_big[:64].cpu().numpy(), y_small.cpu().numpy(), rtol=1e-5, atol=2e-5)
