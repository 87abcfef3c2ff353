"""GPU: the tcgen05 (3xTF32) general product of the update path (csrc/mlb_gemm_tc.cu) against a float64 torch
reference: every operand layout (forward x W^T, input gradient dy W, weight gradient dy^T x), tile tails in M / N / K,
split-K and single-CTA-per-tile shapes, beta accumulation, bias + activation, and the dispatch inside mlb_gemm
(ops.linear / matmul_nn / matmul_tn end on it above the size threshold).
Tolerance: 1e-5 relative (north_star) with the absolute floor an fp32 dot product of length K of unit-normal data needs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# (M, N, K): C4 update shapes (batch 256, state 2816 / 3072, gates 384, hidden 256), QMIX-sized ones, ragged tails
SHAPES = [(256, 384, 3072), (256, 384, 128), (256, 256, 256), (384, 3072, 256), (256, 256, 384), (128, 64, 32),
          (200, 100, 70), (129, 36, 33), (640, 520, 96), (64, 192, 352), (1600, 128, 64), (100, 2816, 256)]


def _tol(K):
    return dict(rtol=1e-5, atol=2e-7 * K + 1e-6)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("trans_a,trans_b", [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tc_layouts_match_float64(M, N, K, trans_a, trans_b):
    from marllb_b200.policy import ops
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + 2 * trans_a + trans_b)
    A = torch.randn((K, M) if trans_a else (M, K), device="cuda", generator=g)
    B = torch.randn((N, K) if trans_b else (K, N), device="cuda", generator=g) * 0.3
    Ad = (A.T if trans_a else A).double()
    Bd = (B.T if trans_b else B).double()
    ref = Ad @ Bd
    sup = ops._L().mlb_gemm_tc_supported(ops._p(A), 1 if trans_a else K, M if trans_a else 1, ops._p(B),
                                         1 if trans_b else N, K if trans_b else 1, ops._p(A), N, M, N, K)
    if (N % 4) or (trans_a and M % 4) or (not trans_a and K % 4) or (trans_b and K % 4) or N < 32 or K < 32:
        assert not sup                                        # strides TMA cannot take: stays on the FFMA kernel
        return
    assert sup
    y = ops.gemm_tc(A, B, trans_a=trans_a, trans_b=trans_b)
    np.testing.assert_allclose(y.cpu().numpy(), ref.cpu().numpy(), **_tol(K))
    # beta accumulation (dW += ...), bias + ReLU
    C0 = torch.randn(M, N, device="cuda", generator=g)
    y2 = ops.gemm_tc(A, B, out=C0.clone(), beta=1.0, trans_a=trans_a, trans_b=trans_b)
    np.testing.assert_allclose(y2.cpu().numpy(), (ref + C0.double()).cpu().numpy(), **_tol(K))
    bias = torch.randn(N, device="cuda", generator=g)
    y3 = ops.gemm_tc(A, B, bias=bias, act=ops.ACT_RELU, trans_a=trans_a, trans_b=trans_b)
    np.testing.assert_allclose(y3.cpu().numpy(), (ref + bias.double()).clamp_min(0).cpu().numpy(), **_tol(K))
    # deterministic (split-K partial sums are added in split order)
    assert torch.equal(y, ops.gemm_tc(A, B, trans_a=trans_a, trans_b=trans_b))


def test_layer_ops_reach_the_tensor_core_kernel_and_agree_with_ffma(monkeypatch):
    """ops.linear / matmul_nn / matmul_tn at C4 update sizes: mlb_gemm's dispatch takes the tcgen05 path; same numbers
    as the FFMA kernel (forced by calling with sizes below / above the threshold is not possible per call, so the FFMA
    result comes from a batched call, which never dispatches)."""
    from marllb_b200.policy import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    M, K, N = 256, 3072, 384
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    b = torch.randn(N, device="cuda", generator=g)
    dy = torch.randn(M, N, device="cuda", generator=g)
    y = ops.linear(x, W, b, ops.ACT_RELU)
    y_ffma = ops.linear(x[None], W[None], b[None], ops.ACT_RELU)[0]
    np.testing.assert_allclose(y.cpu().numpy(), (x.double() @ W.double().T + b.double()).clamp_min(0).cpu().numpy(), **_tol(K))
    np.testing.assert_allclose(y.cpu().numpy(), y_ffma.cpu().numpy(), rtol=2e-5, atol=1e-4)
    dx = ops.matmul_nn(dy, W)
    np.testing.assert_allclose(dx.cpu().numpy(), (dy.double() @ W.double()).cpu().numpy(), **_tol(N))
    dW = torch.randn(N, K, device="cuda", generator=g)
    dW0 = dW.clone()
    ops.matmul_tn(dy, x, out=dW, beta=1.0)
    np.testing.assert_allclose(dW.cpu().numpy(), (dW0.double() + dy.double().T @ x.double()).cpu().numpy(), **_tol(M))
