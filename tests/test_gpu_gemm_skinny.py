"""Streaming kernels for skinny projections (csrc/gemm_skinny.cu) vs fp64: every dispatch class
(small-K, small-N, its transposed form, tall-T with the skinny side as M or N, M-tiny NN), all four
transpositions, alpha/beta/bias, strided operands, ragged row / contraction bounds, and colsum.
Tolerance 2e-6 max-norm relative: these kernels are exact-fp32 FMA chains with tree reductions."""
import pytest
import torch

import gpu_common as G

pytestmark = pytest.mark.gpu
TOL = 2e-6

SHAPES = [
    (6144, 768, 4),    # small-K   (proj_q forward / d(out) of a rank-C fold)
    (300, 772, 16),    # small-K, N % 4 == 0 but not a multiple of the block
    (257, 70, 5),      # small-K, ragged N tail
    (6144, 4, 768),    # small-N   (residual_head)
    (1000, 16, 772),   # small-N, N = 4C
    (513, 20, 776),    # small-N wide variant (C = 5)
    (4, 768, 6144),    # tall-T, skinny M (dW_r)
    (768, 4, 6144),    # tall-T, skinny N (dW_Q)
    (12, 4, 3000),     # both skinny (dW_hh)
    (20, 776, 1500),   # tall-T in two chunks of S (C = 5)
    (4, 768, 768),     # M-tiny (weight-space fold W_r W_o)
]


@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_skinny_auto_backend_matches_fp64(tA, tB, M, N, K):
    from immtsf import ops

    g = torch.Generator().manual_seed(M * 7 + N * 3 + K + 2 * tA + tB)
    A = torch.randn((K, M) if tA else (M, K), generator=g)
    Bm = torch.randn((N, K) if tB else (K, N), generator=g)
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = 0.75 * ((A.double().T if tA else A.double()) @ (Bm.double().T if tB else Bm.double())) + 0.5 * C0.double() + bias.double()
    C = C0.clone().cuda()
    ops.gemm(A.cuda(), Bm.cuda(), C, transA=bool(tA), transB=bool(tB), bias=bias.cuda(), alpha=0.75, beta=0.5)
    G.assert_close("gemm skinny", C.cpu(), ref, TOL)
    C2 = torch.full((M, N), float("nan"), device="cuda")  # beta == 0 must not read C
    ops.gemm(A.cuda(), Bm.cuda(), C2, transA=bool(tA), transB=bool(tB))
    G.assert_close("gemm skinny beta0", C2.cpu(), (ref - 0.5 * C0.double() - bias.double()) / 0.75, TOL)


def test_skinny_strided_views_and_ragged_rows():
    from immtsf import ops

    g = torch.Generator().manual_seed(3)
    M, K, N, m = 640, 768, 4, 300
    big = torch.randn(M, K + 8, generator=g).cuda()
    A = big[:, 4:4 + K]
    A[m:] = 0.0
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    m_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
    out = torch.full((M, N), 7.0, device="cuda")
    ops.gemm(A, W, out, transB=True, bias=b, ragged=m_dev, ragged_dim=1)  # small-N, ragged rows
    ref = A.double().cpu() @ W.double().cpu().T + b.double().cpu()
    G.assert_close("rows<m", out[:m].cpu(), ref[:m], TOL)
    assert (out[m:384] == 0).all() and (out[384:] == 7.0).all()
    # small-K with ragged rows: dE = dG W[:, C:] on a column-offset view of W
    Wc = torch.randn(16, 4 + 768, generator=g).cuda()
    dG = torch.randn(M, 16, generator=g).cuda()
    dE = torch.full((M, 768), 7.0, device="cuda")
    ops.gemm(dG, Wc[:, 4:], dE, ragged=m_dev, ragged_dim=1)
    G.assert_close("smallk rows<m", dE[:m].cpu(), dG[:m].double().cpu() @ Wc[:, 4:].double().cpu(), TOL)
    assert (dE[m:384] == 0).all() and (dE[384:] == 7.0).all()
    # tall-T with a ragged contraction (weight gradient over the live rows only)
    dy = torch.randn(M, N, generator=g).cuda()
    dw = ops.linear_wgrad(dy, A, ragged=m_dev)  # [N, K] = dy^T A
    G.assert_close("wgrad skinny-M", dw.cpu(), dy[:m].double().cpu().T @ A[:m].double().cpu(), TOL)
    dw2 = ops.linear_wgrad(A, dy, ragged=m_dev)  # [K, N] = A^T dy
    G.assert_close("wgrad skinny-N", dw2.cpu(), A[:m].double().cpu().T @ dy[:m].double().cpu(), TOL)


@pytest.mark.parametrize("M,N", [(6144, 768), (2176, 1536), (130, 4), (63, 40), (5000, 772)])
def test_colsum(M, N):
    from immtsf import ops

    g = torch.Generator().manual_seed(M + N)
    X = torch.randn(M, N, generator=g).cuda()
    G.assert_close("colsum", ops.colsum(X).cpu(), X.double().cpu().sum(0), TOL)
    m = M // 3
    m_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
    out = torch.ones(N, device="cuda")
    ops.colsum(X, out=out, ragged=m_dev, beta=2.0)
    G.assert_close("colsum ragged beta", out.cpu(), X[:m].double().cpu().sum(0) + 2.0, TOL)
    a, b = ops.colsum(X), ops.colsum(X)
    assert torch.equal(a, b)  # fixed summation order: bitwise reproducible
