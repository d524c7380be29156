"""tcgen05 3xTF32 GEMM backend vs an fp64 reference (and vs the FFMA backend): all four
transposition cases, tile-unaligned sizes, leading dimensions, bias/alpha/beta, ragged bounds.
Tolerance: 4e-6 max-norm relative (fp32-class accuracy; a single TF32 pass would be ~1e-3)."""
import pytest
import torch

import gpu_common as G

pytestmark = pytest.mark.gpu
TOL = 4e-6


def _ref(A, B, tA, tB):
    return (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())


@pytest.mark.timeout(120)
@pytest.mark.parametrize("tA,tB", [(0, 1), (0, 0), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (256, 256, 768), (200, 96, 72), (6144, 768, 768), (2200, 1152, 768), (768, 1536, 2216)])
def test_gemm_tc_matches_fp64(tA, tB, M, N, K):
    from immtsf import ops

    g = torch.Generator().manual_seed(M + 3 * N + 7 * K + tA * 2 + tB)
    A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
    B = (torch.randn((N, K) if tB else (K, N), generator=g) * 0.3).cuda()
    C = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=ops.BACKEND_TC)
    torch.cuda.synchronize()
    G.assert_close("gemm_tc", C.cpu(), _ref(A.cpu(), B.cpu(), tA, tB), TOL)


@pytest.mark.timeout(120)
def test_gemm_tc_epilogue_and_strides():
    from immtsf import ops

    g = torch.Generator().manual_seed(9)
    M, N, K = 300, 200, 136
    Abig = torch.randn(M, K + 8, generator=g).cuda()
    A = Abig[:, 4:4 + K]  # 16B-aligned column offset, lda = K+8
    W = torch.randn(N, K, generator=g).cuda()
    bias = torch.randn(N, generator=g).cuda()
    Cbig = torch.randn(M, N + 4, generator=g).cuda()
    C0 = Cbig.clone()
    ops.gemm(A, W, Cbig[:, :N], transB=True, bias=bias, alpha=0.5, beta=2.0, backend=ops.BACKEND_TC)
    ref = 0.5 * (A.double().cpu() @ W.double().cpu().T) + 2.0 * C0[:, :N].double().cpu() + bias.double().cpu()
    G.assert_close("gemm_tc epilogue", Cbig[:, :N].cpu(), ref, TOL)
    assert torch.equal(Cbig[:, N:], C0[:, N:])  # columns outside N untouched


@pytest.mark.timeout(120)
def test_gemm_tc_ragged():
    from immtsf import ops

    g = torch.Generator().manual_seed(10)
    M, N, K = 640, 256, 128
    m = 300
    A = torch.randn(M, K, generator=g).cuda()
    A[m:] = 0.0  # producers zero the pad rows
    W = torch.randn(N, K, generator=g).cuda()
    m_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
    out = torch.full((M, N), 7.0, device="cuda")
    ops.gemm(A, W, out, transB=True, ragged=m_dev, ragged_dim=1, backend=ops.BACKEND_TC)
    ref = A.double().cpu() @ W.double().cpu().T
    G.assert_close("rows<m", out[:m].cpu(), ref[:m], TOL)
    # pad rows are zeroed up to the end of the last touched tile (128 rows, or 256 for the CTA-pair variant)
    assert (out[m:384] == 0).all() and (out[512:] == 7.0).all()
    assert (out[384:512] == 0).all() or (out[384:512] == 7.0).all()
    dy = torch.randn(M, N, generator=g).cuda()
    dy[m:] = 0.0
    dw = torch.empty(N, K, device="cuda")
    ops.gemm(dy, A, dw, transA=True, ragged=m_dev, ragged_dim=2, backend=ops.BACKEND_TC)
    G.assert_close("wgrad ragged", dw.cpu(), dy[:m].double().cpu().T @ A[:m].double().cpu(), TOL)
    # K bound of zero: output is bias/beta only
    z = torch.zeros(1, dtype=torch.int32, device="cuda")
    dw2 = torch.full((N, K), 3.0, device="cuda")
    ops.gemm(dy, A, dw2, transA=True, ragged=z, ragged_dim=2, beta=1.0, backend=ops.BACKEND_TC)
    assert (dw2 == 3.0).all()


@pytest.mark.timeout(120)
def test_auto_backend_uses_tc_for_big_and_ffma_for_skinny():
    from immtsf import ops

    g = torch.Generator().manual_seed(11)
    A = torch.randn(512, 768, generator=g).cuda()
    W = torch.randn(768, 768, generator=g).cuda()
    Wc = torch.randn(4, 768, generator=g).cuda()
    big = ops.linear_fwd(A, W, None)
    skinny = ops.linear_fwd(A, Wc, None)
    G.assert_close("auto big", big.cpu(), A.double().cpu() @ W.double().cpu().T, TOL)
    G.assert_close("auto skinny", skinny.cpu(), A.double().cpu() @ Wc.double().cpu().T, TOL)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("bn", ["128", "256", "512"])
def test_gemm_tc_both_tile_widths(bn):
    """The 128x128 (dual TMEM accumulator), 128x256 (setmaxnreg, 8 epilogue warps) and CTA-pair (cta_group::2,
    256x256 per cluster; IMMTSF_TC_BN=512) variants, each forced."""
    import os
    import subprocess
    import sys

    env = dict(os.environ, IMMTSF_TC_BN=bn)
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "tc_bn_check.py")], env=env,
                       capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert f"OK IMMTSF_TC_BN={bn}" in r.stdout


@pytest.mark.timeout(120)
def test_gemm_group_mixed_transpositions_and_shapes():
    """immtsf_gemm_group: four products of different shapes / transpositions in one launch, with beta and C_lo."""
    from immtsf import ops

    g = torch.Generator().manual_seed(31)
    lo = ops.LoCache()
    specs = [(768, 768, 768, 0, 0), (768, 640, 256, 0, 1), (384, 768, 512, 1, 0), (256, 384, 768, 1, 1)]
    probs, refs, los = [], [], []
    for (M, N, K, tA, tB) in specs:
        A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
        B = (torch.randn((N, K) if tB else (K, N), generator=g) * 0.3).cuda()
        C = torch.randn(M, N, generator=g).cuda()
        C_lo = torch.empty(M, N, device="cuda")
        refs.append(0.5 * _ref(A.cpu(), B.cpu(), tA, tB) + 2.0 * C.double().cpu())
        probs.append(dict(A=A, B=B, C=C, transA=bool(tA), transB=bool(tB), alpha=0.5, beta=2.0, emit_lo=C_lo))
        los.append(C_lo)
    ops.gemm_group(probs, lo)
    torch.cuda.synchronize()
    for pr, ref, C_lo in zip(probs, refs, los):
        G.assert_close("gemm_group", pr["C"].cpu(), ref, TOL)
        hi = (pr["C"].view(torch.int32) & -8192).view(torch.float32)
        assert torch.equal(C_lo, pr["C"] - hi)


@pytest.mark.timeout(120)
def test_gemm_emits_lo_of_its_output():
    """immtsf_gemm_ex C_lo (plain, split-K and ragged launches): exactly C - trunc_tf32(C)."""
    from immtsf import ops

    g = torch.Generator().manual_seed(32)
    for (M, N, K, tA, tB, ragged) in [(6144, 768, 768, 0, 1, None), (768, 768, 6144, 1, 0, None), (640, 256, 128, 0, 1, 300)]:
        lo = ops.LoCache()
        A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
        B = torch.randn((N, K) if tB else (K, N), generator=g).cuda()
        C = torch.empty(M, N, device="cuda")
        rg = None if ragged is None else torch.tensor([ragged], dtype=torch.int32, device="cuda")
        ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), ragged=rg, ragged_dim=1 if rg is not None else 0, backend=ops.BACKEND_TC,
                 lo=lo, emit_lo=True)
        C_lo = lo.lo_for(C, None)
        rows = M if ragged is None else 384
        hi = (C[:rows].view(torch.int32) & -8192).view(torch.float32)
        assert torch.equal(C_lo[:rows], C[:rows] - hi)
