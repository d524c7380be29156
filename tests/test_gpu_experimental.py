"""GPU checks of kernel variants that are compiled in but OFF by default because they have not been measured on a B200
yet (written when the round's GPU budget was spent).  They run only with IMMTSF_EXPERIMENTAL=1, so an unmeasured variant
can never turn the default suite red:

    IMMTSF_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q

Each test A/Bs a variant against the default path of the same library inside one process (the switches are read per call).
"""
import os

import pytest
import torch

import gpu_common as G

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("IMMTSF_EXPERIMENTAL") != "1", reason="set IMMTSF_EXPERIMENTAL=1")]


def _recavg_step(B, N, T, d, p, seed=31):
    from immtsf import ops

    notes, tau, t_hat, _, _ = G.synth_batch(B, N, T, d, 1, seed, no_note=B > 2)
    r = ops.csr_build(notes.cuda(), tau.cuda())
    t_hat = t_hat.cuda()
    g = torch.Generator().manual_seed(3)
    ls = torch.tensor(-0.3, device="cuda")
    gamma = (1.0 + 0.1 * torch.randn(d, generator=g)).cuda()
    beta = (0.1 * torch.randn(d, generator=g)).cuda()
    thr, sd = ops.drop_thr(p), 991
    dE = torch.randn(B, T, d, generator=g).cuda()

    def run():
        E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(r.emb_flat, r, t_hat, ls, gamma, beta, T, d, thr, sd, True)
        grads = ops.recavg_pool_bwd(dE.view_as(E_drop), E_raw, mean, rstd, wsum, r.emb_flat, r, t_hat, ls, gamma, T, d, thr, sd)
        torch.cuda.synchronize()
        return E_drop.clone(), E_raw.clone(), [x.clone() for x in grads]

    return run


@pytest.mark.parametrize("fused", ["8", "0"])
@pytest.mark.parametrize("B,N,T,d,p", [(5, 6, 7, 64, 0.1), (64, 16, 24, 768, 0.1), (300, 3, 16, 1024, 0.2), (9, 30, 40, 256, 0.3),
                                         (3, 1, 1, 8, 0.5), (17, 16, 24, 768, 0.0)])
def test_recavg_keep_flags_in_e_raw_lsb(B, N, T, d, p, fused, monkeypatch):
    """IMMTSF_RECAVG_MASKBIT=1: the forward stores every element's dropout keep flag in the mantissa LSB of the saved
    E_raw and the backward reads it back instead of regenerating the Philox mask.  E_drop must be bit-identical, E_raw within
    one ulp, every gradient within 2e-5 of the default path (same mask, x perturbed by <= 1 ulp); with both backward
    variants (one launch / two kernels; T 40 exercises the two-kernel path in both)."""
    monkeypatch.setenv("IMMTSF_RECAVG_FUSED_BWD", fused)
    run = _recavg_step(B, N, T, d, p)
    monkeypatch.setenv("IMMTSF_RECAVG_MASKBIT", "0")
    E0, X0, g0 = run()
    monkeypatch.setenv("IMMTSF_RECAVG_MASKBIT", "1")
    E1, X1, g1 = run()
    assert torch.equal(E0, E1)
    ulp = (X0.view(torch.int32) - X1.view(torch.int32)).abs().max().item()
    assert ulp <= 1, ulp
    if p > 0:
        kept = (X1.view(torch.int32) & 1).bool()
        # a kept element of E_drop is zero only if LayerNorm's output is exactly zero there
        assert torch.equal(kept | (E1 == 0), torch.ones_like(kept)) and (kept & (E1 != 0)).sum() == (E1 != 0).sum()
    for name, ref, got in zip(("dVp", "dgamma", "dbeta", "dlog_sigma"), g0, g1):
        assert torch.isfinite(got).all(), name
        den = max(ref.abs().max().item(), 1e-6)
        err = (got - ref).abs().max().item()
        assert err <= 2e-5 * den, f"{name}: err {err:.3e} vs max {den:.3e}"


@pytest.mark.parametrize("maskbit", ["0", "1"])
@pytest.mark.parametrize("history", [7.0, 0.4, 1.2])
@pytest.mark.parametrize("B,N,T,d,p", [(64, 16, 24, 768, 0.1), (5, 6, 7, 64, 0.1), (9, 30, 24, 256, 0.0), (3, 1, 1, 8, 0.5)])
def test_recavg_fused_bwd_skips_zero_sensitivity_passes(B, N, T, d, p, history, maskbit, monkeypatch):
    """IMMTSF_RECAVG_SKIPQ=1: half passes (4 notes) of the one-launch backward whose c_nt = dw_nt/dlog_sigma are all exactly
    zero (tau_n >= every t_hat_t) skip the Q_n accumulators.  Only exact zeros are skipped, so dV' is bit-identical and
    dlog_sigma differs at most in summation order.  history 7: most notes are newer than the window (c = 0); 0.4 and 1.2:
    most or many notes are older than some query time (little or nothing skipped)."""
    from immtsf import ops

    monkeypatch.setenv("IMMTSF_RECAVG_MASKBIT", maskbit)
    outs = {}
    for notes_per_pass in ("8", "4"):
        monkeypatch.setenv("IMMTSF_RECAVG_FUSED_BWD", notes_per_pass)
        notes, tau, t_hat, _, _ = G.synth_batch(B, N, T, d, 1, 77, history=history, pred=1.0, no_note=B > 2)
        r = ops.csr_build(notes.cuda(), tau.cuda())
        t_hat = t_hat.cuda()
        g = torch.Generator().manual_seed(5)
        ls = torch.tensor(0.2, device="cuda")
        gamma = (1.0 + 0.1 * torch.randn(d, generator=g)).cuda()
        beta = torch.zeros(d, device="cuda")
        thr, sd = ops.drop_thr(p), 17
        E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(r.emb_flat, r, t_hat, ls, gamma, beta, T, d, thr, sd, True)
        dE = torch.randn(B, T, d, generator=g).cuda().view_as(E_drop)
        for skip in ("0", "1"):
            monkeypatch.setenv("IMMTSF_RECAVG_SKIPQ", skip)
            outs[skip] = [x.clone() for x in ops.recavg_pool_bwd(dE, E_raw, mean, rstd, wsum, r.emb_flat, r, t_hat, ls, gamma, T, d, thr, sd)]
        torch.cuda.synchronize()
        assert torch.equal(outs["0"][0], outs["1"][0]), (notes_per_pass, "dVp")
        for name, ref, got in zip(("dgamma", "dbeta"), outs["0"][1:3], outs["1"][1:3]):  # float atomics across CTAs: order varies
            assert (ref - got).abs().max().item() <= 1e-5 * max(ref.abs().max().item(), 1e-6), (notes_per_pass, name)
        ref, got = outs["0"][3], outs["1"][3]
        assert abs(float(ref) - float(got)) <= 1e-6 * max(abs(float(ref)), 1e-6), (notes_per_pass, float(ref), float(got))


@pytest.mark.parametrize("maskbit", ["0", "1"])
@pytest.mark.parametrize("B,N,T,d,p", [(64, 16, 24, 768, 0.1), (5, 6, 7, 64, 0.1), (300, 3, 16, 1024, 0.2), (150, 16, 32, 512, 0.1),
                                         (33, 12, 9, 256, 0.3), (700, 16, 24, 768, 0.0), (3, 1, 1, 8, 0.5)])
def test_recavg_persistent_forward_is_bit_identical(B, N, T, d, p, maskbit, monkeypatch):
    """IMMTSF_RECAVG_FWD_PERSIST=1: resident CTAs walk the (sample, query tile) items with a two-stage ring of bulk copies.
    Same lane ownership and summation order as the staged kernel, so every output is bit-identical (B 700 x 1 tile and
    B 150 x 2 tiles exceed the 296 resident CTAs: several items per CTA, both stages and both mbarrier parities in use;
    d 1024 is outside the variant's range and must fall back)."""
    from immtsf import ops

    monkeypatch.setenv("IMMTSF_RECAVG_MASKBIT", maskbit)
    notes, tau, t_hat, _, _ = G.synth_batch(B, N, T, d, 1, 123, no_note=B > 2)
    r = ops.csr_build(notes.cuda(), tau.cuda())
    t_hat = t_hat.cuda()
    g = torch.Generator().manual_seed(9)
    ls = torch.tensor(0.1, device="cuda")
    gamma = (1.0 + 0.1 * torch.randn(d, generator=g)).cuda()
    beta = (0.1 * torch.randn(d, generator=g)).cuda()
    thr, sd = ops.drop_thr(p), 4242
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("IMMTSF_RECAVG_FWD_PERSIST", mode)
        for save in (True, False):
            o = ops.recavg_pool_fwd(r.emb_flat, r, t_hat, ls, gamma, beta, T, d, thr, sd, save)
            torch.cuda.synchronize()
            outs[(mode, save)] = [None if x is None else x.clone() for x in o]
    for save in (True, False):
        for name, ref, got in zip(("E_drop", "E_raw", "mean", "rstd", "wsum"), outs[("0", save)], outs[("1", save)]):
            assert (ref is None) == (got is None), (save, name)
            if ref is not None:
                assert torch.equal(ref, got), (save, name, (ref - got).abs().max().item())


@pytest.mark.parametrize("maskbit,skipq", [("0", "0"), ("1", "1")])
@pytest.mark.parametrize("B,N,T,d,p", [(5, 6, 7, 64, 0.1), (64, 16, 24, 768, 0.1), (700, 16, 24, 768, 0.1), (300, 3, 16, 1024, 0.2),
                                         (9, 30, 24, 256, 0.0), (3, 1, 1, 8, 0.5), (450, 16, 32, 512, 0.1), (2, 5, 24, 776, 0.1)])
def test_recavg_bwd_pipe_equals_default(B, N, T, d, p, maskbit, skipq, monkeypatch):
    """IMMTSF_RECAVG_BWD_PIPE=1: warp-specialised one-launch backward (8 row warps + 2*NC note warps, two-stage dS ring with
    full / empty mbarriers, 1 CTA per SM) against the default backward on the same inputs.  B 700 and B 450 give every CTA
    several samples (both ring stages, both parities of every mbarrier); T 7 leaves a row warp idle; one sample has no notes."""
    monkeypatch.setenv("IMMTSF_RECAVG_MASKBIT", maskbit)
    monkeypatch.setenv("IMMTSF_RECAVG_SKIPQ", skipq)
    from immtsf import ops

    notes, tau, t_hat, _, _ = G.synth_batch(B, N, T, d, 1, 55, no_note=B > 2)
    r = ops.csr_build(notes.cuda(), tau.cuda())
    t_hat = t_hat.cuda()
    g = torch.Generator().manual_seed(13)
    ls = torch.tensor(-0.2, device="cuda")
    gamma = (1.0 + 0.1 * torch.randn(d, generator=g)).cuda()
    beta = (0.1 * torch.randn(d, generator=g)).cuda()
    thr, sd = ops.drop_thr(p), 808
    E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(r.emb_flat, r, t_hat, ls, gamma, beta, T, d, thr, sd, True)
    dE = torch.randn(B, T, d, generator=g).cuda().view_as(E_drop)
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("IMMTSF_RECAVG_BWD_PIPE", mode)
        outs[mode] = [x.clone() for x in ops.recavg_pool_bwd(dE, E_raw, mean, rstd, wsum, r.emb_flat, r, t_hat, ls, gamma, T, d, thr, sd)]
        torch.cuda.synchronize()
    for name, ref, got in zip(("dVp", "dgamma", "dbeta", "dlog_sigma"), outs["0"], outs["1"]):
        assert torch.isfinite(got).all(), name
        den = max(ref.abs().max().item(), 1e-6)
        err = (got - ref).abs().max().item()
        assert err <= 2e-5 * den, f"{name}: err {err:.3e} vs max {den:.3e}"
