"""CPU checks of two numerical claims behind csrc/recavg.cu's one-launch backward (recavg_bwd_mma_kernel), in numpy fp32
against a float64 evaluation of the reference formula (autograd of fusions/TTF_RecAvg.py:100-106):

  * rows phase: the re-associated LayerNorm backward  dS = sc*g + kx*x + k0  (two FMAs per element, x^ = x*rstd - mean*rstd)
    equals  rstd/den * (g - mean(g) - x^ * mean(g x^))  far inside the 5e-5 gradient tolerance, also for rows whose mean is
    large against their spread;
  * note phase: the 3xTF32 split of the tensor-core product (lo*hi + hi*lo + hi*hi with hi = the operand truncated to 10
    mantissa bits, lo = the exact remainder, fp32 accumulation) reproduces the fp32 contraction  dV'_n = sum_t w_nt dS_t.
"""
import numpy as np
import pytest

F = np.float32


def rows_phase_reference(dy, x, mean, rstd, wsum, gamma, ks):
    dy, x, mean, rstd, wsum, gamma, ks = (a.astype(np.float64) for a in (dy, x, mean, rstd, wsum, gamma, ks))
    dye = dy * ks
    h = (x - mean[:, None]) * rstd[:, None]
    g = dye * gamma[None, :]
    d = x.shape[1]
    m1 = g.sum(1) / d
    m2 = (g * h).sum(1) / d
    den = np.maximum(wsum, 1e-6)
    return (rstd / den)[:, None] * (g - m1[:, None] - h * m2[:, None]), (dye * h).sum(0), dye.sum(0)


def rows_phase_kernel(dy, x, mean, rstd, wsum, gamma, ks):
    """The kernel's arithmetic, every operation rounded to fp32 (fused multiply-adds evaluated in float64 and rounded once)."""
    fma = lambda a, b, c: (a.astype(np.float64) * np.float64(b) + np.float64(c)).astype(F)
    rs, nmr = rstd[:, None], (-mean * rstd).astype(F)[:, None]
    dye = (dy * ks).astype(F)
    h = (x.astype(np.float64) * rs + nmr).astype(F)
    g = (dye * gamma[None, :]).astype(F)
    d = F(x.shape[1])
    m1 = (g.sum(1, dtype=F) / d).astype(F)
    s2 = (g.astype(np.float64) * h).sum(1).astype(F)
    m2 = (s2 / d).astype(F)
    den = np.maximum(wsum, F(1e-6))
    sc = (rstd / den).astype(F)
    kx = (-sc * m2 * rstd).astype(F)
    k0 = (-sc * (m1 + m2 * nmr[:, 0])).astype(F)
    inner = (x.astype(np.float64) * kx[:, None] + k0[:, None]).astype(F)
    dS = (g.astype(np.float64) * sc[:, None] + inner).astype(F)
    return dS, (dye.astype(np.float64) * h).sum(0).astype(F), dye.sum(0, dtype=F)


@pytest.mark.parametrize("rows,d,p,scale,shift", [(96, 768, 0.1, 1.0, 0.3), (64, 64, 0.5, 1e-3, 1e-3), (32, 1024, 0.2, 50.0, 10.0),
                                                  (48, 768, 0.1, 1.0, 20.0)])
def test_two_fma_form_of_the_layernorm_backward(rows, d, p, scale, shift):
    rng = np.random.default_rng(7)
    x = (rng.standard_normal((rows, d)) * scale + shift).astype(F)
    dy = rng.standard_normal((rows, d)).astype(F)
    gamma = (1.0 + 0.1 * rng.standard_normal(d)).astype(F)
    wsum = (rng.random(rows) * 8 + 0.5).astype(F)
    mean = x.mean(1, dtype=np.float64).astype(F)
    rstd = (1.0 / np.sqrt(x.astype(np.float64).var(1) + 1e-5)).astype(F)
    ks = np.where(rng.random((rows, d)) >= p, F(1.0 / (1.0 - p)), F(0)).astype(F)
    ref = rows_phase_reference(dy, x, mean, rstd, wsum, gamma, ks)
    got = rows_phase_kernel(dy, x, mean, rstd, wsum, gamma, ks)
    for name, a, b in zip(("dS", "dgamma", "dbeta"), ref, got):
        err = np.abs(a - b.astype(np.float64)).max() / max(np.abs(a).max(), 1e-30)
        # |mean| / spread = 20 costs a factor ~20 in absolute error of x^ (one rounding of mean*rstd): still 5x inside 5e-5
        assert err <= (1e-5 if shift / scale > 5 else 2e-6), (name, err)


def test_3xtf32_split_reproduces_the_fp32_contraction():
    rng = np.random.default_rng(3)
    T, N, d = 24, 16, 256
    w = np.exp(-rng.random((N, T)) * 6).astype(F)          # recency weights in (0, 1]
    dS = (rng.standard_normal((T, d)) * 0.05).astype(F)
    trunc = lambda a: (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(F)
    w_hi, dS_hi = trunc(w), trunc(dS)
    w_lo, dS_lo = (w - w_hi).astype(F), (dS - dS_hi).astype(F)
    assert np.array_equal((w_hi.astype(np.float64) + w_lo), w.astype(np.float64))  # the split is exact
    # the tensor core also truncates the lo operands to TF32; products are exact in fp32, accumulation in fp32
    acc = (trunc(w_lo).astype(np.float64) @ dS_hi + w_hi.astype(np.float64) @ trunc(dS_lo) + w_hi.astype(np.float64) @ dS_hi).astype(F)
    ref = w.astype(np.float64) @ dS.astype(np.float64)
    err = np.abs(acc - ref).max() / np.abs(ref).max()
    single = np.abs((w_hi.astype(np.float64) @ dS_hi) - ref).max() / np.abs(ref).max()
    assert err <= 1e-6, err          # fp32-exact for the purposes of the 2e-5 / 5e-5 test bars
    assert single >= 1e-5, single    # a single TF32 pass would not be
