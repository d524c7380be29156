"""CPU check of the numerical claim behind the (default-off) IMMTSF_RECAVG_MASKBIT variant of csrc/recavg.cu: carrying the
dropout keep flag in the mantissa LSB of the saved E_raw perturbs the LayerNorm-backward result by far less than the
gradient tolerance of the parity tests (5e-5 max-norm relative).  Pure numpy fp32 restatement of the rows phase
(recavg_bwd_fused_kernel / recavg_bwd_rows_*: dy*keep, x^ = (x - mean) * rstd, dgamma, dbeta, dS, d(den))."""
import numpy as np
import pytest


def rows_phase(dy, x, mean, rstd, wsum, gamma, keep_scale, eps=np.float32(1e-5)):
    f = np.float32
    dye = dy * keep_scale
    h = (x - mean[:, None]) * rstd[:, None]
    dgamma = (dye * h).sum(0, dtype=f)
    dbeta = dye.sum(0, dtype=f)
    g = dye * gamma[None, :]
    d = f(x.shape[1])
    m1 = g.sum(1, dtype=f) / d
    s2 = (g * h).sum(1, dtype=f)
    m2 = s2 / d
    den = np.maximum(wsum, f(1e-6))
    dS = (rstd / den)[:, None] * (g - m1[:, None] - h * m2[:, None])
    dden = np.where(wsum >= f(1e-6), -(s2 * eps * rstd * rstd) / den, f(0))
    return dS.astype(f), dgamma, dbeta, dden.astype(f)


@pytest.mark.parametrize("rows,d,p,scale", [(96, 768, 0.1, 1.0), (64, 64, 0.5, 1e-3), (32, 1024, 0.2, 50.0)])
def test_keep_flag_in_lsb_is_far_inside_the_gradient_tolerance(rows, d, p, scale):
    rng = np.random.default_rng(7)
    f = np.float32
    x = (rng.standard_normal((rows, d)) * scale + 0.3 * scale).astype(f)
    dy = rng.standard_normal((rows, d)).astype(f)
    gamma = (1.0 + 0.1 * rng.standard_normal(d)).astype(f)
    wsum = (rng.random(rows) * 8 + 0.5).astype(f)
    mean = x.mean(1, dtype=f)
    rstd = (1.0 / np.sqrt(x.var(1, dtype=f) + f(1e-5))).astype(f)
    keep = rng.random((rows, d)) >= p
    inv_keep = f(1.0 / (1.0 - p))
    ks = np.where(keep, inv_keep, f(0)).astype(f)
    # forward side: tag_keep(x, ks) = (bits & ~1) | keep
    bits = x.view(np.uint32)
    tagged = ((bits & np.uint32(0xFFFFFFFE)) | keep.astype(np.uint32)).view(f)
    assert np.abs(tagged.view(np.int32) - x.view(np.int32)).max() <= 1
    # backward side: keep_of(x) reads the flag back
    ks_back = np.where((tagged.view(np.uint32) & 1) == 1, inv_keep, f(0)).astype(f)
    assert np.array_equal(ks_back, ks)
    ref = rows_phase(dy, x, mean, rstd, wsum, gamma, ks)
    got = rows_phase(dy, tagged, mean, rstd, wsum, gamma, ks_back)
    for name, a, b in zip(("dS", "dgamma", "dbeta", "dden"), ref, got):
        den = max(np.abs(a).max(), 1e-20)
        err = np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / den
        assert err <= 2e-6, (name, err)  # 25x inside the 5e-5 gradient tolerance
