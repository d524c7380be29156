"""CPU model of the tensor-core form of TTF_RecAvg's pooling (imm-tsf_b200/immtsf/ops.py: _recavg_pool_fwd_tc / _bwd_tc around
csrc/recavg_tc.cu) against the oracle (oracle/immtsf_oracle.py: ttf_recavg, fusions/TTF_RecAvg.py:94-102) in float64:

  forward   E_raw[b] = Wn[b] V'pad[b],   Wn = w / clamp_min(sum_n w, 1e-6), zero beyond the sample's notes
  backward  dV'pad[b] = Wn[b]^T dE_raw[b];   dlog_sigma = sum dE_raw * (R - csum * E_raw),  R = Cn V'pad,  Cn = Wn 2 (delta/sigma)^2

i.e. the algebra the GPU path composes from batched products -- checked here with autograd of the reference formula, including
samples without notes, ragged counts and a note count that is not a multiple of 4 (the padded contraction dimension)."""
import pytest
import torch

from oracle import immtsf_oracle as O


def tc_model(V, tau, counts, t_hat, log_sigma, Np):
    """What recavg_weights + csr_to_padded + the three batched products compute (dense torch, float64)."""
    B, N, d = V.shape
    T = t_hat.shape[1]
    n = torch.arange(Np)[None, :]
    valid = (n < counts[:, None]).to(V.dtype)  # [B, Np]
    tau_p = torch.zeros(B, Np, dtype=V.dtype)
    tau_p[:, :N] = tau
    Vpad = torch.zeros(B, Np, d, dtype=V.dtype)
    Vpad[:, :N] = V
    Vpad = Vpad * valid[:, :, None]
    r = (t_hat[:, :, None] - tau_p[:, None, :]).clamp_min(0) * (1.0 / torch.exp(log_sigma))  # [B, T, Np]
    w = torch.exp(-(r * r)) * valid[:, None, :]
    wsum = w.sum(2)
    Wn = w / wsum.clamp_min(1e-6)[:, :, None]
    Cn = Wn * 2.0 * r * r
    csum = Cn.sum(2)
    E_raw = Wn @ Vpad

    def backward(dE_raw):
        dVpad = Wn.transpose(1, 2) @ dE_raw
        R = Cn @ Vpad
        dls = (dE_raw * (R - csum[:, :, None] * E_raw)).sum()
        return dVpad * valid[:, :, None], dls

    return E_raw, wsum, backward


@pytest.mark.parametrize("B,N,T,d,seed", [(4, 7, 5, 12, 0), (3, 70, 40, 8, 1), (2, 1, 1, 4, 2)])
def test_tensor_core_form_equals_the_reference_formula(B, N, T, d, seed):
    g = torch.Generator().manual_seed(seed)
    f64 = torch.float64
    counts = torch.randint(0 if B > 2 else 1, N + 1, (B,), generator=g)
    counts[0] = N
    if B > 2:
        counts[1] = 0  # a sample without notes
    V = torch.randn(B, N, d, generator=g, dtype=f64) * (torch.arange(N)[None, :, None] < counts[:, None, None])
    tau = torch.rand(B, N, generator=g, dtype=f64) * 3.0 * (torch.arange(N)[None, :] < counts[:, None])
    t_hat = torch.rand(B, T, generator=g, dtype=f64).sort(dim=1)[0] * 4.0
    ls = torch.tensor(0.3, dtype=f64, requires_grad=True)
    Vr = V.clone().requires_grad_(True)
    # the oracle's pooling lines (TTF_RecAvg.py:94-102), no projection / LayerNorm
    mask = O.note_mask_from_content(V)
    delta = (t_hat[:, None] - tau[:, :, None]).clamp_min(0)
    w = torch.exp(-((delta / ls.exp()) ** 2)) * mask.to(f64)[:, :, None]
    E_ref = torch.einsum("bnt,bnd->btd", w, Vr) / w.sum(dim=1).clamp_min(1e-6).unsqueeze(-1)
    dE = torch.randn(B, T, d, generator=g, dtype=f64)
    (E_ref * dE).sum().backward()
    Np = (N + 3) // 4 * 4
    E_raw, wsum, bwd = tc_model(V, tau, counts, t_hat, ls.detach(), Np)
    dVpad, dls = bwd(dE)
    assert torch.allclose(E_raw, E_ref.detach(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(wsum, w.detach().sum(dim=1), rtol=1e-12, atol=1e-14)
    assert torch.allclose(dVpad[:, :N], Vr.grad, rtol=1e-11, atol=1e-12)
    assert torch.allclose(dls, ls.grad, rtol=1e-9, atol=1e-12)


def test_selection_thresholds(monkeypatch):
    """ops.recavg_tc_ok: streaming kernels at Time-IMM sizes, tcgen05 form from the measured cross-over (forward before backward)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "imm-tsf_b200"))
    from immtsf import ops

    monkeypatch.delenv("IMMTSF_RECAVG_TC", raising=False)
    monkeypatch.delenv("IMMTSF_GEMM", raising=False)

    class R:  # the two fields the selection reads
        def __init__(self, B, N):
            self.B, self.N = B, N

    Vp = torch.zeros(8, 768)
    pick = lambda N, T, bwd=False: ops.recavg_tc_ok(R(64, N), Vp, T, 768, backward=bwd)
    assert not pick(16, 24) and not pick(64, 64) and not pick(256, 64) and not pick(64, 256)
    assert pick(256, 256) and pick(1024, 64) and pick(1024, 256)
    assert not pick(256, 256, True) and not pick(1024, 64, True) and pick(1024, 256, True)
    monkeypatch.setenv("IMMTSF_RECAVG_TC", "1")
    assert pick(5, 7) and pick(5, 7, True)
    monkeypatch.setenv("IMMTSF_RECAVG_TC", "0")
    assert not pick(1024, 256)
