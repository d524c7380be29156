"""RecAvg pooling kernels in isolation vs an fp64 torch restatement of the same lines (TTF_RecAvg.py:94-106),
given the same upstream gradient: isolates the kernels' own error in every output (not a test; a diagnostic)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from immtsf import ops
import gpu_common as G, philox_ref

for (B, N, T, d, p) in [(32, 16, 24, 768, 0.0), (32, 16, 24, 768, 0.1), (8, 40, 24, 768, 0.1), (32, 16, 24, 4096, 0.1)]:
    notes, tau, t_hat, _, _ = G.synth_batch(B, N, T, d, 4, 31)
    r = ops.csr_build(notes.cuda(), tau.cuda())
    g = torch.Generator().manual_seed(3)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).cuda(); beta = (0.1 * torch.randn(d, generator=g)).cuda()
    ls = torch.tensor(-1.2).cuda()
    seed, thr = 12345, ops.drop_thr(p)
    E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(r.emb_flat, r, t_hat.cuda(), ls, gamma, beta, T, d, thr, seed, True)
    dE = torch.randn(B, T, d, generator=g).cuda()
    dVp, dgam, dbet, dls = ops.recavg_pool_bwd(dE, E_raw, mean, rstd, wsum, r.emb_flat, r, t_hat.cuda(), ls, gamma, T, d, thr, seed)
    torch.cuda.synchronize()
    # fp64 reference
    V = notes.double().requires_grad_(True)
    mask = (notes.abs().sum(2) > 0).double()
    lsr = torch.tensor(-1.2, dtype=torch.float64, requires_grad=True)
    gr, br = gamma.double().cpu().requires_grad_(True), beta.double().cpu().requires_grad_(True)
    delta = (t_hat.double()[:, None] - tau.double()[:, :, None]).clamp_min(0)
    w = torch.exp(-((delta / lsr.exp()) ** 2)) * mask[:, :, None]
    Er = torch.einsum("bnt,bnd->btd", w, V) / w.sum(1).clamp_min(1e-6).unsqueeze(-1)
    mu = Er.mean(-1, keepdim=True); var = ((Er - mu) ** 2).mean(-1, keepdim=True)
    En = (Er - mu) / torch.sqrt(var + 1e-5) * gr + br
    pe = philox_ref.realised_p(p)
    keep = torch.from_numpy(philox_ref.keep_mask(seed, 1, np.arange(B * T * d, dtype=np.uint64), p).reshape(B, T, d)).double() if p > 0 else torch.ones(B, T, d, dtype=torch.float64)
    Ed = En * keep / (1 - pe)
    (Ed * dE.double().cpu()).sum().backward()
    rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
    dV_ref = V.grad[mask.bool()]
    sumN = int(mask.sum())
    print(f"B{B} N{N} T{T} d{d} p{p}: E_drop {rel(E_drop, Ed.detach()):.2e} dV' {rel(dVp[:sumN], dV_ref):.2e} dgamma {rel(dgam, gr.grad):.2e} "
          f"dbeta {rel(dbet, br.grad):.2e} dls kernel {dls.item():.8f} ref {lsr.grad.item():.8f} abs err {abs(dls.item() - lsr.grad.item()):.2e}")
