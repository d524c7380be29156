"""Shared pieces of the GPU parity tests (test infrastructure)."""
from types import SimpleNamespace

import numpy as np
import torch

import philox_ref
from oracle import immtsf_oracle as O

SITE_TTF_DROPOUT, SITE_TTF_ATTN, SITE_MMF_DROPOUT, SITE_MMF_ATTN = 1, 2, 3, 4


def make_args(ttf, mmf, alias, d_txt, C, H, kappa, dropout):
    return SimpleNamespace(TTF_module=ttf, MMF_module=mmf, llm_model_fusion=alias, llm_layers_fusion=1, max_length=1024,
                           device="cuda", use_text_embeddings=True, recency_sigma=1.0, dropout=dropout, d_txt=d_txt,
                           n_heads_fusion=H, C=C, kappa=kappa)


def build_model(cfg, d_model, params=None, dropout=0.0, seed=0):
    import fusions.load_llm as L
    from fusions.FusionModel import FusionModel

    alias = f"SYN{d_model}"
    L.register_d_model(alias, d_model)
    torch.manual_seed(seed)
    fm = FusionModel(make_args(cfg["ttf"], cfg["mmf"], alias, cfg["d_txt"], cfg["C"], cfg["H"], cfg["kappa"], dropout))
    if params is not None:
        fm.load_state_dict(params, strict=True)
    return fm.cuda()


def synth_batch(B, N, T, d_model, C, seed, history=7.0, pred=7.0, no_note=False, t1d=False, full=False):
    """Time-IMM-shaped synthetic batch: ragged N_i ~ U{1..N}, one sample full, zero tail
    padding, tau in [0,history) unsorted, t_hat in [h/(h+p),1) sorted then zero padded."""
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, N + 1, (B,), generator=g)
    if full:
        counts[:] = N
    counts[0] = N
    if B > 1 and not full:
        counts[1] = 1
    if no_note:
        counts[B - 1] = 0
    notes = torch.zeros(B, N, d_model)
    tau = torch.zeros(B, N)
    for b in range(B):
        n = int(counts[b])
        notes[b, :n] = torch.randn(n, d_model, generator=g)
        tau[b, :n] = torch.rand(n, generator=g) * history
    lo = history / (history + pred)
    if t1d:
        t_hat = torch.sort(lo + torch.rand(T, generator=g) * (1 - lo))[0]
    else:
        t_hat = torch.zeros(B, T)
        for b in range(B):
            tl = T if b == 0 else int(torch.randint((T + 2) // 3, T + 1, (1,), generator=g))
            t_hat[b, :tl] = torch.sort(lo + torch.rand(tl, generator=g) * (1 - lo))[0]
    Y = torch.randn(B, T, C, generator=g)
    G = torch.randn(B, T, C, generator=g)
    return notes, tau, t_hat, Y, G


def randomise_(fm, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fm.named_parameters():
            if name.endswith("log_recency_sigma"):
                p.copy_(torch.tensor(-1.2))
            elif p.dim() >= 2:
                p.add_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
            else:
                p.add_((torch.randn(p.shape, generator=g) * 0.1).to(p.device))


def oracle_masks(cfg, notes, T, C, d, p, seed):
    """Rebuild on the host the dropout masks the kernels will draw for `seed`."""
    B, N, _ = notes.shape
    H = cfg["H"]
    m = {}
    if p <= 0:
        return m
    m["ttf.dropout"] = torch.from_numpy(
        philox_ref.keep_mask(seed, SITE_TTF_DROPOUT, np.arange(B * T * d, dtype=np.uint64), p).reshape(B, T, d))
    if cfg["ttf"].startswith("TTF_T2V_XAttn"):
        mask = O.note_mask_from_content(notes)
        rank = (torch.cumsum(mask.to(torch.int64), dim=1) - 1).clamp_min(0).numpy().astype(np.uint64)  # [B,N]
        b_ = np.arange(B, dtype=np.uint64)[:, None, None, None]
        t_ = np.arange(T, dtype=np.uint64)[None, :, None, None]
        h_ = np.arange(H, dtype=np.uint64)[None, None, :, None]
        idx = ((b_ * np.uint64(T) + t_) * np.uint64(H) + h_) * np.uint64(N) + rank[:, None, None, :]
        m["ttf.attn_dropout"] = torch.from_numpy(philox_ref.keep_mask(seed, SITE_TTF_ATTN, idx, p))
    m["mmf.dropout"] = torch.from_numpy(
        philox_ref.keep_mask(seed, SITE_MMF_DROPOUT, np.arange(B * T * C, dtype=np.uint64), p).reshape(B, T, C))
    if cfg["mmf"] == "MMF_XAttn_Add":
        m["mmf.attn_dropout"] = torch.from_numpy(
            philox_ref.keep_mask(seed, SITE_MMF_ATTN, np.arange(B * H * T * T, dtype=np.uint64), p).reshape(B, H, T, T))
    return m


def oracle_run(cfg, params, notes, tau, t_hat, Y, G, dtype=torch.float64, p=0.0, masks=None, grads=True):
    p = philox_ref.realised_p(p)  # the kernels quantise the drop rate to 16 bits and rescale by the realised rate
    P = {k: v.detach().cpu().to(dtype).clone().requires_grad_(grads) for k, v in params.items()}
    Yr = Y.to(dtype).clone().requires_grad_(grads)
    mk = None if masks is None else {k: v.to(dtype) for k, v in masks.items()}
    Yo, E, M = O.fusion_forward(P, cfg["ttf"], cfg["mmf"], notes.to(dtype), tau.to(dtype), t_hat.to(dtype), Yr,
                                n_heads=cfg["H"], kappa=cfg["kappa"], p=p, masks=mk, faithful_expand=False,
                                return_intermediate=True)
    out = {"Y_out": Yo.detach(), "E_txt": E.detach(), "M_txt": M}
    if grads:
        (Yo * G.to(dtype)).sum().backward()
        out["dY"] = Yr.grad
        out["grads"] = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    return out


def gpu_run(fm, notes, tau, t_hat, Y, G, train, grads=True):
    fm.train(train)
    fm.zero_grad(set_to_none=True)
    Yc = Y.cuda().clone().requires_grad_(grads)
    with torch.set_grad_enabled(grads):
        Yo = fm(notes.cuda(), tau.cuda(), t_hat.cuda(), Yc)
    out = {"Y_out": Yo.detach().cpu()}
    if grads:
        (Yo * G.cuda()).sum().backward()
        out["dY"] = Yc.grad.cpu()
        out["grads"] = {k: (p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu())
                        for k, p in fm.named_parameters()}
    torch.cuda.synchronize()
    return out


def assert_close(name, got, ref, rtol, floor=0.0):
    """max-norm relative: ||got-ref||_inf <= rtol * max(||ref||_inf, floor)."""
    got = got.double()
    ref = torch.as_tensor(ref).double()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    err = (got - ref).abs().max().item() if got.numel() else 0.0
    den = max(ref.abs().max().item() if ref.numel() else 0.0, floor)
    assert err <= rtol * den + 1e-30, f"{name}: err {err:.3e} > {rtol:.1e} * {den:.3e}"
    return err / den if den > 0 else err
