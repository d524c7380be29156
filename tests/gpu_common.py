"""Shared pieces of the GPU parity tests (test infrastructure)."""
import numpy as np
import torch

import philox_ref
from oracle import immtsf_oracle as O

SITE_TTF_DROPOUT, SITE_TTF_ATTN, SITE_MMF_DROPOUT, SITE_MMF_ATTN = 1, 2, 3, 4


from immtsf.synth import build_model, make_args, randomise_, synth_batch  # noqa: E402,F401  (product-side generators)


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
