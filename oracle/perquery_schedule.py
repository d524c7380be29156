"""Schedule model of the B200 per-(note, query) Time2Vec attention (test infrastructure, CPU, torch).

The reference (fusions/TTF_T2V_XAttn_old.py:119-143) materialises [V ; Time2Vec(lag)] for every (sample, query, note),
runs KV_proj on B*T*N rows and lets nn.MultiheadAttention project K and V again.  The CUDA path never forms a per-pair
vector of width d.  With  X_nt = A_n + W_phi phi_nt,  A_n = W_a V'_n + b_kv  (W_kv = [W_a | W_phi]):

    score_{h,n,t} = q_h . (W_k[h] X_nt + b_k[h]) = u_h . A_n + g_h . phi_nt + const_h      u_h = W_k[h]^T q_h, g_h = W_phi^T u_h
    o_{h,t}       = sum_n P~ (W_v[h] X_nt + b_v[h]) = W_v[h] (Z_{h,t} + W_phi Phi_{h,t}) + sp_{h,t} b_v[h]
    Z = sum_n P~ A_n   Phi = sum_n P~ phi_nt   sp = sum_n P~

so the fused kernel (csrc/t2v_perquery.cu) needs only A [sumN, d], the per-note score base a = A U^T [sumN, H] and
g [H, d_tau]; it evaluates sin() on the fly, does the softmax over each ragged segment and returns Z, Phi, sp.

This file states (1) the CONTRACT of the two kernels in plain torch (`pool_fwd`, `pool_bwd`, hand-written backward --
the formulas the CUDA kernel implements) and (2) the host composition around them (`forward`, `backward`), step for
step what immtsf/functional.py: T2VPerQueryFn does with GEMM calls.  tests/test_perquery_cpu.py checks both against
the oracle (autograd over the reference restatement) and the golden vectors; the GPU tests check the kernels against
(1) in isolation.  const_h (= q_h . b_k[h]) is constant over n and cancels in the softmax: the gradient of the key bias
is exactly zero, which is what autograd returns up to rounding.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

Tensor = torch.Tensor


def t2v(delta: Tensor, w_lin, b_lin, w_per, b_per) -> Tensor:
    """delta [...] -> [..., d_tau]: [w0*x + b0 ; sin(w_k x + b_k)]."""
    lin = delta.unsqueeze(-1) * w_lin.reshape(1) + b_lin.reshape(1)
    per = torch.sin(delta.unsqueeze(-1) * w_per.reshape(-1) + b_per.reshape(-1))
    return torch.cat([lin, per], dim=-1)


# ----------------------------------------------------------------------------- kernel contracts (padded layout)
def pool_fwd(A, a_sc, g, tau, t_hat, mask, w_lin, b_lin, w_per, b_per, keep_scale=None):
    """A [B,N,d], a_sc [B,N,H], g [H,dt], tau [B,N], t_hat [B,T], mask [B,N] bool, keep_scale [B,T,H,N] (0 or 1/(1-p)).
    Returns Z [B,T,H,d], Phi [B,T,H,dt], sp [B,T,H], P [B,T,H,N] (softmax, 0 on masked notes / no-note samples)."""
    delta = (t_hat[:, :, None] - tau[:, None, :]).clamp_min(0)  # [B,T,N]
    phi = t2v(delta, w_lin, b_lin, w_per, b_per)  # [B,T,N,dt]
    s = a_sc.permute(0, 2, 1)[:, None] + torch.einsum("btnk,hk->bthn", phi, g)
    s = s.masked_fill(~mask[:, None, None, :], float("-inf"))
    mx = s.max(dim=-1, keepdim=True).values
    mx = torch.where(torch.isfinite(mx), mx, torch.zeros_like(mx))
    e = torch.exp(s - mx)
    den = e.sum(-1, keepdim=True)
    P = torch.where(den > 0, e / den.clamp_min(1e-300), torch.zeros_like(e))
    Pt = P if keep_scale is None else P * keep_scale
    Z = torch.einsum("bthn,bnd->bthd", Pt, A)
    Phi = torch.einsum("bthn,btnk->bthk", Pt, phi)
    return Z, Phi, Pt.sum(-1), P


def pool_bwd(dZ, dPhi, dsp, A, g, P, tau, t_hat, w_lin, b_lin, w_per, b_per, keep_scale=None):
    """Hand-written backward of pool_fwd.  Returns dA [B,N,d], da [B,N,H], dw [dt], db [dt] (k = 0: the linear unit,
    k >= 1: periodic unit k-1), dg [H,dt]."""
    delta = (t_hat[:, :, None] - tau[:, None, :]).clamp_min(0)
    arg_per = delta.unsqueeze(-1) * w_per.reshape(-1) + b_per.reshape(-1)
    phi = torch.cat([delta.unsqueeze(-1) * w_lin.reshape(1) + b_lin.reshape(1), torch.sin(arg_per)], dim=-1)
    dpre_scale = torch.cat([torch.ones_like(delta).unsqueeze(-1), torch.cos(arg_per)], dim=-1)  # d phi / d arg
    ks = torch.ones_like(P) if keep_scale is None else keep_scale
    Pt = P * ks
    dPt = torch.einsum("bthd,bnd->bthn", dZ, A) + torch.einsum("bthk,btnk->bthn", dPhi, phi) + dsp.unsqueeze(-1)
    dP = dPt * ks
    D = (P * dP).sum(-1, keepdim=True)
    dS = P * (dP - D)
    da = dS.sum(dim=1).permute(0, 2, 1)  # [B,N,H]
    dg = torch.einsum("bthn,btnk->hk", dS, phi)
    dphi = torch.einsum("bthn,bthk->btnk", Pt, dPhi) + torch.einsum("bthn,hk->btnk", dS, g)
    dpre = dphi * dpre_scale
    dw = (dpre * delta.unsqueeze(-1)).sum(dim=(0, 1, 2))
    db = dpre.sum(dim=(0, 1, 2))
    dA = torch.einsum("bthn,bthd->bnd", Pt, dZ)
    return dA, da, dw, db, dg


# ----------------------------------------------------------------------------- host composition
def _ln_fwd(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + eps)
    return (x - mu) * rstd * w + b, mu, rstd


def _ln_bwd(dy, x, w, mu, rstd):
    xh = (x - mu) * rstd
    dg, db = (dy * xh).reshape(-1, x.shape[-1]).sum(0), dy.reshape(-1, x.shape[-1]).sum(0)
    dxh = dy * w
    dx = rstd * (dxh - dxh.mean(-1, keepdim=True) - xh * (dxh * xh).mean(-1, keepdim=True))
    return dx, dg, db


def forward(P: Dict[str, Tensor], notes, tau, t_hat, H: int, p: float = 0.0, masks: Optional[Dict[str, Tensor]] = None,
            prefix: str = "ttf."):
    """Returns (E_txt [B,T,d], M_txt [B,1], ctx)."""
    g_ = lambda k: P[prefix + k]
    masks = masks or {}
    mask = notes.abs().sum(2) > 0
    M_txt = mask.any(1, keepdim=True)
    B, N, _ = notes.shape
    if t_hat.dim() == 1:
        t_hat = t_hat.unsqueeze(0).repeat(B, 1)
    T = t_hat.shape[1]
    has_in = prefix + "input_proj.weight" in P
    Vp = notes @ g_("input_proj.weight").T + g_("input_proj.bias") if has_in else notes
    d = Vp.shape[-1]
    hd = d // H
    W_kv, b_kv = g_("KV_proj.weight"), g_("KV_proj.bias")
    W_a, W_phi = W_kv[:, :d], W_kv[:, d:]
    in_w, in_b = g_("attn.in_proj_weight"), g_("attn.in_proj_bias")
    W_q, W_k, W_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:]
    b_q, b_v = in_b[:d], in_b[2 * d:]
    Qp = g_("Q_param").reshape(d)
    A = (Vp @ W_a.T + b_kv) * mask.unsqueeze(-1)  # rows of masked notes are never read
    scale = math.sqrt(1.0 / hd)
    q = (Qp @ W_q.T + b_q) * scale
    Qblk = torch.zeros(H, d, dtype=q.dtype)
    for h in range(H):
        Qblk[h, h * hd:(h + 1) * hd] = q[h * hd:(h + 1) * hd]
    U = Qblk @ W_k  # [H,d]
    a_sc = A @ U.T  # [B,N,H]
    g = U @ W_phi  # [H,dt]
    ks = None
    if p > 0 and masks.get(prefix + "attn_dropout") is not None:
        ks = masks[prefix + "attn_dropout"].to(A.dtype) / (1.0 - p)  # [B,T,H,N]
    tw = (g_("time2vec.linear.weight"), g_("time2vec.linear.bias"), g_("time2vec.periodic.weight"), g_("time2vec.periodic.bias"))
    Z, Phi, sp, Pm = pool_fwd(A, a_sc, g, tau, t_hat, mask, *tw, keep_scale=ks)
    XZ = Z + Phi @ W_phi.T  # [B,T,H,d]
    O = torch.empty(B, T, d, dtype=A.dtype)
    for h in range(H):
        hs = slice(h * hd, (h + 1) * hd)
        O[:, :, hs] = XZ[:, :, h] @ W_v[hs].T + sp[:, :, h, None] * b_v[hs]
    attn_out = O @ g_("attn.out_proj.weight").T + g_("attn.out_proj.bias")
    valid = M_txt.view(B, 1, 1).to(A.dtype)
    z = attn_out * valid + Qp
    y0, mu, rstd = _ln_fwd(z, g_("layer_norm.weight"), g_("layer_norm.bias"))
    ks2 = None
    if p > 0 and masks.get(prefix + "dropout") is not None:
        ks2 = masks[prefix + "dropout"].to(A.dtype) / (1.0 - p)
    y = y0 if ks2 is None else y0 * ks2
    E = y @ g_("proj_out.weight").T + g_("proj_out.bias")
    ctx = dict(mask=mask, Vp=Vp, A=A, U=U, Qblk=Qblk, g=g, Z=Z, Phi=Phi, sp=sp, Pm=Pm, XZ=XZ, O=O, z=z, mu=mu, rstd=rstd,
               y=y, ks=ks, ks2=ks2, valid=valid, t_hat=t_hat, scale=scale, has_in=has_in, H=H)
    return E, M_txt, ctx


def backward(P: Dict[str, Tensor], notes, tau, ctx, dE, prefix: str = "ttf."):
    """Gradient of every parameter (state_dict names) for upstream dE [B,T,d]; the GEMM sequence of T2VPerQueryFn.backward."""
    g_ = lambda k: P[prefix + k]
    c = ctx
    B, T, d = dE.shape
    H = c["H"]
    hd = d // H
    dt = d // 2
    W_kv = g_("KV_proj.weight")
    W_a, W_phi = W_kv[:, :d], W_kv[:, d:]
    in_w, in_b = g_("attn.in_proj_weight"), g_("attn.in_proj_bias")
    W_q, W_k, W_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:]
    b_v = in_b[2 * d:]
    Qp = g_("Q_param").reshape(d)
    W_o = g_("attn.out_proj.weight")
    G = {}
    dE2 = dE.reshape(B * T, d)
    G["proj_out.weight"] = dE2.T @ c["y"].reshape(B * T, d)
    G["proj_out.bias"] = dE2.sum(0)
    dy = dE @ g_("proj_out.weight")
    if c["ks2"] is not None:
        dy = dy * c["ks2"]
    dz, G["layer_norm.weight"], G["layer_norm.bias"] = _ln_bwd(dy, c["z"], g_("layer_norm.weight"), c["mu"], c["rstd"])
    dres = dz.reshape(B * T, d).sum(0)
    dx = dz * c["valid"]
    dx2 = dx.reshape(B * T, d)
    G["attn.out_proj.weight"] = dx2.T @ c["O"].reshape(B * T, d)
    G["attn.out_proj.bias"] = dx2.sum(0)
    dO = dx @ W_o
    d_in_w, d_in_b = torch.zeros_like(in_w), torch.zeros_like(in_b)
    dXZ = torch.empty_like(c["XZ"])
    dsp = torch.empty_like(c["sp"])
    for h in range(H):
        hs = slice(h * hd, (h + 1) * hd)
        dO_h = dO[:, :, hs].reshape(B * T, hd)
        dXZ[:, :, h] = (dO_h @ W_v[hs]).reshape(B, T, d)
        d_in_w[2 * d + h * hd:2 * d + (h + 1) * hd] = dO_h.T @ c["XZ"][:, :, h].reshape(B * T, d)
        d_in_b[2 * d + h * hd:2 * d + (h + 1) * hd] = dO_h.T @ c["sp"][:, :, h].reshape(B * T)
        dsp[:, :, h] = (dO_h @ b_v[hs]).reshape(B, T)
    dZ = dXZ
    dPhi = dXZ @ W_phi
    dW_phi = dXZ.reshape(-1, d).T @ c["Phi"].reshape(-1, dt)
    tw = (g_("time2vec.linear.weight"), g_("time2vec.linear.bias"), g_("time2vec.periodic.weight"), g_("time2vec.periodic.bias"))
    dA, da, dw, db, dg = pool_bwd(dZ, dPhi, dsp, c["A"], c["g"], c["Pm"], tau, c["t_hat"], *tw, keep_scale=c["ks"])
    G["time2vec.linear.weight"], G["time2vec.linear.bias"] = dw[:1].reshape(1, 1), db[:1]
    G["time2vec.periodic.weight"], G["time2vec.periodic.bias"] = dw[1:].reshape(-1, 1), db[1:]
    A2, da2 = c["A"].reshape(-1, d), da.reshape(-1, H)
    dU = da2.T @ A2 + dg @ W_phi.T
    dA = dA + da @ c["U"]
    dA = dA * c["mask"].unsqueeze(-1)
    dW_phi = dW_phi + c["U"].T @ dg
    d_in_w[d:2 * d] = c["Qblk"].T @ dU
    dQblk = dU @ W_k.T
    dq = torch.cat([dQblk[h, h * hd:(h + 1) * hd] for h in range(H)])
    dq_pre = dq * c["scale"]
    d_in_w[:d] = dq_pre[:, None] * Qp[None, :]
    d_in_b[:d] = dq_pre
    G["Q_param"] = (dq_pre @ W_q + dres).reshape(1, 1, d)
    G["attn.in_proj_weight"], G["attn.in_proj_bias"] = d_in_w, d_in_b
    dA2 = dA.reshape(-1, d)
    dW_a = dA2.T @ c["Vp"].reshape(-1, d)
    G["KV_proj.weight"] = torch.cat([dW_a, dW_phi], dim=1)
    G["KV_proj.bias"] = dA2.sum(0)
    if c["has_in"]:
        dVp = dA2 @ W_a
        G["input_proj.weight"] = dVp.T @ notes.reshape(-1, notes.shape[-1])
        G["input_proj.bias"] = dVp.sum(0)
    return {prefix + k: v for k, v in G.items()}
