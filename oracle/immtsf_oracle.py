"""CPU oracle for the IMM-TSF text->time-series fusion hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the CPU baseline.
The product path (``imm-tsf_b200/``) never imports this file and has no CPU
fallback.

What it is: a plain-tensor restatement (torch CPU ops: matmul, exp, sin,
tanh, sigmoid, softmax -- no ``nn.MultiheadAttention``, no ``nn.GRU``, no
``nn.LayerNorm``) of the four reference modules and their composition, one
function per reference ``forward``.  It follows the reference line by line,
including the ``T_f``-fold K/V expansion of the active TTF_T2V_XAttn, so its
run time on host cores is an honest stand-in for the reference's CPU path.
Gradients come from autograd over this restatement -- which is exactly how
the reference obtains them (``loss.backward()`` at main.py:1097).

Parameters are passed as a flat ``dict`` keyed by the reference's
``state_dict`` names (``ttf.input_proj.weight`` ...).

Parity pin: the reference ships no tests or golden vectors.  This oracle is
pinned against outputs of the reference itself, generated in the build
container by ``oracle/make_golden.py`` (imports /root/reference/fusions) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the
oracle against them.

Dropout: the reference draws masks from torch's RNG stream, which cannot be
reproduced by a CUDA kernel.  Every dropout site therefore takes an optional
explicit keep-mask (``masks`` dict); with ``p == 0`` / eval no mask is needed.
Site names: ``ttf.dropout`` [B,T,d], ``ttf.attn_dropout`` [B,T,H,N],
``mmf.dropout`` [B,T,C], ``mmf.attn_dropout`` [B,H,T,T].
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------
# primitives (restated; semantics of torch 2.7 nn.Linear / LayerNorm / GRU / MHA)
# --------------------------------------------------------------------------
def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.transpose(-1, -2)
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)  # biased, as nn.LayerNorm
    return (x - mu) / torch.sqrt(var + eps) * w + b


def apply_dropout(x: Tensor, p: float, keep: Optional[Tensor]) -> Tensor:
    """keep: 0/1 mask of x's shape (or broadcastable); None => identity."""
    if keep is None or p <= 0.0:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


def note_mask_from_content(V: Tensor) -> Tensor:
    """fusions/TTF_RecAvg.py:69, fusions/TTF_T2V_XAttn.py:107."""
    return (V.abs().sum(dim=2) > 0).to(torch.bool)


def mha(
    query: Tensor,  # [Bq, L, E]
    key: Tensor,  # [Bq, S, E]
    value: Tensor,  # [Bq, S, E]
    in_w: Tensor,  # [3E, E]
    in_b: Tensor,  # [3E]
    out_w: Tensor,
    out_b: Tensor,
    n_heads: int,
    key_padding_mask: Optional[Tensor],  # [Bq, S] bool, True = ignore
    p: float = 0.0,
    keep: Optional[Tensor] = None,  # [Bq, H, L, S]
) -> Tensor:
    """torch.nn.functional.multi_head_attention_forward, need_weights=True branch
    (the one nn.MultiheadAttention takes at fusions/TTF_T2V_XAttn.py:161-166 and
    fusions/MMF_XAttn_Add.py:76): separate/packed in-projection, q scaled by
    hd^-1/2 BEFORE q k^T, additive -inf key-padding mask, softmax, dropout on the
    weights, weighted sum, out-projection."""
    Bq, L, E = query.shape
    S = key.shape[1]
    hd = E // n_heads
    q = linear(query, in_w[:E], in_b[:E])
    k = linear(key, in_w[E : 2 * E], in_b[E : 2 * E])
    v = linear(value, in_w[2 * E :], in_b[2 * E :])
    q = q.reshape(Bq, L, n_heads, hd).permute(0, 2, 1, 3)  # [Bq,H,L,hd]
    k = k.reshape(Bq, S, n_heads, hd).permute(0, 2, 1, 3)
    v = v.reshape(Bq, S, n_heads, hd).permute(0, 2, 1, 3)
    q = q * math.sqrt(1.0 / float(hd))
    s = q @ k.transpose(-1, -2)  # [Bq,H,L,S]
    if key_padding_mask is not None:
        neg = torch.zeros(Bq, 1, 1, S, dtype=s.dtype, device=s.device)
        neg = neg.masked_fill(key_padding_mask.view(Bq, 1, 1, S), float("-inf"))
        s = s + neg
    a = torch.softmax(s, dim=-1)
    a = apply_dropout(a, p, keep)
    o = a @ v  # [Bq,H,L,hd]
    o = o.permute(0, 2, 1, 3).reshape(Bq, L, E)
    return linear(o, out_w, out_b)


def gru(x: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """nn.GRU(batch_first=True, 1 layer, h0 = 0), gate order r,z,n
    (fusions/MMF_GR_Add.py:20-22,46)."""
    B, T, _ = x.shape
    Hd = w_hh.shape[1]
    h = torch.zeros(B, Hd, dtype=x.dtype, device=x.device)
    gi_all = linear(x, w_ih, b_ih)  # [B,T,3H]
    outs = []
    for t in range(T):
        gi = gi_all[:, t]
        gh = linear(h, w_hh, b_hh)
        i_r, i_z, i_n = gi.chunk(3, dim=1)
        h_r, h_z, h_n = gh.chunk(3, dim=1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        h = (1.0 - z) * n + z * h
        outs.append(h)
    return torch.stack(outs, dim=1)


# --------------------------------------------------------------------------
# the four modules
# --------------------------------------------------------------------------
def _fix_t_hat(t_hat: Tensor, B: int) -> Tensor:
    """fusions/TTF_RecAvg.py:86-91, fusions/TTF_T2V_XAttn.py:128-133."""
    if t_hat.dim() == 1:
        return t_hat.unsqueeze(0).repeat(B, 1)
    if t_hat.shape[0] != B:
        raise ValueError(f"Expected t_hat shape (B, T_f) or (T_f,), got {t_hat.shape}")
    return t_hat


def ttf_recavg(
    P: Params,
    notes: Tensor,
    tau: Tensor,
    t_hat: Tensor,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    prefix: str = "ttf.",
) -> Tuple[Tensor, Tensor]:
    """fusions/TTF_RecAvg.py:54-112."""
    masks = masks or {}
    V = notes
    note_mask = note_mask_from_content(V)  # :69
    if torch.isnan(V).any():  # :75
        raise ValueError("Input embeddings V contain NaN values.")
    if prefix + "input_proj.weight" in P:  # :79-80
        V = linear(V, P[prefix + "input_proj.weight"], P[prefix + "input_proj.bias"])
    B = V.shape[0]
    t_hat = _fix_t_hat(t_hat, B)
    delta = (t_hat[:, None] - tau[:, :, None]).clamp_min(0)  # :94  [B,N,T]
    sigma = P[prefix + "log_recency_sigma"].exp()  # :95
    w = torch.exp(-((delta / sigma) ** 2))  # :96
    w = w * note_mask.to(w.dtype)[:, :, None]  # :97
    E_wsum = torch.einsum("bnt,bnd->btd", w, V)  # :100
    denom = w.sum(dim=1).clamp_min(1e-6)  # :101
    E_raw = E_wsum / denom.unsqueeze(-1)  # :102
    E_norm = layer_norm(E_raw, P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"])
    E_drop = apply_dropout(E_norm, p, masks.get(prefix + "dropout"))  # :106
    E_txt = linear(E_drop, P[prefix + "proj.weight"], P[prefix + "proj.bias"])  # :109
    M_txt = note_mask.any(dim=1, keepdim=True)  # :110
    return E_txt, M_txt


def time2vec(P: Params, x: Tensor, prefix: str) -> Tensor:
    """fusions/TTF_T2V_XAttn.py:20-24; x: [...,1]."""
    lin = linear(x, P[prefix + "linear.weight"], P[prefix + "linear.bias"])
    per = torch.sin(linear(x, P[prefix + "periodic.weight"], P[prefix + "periodic.bias"]))
    return torch.cat([lin, per], dim=-1)


def ttf_t2v_xattn(
    P: Params,
    notes: Tensor,
    tau: Tensor,
    t_hat: Tensor,
    n_heads: int = 1,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    prefix: str = "ttf.",
    faithful_expand: bool = True,
) -> Tuple[Tensor, Tensor]:
    """fusions/TTF_T2V_XAttn.py:93-184.

    faithful_expand=True materialises K/V once per (sample, query) exactly like
    :151-159 (what the reference pays for on CPU).  False shares K/V across the
    T_f axis -- same numbers, used only to keep big parity cases cheap."""
    masks = masks or {}
    V = notes
    note_mask = note_mask_from_content(V)  # :107
    if torch.isnan(V).any():  # :116
        raise ValueError("Input embeddings V contain NaN values.")
    if prefix + "input_proj.weight" in P:  # :120-121
        V = linear(V, P[prefix + "input_proj.weight"], P[prefix + "input_proj.bias"])
    M_txt = note_mask.any(dim=1, keepdim=True)  # :124
    B, N, d = V.shape
    t_hat = _fix_t_hat(t_hat, B)
    T = t_hat.shape[1]
    tau_feat = time2vec(P, tau.unsqueeze(-1), prefix + "time2vec.")  # :136
    V_fused = torch.cat([V, tau_feat], dim=-1)  # :139
    KV = linear(V_fused, P[prefix + "KV_proj.weight"], P[prefix + "KV_proj.bias"])  # :140
    Qp = P[prefix + "Q_param"]
    mask_pad = ~note_mask  # :146
    keep = masks.get(prefix + "attn_dropout")  # [B,T,H,N]
    aw = (
        P[prefix + "attn.in_proj_weight"],
        P[prefix + "attn.in_proj_bias"],
        P[prefix + "attn.out_proj.weight"],
        P[prefix + "attn.out_proj.bias"],
    )
    if faithful_expand:
        Q_flat = Qp.expand(B, T, d).reshape(B * T, 1, d)  # :143,150
        KV_flat = KV.unsqueeze(1).expand(-1, T, -1, -1).reshape(B * T, N, d)  # :151-154
        mp_flat = mask_pad.unsqueeze(1).expand(-1, T, -1).reshape(B * T, N)  # :155-159
        kflat = None if keep is None else keep.reshape(B * T, n_heads, 1, N)
        attn_out = mha(Q_flat, KV_flat, KV_flat, *aw, n_heads, mp_flat, p, kflat)  # :161-166
        E_attn = attn_out.reshape(B, T, d)  # :167
    else:
        Q_b = Qp.expand(B, T, d)
        kb = None if keep is None else keep.permute(0, 2, 1, 3)  # [B,H,T,N]
        E_attn = mha(Q_b, KV, KV, *aw, n_heads, mask_pad, p, kb)
    mask = M_txt.view(B, 1, 1).expand(B, T, d)  # :171
    E_attn = torch.where(mask, E_attn, torch.zeros_like(E_attn))  # :173
    E_resid = layer_norm(  # :177-178
        E_attn + Qp.expand(B, T, d), P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"]
    )
    E_drop = apply_dropout(E_resid, p, masks.get(prefix + "dropout"))  # :179
    E_txt = linear(E_drop, P[prefix + "proj_out.weight"], P[prefix + "proj_out.bias"])  # :182
    return E_txt, M_txt


def mmf_gr_add(
    P: Params,
    Y_ts: Tensor,
    E_txt: Tensor,
    M_txt: Tensor,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    prefix: str = "mmf.",
) -> Tensor:
    """fusions/MMF_GR_Add.py:31-61."""
    masks = masks or {}
    B, T, C = Y_ts.shape
    x = torch.cat([Y_ts, E_txt], dim=-1)  # :43
    h = gru(  # :46
        x,
        P[prefix + "gru.weight_ih_l0"],
        P[prefix + "gru.weight_hh_l0"],
        P[prefix + "gru.bias_ih_l0"],
        P[prefix + "gru.bias_hh_l0"],
    )
    delta_y = linear(h, P[prefix + "residual_head.weight"], P[prefix + "residual_head.bias"])  # :47
    delta_norm = layer_norm(delta_y, P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"])
    delta_drop = apply_dropout(delta_norm, p, masks.get(prefix + "dropout"))  # :51
    g = torch.sigmoid(linear(x, P[prefix + "gate_net.weight"], P[prefix + "gate_net.bias"]))  # :54-55
    mask = M_txt.view(B, 1, 1).expand(-1, T, C)  # :56
    g = torch.where(mask, g, torch.ones_like(g))  # :57
    return g * Y_ts + (1 - g) * (Y_ts + delta_drop)  # :60


def mmf_xattn_add(
    P: Params,
    Y_ts: Tensor,
    E_txt: Tensor,
    M_txt: Tensor,
    n_heads: int = 1,
    kappa: float = 1.0,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    prefix: str = "mmf.",
) -> Tensor:
    """fusions/MMF_XAttn_Add.py:56-103."""
    masks = masks or {}
    B, T, C = Y_ts.shape
    Q = linear(Y_ts, P[prefix + "proj_q.weight"])  # :68
    K = linear(E_txt, P[prefix + "proj_k.weight"])  # :69
    V = linear(E_txt, P[prefix + "proj_v.weight"])  # :70
    d_attn = Q.shape[-1]
    key_pad = (~M_txt).view(B, 1).expand(-1, T)  # :73
    attn_out = mha(  # :76
        Q,
        K,
        V,
        P[prefix + "attn.in_proj_weight"],
        P[prefix + "attn.in_proj_bias"],
        P[prefix + "attn.out_proj.weight"],
        P[prefix + "attn.out_proj.bias"],
        n_heads,
        key_pad,
        p,
        masks.get(prefix + "attn_dropout"),
    )
    mask_attn = M_txt.view(B, 1, 1).expand(-1, T, d_attn)  # :79
    attn_out = torch.where(mask_attn, attn_out, torch.zeros_like(attn_out))  # :80
    delta_y = linear(attn_out, P[prefix + "residual_head.weight"], P[prefix + "residual_head.bias"])
    if torch.isnan(delta_y).any():  # :84-91
        raise ValueError("delta_y contains NaN values.")
    delta_norm = layer_norm(delta_y, P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"])
    delta_drop = apply_dropout(delta_norm, p, masks.get(prefix + "dropout"))  # :95
    mask = M_txt.view(B, 1, 1).expand(-1, T, C)  # :98
    delta_drop = torch.where(mask, delta_drop, torch.zeros_like(delta_drop))  # :99
    return (Y_ts + kappa * delta_drop) / (1.0 + kappa)  # :102


# --------------------------------------------------------------------------
# composition (fusions/FusionModel.py:98-113)
# --------------------------------------------------------------------------
def fusion_forward(
    P: Params,
    ttf_name: str,
    mmf_name: str,
    notes: Tensor,
    tau: Tensor,
    t_hat: Tensor,
    Y_ts: Tensor,
    n_heads: int = 1,
    kappa: float = 0.5,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    faithful_expand: bool = True,
    return_intermediate: bool = False,
):
    if torch.isnan(Y_ts).any():  # :103
        raise ValueError("Y_ts contains NaN values.")
    if ttf_name == "TTF_RecAvg":
        E_txt, M_txt = ttf_recavg(P, notes, tau, t_hat, p, masks)
    elif ttf_name == "TTF_T2V_XAttn":
        E_txt, M_txt = ttf_t2v_xattn(P, notes, tau, t_hat, n_heads, p, masks, faithful_expand=faithful_expand)
    elif ttf_name == "TTF_T2V_XAttn_old":
        E_txt, M_txt = ttf_t2v_xattn_perquery(P, notes, tau, t_hat, n_heads, p, masks)
    else:
        raise KeyError(ttf_name)
    if torch.isnan(E_txt).any():  # :107
        raise ValueError("E_txt contains NaN values.")
    if mmf_name == "MMF_GR_Add":
        Y_out = mmf_gr_add(P, Y_ts, E_txt, M_txt, p, masks)
    elif mmf_name == "MMF_XAttn_Add":
        Y_out = mmf_xattn_add(P, Y_ts, E_txt, M_txt, n_heads, kappa, p, masks)
    else:
        raise KeyError(mmf_name)
    if torch.isnan(Y_out).any():  # :111
        raise ValueError("Y_out contains NaN values.")
    if return_intermediate:
        return Y_out, E_txt, M_txt
    return Y_out


# --------------------------------------------------------------------------
# ragged (CSR) layout -- integer oracle for the pad->CSR adapter
# --------------------------------------------------------------------------
def csr_from_padded(notes: Tensor):
    """Bit-exact definition of the ragged layout: ``offsets[B+1]`` (int32
    exclusive scan of per-sample valid-note counts), ``rows[sumN]`` (flat index
    b*N_max+n of every valid note, sample-major, original note order),
    ``seg_id[sumN]`` (owning sample).  Valid == (V.abs().sum(2) > 0), the
    reference's own mask (fusions/TTF_RecAvg.py:69)."""
    mask = note_mask_from_content(notes)
    B, N = mask.shape
    counts = mask.sum(dim=1).to(torch.int32)
    offsets = torch.zeros(B + 1, dtype=torch.int32)
    offsets[1:] = torch.cumsum(counts, 0).to(torch.int32)
    flat = torch.nonzero(mask.reshape(-1), as_tuple=False).reshape(-1).to(torch.int32)
    seg = (flat // N).to(torch.int32)
    return offsets, flat, seg, mask


# --------------------------------------------------------------------------
# parameter construction with the reference's shapes (values are caller's job)
# --------------------------------------------------------------------------
def param_shapes(ttf_name: str, mmf_name: str, d_model: int, d_txt: Optional[int], C: int):
    """state_dict name -> shape, SURVEY.md section 8 a8 (probed from the reference)."""
    d = d_txt if d_txt is not None else d_model
    s: Dict[str, Tuple[int, ...]] = {}
    if ttf_name == "TTF_RecAvg":
        s["ttf.log_recency_sigma"] = ()
        if d_txt is not None:
            s["ttf.input_proj.weight"] = (d, d_model)
            s["ttf.input_proj.bias"] = (d,)
        s["ttf.proj.weight"] = (d, d)
        s["ttf.proj.bias"] = (d,)
        s["ttf.layer_norm.weight"] = (d,)
        s["ttf.layer_norm.bias"] = (d,)
    else:
        dt = d // 2
        s["ttf.Q_param"] = (1, 1, d)
        if d_txt is not None:
            s["ttf.input_proj.weight"] = (d, d_model)
            s["ttf.input_proj.bias"] = (d,)
        s["ttf.time2vec.linear.weight"] = (1, 1)
        s["ttf.time2vec.linear.bias"] = (1,)
        s["ttf.time2vec.periodic.weight"] = (dt - 1, 1)
        s["ttf.time2vec.periodic.bias"] = (dt - 1,)
        s["ttf.KV_proj.weight"] = (d, d + dt)
        s["ttf.KV_proj.bias"] = (d,)
        s["ttf.attn.in_proj_weight"] = (3 * d, d)
        s["ttf.attn.in_proj_bias"] = (3 * d,)
        s["ttf.attn.out_proj.weight"] = (d, d)
        s["ttf.attn.out_proj.bias"] = (d,)
        s["ttf.layer_norm.weight"] = (d,)
        s["ttf.layer_norm.bias"] = (d,)
        s["ttf.proj_out.weight"] = (d, d)
        s["ttf.proj_out.bias"] = (d,)
    if mmf_name == "MMF_GR_Add":
        s["mmf.gru.weight_ih_l0"] = (3 * C, C + d)
        s["mmf.gru.weight_hh_l0"] = (3 * C, C)
        s["mmf.gru.bias_ih_l0"] = (3 * C,)
        s["mmf.gru.bias_hh_l0"] = (3 * C,)
        s["mmf.residual_head.weight"] = (C, C)
        s["mmf.residual_head.bias"] = (C,)
        s["mmf.gate_net.weight"] = (C, C + d)
        s["mmf.gate_net.bias"] = (C,)
        s["mmf.layer_norm.weight"] = (C,)
        s["mmf.layer_norm.bias"] = (C,)
    else:
        s["mmf.proj_q.weight"] = (d, C)
        s["mmf.proj_k.weight"] = (d, d)
        s["mmf.proj_v.weight"] = (d, d)
        s["mmf.attn.in_proj_weight"] = (3 * d, d)
        s["mmf.attn.in_proj_bias"] = (3 * d,)
        s["mmf.attn.out_proj.weight"] = (d, d)
        s["mmf.attn.out_proj.bias"] = (d,)
        s["mmf.residual_head.weight"] = (C, d)
        s["mmf.residual_head.bias"] = (C,)
        s["mmf.layer_norm.weight"] = (C,)
        s["mmf.layer_norm.bias"] = (C,)
    return s


# ------------------------------------------------------------------ masked MSE
def masked_mse(truth: Tensor, pred: Tensor, mask: Tensor, reduce: str = "mean", count: Optional[Tensor] = None):
    """lib/evaluation.py:17-69 `compute_error(truth, pred_y, mask, "MSE", reduce)` restated (n_traj_samples = 1).
    error = (truth - pred)^2 * mask (:27-30); per-variable sums over every (sample, time) (:51-52);
    'mean': sum_c [err_c / (count_c + 1e-8)] / count_nonzero(count) (:57-61); 'sum': (err, count) (:65-67).
    count (optional): the GLOBAL per-variable counts of a batch-sharded run -- the returned value is then this shard's
    share of the global loss (shares add up to the single-process loss; immtsf/dp.py)."""
    C = pred.shape[-1]
    err = ((truth - pred) ** 2) * mask
    err_c = err.reshape(-1, C).sum(dim=0)
    cnt_c = mask.reshape(-1, C).sum(dim=0)
    if reduce == "sum":
        return err_c, cnt_c
    if count is None:
        count = cnt_c
    n_avail = torch.count_nonzero(count)
    return (err_c / (count + 1e-8)).sum() / n_avail


# ------------------------------------------------------------------ per-(note, query) Time2Vec attention
def ttf_t2v_xattn_perquery(
    P: Params,
    notes: Tensor,
    tau: Tensor,
    t_hat: Tensor,
    n_heads: int = 1,
    p: float = 0.0,
    masks: Optional[Dict[str, Tensor]] = None,
    prefix: str = "ttf.",
) -> Tuple[Tensor, Tensor]:
    """fusions/TTF_T2V_XAttn_old.py:82-161: the variant in which the query time matters.  Time2Vec encodes the
    clamped lag max(t_hat - tau, 0) of every (note, query) pair (:120-121), so keys/values differ per query time and
    the attention is a dense [T_f x N] problem per sample.  `input_proj` (absent from the _old file, present in the
    active module fusions/TTF_T2V_XAttn.py:120-121) is applied when the parameter exists, so the variant plugs into FusionModel(d_txt=...).

    masks: ``ttf.attn_dropout`` [B,T,H,N], ``ttf.dropout`` [B,T,d]."""
    masks = masks or {}
    V = notes
    note_mask = note_mask_from_content(V)  # :95
    if torch.isnan(V).any():  # :104
        raise ValueError("Input embeddings V contain NaN values.")
    if prefix + "input_proj.weight" in P:
        V = linear(V, P[prefix + "input_proj.weight"], P[prefix + "input_proj.bias"])
    M_txt = note_mask.any(dim=1, keepdim=True)  # :108
    B, N, d = V.shape
    t_hat = _fix_t_hat(t_hat, B)  # :112-117
    T = t_hat.shape[1]
    delta = (t_hat[:, None, :] - tau[:, :, None]).clamp_min(0)  # :120  [B,N,T]
    phi = time2vec(P, delta.unsqueeze(-1), prefix + "time2vec.")  # :121  [B,N,T,d_tau]
    V_exp = V.unsqueeze(2).expand(-1, -1, T, -1)  # :126
    KV = torch.cat([V_exp, phi], dim=-1)  # :127
    KV = KV.permute(0, 2, 1, 3).reshape(B * T, N, KV.shape[-1])  # :128
    KVp = linear(KV, P[prefix + "KV_proj.weight"], P[prefix + "KV_proj.bias"])  # :129
    Qp = P[prefix + "Q_param"]
    Q = Qp.expand(B, T, d).reshape(B * T, 1, d)  # :132
    mp_flat = (~note_mask).repeat_interleave(T, dim=0)  # :135
    keep = masks.get(prefix + "attn_dropout")
    kflat = None if keep is None else keep.reshape(B * T, n_heads, 1, N)
    attn_out = mha(  # :138-143
        Q, KVp, KVp,
        P[prefix + "attn.in_proj_weight"], P[prefix + "attn.in_proj_bias"],
        P[prefix + "attn.out_proj.weight"], P[prefix + "attn.out_proj.bias"],
        n_heads, mp_flat, p, kflat,
    )
    E_attn = attn_out.reshape(B, T, d)  # :144
    mask = M_txt.view(B, 1, 1).expand(B, T, d)  # :148
    E_attn = torch.where(mask, E_attn, torch.zeros_like(E_attn))  # :150
    E_resid = layer_norm(  # :154-155
        E_attn + Qp.expand(B, T, d), P[prefix + "layer_norm.weight"], P[prefix + "layer_norm.bias"]
    )
    E_drop = apply_dropout(E_resid, p, masks.get(prefix + "dropout"))  # :156
    E_txt = linear(E_drop, P[prefix + "proj_out.weight"], P[prefix + "proj_out.bias"])  # :159
    return E_txt, M_txt


# ------------------------------------------------------------------ text store / chunk windows
def chunk_windows(tt: Tensor, mask: Tensor, history: float, pred_window: float, stride: float):
    """lib/parse_datasets.py:176-227 restated for ONE record, numeric side only: the window starts `st` for which the
    reference forms a chunk BEFORE looking at the texts (>= 2 observations in [st, st+total), at least one observed
    value in the history part and one in the prediction part).  Python floats throughout, as in the reference."""
    total = history + pred_window
    t_max = tt.max().item()
    st = tt.min().item()
    out = []
    while st + total <= t_max:  # :183
        idx = ((tt >= st) & (tt < st + total)).nonzero(as_tuple=False).squeeze(1)  # :184-186
        if idx.numel() >= 2:  # :187
            sub_tt = tt[idx] - st
            sub_mask = mask[idx]
            hist_mask = sub_mask[sub_tt < history]  # :197
            pred_mask = sub_mask[sub_tt >= history]  # :198
            if hist_mask.sum() == 0 or pred_mask.sum() == 0:  # :201-203
                st += stride
                continue
            out.append(st)
        st += stride  # :227
    return out


def select_window_notes(record_texts, st: float, history: float):
    """lib/parse_datasets.py:203-209: notes of the history window only, in the order of the record's list (= file
    order), times made relative to the window start.  record_texts: list of (t, payload)."""
    hist_end = st + history
    return [(t - st, payload) for (t, payload) in record_texts if st <= t < hist_end]


def collate_text(raws):
    """The text part of multimodal_collate, lib/parse_datasets.py:781-819: raws[b] = list of (t, embedding row).
    Returns tau [B, N_max] and notes_embeddings [B, N_max, d] (zero tail padding)."""
    from torch.nn.utils.rnn import pad_sequence

    time_seqs = [torch.tensor([t for (t, _) in seq], dtype=torch.float32) for seq in raws]  # :786-790
    tau = pad_sequence(time_seqs, batch_first=True, padding_value=0.0)  # :792
    d_txt = None
    for seq in raws:  # :802-806
        if seq:
            d_txt = seq[0][1].size(-1)
            break
    if d_txt is None:
        return tau, torch.zeros((len(raws), 0, 0))  # :809
    emb_seqs = [torch.stack([e for (_, e) in seq], dim=0) if seq else torch.zeros((0, d_txt)) for seq in raws]  # :811-816
    return tau, pad_sequence(emb_seqs, batch_first=True, padding_value=0.0)  # :817-819
