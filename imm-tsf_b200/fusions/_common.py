"""Shared host logic of the drop-in fusion modules."""
from __future__ import annotations

import torch

from immtsf import ops, runtime


def require_cuda(t: torch.Tensor, who: str):
    if not isinstance(t, torch.Tensor):
        raise NotImplementedError(
            f"{who}: raw note strings (use_text_embeddings=False) are not supported by the B200 path; "
            "pass precomputed embeddings [B, N_max, d_model]."
        )
    if not t.is_cuda:
        raise RuntimeError(f"{who}: inputs must be CUDA tensors -- the immtsf path has no CPU fallback")


def as_f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


def fix_t_hat(t_hat: torch.Tensor, B: int):
    """Reference: fusions/TTF_RecAvg.py:86-91 / fusions/TTF_T2V_XAttn.py:128-133.
    A 1-D t_hat is shared by all samples (no repeat: kernels take a batch stride of 0)."""
    if not t_hat.is_cuda:  # a host pointer handed to a kernel is an asynchronous illegal address, not an exception
        raise RuntimeError("t_hat must be a CUDA tensor on the device of the notes -- the immtsf path has no CPU fallback")
    if t_hat.dim() == 1:
        return as_f32(t_hat).contiguous(), t_hat.shape[0]
    if t_hat.shape[0] != B:
        raise ValueError(f"Expected t_hat shape (B, T_f) or (T_f,), got {t_hat.shape}")
    return as_f32(t_hat).contiguous(), t_hat.shape[1]


def m_txt_bool(r: ops.RaggedNotes) -> torch.Tensor:
    """M_txt [B, 1] bool (TTF_RecAvg.py:110): the kernels' 0/1 bytes reinterpreted, no conversion launch."""
    return r.m_txt[: r.B].view(r.B, 1).view(torch.bool)


def m_txt_u8(M_txt: torch.Tensor, B: int) -> torch.Tensor:
    if M_txt.dtype == torch.bool and M_txt.is_contiguous():
        return M_txt.reshape(B).view(torch.uint8)  # same bytes
    return M_txt.reshape(B).to(torch.uint8).contiguous()


def dropout_args(module_p: float, training: bool, module=None):
    """(thr, seed) of a module's dropout sites.  The kernels take ONE LayerNorm epsilon (ops.LN_EPS, the nn.LayerNorm
    default the reference constructs with) and ONE drop rate per module (the reference passes the same `dropout` to
    nn.Dropout and nn.MultiheadAttention): a module whose containers were edited to anything else would silently diverge
    from the reference, so that is an error here."""
    if module is not None:
        ln, attn = getattr(module, "layer_norm", None), getattr(module, "attn", None)
        if ln is not None and ln.eps != ops.LN_EPS:
            raise ValueError(f"layer_norm.eps = {ln.eps}: the immtsf kernels are built for eps = {ops.LN_EPS}")
        if attn is not None and float(attn.dropout) != float(module_p):
            raise ValueError(f"attn.dropout = {attn.dropout} differs from dropout.p = {module_p}: one drop rate per module is supported")
    thr = ops.drop_thr(module_p) if training else 0
    seed = runtime.SEEDS.next() if thr else 0
    return thr, seed
