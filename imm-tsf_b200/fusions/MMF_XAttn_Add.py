"""MMF_XAttn_Add -- cross-attention (forecast queries over the time-aligned
text embeddings) residual add with a fixed kappa blend, on the immtsf sm_100a
kernels.  Same constructor, parameter names/shapes, forward signature and
results as the reference (fusions/MMF_XAttn_Add.py:10-103)."""
from __future__ import annotations

import torch
import torch.nn as nn

from immtsf import functional as F_, ops, runtime
from fusions import _common as cm


class MMF_XAttn_Add(nn.Module):
    def __init__(self, d_txt: int, C: int, d_attn: int, n_heads_fusion: int = 1, dropout: float = 0.1,
                 kappa: float = 1.0):
        super().__init__()
        self.C = C
        self.d_attn = d_attn
        self.kappa = kappa
        self.n_heads = n_heads_fusion
        self.proj_q = nn.Linear(C, d_attn, bias=False)
        self.proj_k = nn.Linear(d_txt, d_attn, bias=False)
        self.proj_v = nn.Linear(d_txt, d_attn, bias=False)
        # parameter container only; the attention itself is immtsf_xattn_core_*
        self.attn = nn.MultiheadAttention(embed_dim=d_attn, num_heads=n_heads_fusion, dropout=dropout, batch_first=True)
        self.residual_head = nn.Linear(d_attn, C)
        self.layer_norm = nn.LayerNorm(C)
        self.dropout = nn.Dropout(dropout)

    def forward_flags(self, Y_ts, E_txt, M_txt, flags, final_proj=None):
        """final_proj = (W_p, b_p): E_txt is handed in WITHOUT the producer's last projection (E_txt_true = E_txt W_p^T +
        b_p); only valid on the rank path, which folds it into its skinny operand."""
        cm.require_cuda(Y_ts, "MMF_XAttn_Add")
        B, T, C = Y_ts.shape
        thr, seed = cm.dropout_args(self.dropout.p, self.training)
        at = self.attn
        params = (self.proj_q.weight, self.proj_k.weight, self.proj_v.weight, at.in_proj_weight, at.in_proj_bias,
                  at.out_proj.weight, at.out_proj.bias, self.residual_head.weight, self.residual_head.bias,
                  self.layer_norm.weight, self.layer_norm.bias)
        save = F_._need_save(Y_ts, E_txt, *params)
        own_flags = flags if flags is not None else runtime.new_flags(Y_ts.device)
        # Time-IMM shapes (T <= 32, few channels): the rank-(2C+1) form -- one skinny pass over E_txt, no tensor of width d
        fn = F_.XAttnAddRankFn if self.rank_path(T) else F_.XAttnAddFn
        extra = ()
        if fn is F_.XAttnAddRankFn:
            extra = tuple(final_proj) if final_proj is not None else (None, None)
        elif final_proj is not None:
            raise RuntimeError("MMF_XAttn_Add: a deferred projection needs the rank path")
        out = fn.apply(cm.as_f32(Y_ts), cm.as_f32(E_txt), cm.m_txt_u8(M_txt, B), self.n_heads, float(self.kappa),
                       thr, seed, save, own_flags, *params, *extra)
        if flags is None:  # standalone call: keep the reference's "delta_y contains NaN" ValueError (:84-91)
            if runtime.nan_check_enabled() and own_flags.tolist()[ops.FLAG_OUT]:
                raise ValueError("delta_y contains NaN values.")
        return out

    def rank_path(self, T: int) -> bool:
        """Whether forward will use the rank-(2C+1) form (csrc/xattn_rank.cu) for T query times."""
        return ops.xattn_rank_ok(T, self.n_heads, self.d_attn, self.C)

    def forward(self, Y_ts, E_txt, M_txt):
        return self.forward_flags(Y_ts, E_txt, M_txt, None)
