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

    def _params(self):
        at = self.attn
        return (self.proj_q.weight, self.proj_k.weight, self.proj_v.weight, at.in_proj_weight, at.in_proj_bias,
                at.out_proj.weight, at.out_proj.bias, self.residual_head.weight, self.residual_head.bias,
                self.layer_norm.weight, self.layer_norm.bias)

    def rank_weights(self, final_proj=None, side: bool = False, flags=None):
        """Weight-space half of the rank form (functional.XAttnRankWeightsFn): (Wr, br, bo_f).  final_proj = (W_p, b_p)
        folds the producer's deferred last projection in.  side=True runs it on ops.side_stream() -- it depends on
        parameters only, so FusionModel starts it before the TTF forward; join with `wait_rank_weights`."""
        W_p, b_p = final_proj if final_proj is not None else (None, None)
        args = (self.n_heads, self.C) + self._params()[:9] + (W_p, b_p, self.layer_norm.weight, self.layer_norm.bias)
        def work():
            if flags is not None and W_p is not None:  # NaN guard of the folded projection (FusionModel.py:107-108)
                ops.nan_check(W_p, flags, ops.FLAG_E)
                ops.nan_check(b_p, flags, ops.FLAG_E)
            return F_.XAttnRankWeightsFn.apply(*args)

        if not side:
            return work()
        cur = torch.cuda.current_stream()
        st = ops.side_stream(cur.device)
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            out = work()
        return out

    @staticmethod
    def wait_rank_weights(weights):
        """Make the current stream wait for rank_weights(side=True) and tell the allocator that it uses the results."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(ops.side_stream(cur.device))
        for t in weights:
            t.record_stream(cur)

    def forward_flags(self, Y_ts, E_txt, M_txt, flags, final_proj=None, rank_weights=None):
        """final_proj = (W_p, b_p): E_txt is handed in WITHOUT the producer's last projection (E_txt_true = E_txt W_p^T +
        b_p); only valid on the rank path, which folds it into its skinny operand.  rank_weights: the result of
        rank_weights(...) when the caller already computed it (FusionModel, on the side stream)."""
        cm.require_cuda(Y_ts, "MMF_XAttn_Add")
        B, T, C = Y_ts.shape
        thr, seed = cm.dropout_args(self.dropout.p, self.training, self)
        params = self._params()
        save = F_._need_save(Y_ts, E_txt, *params)
        # flags: None = standalone call (own flags, delta_y ValueError below); False = the caller does not want NaN flags
        own_flags = None if flags is False else (flags if flags is not None else runtime.new_flags(Y_ts.device))
        # Time-IMM shapes (T <= 32, few channels): the rank-(2C+1) form -- one skinny pass over E_txt, no tensor of width d
        if self.rank_path(T):
            Wr, br, bo_f, gamma, beta = rank_weights if rank_weights is not None else self.rank_weights(final_proj)
            out = F_.XAttnRankDataFn.apply(cm.as_f32(Y_ts), cm.as_f32(E_txt), cm.m_txt_u8(M_txt, B), self.n_heads,
                                           float(self.kappa), thr, seed, save, own_flags, self.d_attn, Wr, br, bo_f, gamma, beta)
        else:
            if final_proj is not None or rank_weights is not None:
                raise RuntimeError("MMF_XAttn_Add: a deferred projection needs the rank path")
            out = F_.XAttnAddFn.apply(cm.as_f32(Y_ts), cm.as_f32(E_txt), cm.m_txt_u8(M_txt, B), self.n_heads, float(self.kappa),
                                      thr, seed, save, own_flags, *params)
        if flags is None and own_flags is not None:  # standalone call: keep the reference's "delta_y contains NaN" ValueError (:84-91)
            if runtime.nan_check_enabled() and own_flags.tolist()[ops.FLAG_OUT]:
                raise ValueError("delta_y contains NaN values.")
        return out

    def rank_path(self, T: int) -> bool:
        """Whether forward will use the rank-(2C+1) form (csrc/xattn_rank.cu) for T query times."""
        return ops.xattn_rank_ok(T, self.n_heads, self.d_attn, self.C)

    def forward(self, Y_ts, E_txt, M_txt):
        return self.forward_flags(Y_ts, E_txt, M_txt, None)
