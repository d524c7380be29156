"""TTF_T2V_XAttn -- Time2Vec-encoded note timestamps + masked cross-attention
of a learned query over each sample's variable-length note set, on the immtsf
sm_100a kernels.

Same constructor, parameter names/shapes (incl. the nn.MultiheadAttention
packed in_proj_weight), forward signature and results as the reference module
(fusions/TTF_T2V_XAttn.py:7-184).  What differs is the schedule: K/V are
projected once per note (not once per (note, query) as :151-166 does), and in
eval / dropout-0 mode the T_f identical output rows are computed once and
returned as a broadcast view."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from immtsf import functional as F_, ops, runtime
from fusions.load_llm import get_d_model
from fusions import _common as cm


class Time2Vec(nn.Module):
    """Parameter container with the reference's layout (fusions/TTF_T2V_XAttn.py:7-24);
    the features are computed by immtsf_time2vec_fwd straight into the [V';phi] buffer."""

    def __init__(self, d_tau: int):
        super().__init__()
        assert d_tau > 1, "d_tau must be > 1"
        self.linear = nn.Linear(1, 1)
        self.periodic = nn.Linear(1, d_tau - 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        raise RuntimeError("Time2Vec is evaluated inside TTF_T2V_XAttn's fused CUDA path")


class TTF_T2V_XAttn(nn.Module):
    def __init__(self, llm_model_fusion: str, llm_layers_fusion: int, max_length: int = 1024, device: str = "cpu",
                 use_text_embeddings: bool = True, n_heads_fusion: int = 1, dropout: float = 0.1,
                 d_txt: int | None = 768):
        super().__init__()
        self.use_text_embeddings = use_text_embeddings
        if not use_text_embeddings:
            raise NotImplementedError("TTF_T2V_XAttn (B200): only precomputed text embeddings are supported")
        d_model = get_d_model(llm_model_fusion)
        if d_txt is not None:
            self.input_proj = nn.Linear(d_model, d_txt)
            self.d_txt = d_txt
        else:
            self.input_proj = None
            self.d_txt = d_model
        self.d_tau = self.d_txt // 2
        self.max_length = max_length
        self.n_heads = n_heads_fusion
        self.time2vec = Time2Vec(self.d_tau)
        self.KV_proj = nn.Linear(self.d_txt + self.d_tau, self.d_txt)
        # parameter container only (in_proj_weight [3d,d], in_proj_bias, out_proj.*); never called
        self.attn = nn.MultiheadAttention(embed_dim=self.d_txt, num_heads=n_heads_fusion, dropout=dropout, batch_first=True)
        self.layer_norm = nn.LayerNorm(self.d_txt)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = nn.Linear(self.d_txt, self.d_txt)
        self.Q_param = nn.Parameter(torch.randn(1, 1, self.d_txt))

    def final_proj(self):
        """The last linear map of forward (TTF_T2V_XAttn.py:182), which a consumer may fold into its own operand."""
        return self.proj_out.weight, self.proj_out.bias

    def can_defer(self) -> bool:
        """Only when every (sample, query time) row is distinct (train mode with attention dropout); in eval the T_f rows
        of a sample are one broadcast row."""
        return self.training and ops.drop_thr(self.dropout.p) != 0

    def _collapsed(self, thr: int) -> bool:
        """One head in train mode with attention dropout: the collapsed schedule (functional.T2VXAttnFoldFn)."""
        return thr != 0 and self.n_heads == 1 and os.environ.get("IMMTSF_T2V_COLLAPSE", "1") != "0"

    def csr_pad_cols(self) -> int:
        """Extra columns the pad -> CSR gather should leave to the right of the compacted notes: the collapsed schedule appends the
        Time2Vec features there and multiplies the whole buffer (ops.csr_build(pad_cols=...))."""
        thr = ops.drop_thr(self.dropout.p) if self.training else 0
        return self.d_tau if self._collapsed(thr) else 0

    def dp_prereduced_params(self):
        """The collapsed schedule all-reduces the sufficient statistics of every gradient of this module itself."""
        thr = ops.drop_thr(self.dropout.p) if self.training else 0
        return list(self.parameters()) if self._collapsed(thr) else []

    def forward_ragged(self, r: ops.RaggedNotes, t_hat: torch.Tensor, defer: bool = False):
        """defer=True: return dropout(LN(attn + Q)) without proj_out; the caller applies final_proj()."""
        _, T = cm.fix_t_hat(t_hat, r.B)  # only the length of t_hat matters (reference :143,150)
        thr, seed = cm.dropout_args(self.dropout.p, self.training, self)
        ip, t2v, at = self.input_proj, self.time2vec, self.attn
        params = (self.Q_param, ip.weight if ip is not None else None, ip.bias if ip is not None else None,
                  t2v.linear.weight, t2v.linear.bias, t2v.periodic.weight, t2v.periodic.bias,
                  self.KV_proj.weight, self.KV_proj.bias, at.in_proj_weight, at.in_proj_bias,
                  at.out_proj.weight, at.out_proj.bias, self.layer_norm.weight, self.layer_norm.bias,
                  self.proj_out.weight, self.proj_out.bias)
        save = F_._need_save(*params)
        # one head in train mode: the collapsed schedule (key projection = one vector, input_proj folded into KV_proj)
        fn = F_.T2VXAttnFoldFn if self._collapsed(thr) else F_.T2VXAttnFn
        E_txt = fn.apply(r, T, self.n_heads, thr, seed, save, bool(defer), *params)
        return E_txt, cm.m_txt_bool(r)

    def forward(self, notes_input, tau: torch.Tensor, t_hat: torch.Tensor):
        cm.require_cuda(notes_input, "TTF_T2V_XAttn")
        r = ops.csr_build(cm.as_f32(notes_input), cm.as_f32(tau), pad_cols=self.csr_pad_cols())
        out = self.forward_ragged(r, t_hat)
        runtime.raise_on_flags(r.flags, (ops.FLAG_V,))
        return out
