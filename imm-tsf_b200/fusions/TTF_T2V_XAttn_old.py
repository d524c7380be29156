"""TTF_T2V_XAttn, per-(note, query) variant -- the semantics of the reference's fusions/TTF_T2V_XAttn_old.py:27-161:
Time2Vec encodes the clamped lag max(t_hat - tau, 0) of EVERY (note, query) pair, so the query time matters and the
attention is a dense [T_f x N_i] problem per sample (in the active module only the length of t_hat is used).

Same parameter names / shapes and results as the reference class (pinned by tests/golden/pq_*.npz, generated from that
class).  Two additions so that it plugs into FusionModel like the active module: the optional `d_txt` / `input_proj`
(fusions/TTF_T2V_XAttn.py:59-69, 120-121; `d_txt=None` reproduces the _old constructor) and the registry name
"TTF_T2V_XAttn_old".  The schedule differs completely (immtsf/functional.py: T2VPerQueryFn): nothing of width d is
ever formed per (note, query) pair."""
from __future__ import annotations

import torch
import torch.nn as nn

from immtsf import functional as F_, ops, runtime
from fusions.load_llm import get_d_model
from fusions import _common as cm
from fusions.TTF_T2V_XAttn import Time2Vec


class TTF_T2V_XAttn(nn.Module):
    def __init__(self, llm_model_fusion: str, llm_layers_fusion: int, max_length: int = 1024, device: str = "cpu",
                 use_text_embeddings: bool = True, n_heads_fusion: int = 1, dropout: float = 0.1,
                 d_txt: int | None = None):
        super().__init__()
        self.use_text_embeddings = use_text_embeddings
        if not use_text_embeddings:
            raise NotImplementedError("TTF_T2V_XAttn_old (B200): only precomputed text embeddings are supported")
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
        if self.d_txt % n_heads_fusion != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")  # nn.MultiheadAttention's own check
        self.time2vec = Time2Vec(self.d_tau)
        self.KV_proj = nn.Linear(self.d_txt + self.d_tau, self.d_txt)
        # parameter container only (packed in_proj_weight [3d,d], in_proj_bias, out_proj.*); never called
        self.attn = nn.MultiheadAttention(embed_dim=self.d_txt, num_heads=n_heads_fusion, dropout=dropout, batch_first=True)
        self.layer_norm = nn.LayerNorm(self.d_txt)
        self.dropout = nn.Dropout(dropout)
        self.proj_out = nn.Linear(self.d_txt, self.d_txt)
        self.Q_param = nn.Parameter(torch.randn(1, 1, self.d_txt))

    def final_proj(self):
        return self.proj_out.weight, self.proj_out.bias

    def can_defer(self) -> bool:
        return False

    def forward_ragged(self, r: ops.RaggedNotes, t_hat: torch.Tensor, defer: bool = False):
        assert not defer
        t_hat, T = cm.fix_t_hat(t_hat, r.B)
        thr, seed = cm.dropout_args(self.dropout.p, self.training, self)
        ip, t2v, at = self.input_proj, self.time2vec, self.attn
        params = (self.Q_param, ip.weight if ip is not None else None, ip.bias if ip is not None else None,
                  t2v.linear.weight, t2v.linear.bias, t2v.periodic.weight, t2v.periodic.bias,
                  self.KV_proj.weight, self.KV_proj.bias, at.in_proj_weight, at.in_proj_bias,
                  at.out_proj.weight, at.out_proj.bias, self.layer_norm.weight, self.layer_norm.bias,
                  self.proj_out.weight, self.proj_out.bias)
        save = F_._need_save(*params)
        E_txt = F_.T2VPerQueryFn.apply(r, t_hat, T, self.n_heads, thr, seed, save, *params)
        return E_txt, cm.m_txt_bool(r)

    def forward(self, notes_input, tau: torch.Tensor, t_hat: torch.Tensor):
        cm.require_cuda(notes_input, "TTF_T2V_XAttn_old")
        r = ops.csr_build(cm.as_f32(notes_input), cm.as_f32(tau))
        out = self.forward_ragged(r, t_hat.to(notes_input.device))
        runtime.raise_on_flags(r.flags, (ops.FLAG_V,))
        return out
