"""TTF_RecAvg -- recency-weighted pooling of ragged, timestamped text
embeddings at each forecast query time, on the immtsf sm_100a kernels.

Same constructor, parameter names/shapes, forward signature, return values and
error behaviour as the reference module (fusions/TTF_RecAvg.py:8-112); the
forward runs: pad->CSR (csrc/csr.cu), input_proj (csrc/gemm_*.cu), fused
recency-pool + LayerNorm + dropout (csrc/recavg.cu), proj."""
from __future__ import annotations

import torch
import torch.nn as nn

from immtsf import functional as F_, ops, runtime
from fusions.load_llm import get_d_model
from fusions import _common as cm


class TTF_RecAvg(nn.Module):
    def __init__(self, llm_model_fusion: str, llm_layers_fusion: int, max_length: int = 1024, device: str = "cpu",
                 use_text_embeddings: bool = True, recency_sigma: float = 1.0, dropout: float = 0.1,
                 d_txt: int | None = 768):
        super().__init__()
        self.use_text_embeddings = use_text_embeddings
        if not use_text_embeddings:
            raise NotImplementedError("TTF_RecAvg (B200): only precomputed text embeddings are supported")
        d_model = get_d_model(llm_model_fusion)
        if d_txt is not None:
            self.input_proj = nn.Linear(d_model, d_txt)
            self.d_txt = d_txt
        else:
            self.input_proj = None
            self.d_txt = d_model
        self.max_length = max_length
        assert recency_sigma > 0, "recency_sigma must be > 0"
        self.log_recency_sigma = nn.Parameter(torch.log(torch.tensor(recency_sigma)))
        self.proj = nn.Linear(self.d_txt, self.d_txt)
        self.layer_norm = nn.LayerNorm(self.d_txt)
        self.dropout = nn.Dropout(dropout)

    # -- ragged entry point (shared CSR + NaN flags when called from FusionModel)
    def final_proj(self):
        """The last linear map of forward (TTF_RecAvg.py:109), which a consumer may fold into its own operand."""
        return self.proj.weight, self.proj.bias

    def can_defer(self) -> bool:
        return True

    def forward_ragged(self, r: ops.RaggedNotes, t_hat: torch.Tensor, defer: bool = False):
        """defer=True: return dropout(LN(E_raw)) without `proj`; the caller applies final_proj() (FusionModel fuses it
        into the rank form of MMF_XAttn_Add)."""
        t_hat, T = cm.fix_t_hat(t_hat, r.B)
        thr, seed = cm.dropout_args(self.dropout.p, self.training, self)
        ip = self.input_proj
        params = (self.log_recency_sigma, ip.weight if ip is not None else None, ip.bias if ip is not None else None,
                  self.layer_norm.weight, self.layer_norm.bias, self.proj.weight, self.proj.bias)
        save = F_._need_save(*params)
        E_txt = F_.RecAvgFn.apply(r, t_hat, T, thr, seed, save, bool(defer), *params)
        return E_txt, cm.m_txt_bool(r)

    def forward(self, notes_input, tau: torch.Tensor, t_hat: torch.Tensor):
        cm.require_cuda(notes_input, "TTF_RecAvg")
        r = ops.csr_build(cm.as_f32(notes_input), cm.as_f32(tau))
        if t_hat.dim() != 1 and t_hat.shape[0] != r.B:  # shape error first, like the reference's control flow allows
            raise ValueError(f"Expected t_hat shape (B, T_f) or (T_f,), got {t_hat.shape}")
        out = self.forward_ragged(r, t_hat)
        runtime.raise_on_flags(r.flags, (ops.FLAG_V,))
        return out
