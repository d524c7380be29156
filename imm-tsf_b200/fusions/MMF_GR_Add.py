"""MMF_GR_Add -- GRU-gated residual add of the text signal into the forecast,
on the immtsf sm_100a kernels.  Same constructor, parameter names/shapes
(nn.GRU's weight_ih_l0 ... layout, gate order r,z,n), forward signature and
results as the reference (fusions/MMF_GR_Add.py:9-61)."""
from __future__ import annotations

import torch
import torch.nn as nn

from immtsf import functional as F_, ops, runtime
from fusions import _common as cm


class MMF_GR_Add(nn.Module):
    def __init__(self, d_txt: int, C: int, hidden_dim: int, dropout: float = 0.1):
        super().__init__()
        self.C = C
        self.d_txt = d_txt
        if hidden_dim != C:
            raise NotImplementedError("MMF_GR_Add (B200): hidden_dim must equal C (FusionModel.py:84 always passes C)")
        self.gru = nn.GRU(input_size=C + d_txt, hidden_size=hidden_dim, batch_first=True)  # parameter container
        self.residual_head = nn.Linear(hidden_dim, C)
        self.gate_net = nn.Linear(C + d_txt, C)
        self.layer_norm = nn.LayerNorm(C)
        self.dropout = nn.Dropout(dropout)

    def _apply(self, fn, *a, **k):
        # nn.GRU re-flattens its weights for cuDNN on every device move; harmless here but not needed
        return super()._apply(fn, *a, **k)

    def forward_flags(self, Y_ts, E_txt, M_txt, flags):
        cm.require_cuda(Y_ts, "MMF_GR_Add")
        B, T, C = Y_ts.shape
        thr, seed = cm.dropout_args(self.dropout.p, self.training, self)
        g = self.gru
        params = (g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0, self.residual_head.weight,
                  self.residual_head.bias, self.gate_net.weight, self.gate_net.bias, self.layer_norm.weight,
                  self.layer_norm.bias)
        save = F_._need_save(Y_ts, E_txt, *params)
        return F_.GRAddFn.apply(cm.as_f32(Y_ts), cm.as_f32(E_txt), cm.m_txt_u8(M_txt, B), thr, seed, save, flags, *params)

    def forward(self, Y_ts, E_txt, M_txt):
        return self.forward_flags(Y_ts, E_txt, M_txt, None)
