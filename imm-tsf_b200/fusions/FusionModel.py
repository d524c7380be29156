"""FusionModel -- the plugin surface of the hot path, unchanged for callers
(reference: fusions/FusionModel.py:24-113): `FusionModel(args)` with the same
Namespace attributes, `forward(notes_input, tau, t_hat, Y_ts) -> Y_out`, the
same `--TTF_module` / `--MMF_module` registries, the same state_dict.

Differences in schedule only: the padded batch is converted to the ragged CSR
layout once and shared by TTF and MMF; the reference's three isnan().any()
host syncs (:103-112) plus the ones inside the TTF/MMF modules are folded into
device flags read once at the end of forward (still raising ValueError)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from immtsf import ops, runtime
from fusions import _common as cm
from fusions.TTF_RecAvg import TTF_RecAvg
from fusions.TTF_T2V_XAttn import TTF_T2V_XAttn
from fusions.TTF_T2V_XAttn_old import TTF_T2V_XAttn as TTF_T2V_XAttn_old
from fusions.MMF_GR_Add import MMF_GR_Add
from fusions.MMF_XAttn_Add import MMF_XAttn_Add

# "TTF_T2V_XAttn_old": the per-(note, query) variant (reference file of that name; not in main.py's argparse choices --
# pass the class or this name through args.TTF_module)
_TTF_CLASSES = {"TTF_RecAvg": TTF_RecAvg, "TTF_T2V_XAttn": TTF_T2V_XAttn, "TTF_T2V_XAttn_old": TTF_T2V_XAttn_old}
_MMF_CLASSES = {"MMF_GR_Add": MMF_GR_Add, "MMF_XAttn_Add": MMF_XAttn_Add}


class FusionModel(nn.Module):
    def __init__(self, args):
        super().__init__()
        TTF_ref, MMF_ref = args.TTF_module, args.MMF_module
        TTF_cls = _TTF_CLASSES.get(TTF_ref, TTF_ref) if isinstance(TTF_ref, str) else TTF_ref
        MMF_cls = _MMF_CLASSES.get(MMF_ref, MMF_ref) if isinstance(MMF_ref, str) else MMF_ref
        print(f"Using TTF module: {args.TTF_module}")
        print(f"Using MMF module: {args.MMF_module}")
        common = dict(max_length=args.max_length, device=args.device, use_text_embeddings=args.use_text_embeddings,
                      dropout=args.dropout, d_txt=args.d_txt)
        if TTF_cls is TTF_RecAvg:
            self.ttf = TTF_cls(args.llm_model_fusion, args.llm_layers_fusion, recency_sigma=args.recency_sigma, **common)
        else:
            self.ttf = TTF_cls(args.llm_model_fusion, args.llm_layers_fusion, n_heads_fusion=args.n_heads_fusion, **common)
        d_txt = self.ttf.d_txt
        if MMF_cls is MMF_GR_Add:
            self.mmf = MMF_cls(d_txt=d_txt, C=args.C, hidden_dim=args.C, dropout=args.dropout)
        else:
            self.mmf = MMF_cls(d_txt=d_txt, C=args.C, d_attn=d_txt, n_heads_fusion=args.n_heads_fusion,
                               dropout=args.dropout, kappa=args.kappa)

    def enable_graphs(self, on: bool = True):
        """Transparent CUDA-graph replay of forward and backward, one captured pair per input shape (immtsf/autograph.py): the
        unmodified caller (lib/evaluation.py:95-100 + loss.backward(), main.py:1097) runs at graph speed.  Also IMMTSF_AUTOGRAPH=1."""
        self._autograph_on = bool(on)  # (captured graphs are kept when switched off; clear_graphs() drops them)
        return self

    def clear_graphs(self):
        self._autograph = None

    def forward(self, notes_input, tau, t_hat, Y_ts):
        cm.require_cuda(notes_input, "FusionModel")
        on = getattr(self, "_autograph_on", None)
        if on is None:
            on = self._autograph_on = os.environ.get("IMMTSF_AUTOGRAPH", "0") == "1"
        if (on and hasattr(self.ttf, "forward_ragged") and hasattr(self.mmf, "forward_flags") and not torch.cuda.is_current_stream_capturing()
                and notes_input.dim() == 3 and Y_ts.is_cuda and t_hat.is_cuda and tau.is_cuda):
            if getattr(self, "_autograph", None) is None:
                from immtsf import autograph

                self._autograph = autograph.GraphCache(self)
            return self._autograph(notes_input, tau, t_hat, Y_ts)
        return self._forward_eager(notes_input, tau, t_hat, Y_ts)

    def _forward_eager(self, notes_input, tau, t_hat, Y_ts):
        if not (hasattr(self.ttf, "forward_ragged") and hasattr(self.mmf, "forward_flags")):
            # a user-supplied TTF/MMF class: plain composition, reference order of checks
            if torch.isnan(Y_ts).any():
                raise ValueError("Y_ts contains NaN values.")
            E_txt, M_txt = self.ttf(notes_input, tau, t_hat)
            if torch.isnan(E_txt).any():
                raise ValueError("E_txt contains NaN values.")
            Y_out = self.mmf(Y_ts, E_txt, M_txt)
            if torch.isnan(Y_out).any():
                raise ValueError("Y_out contains NaN values.")
            return Y_out
        flags = runtime.new_flags(Y_ts.device)
        pad = self.ttf.csr_pad_cols() if hasattr(self.ttf, "csr_pad_cols") else 0
        r = ops.csr_build(cm.as_f32(notes_input), cm.as_f32(tau), flags, pad_cols=pad)
        return self.forward_csr(r, t_hat, Y_ts)

    def _schedule(self, T):
        """(rank, defer): MMF_XAttn_Add in its rank form for T query times / TTF final projection folded into it."""
        rank = isinstance(self.mmf, MMF_XAttn_Add) and self.mmf.rank_path(T)
        defer = rank and self.ttf.can_defer() and os.environ.get("IMMTSF_FUSE_PROJ", "1") != "0"
        return rank, defer

    def dp_prereduced_params(self, T):
        """Parameters whose gradients are born all-reduced under in-graph data parallelism (runtime.GraphedStep(allreduce_group=
        ...)): their Functions all-reduce the sufficient statistics of the gradients (ops.dp_allreduce) before the weight-space
        un-folds -- the rank form's [dWr | dbr | d bo | dgamma | dbeta], the collapsed T2V schedule's packed weight gradients."""
        rank, defer = self._schedule(T)
        ps = []
        if rank:
            ps += list(self.mmf._params())
            if defer:
                ps += list(self.ttf.final_proj())
        if hasattr(self.ttf, "dp_prereduced_params"):
            have = {id(p) for p in ps}
            ps += [p for p in self.ttf.dp_prereduced_params() if id(p) not in have]
        return ps

    def forward_csr(self, r, t_hat, Y_ts):
        """Same as forward() for callers that already hold the ragged layout (immtsf.collate.ragged_collate):
        no padded tensor, no content-mask pass, no compaction copy."""
        flags = r.flags
        self._last_flags = flags  # read by runtime.GraphedStep.check_nan()
        check = runtime.nan_flags_enabled()
        Y32 = cm.as_f32(Y_ts)
        if check:
            ops.nan_check(Y32, flags, ops.FLAG_Y)
        # one operand-split cache for the whole step: E_txt (and dE_txt in backward) are handed from one module's GEMM
        # epilogue to the other module's product together with their tcgen05 lo operand
        T = t_hat.shape[-1]
        # rank form of MMF_XAttn_Add: E_txt only enters through one skinny product, so the TTF's final projection is folded
        # into that operand in weight space and E_txt [B, T, d] is never materialised (IMMTSF_FUSE_PROJ=0 keeps it)
        rank, defer = self._schedule(T)
        x_cat, d_txt = None, self.ttf.d_txt
        if isinstance(self.mmf, MMF_GR_Add):
            # MMF_GR_Add reads x = [E ; Y]: the TTF's final projection writes E_txt straight into that buffer
            C = Y32.shape[-1]
            x_cat = torch.empty(Y32.shape[0] * T, ops.round_up(C + d_txt, 4), dtype=torch.float32, device=Y32.device)
        ops.begin_step(e_txt_feeds_tc=isinstance(self.mmf, MMF_XAttn_Add) and not rank, x_cat=x_cat, d_txt=d_txt)
        try:
            return self._forward_step(r, t_hat, Y32, flags, check, defer, rank)
        finally:
            ops.end_step()

    def _forward_step(self, r, t_hat, Y32, flags, check, defer, rank):
        # the weight-space half of the rank form depends on parameters only: it runs on a side stream beside the TTF
        # forward (and, through autograd, its backward beside the TTF backward).  IMMTSF_SIDE_STREAM=0 keeps one stream.
        side = os.environ.get("IMMTSF_SIDE_STREAM", "1") != "0"
        if defer:
            W_p, b_p = self.ttf.final_proj()
            # E_txt_true = E_txt W_p^T + b_p has a NaN iff one of its three factors has one: W_p and b_p are checked beside the
            # weight-space work, E_txt by the rank form's data kernel while it reads it
            wts = self.mmf.rank_weights((W_p, b_p), side=side, flags=flags if check else None)
            E_txt, M_txt = self.ttf.forward_ragged(r, t_hat, defer=True)
            if side:
                self.mmf.wait_rank_weights(wts)
            Y_out = self.mmf.forward_flags(Y32, E_txt, M_txt, flags if check else False, final_proj=(W_p, b_p), rank_weights=wts)
            runtime.raise_on_flags(flags)
            return Y_out
        wts = self.mmf.rank_weights(None, side=side) if rank else None
        E_txt, M_txt = self.ttf.forward_ragged(r, t_hat)
        if rank and side:
            self.mmf.wait_rank_weights(wts)
        if check and not rank:
            # the broadcast view of T2V eval mode has B distinct rows: check those only
            ops.nan_check(E_txt[:, :1] if E_txt.stride(1) == 0 else E_txt, flags, ops.FLAG_E)
        if rank:  # (the rank form's data kernel checks E_txt while it reads it)
            Y_out = self.mmf.forward_flags(Y32, E_txt, M_txt, flags if check else False, rank_weights=wts)
        else:
            Y_out = self.mmf.forward_flags(Y32, E_txt, M_txt, flags)
        runtime.raise_on_flags(flags)
        return Y_out
