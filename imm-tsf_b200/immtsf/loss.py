"""Fused masked-MSE training loss: lib/evaluation.py:17-69 `compute_error(truth, pred, mask,
"MSE", "mean")` as compute_all_losses calls it (:107-113), with the all-zero-mask check of :128-132 as a device flag.

    loss = masked_mse(pred, truth, mask)                      # pred/truth/mask [B, T, C] on the GPU
    loss = masked_mse(pred, truth, mask, group=dist.group.WORLD)   # batch-sharded: this rank's SHARE of the global loss

Three launches forward (partial sums with an in-kernel ordered reduction, [all-reduce of the C counts], finalize) and one
backward, instead of the reference's ~12 elementwise / reduction launches and B host syncs.  Deterministic.  Under
sharding only the per-variable COUNTS are all-reduced (C floats): every rank then back-propagates its share of the exact
single-process loss, so the gradient all-reduce stays a plain SUM (immtsf/dp.py)."""
from __future__ import annotations

import torch

from . import _lib, ops

_TICKET = {}


def _ticket(dev) -> torch.Tensor:
    t = _TICKET.get(dev)
    if t is None:
        t = _TICKET[dev] = torch.zeros(1, dtype=torch.int32, device=dev)
    return t


class MaskedMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, truth, mask, group, empty_flag):
        for t, nm in ((pred, "pred"), (truth, "truth"), (mask, "mask")):
            ops._chk(t, nm)
        ctx.in_shape = pred.shape  # the gradient goes back in the caller's layout ([1, B, T, C] included)
        if pred.dim() == 4 and pred.shape[0] == 1:  # [n_traj_samples = 1, B, T, C] (lib/evaluation.py:21-23)
            pred = pred[0]
        if pred.dim() != 3 or truth.shape != pred.shape or mask.shape != pred.shape:
            raise ValueError(f"masked_mse: pred/truth/mask must be [B, T, C]; got {tuple(pred.shape)}, {tuple(truth.shape)}, {tuple(mask.shape)}")
        B, T, C = pred.shape
        p, t, m = pred.contiguous(), truth.contiguous(), mask.contiguous()
        dev = p.device
        lib = _lib.load()
        ws = ops._workspace(dev, lib.immtsf_masked_mse_workspace_bytes(C))
        err_cnt = torch.empty(2 * C, dtype=torch.float32, device=dev)
        _lib.call("immtsf_masked_mse_partial", ops._p(p), ops._p(t), ops._p(m), B * T, T, C, ops._p(err_cnt), ops._p(empty_flag),
                  ops._p(_ticket(dev)), ops._p(ws), ws.numel(), ops._stream())
        cnt = err_cnt[C:]
        if group is not None:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
                cnt = cnt.clone()
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scale = torch.empty(C, dtype=torch.float32, device=dev)
        _lib.call("immtsf_masked_mse_finalize", ops._p(err_cnt), ops._p(cnt), C, ops._p(loss), ops._p(scale), ops._stream())
        ctx.save_for_backward(p, t, m, scale)
        ctx.shape = pred.shape
        return loss

    @staticmethod
    def backward(ctx, gloss):
        p, t, m, scale = ctx.saved_tensors
        B, T, C = ctx.shape
        g = gloss.contiguous().to(torch.float32)
        dpred = torch.empty_like(p)
        _lib.call("immtsf_masked_mse_bwd", ops._p(p), ops._p(t), ops._p(m), B * T, C, ops._p(scale), ops._p(g), ops._p(dpred), ops._stream())
        return dpred.view(ctx.in_shape), None, None, None, None


def masked_mse(pred: torch.Tensor, truth: torch.Tensor, mask: torch.Tensor, group=None, empty_flag: torch.Tensor = None):
    """Reference loss (lib/evaluation.py:17-69, "MSE", "mean").  empty_flag: optional int32[1] device tensor that is set
    to 1 when some sample's mask is all zero (the reference raises ValueError there, :128-132; read it when convenient)."""
    return MaskedMSEFn.apply(pred, truth, mask, group, empty_flag)
