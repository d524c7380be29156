"""Batch-sharded data parallelism for the fusion path.

Every sample is independent through all four modules, so forward / eval needs no
communication.  Training needs exactly one collective per step: a SUM all-reduce
of the gradients (NCCL over NVLink on GPUs, gloo in the CPU tests), issued on
one flat bucket so launch latency is paid once.  To keep the reference's loss
(per-variable masked MSE averaged over the variables that have data,
lib/evaluation.py:17-69) EXACT under sharding, the per-variable numerator and
count are all-reduced in the forward (2*C floats) and each rank back-propagates
its share; the gradient all-reduce is then a plain SUM (no division by world
size)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous split of n samples: the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank: int, world: int):
    """Slice every [B, ...] tensor (1-D t_hat is shared and passed through)."""
    B = tensors[0].shape[0]
    lo, hi = shard_bounds(B, rank, world)
    return [t if (t.dim() == 1 and t.shape[0] != B) else t[lo:hi] for t in tensors]


def allreduce_grads(params: Iterable[torch.nn.Parameter], group=None, average: bool = False,
                    flat: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """One SUM all-reduce over a flat bucket of every existing .grad; results are copied back in place.
    flat: a buffer the gradients already live in as views (runtime.GraphedStep(flat_grads=True)) -- reduced in
    place, no gather / scatter copies."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    if flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
        return flat
    ps: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
    for p in ps:
        if p.grad is None:  # a rank whose shard produced no gradient still has to take part
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat


def masked_mse_exact(pred: torch.Tensor, truth: torch.Tensor, mask: torch.Tensor, group=None) -> torch.Tensor:
    """Reference loss (lib/evaluation.py compute_error(..., 'MSE', 'mean')) made exact under batch sharding.
    pred/truth/mask: [B_local, T, C].  Returns this rank's SHARE of the global loss: summing the
    returned values over ranks gives the single-process loss, and so do the gradients after a SUM all-reduce."""
    if pred.is_cuda:  # the fused kernels (immtsf/loss.py); same value up to the reference's 1e-8 in the denominator
        from . import loss as _loss

        return _loss.masked_mse(pred, truth, mask, group=group)
    C = pred.shape[-1]
    err = ((pred - truth) ** 2) * mask
    numer = err.reshape(-1, C).sum(0)  # differentiable, local
    count = mask.reshape(-1, C).sum(0).detach().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    has = count > 0
    nvars = has.sum().clamp_min(1)
    per_var = torch.where(has, numer / count.clamp_min(1), torch.zeros_like(numer))
    return per_var.sum() / nvars
