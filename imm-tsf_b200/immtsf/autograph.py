"""Transparent CUDA-graph replay behind the plugin surface: `fusion(notes, tau, t_hat, Y_ts)` ... `loss.backward()` as the
reference's own loop calls them (lib/evaluation.py:95-100, main.py:1097), at graph speed.

The eager path issues ~110 kernels of a few microseconds per cfg2 step from Python (one host call each): the GPU waits for
the host.  runtime.GraphedStep removes that for a bespoke fixed-shape loop that owns the loss; an UNMODIFIED caller computes
its own loss between forward and backward and feeds batches whose padded shapes vary.  So, per input shape, this module
captures TWO graphs once -- the fusion forward, and the backward of the autograd graph that forward built (the same
mechanism as torch.cuda.make_graphed_callables, with the path's own needs: device-resident dropout seed offset advanced
inside the forward graph, NaN flags read after the replay, side lanes captured as parallel branches) -- and wraps the pair
in one autograd.Function:

    fusion.enable_graphs()            # or IMMTSF_AUTOGRAPH=1 (tools/run_with_immtsf.py sets it)
    Y_out = fusion(notes, tau, t_hat, Y_ts)      # copies into static inputs, replays the forward graph
    loss_fn(Y_out).backward()                    # copies dY_out into a static buffer, replays the backward graph

Padded note counts are bucketed (N_max rounded up to a multiple of 8 with zero rows, which the content mask ignores: results
unchanged) so that a handful of graphs serves a whole epoch; the cache is LRU-bounded.  Constraints, as for any CUDA graph:
the returned Y_out and the gradients alias static buffers that the next call with the same shape overwrites, and the backward
of a call must run before the next graphed forward (it regenerates that forward's dropout masks from the seed offset)."""
from __future__ import annotations

import collections
import os

import torch

from . import _lib, ops, runtime

N_BUCKET = 8
MAX_GRAPHS = int(os.environ.get("IMMTSF_AUTOGRAPH_MAX", "24"))

_SEED_OFFSET = {}


def _seed_offset(dev) -> torch.Tensor:
    key = torch.device(dev).index
    t = _SEED_OFFSET.get(key)
    if t is None:
        t = _SEED_OFFSET[key] = torch.zeros(1, dtype=torch.int64, device=dev)
    return t


class _Pair:
    """Forward + backward graph of one (shapes, mode) key."""

    def __init__(self, fusion, inputs, need_grad: bool):
        dev = inputs[0].device
        self.fusion, self.need_grad = fusion, need_grad
        self.static_in = [t.detach().clone() for t in inputs]
        self.params = [p for p in fusion.parameters() if p.requires_grad] if need_grad else []
        self.seed_offset = _seed_offset(dev)
        st = runtime.capture_stream(dev)
        st.wait_stream(torch.cuda.current_stream())
        _lib.call("immtsf_set_seed_offset_ptr", self.seed_offset.data_ptr())
        anomaly = torch.is_anomaly_enabled()
        torch.autograd.set_detect_anomaly(False)  # (main.py:1079 trains under anomaly mode: its NaN probes are host syncs)
        try:
            with torch.cuda.stream(st):
                for _ in range(2):  # warm-up on the capture stream: lazy allocations (workspaces, lanes) happen here
                    self._run_eager()
            torch.cuda.current_stream().wait_stream(st)
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            self.fwd, self.bwd = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(self.fwd, pool=pool, stream=st):
                _lib.call("immtsf_seed_advance", self.seed_offset.data_ptr(), 1, ops._stream())
                self.y_req = self.static_in[3].requires_grad_(need_grad)
                with torch.set_grad_enabled(need_grad):
                    self.out = fusion._forward_eager(self.static_in[0], self.static_in[1], self.static_in[2], self.y_req)
            self.flags = getattr(fusion, "_last_flags", None)
            if need_grad:
                self.gout = torch.empty_like(self.out)
                self.bwd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.bwd, pool=pool, stream=st):
                    self.grads = torch.autograd.grad(self.out, [self.y_req] + self.params, self.gout, allow_unused=True)
        finally:
            torch.autograd.set_detect_anomaly(anomaly)
            _lib.call("immtsf_set_seed_offset_ptr", None)

    def _run_eager(self):
        y = self.static_in[3].detach().requires_grad_(self.need_grad)
        with torch.set_grad_enabled(self.need_grad):
            out = self.fusion._forward_eager(self.static_in[0], self.static_in[1], self.static_in[2], y)
        if self.need_grad:
            torch.autograd.grad(out, [y] + self.params, torch.ones_like(out), allow_unused=True)

    @torch.no_grad()
    def load(self, inputs):
        for s, t in zip(self.static_in, inputs):
            if t.shape == s.shape:
                s.copy_(t, non_blocking=True)
            else:  # notes / tau with fewer padded rows than the bucket: zero rows are not notes
                s.zero_()
                s[:, : t.shape[1]].copy_(t, non_blocking=True)


class _GraphedFusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pair, Y_ts, *params):
        _lib.call("immtsf_set_seed_offset_ptr", pair.seed_offset.data_ptr())
        try:
            pair.fwd.replay()
        finally:
            _lib.call("immtsf_set_seed_offset_ptr", None)
        ctx.pair = pair
        return pair.out.detach()

    @staticmethod
    def backward(ctx, gout):
        pair = ctx.pair
        pair.gout.copy_(gout, non_blocking=True)
        pair.bwd.replay()
        return (None,) + tuple(g.detach() if g is not None else None for g in pair.grads)


class GraphCache:
    def __init__(self, fusion):
        self.fusion = fusion
        self.pairs = collections.OrderedDict()
        self.captures = 0

    def __call__(self, notes, tau, t_hat, Y_ts):
        fm = self.fusion
        need_grad = torch.is_grad_enabled() and (Y_ts.requires_grad or any(p.requires_grad for p in fm.parameters()))
        B, N, dm = notes.shape
        Nb = max((N + N_BUCKET - 1) // N_BUCKET * N_BUCKET, N_BUCKET)
        key = (B, Nb, dm, tuple(t_hat.shape), tuple(Y_ts.shape), fm.training, need_grad, notes.dtype, Y_ts.dtype)
        pair = self.pairs.get(key)
        inputs = (notes, tau, t_hat, Y_ts.detach())
        if pair is None:
            ex = [torch.zeros(B, Nb, dm, dtype=notes.dtype, device=notes.device), torch.zeros(B, Nb, dtype=tau.dtype, device=tau.device),
                  t_hat, Y_ts.detach()]
            ex[0][:, :N].copy_(notes)
            ex[1][:, :N].copy_(tau)
            pair = self.pairs[key] = _Pair(fm, ex, need_grad)
            self.captures += 1
            while len(self.pairs) > MAX_GRAPHS:
                self.pairs.popitem(last=False)
        else:
            self.pairs.move_to_end(key)
        pair.load(inputs)
        if need_grad:
            out = _GraphedFusionFn.apply(pair, Y_ts, *pair.params)
        else:
            _lib.call("immtsf_set_seed_offset_ptr", pair.seed_offset.data_ptr())
            try:
                pair.fwd.replay()
            finally:
                _lib.call("immtsf_set_seed_offset_ptr", None)
            out = pair.out.detach()
        if pair.flags is not None and runtime.nan_flags_enabled() and os.environ.get("IMMTSF_AUTOGRAPH_CHECK", "1") != "0":
            runtime.raise_on_flags(pair.flags)  # the reference's ValueErrors: ONE device->host read per forward (it does five)
        return out
