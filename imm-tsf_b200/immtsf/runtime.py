"""Small runtime pieces shared by the fusion modules: the deferred NaN guard
and dropout seeding.

The reference raises ValueError on NaN at five places (fusions/FusionModel.py:
103-112, fusions/TTF_*.py:75/116, fusions/MMF_XAttn_Add.py:84-91), each one a
device->host sync.  Here every kernel that already touches the data sets a
device flag and the flags are read ONCE per forward (one sync), keeping the
ValueError convention.  IMMTSF_NAN_CHECK=0 (or CUDA-graph capture) skips the
read entirely."""
from __future__ import annotations

import os

import torch

from . import ops

_MESSAGES = {
    ops.FLAG_V: "Input embeddings V contain NaN values.",
    ops.FLAG_Y: "Y_ts contains NaN values.",
    ops.FLAG_E: "E_txt contains NaN values.",
    ops.FLAG_OUT: "Y_out contains NaN values.",
}


def nan_flags_enabled() -> bool:
    """Whether the flag-setting check kernels are launched at all (IMMTSF_NAN_CHECK=0 removes them)."""
    return os.environ.get("IMMTSF_NAN_CHECK", "1") != "0"


def nan_check_enabled() -> bool:
    """Whether the flags are read back (a device->host sync) at the end of forward; never inside a graph capture."""
    return nan_flags_enabled() and not torch.cuda.is_current_stream_capturing()


def new_flags(device) -> torch.Tensor:
    return torch.zeros(4, dtype=torch.int32, device=device)


def raise_on_flags(flags: torch.Tensor, slots=(ops.FLAG_Y, ops.FLAG_V, ops.FLAG_E, ops.FLAG_OUT)):
    """One device->host read; raises the reference's ValueError for the first set flag."""
    if not nan_check_enabled():
        return
    host = flags.tolist()
    for s in slots:
        if host[s]:
            raise ValueError(_MESSAGES[s])


class SeedSource:
    """Dropout seeds.  Default: drawn from torch's CPU generator per call.  A fixed
    seed can be pinned (tests, CUDA graphs replaying one captured step)."""

    def __init__(self):
        self.fixed = None

    def next(self) -> int:
        return self.fixed if self.fixed is not None else ops.new_seed()


SEEDS = SeedSource()


_CAPTURE = {}
_NVLS = {}


def nvls_comm(group):
    """The symmetric-memory arena + in-switch all-reduce of a process group (immtsf/nvls.py), created once; None when the
    fabric has no multicast (NCCL then carries the statistics)."""
    from . import nvls

    key = getattr(group, "group_name", id(group))
    if key not in _NVLS:
        _NVLS[key] = nvls.NvlsComm.create(group)
    return _NVLS[key]



def capture_stream(dev) -> "torch.cuda.Stream":
    dev = torch.device(dev)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _CAPTURE.get(key)
    if st is None:
        # HIGH priority: the captured kernel nodes inherit it, so the data chain of the step (which runs on this stream) gets SM
        # slots before the weight-space work on the side lanes (default priority) whenever both have blocks pending -- on the
        # cfg2 timeline a one-block kernel of the data chain waited 16 us behind a lane's 300-CTA skinny product
        st = _CAPTURE[key] = torch.cuda.Stream(device=dev, priority=int(os.environ.get("IMMTSF_CAPTURE_PRIORITY", "-1")))
    return st


class GraphedStep:
    """One training step of the fusion path -- FusionModel forward, loss, backward (every parameter gradient and
    dY_ts) -- captured ONCE in a CUDA graph and replayed with a single launch.

    At Time-IMM batch sizes a step is ~140 kernel launches of a few microseconds each, so the eager path is bound
    by host launch overhead, not by the GPU.  Capture removes that.  What stays dynamic: the ragged content (the
    number of valid notes per sample lives on the device), the parameter values, and the dropout masks (the
    by-value seeds are frozen in the graph, so the graph itself bumps a device-resident seed offset that every
    kernel adds -- include/immtsf.h immtsf_set_seed_offset_ptr).  What is static: tensor shapes; build one
    GraphedStep per (B, N_max, T) shape.

        step = GraphedStep(fusion, example=(notes, tau, t_hat, Y_ts), loss_fn=lambda out, tgt, m: ..., extras=(tgt, m))
        loss = step(notes, tau, t_hat, Y_ts, tgt, m)   # host (pinned) or device tensors; returns the static loss
        step.dY_ts, step.Y_out, [p.grad for p in fusion.parameters()]   # refreshed by every call

    The reference's NaN ValueErrors need a device->host read, which a graph cannot contain: call check_nan() after a
    step when that guard is wanted (one sync)."""

    def __init__(self, fusion, example, loss_fn=None, extras=(), warmup: int = 3, flat_grads: bool = False,
                 allreduce_group=None):
        """allreduce_group: a torch.distributed process group (or True for the default group).  The data-parallel
        gradient all-reduce is then captured INSIDE the graph.  With the rank form of MMF_XAttn_Add the
        MMF weight gradients (and the folded TTF projection's) are not communicated at all: their three small upstream
        tensors are all-reduced in the middle of backward and the weight-space backward runs on the reduced values.
        Implies flat_grads."""
        from . import _lib

        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device -- the immtsf path has no CPU fallback")
        self.fusion = fusion
        self.loss_fn = loss_fn if loss_fn is not None else (lambda out, *_: out.square().mean())
        dev = next(fusion.parameters()).device
        self.static_in = [t.detach().to(dev, copy=True) for t in example]
        self.static_in[3].requires_grad_(True)
        self.static_extra = [t.detach().to(dev, copy=True) for t in extras]
        self.params = [p for p in fusion.parameters() if p.requires_grad]
        self.group = None
        if allreduce_group is not None and allreduce_group is not False:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.group = dist.group.WORLD if allreduce_group is True else allreduce_group
                if dist.get_world_size(self.group) == 1:
                    self.group = None
        # in-graph data parallelism reduces the gradient tensors where autograd put them, as ONE coalesced NCCL launch
        # (no flat bucket: its zeroing and its accumulate kernel per parameter cost 31 us per cfg2 step); IMMTSF_DP_FLAT=1
        # brings the bucket back
        self.coalesced = self.group is not None and not flat_grads and os.environ.get("IMMTSF_DP_FLAT", "0") != "1"
        flat_grads = flat_grads or (self.group is not None and not self.coalesced)
        # In-graph data parallelism: parameters whose gradients come out of the weight-space backward of the rank form
        # (functional.XAttnRankWeightsFn) are functions of three small upstream tensors; those are all-reduced instead
        # (28 KB instead of 14 MB at cfg2), so these parameters' gradients are born reduced and sit in front of the flat
        # bucket, outside the all-reduce that closes the step.
        pre = []
        if self.group is not None and hasattr(fusion, "dp_prereduced_params"):
            pre = fusion.dp_prereduced_params(self.static_in[2].shape[-1])
        pre_ids = self._pre_ids = {id(p) for p in pre}
        self.n_total = sum(p.numel() for p in self.params)
        self.params.sort(key=lambda p: 0 if id(p) in pre_ids else 1)
        self.n_first = sum(p.numel() for p in self.params if id(p) in pre_ids)
        self.seed_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        self._lib = _lib
        _lib.call("immtsf_set_seed_offset_ptr", self.seed_offset.data_ptr())
        # ONE capture stream per device for every GraphedStep of the process: the warm-up runs on it, so the per-stream
        # GEMM workspaces (ops._workspace is keyed by stream) exist before the capture begins and are shared by all graphs
        side = capture_stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        ops.DP_GROUP = self.group if self.n_first > 0 else None
        ops.DP_NVLS = nvls_comm(self.group) if ops.DP_GROUP is not None else None
        with torch.cuda.stream(side):  # lazy allocations (workspaces, side streams) happen here, outside the graph's pool
            for _ in range(max(warmup, 1)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # flat_grads: every parameter gradient is a view into ONE buffer, zeroed inside the graph and accumulated
        # in place by autograd, so the data-parallel all-reduce (dp.allreduce_grads(flat=...)) needs no gather copy
        self.flat_grads = None
        if flat_grads:
            total = sum(p.numel() for p in self.params)
            self.flat_grads = torch.zeros(total, dtype=torch.float32, device=dev)
            off = 0
            for p in self.params:
                p.grad = self.flat_grads[off:off + p.numel()].view_as(p)
                off += p.numel()
        else:
            for p in self.params:
                p.grad = None
        self.static_in[3].grad = None
        self.graph = torch.cuda.CUDAGraph()
        self._works = []
        self._stage = self._copy_stream = self._staged = self._consumed = None
        try:
            # NCCL's watchdog thread may touch CUDA while we capture: thread-local capture mode tolerates that
            mode = dict(capture_error_mode="thread_local") if self.group is not None else {}
            ops.DP_STATS["floats"] = ops.DP_STATS["calls"] = ops.DP_STATS["nvls"] = 0
            with torch.cuda.graph(self.graph, stream=side, **mode):
                _lib.call("immtsf_seed_advance", self.seed_offset.data_ptr(), 1, ops._stream())
                if self.flat_grads is not None:
                    self.flat_grads.zero_()
                out, loss = self._eager()
                if self.group is not None:
                    self._reduce_rest()
        finally:
            ops.DP_GROUP = None
            self.nvls, ops.DP_NVLS = ops.DP_NVLS, None
        # what one replay sends through the all-reduces captured inside the Functions (floats, collectives)
        self.dp_floats_per_step, self.dp_calls_per_step = ops.DP_STATS["floats"], ops.DP_STATS["calls"]
        self.dp_nvls_calls_per_step = ops.DP_STATS.get("nvls", 0)
        self.Y_out, self.loss = out.detach(), loss.detach()
        self.flags = getattr(fusion, "_last_flags", None)

    def _reduce_rest(self):
        """The all-reduce that closes the step: every gradient that was not born reduced."""
        import torch.distributed as dist

        if self.coalesced:
            todo = [p.grad for p in self.params if id(p) not in self._pre_ids and p.grad is not None]
            if todo:
                ops.DP_STATS["floats"] += sum(g.numel() for g in todo)
                ops.DP_STATS["calls"] += 1
                with dist._coalescing_manager(group=self.group, device=todo[0].device):
                    for g in todo:
                        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        elif self.n_first < self.flat_grads.numel():
            ops.DP_STATS["floats"] += self.flat_grads.numel() - self.n_first
            ops.DP_STATS["calls"] += 1
            dist.all_reduce(self.flat_grads[self.n_first:], op=dist.ReduceOp.SUM, group=self.group)

    def _eager(self):
        fwd = getattr(self.fusion, "_forward_eager", self.fusion)  # (never through the transparent graph cache)
        out = fwd(self.static_in[0], self.static_in[1], self.static_in[2], self.static_in[3])
        loss = self.loss_fn(out, *self.static_extra)
        loss.backward()
        return out, loss

    @property
    def dY_ts(self):
        return self.static_in[3].grad

    def prefetch(self, notes, tau, t_hat, Y_ts, *extras):
        """Start the host->device copy of the NEXT step's inputs (pinned host tensors) on a copy stream into staging
        buffers, so that it overlaps the step that is running now; step_prefetched() then consumes them.  This is the
        double-buffered input pipeline of a training loop: the graph's static inputs cannot be overwritten while a
        replay may still read them, so the DMA targets a second set of buffers."""
        if self._stage is None:
            self._stage = [torch.empty_like(t) for t in self.static_in + self.static_extra]
            self._copy_stream = torch.cuda.Stream()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        cs = self._copy_stream
        cs.wait_event(self._consumed)  # the previous staged batch has been moved into the static inputs
        with torch.cuda.stream(cs), torch.no_grad():
            for s, t in zip(self._stage, (notes, tau, t_hat, Y_ts) + tuple(extras)):
                s.copy_(t, non_blocking=True)
            self._staged = torch.cuda.Event(enable_timing=True)  # one per batch: callers may keep it to time the copy
            self._staged.record(cs)

    def prefetch_done(self) -> bool:
        """True when the last prefetch() has fully landed on the device (no host wait)."""
        return self._staged is not None and self._staged.query()

    def step_prefetched(self):
        """One step on the inputs of the last prefetch(): device-to-device move into the static inputs + one replay."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        with torch.no_grad():
            for s, t in zip(self.static_in + self.static_extra, self._stage):
                s.copy_(t, non_blocking=True)
        self._consumed.record(cur)
        self.graph.replay()
        return self.loss

    def __call__(self, notes, tau, t_hat, Y_ts, *extras):
        with torch.no_grad():
            for s, t in zip(self.static_in, (notes, tau, t_hat, Y_ts)):
                if t is not s:
                    s.copy_(t, non_blocking=True)
            for s, t in zip(self.static_extra, extras):
                if t is not s:
                    s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.loss

    def check_comm(self):
        """Raise if a rank never arrived at a barrier of the in-switch all-reduce (one device->host read)."""
        if getattr(self, "nvls", None) is not None:
            self.nvls.check()

    def check_nan(self):
        """The reference's ValueErrors (fusions/FusionModel.py:103-112) for the last replay; one device->host read."""
        if self.flags is None:
            return
        host = self.flags.tolist()
        for s in (ops.FLAG_Y, ops.FLAG_V, ops.FLAG_E, ops.FLAG_OUT):
            if host[s]:
                raise ValueError(_MESSAGES[s])

    def close(self):
        self._lib.call("immtsf_set_seed_offset_ptr", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
