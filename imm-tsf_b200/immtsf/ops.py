"""Tensor-level wrappers over the C ABI: argument checking, pointer extraction,
current-stream plumbing.  PyTorch is used only for device memory and streams."""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

BACKEND_AUTO, BACKEND_FFMA, BACKEND_TC = 0, 1, 2
SITE_TTF_DROPOUT, SITE_TTF_ATTN, SITE_MMF_DROPOUT, SITE_MMF_ATTN = 1, 2, 3, 4
FLAG_V, FLAG_Y, FLAG_E, FLAG_OUT = 0, 1, 2, 3
LN_EPS = 1e-5

# bench.py sets this to a list to collect (label, flops, start_event, end_event) per GEMM launch
PROFILE = None


def gemm_backend() -> int:
    v = os.environ.get("IMMTSF_GEMM", "auto").lower()
    return {"auto": BACKEND_AUTO, "ffma": BACKEND_FFMA, "tc": BACKEND_TC}.get(v, BACKEND_AUTO)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, name: str, dtype=torch.float32):
    if not t.is_cuda:
        raise _lib.ImmtsfError(f"{name}: tensor must live on a CUDA device (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.ImmtsfError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t


def drop_thr(p: float) -> int:
    """floor(p * 2^16) clamped to 16 bits (csrc/common.cuh: 16 random bits per element); 0 disables dropout."""
    if p <= 0.0:
        return 0
    return min(int(p * 65536.0), 65535)


def new_seed() -> int:
    """63-bit seed drawn from torch's CPU generator (so torch.manual_seed governs it)."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# ------------------------------------------------------------------ CSR layout
@dataclass
class RaggedNotes:
    """Ragged (CSR) view of a padded note batch.  sumN = offsets[B] stays on the device."""
    B: int
    N: int
    d_m: int
    M_alloc: int
    note_mask: torch.Tensor  # [B*N] u8
    offsets: torch.Tensor  # [B+1] i32
    rows: torch.Tensor  # [B*N] i32
    seg: torch.Tensor  # [B*N] i32
    emb_flat: torch.Tensor  # [M_alloc, d_m] (a view of emb_wide's left columns when the consumer asked for a wider buffer)
    tau_flat: torch.Tensor  # [M_alloc]
    m_txt: torch.Tensor  # [B] u8
    flags: torch.Tensor  # [4] i32
    emb_wide: Optional[torch.Tensor] = None  # [M_alloc, d_m + pad_cols]: compacted rows + room for the consumer's extra columns
    emb_wide_lo: Optional[torch.Tensor] = None  # its tcgen05 lo operand (the left d_m columns are written by the gather)

    @property
    def m_dev(self) -> torch.Tensor:
        return self.offsets[self.B:]


def csr_build(notes: torch.Tensor, tau: torch.Tensor, flags: Optional[torch.Tensor] = None, pad_cols: int = 0) -> RaggedNotes:
    """pad_cols > 0: the compacted rows are written into the left columns of a [M_alloc, d_m + pad_cols] buffer (emb_wide) together
    with their tcgen05 lo operand -- the collapsed T2V schedule appends the Time2Vec features there and uses the buffer as the
    operand of its first product, so the notes are neither copied nor split again."""
    _chk(notes, "notes"), _chk(tau, "tau")
    if notes.dim() != 3 or tau.dim() != 2 or tau.shape[0] != notes.shape[0] or tau.shape[1] != notes.shape[1]:
        raise ValueError(f"notes must be [B,N,d_model] and tau [B,N]; got {tuple(notes.shape)} and {tuple(tau.shape)}")
    notes, tau = notes.contiguous(), tau.contiguous()
    B, N, d_m = notes.shape
    dev = notes.device
    M_alloc = max(round_up(B * N, 128), 128)
    i32, u8 = torch.int32, torch.uint8
    wide = wide_lo = None
    if pad_cols > 0 and d_m > 0 and d_m % 4 == 0:
        wide = torch.empty(M_alloc, d_m + pad_cols, dtype=torch.float32, device=dev)
        wide_lo = torch.empty(M_alloc, round_up(d_m + pad_cols, 4), dtype=torch.float32, device=dev)
        emb = wide[:, :d_m]
    else:
        emb = torch.empty(M_alloc, max(d_m, 1), dtype=torch.float32, device=dev)
    r = RaggedNotes(
        B, N, d_m, M_alloc,
        torch.empty(max(B * N, 1), dtype=u8, device=dev), torch.empty(B + 1, dtype=i32, device=dev),
        torch.empty(max(B * N, 1), dtype=i32, device=dev), torch.empty(max(B * N, 1), dtype=i32, device=dev),
        emb,
        torch.empty(M_alloc, dtype=torch.float32, device=dev), torch.empty(max(B, 1), dtype=u8, device=dev),
        flags if flags is not None else torch.zeros(4, dtype=i32, device=dev), wide, wide_lo,
    )
    _lib.call("immtsf_csr_build_ex", _p(notes), _p(tau), B, N, d_m, _p(r.note_mask), _p(r.offsets), _p(r.rows), _p(r.seg),
              _p(r.emb_flat), r.emb_flat.stride(0), _p(wide_lo), wide_lo.stride(0) if wide_lo is not None else 0, _p(r.tau_flat),
              _p(r.m_txt), _p(r.flags), M_alloc, _stream())
    return r


def nan_check(x: torch.Tensor, flags: torch.Tensor, slot: int):
    x = _chk(x, "nan_check").contiguous()
    _lib.call("immtsf_nan_check", _p(x), x.numel(), _p(flags), slot, _stream())


_TICKET = {}


def ticket(dev) -> torch.Tensor:
    """One zero-initialised device counter per device for the kernels that end with a last-CTA ordered reduction (they leave
    it at zero).  Launches that use it are ordered on a stream or run in different steps."""
    key = torch.device(dev).index
    t = _TICKET.get(key)
    if t is None:
        t = _TICKET[key] = torch.zeros(4, dtype=torch.int32, device=dev)
    return t


_ZEROS = {}


def zeros_cached(dev, rows: int, cols: int) -> torch.Tensor:
    """A read-only block of zeros (source of pad columns in immtsf_multi_split copies); grown on demand, never written."""
    key = (torch.device(dev).index, cols)
    z = _ZEROS.get(key)
    if z is None or z.shape[0] < rows:
        z = _ZEROS[key] = torch.zeros(rows, cols, dtype=torch.float32, device=dev)
    return z


def zero_pad_rows(X: torch.Tensor, ncols: int, m_dev: torch.Tensor, M_alloc: int):
    _lib.call("immtsf_zero_pad_rows", _p(X), X.stride(0), ncols, _p(m_dev), M_alloc, _stream())


# ------------------------------------------------------------------ GEMM
_WS = {}


def _workspace(dev, nbytes: int) -> torch.Tensor:
    """Per-device scratch for the tcgen05 backend's operand split.  Grows on demand; reuse across GEMMs is
    safe because every launch is ordered on the current stream."""
    key = (dev, torch.cuda.current_stream().cuda_stream)  # kernels on different streams may run concurrently
    w = _WS.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(nbytes, (64 << 20) if len(_WS) == 0 else (16 << 20)), dtype=torch.uint8, device=dev)
        _WS[key] = w
    return w


def _mat(t: torch.Tensor, name: str):
    _chk(t, name)
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.ImmtsfError(f"{name}: need a 2-D row-major matrix (unit column stride), got strides {t.stride()}")
    return t


class LoCache(dict):
    """Per-step cache of tcgen05 operand splits (lo = x - trunc_tf32(x)), keyed by the operand view.  Nearly every
    operand of the path is used by two products (forward + weight gradient, or dgrad + wgrad); with a cache it is
    split once.  Entries keep the source tensor alive so that a recycled allocation can never alias a stale entry."""

    def lo_for(self, t: torch.Tensor, ragged: Optional[torch.Tensor]):
        key = (t.data_ptr(), tuple(t.shape), t.stride(0), ragged is not None)
        e = self.get(key)
        if e is None:
            rows, cols = t.shape
            lo = torch.empty(rows, round_up(cols, 4), dtype=torch.float32, device=t.device)
            _lib.call("immtsf_split_lo", _p(t), t.stride(0), rows, cols, _p(lo), lo.stride(0), _p(ragged), _stream())
            e = (t, lo)
            self[key] = e
        return e[1]

    def transposed(self, w: torch.Tensor) -> torch.Tensor:
        """w^T (contiguous) with its lo registered, made once per step per weight (immtsf_transpose_split): the data-gradient
        products read it as a K-major operand."""
        key = (w.data_ptr(), tuple(w.shape), w.stride(0), "T")
        e = self.get(key)
        if e is None:
            rows, cols = w.shape
            wT = torch.empty(cols, rows, dtype=torch.float32, device=w.device)
            wT_lo = torch.empty(cols, round_up(rows, 4), dtype=torch.float32, device=w.device)
            _lib.call("immtsf_transpose_split", _p(w), w.stride(0), rows, cols, _p(wT), wT.stride(0), _p(wT_lo), wT_lo.stride(0), _stream())
            self.put(wT, wT_lo)
            e = (w, wT)
            self[key] = e
        return e[1]

    def put(self, t: torch.Tensor, lo: torch.Tensor):
        """Register a lo that a producer kernel wrote together with t (GEMM epilogue, immtsf_gemm_ex C_lo).  Valid
        for ragged and non-ragged consumers alike: rows past the ragged bound are never live in a product."""
        for rg in (False, True):
            self[(t.data_ptr(), tuple(t.shape), t.stride(0), rg)] = (t, lo)


class StepCtx:
    """What TTF and MMF share within one FusionModel step: the operand-split cache (so that E_txt / dE_txt, produced by
    one module's GEMM epilogue together with their lo, are not split again by the other) and whether the MMF module
    consumes E_txt with a tcgen05 product at all."""

    def __init__(self, lo=None, e_txt_feeds_tc=False):
        self.lo = lo if lo is not None else LoCache()
        self.e_txt_feeds_tc = e_txt_feeds_tc
        # MMF_GR_Add reads x = [E ; Y] (MMF_GR_Add.py:43): FusionModel allocates that buffer BEFORE the TTF forward and the
        # TTF's final projection writes E_txt straight into its left columns (e_out_rows x d view, leading dimension = the
        # buffer's), so E_txt is never copied.  x_cat is the whole buffer.
        self.x_cat = None
        self.e_out = None

    def final_out(self, rows: int, d: int):
        """Destination of a TTF's final projection when the consumer wants E_txt in place (else None)."""
        e = self.e_out
        return e if e is not None and tuple(e.shape) == (rows, d) else None


_STEP: Optional[StepCtx] = None


def begin_step(e_txt_feeds_tc: bool, x_cat=None, d_txt: int = 0) -> StepCtx:
    global _STEP
    if DP_NVLS is not None:
        DP_NVLS.reset()  # the statistics buffers of a step sit at the same arena offsets in every step
    _STEP = StepCtx(e_txt_feeds_tc=e_txt_feeds_tc)
    if x_cat is not None:
        _STEP.x_cat, _STEP.e_out = x_cat, x_cat[:, :d_txt]
    return _STEP


def end_step():
    global _STEP
    _STEP = None


_SIDE = {}
# process group for in-graph data parallelism (runtime.GraphedStep(allreduce_group=...)): while set, the weight-space
# backward of the rank form all-reduces its upstream gradients and so produces already-reduced parameter gradients
DP_GROUP = None
RANK_PACK = None  # hand-off of the rank form's packed upstream gradients from the data Function to the weights Function
DP_STATS = {"floats": 0, "calls": 0}  # what dp_allreduce moved since it was last cleared (runtime.GraphedStep: per captured step)


def dp_allreduce(buf: torch.Tensor) -> torch.Tensor:
    """SUM all-reduce of `buf` over DP_GROUP on the current stream (no-op without a group).  The autograd Functions call it
    on the SUFFICIENT STATISTICS of their parameter gradients -- the packed outputs of the weight-gradient products before
    the weight-space un-folds, which are linear in them with replicated parameters -- so the parameter gradients are born
    reduced, fewer floats cross NVLink than the parameters have, and no collective is left for the end of the step."""
    if DP_GROUP is None:
        return buf
    DP_STATS["floats"] += buf.numel()
    DP_STATS["calls"] += 1
    if DP_NVLS is not None and DP_NVLS.owns(buf):
        DP_NVLS.all_reduce(buf)  # hand-written in-switch all-reduce (csrc/nvls.cu)
        DP_STATS["nvls"] = DP_STATS.get("nvls", 0) + 1
        return buf
    import torch.distributed as dist

    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=DP_GROUP)
    return buf


DP_NVLS = None  # immtsf.nvls.NvlsComm of DP_GROUP when the fabric has NVSwitch multicast


def dp_pack(n: int, device) -> torch.Tensor:
    """Buffer for `n` floats of gradient statistics: a piece of the symmetric arena when the in-switch all-reduce is
    available (the producing kernels then write straight into memory every peer has mapped), else plain device memory."""
    if DP_GROUP is not None and DP_NVLS is not None:
        t = DP_NVLS.alloc(n)
        if t is not None:
            return t
    return torch.empty(n, dtype=torch.float32, device=device)


def side_stream(dev) -> "torch.cuda.Stream":
    """Side stream for work that depends on parameters only (the weight-space half of the rank form of MMF_XAttn_Add):
    it runs beside the TTF forward, and autograd runs its backward there beside the TTF backward."""
    st = _SIDE.get(dev)
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _SIDE[dev] = st
    return st


def side_streams_enabled() -> bool:
    return os.environ.get("IMMTSF_SIDE_STREAM", "1") != "0"


class Fork:
    """Work that is off the critical path on side streams ("lanes").  In a backward pass weight and bias gradients are not
    needed until the end of the step, while the data-gradient chain (dgrad products, attention / LayerNorm backward) is the
    critical path and its ragged products leave a third of the SMs idle; in a forward pass everything that depends on
    parameters only (operand splits, weight-space folds, the query path) can run beside the data chain.  `run` enqueues a
    closure on a lane after everything issued so far on the current stream, `mark` / `wait` order the current stream after
    a point of a lane, `join` after all lanes.  Captured in a CUDA graph the lanes are parallel branches (the timeline of
    round 1's single wgrad lane showed 265 us of serial weight-gradient work after a 95 us data chain).  Operands read on
    several streams must have their lo split BEFORE the fork."""

    def __init__(self, device, enabled: bool = True, lanes: int = 1, name: str = "wgrad"):
        self.cur = torch.cuda.current_stream(device)
        self.lanes = []
        if enabled and side_streams_enabled():
            for i in range(lanes):
                key = (device, name, i)
                st = _SIDE.get(key)
                if st is None:
                    st = _SIDE[key] = torch.cuda.Stream(device=device)
                self.lanes.append(st)
        self.side = self.lanes[0] if self.lanes else None
        self._used = set()

    def run(self, fn, *reads, lane: int = 0, after_current: bool = True):
        if not self.lanes:
            return fn()
        st = self.lanes[lane % len(self.lanes)]
        if after_current or st not in self._used:
            st.wait_stream(self.cur)
        self._used.add(st)
        for t in reads:
            if t is not None:
                t.record_stream(st)
        with torch.cuda.stream(st):
            return fn()

    def mark(self, lane: int = 0):
        """An event at the current end of a lane (None when the fork is disabled)."""
        if not self.lanes:
            return None
        ev = torch.cuda.Event()
        ev.record(self.lanes[lane % len(self.lanes)])
        return ev

    def lane_wait(self, lane: int, ev):
        """Order lane `lane` after event `ev` (an event of another lane)."""
        if ev is not None and self.lanes:
            self.lanes[lane % len(self.lanes)].wait_event(ev)

    def wait(self, ev, *tensors):
        if ev is not None:
            self.cur.wait_event(ev)
            for t in tensors:
                if t is not None:
                    t.record_stream(self.cur)

    def join(self, *tensors):
        if not self.lanes:
            return
        for st in self.lanes:
            if st in self._used:
                self.cur.wait_stream(st)
        for t in tensors:
            if t is not None:
                t.record_stream(self.cur)


def step_ctx() -> StepCtx:
    """The step context opened by FusionModel.forward, or a private one for a module called on its own."""
    return _STEP if _STEP is not None else StepCtx()


def weight_los(lo: "LoCache", weights, extra_tasks=()):
    """lo operands of a module's weight matrices in ONE launch (immtsf_multi_split) instead of one split per matrix.
    weights: list of (W, [row slices that are used as operands on their own]).  extra_tasks: (src, hi, lo) copies that
    ride along.  Returns the lo buffers in order."""
    tasks, out = list(extra_tasks), []
    for w, slices in weights:
        buf = torch.empty(w.shape[0], round_up(w.shape[1], 4), dtype=torch.float32, device=w.device)
        tasks.append((w, None, buf))
        lo.put(w, buf)
        for sl in slices:
            lo.put(w[sl], buf[sl])
        out.append(buf)
    multi_split(tasks)
    return out


def multi_split(tasks):
    """tasks: list of (src, hi_dst or None, lo_dst or None), 2-D row-major views (1-D = one row).  One launch per 16
    tasks: hi_dst gets a copy of src, lo_dst gets src - trunc_tf32(src)."""
    import ctypes as C

    as2d = lambda t: t if t is None or t.dim() == 2 else t.view(1, -1)
    for i0 in range(0, len(tasks), 16):
        chunk = [(as2d(s), as2d(h), as2d(l)) for s, h, l in tasks[i0:i0 + 16]]
        n = len(chunk)
        for s, h, l in chunk:
            _mat(s, "multi_split src")
            assert (h is None or (h.stride(1) == 1 and h.shape == s.shape)) and (l is None or (l.stride(1) == 1 and l.shape[0] == s.shape[0]))
        PA, IA = C.c_void_p * n, C.c_int * n
        _lib.call("immtsf_multi_split", n, PA(*[s.data_ptr() for s, _, _ in chunk]), IA(*[s.stride(0) for s, _, _ in chunk]),
                  IA(*[s.shape[0] for s, _, _ in chunk]), IA(*[s.shape[1] for s, _, _ in chunk]),
                  PA(*[_p(h) for _, h, _ in chunk]), IA(*[h.stride(0) if h is not None else 0 for _, h, _ in chunk]),
                  PA(*[_p(l) for _, _, l in chunk]), IA(*[l.stride(0) if l is not None else 0 for _, _, l in chunk]), _stream())


def gemm(A, B, C, transA=False, transB=False, bias=None, alpha=1.0, beta=0.0, ragged=None, ragged_dim=0, backend=None,
         lo: Optional[LoCache] = None, emit_lo=False):
    """C[M,N] = alpha*op(A)*op(B) + beta*C + bias (shapes per include/immtsf.h).
    emit_lo (needs lo): the epilogue also writes C - trunc_tf32(C) and registers it in the cache, for a C that is an
    operand of a later tcgen05 product."""
    _mat(A, "A"), _mat(B, "B"), _mat(C, "C")
    M, N = C.shape
    K = A.shape[0] if transA else A.shape[1]
    ea = (K, M) if transA else (M, K)
    eb = (N, K) if transB else (K, N)
    if tuple(A.shape) != ea or tuple(B.shape) != eb:
        raise _lib.ImmtsfError(f"gemm: shape mismatch A{tuple(A.shape)} (want {ea}) B{tuple(B.shape)} (want {eb}) C{tuple(C.shape)}")
    if bias is not None:
        _chk(bias, "bias")
        assert bias.numel() == N and bias.is_contiguous()
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    be = gemm_backend() if backend is None else backend
    lib = _lib.load()
    ws, ws_bytes = None, 0
    if be != BACKEND_FFMA:
        ws_bytes = lib.immtsf_gemm_workspace_bytes(int(transA), int(transB), M, N, K)
        ws = _workspace(C.device, ws_bytes)
        ws_bytes = ws.numel()
    plan = lib.immtsf_gemm_plan(int(transA), int(transB), M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(C), C.stride(0), be)
    A_lo = B_lo = None
    if lo is not None and plan == 2:
        a_ragged = ragged is not None and ((ragged_dim == 1 and not transA) or (ragged_dim == 2 and transA))
        b_ragged = ragged is not None and (ragged_dim == 2 and not transB)
        A_lo = lo.lo_for(A, ragged if a_ragged else None)
        B_lo = lo.lo_for(B, ragged if b_ragged else None)
    C_lo = None
    if isinstance(emit_lo, torch.Tensor):
        C_lo = emit_lo  # caller-owned destination (a slice of a packed operand's lo); written by any backend
    elif emit_lo and lo is not None and plan == 2:
        # (ragged rows: pad rows inside touched tiles get lo = 0 like C; rows past them are never live in a product)
        C_lo = torch.empty(M, round_up(N, 4), dtype=torch.float32, device=C.device)
    _lib.call("immtsf_gemm_ex", int(transA), int(transB), M, N, K, float(alpha), _p(A), A.stride(0), _p(A_lo),
              A_lo.stride(0) if A_lo is not None else 0, _p(B), B.stride(0), _p(B_lo), B_lo.stride(0) if B_lo is not None else 0,
              float(beta), _p(C), C.stride(0), _p(C_lo), C_lo.stride(0) if C_lo is not None else 0, _p(bias), _p(ragged),
              ragged_dim, be, _p(ws), ws_bytes, _stream())
    if C_lo is not None and lo is not None:
        lo.put(C, C_lo)
    if prof is not None:
        ev1.record()
        prof.append(("gemm", (M, N, K, ragged_dim), ev0, ev1, plan))
    return C


def gemm_group(problems, lo: LoCache):
    """Independent products in one launch (immtsf_gemm_group).  problems: dicts with A, B, C and optional transA,
    transB, alpha, beta, emit_lo (a tensor: destination of C's lo).  Falls back to one immtsf_gemm_ex per product when
    a shape is not eligible for the tcgen05 kernel (tiny test sizes)."""
    import ctypes as C_

    lib = _lib.load()
    norm = []
    for pr in problems:
        A, B, C = _mat(pr["A"], "A"), _mat(pr["B"], "B"), _mat(pr["C"], "C")
        tA, tB = bool(pr.get("transA", False)), bool(pr.get("transB", False))
        M, N = C.shape
        K = A.shape[0] if tA else A.shape[1]
        assert tuple(A.shape) == ((K, M) if tA else (M, K)) and tuple(B.shape) == ((N, K) if tB else (K, N)), "gemm_group: shape mismatch"
        norm.append((A, B, C, tA, tB, M, N, K, float(pr.get("alpha", 1.0)), float(pr.get("beta", 0.0)), pr.get("emit_lo")))
    ok = all(lib.immtsf_gemm_plan(int(tA), int(tB), M, N, K, _p(A), A.stride(0), _p(B), B.stride(0), _p(C), C.stride(0),
                                  gemm_backend()) == 2 for (A, B, C, tA, tB, M, N, K, _, _, _) in norm)
    if not ok or gemm_backend() == BACKEND_FFMA:
        for (A, B, C, tA, tB, M, N, K, al, be, el) in norm:
            gemm(A, B, C, transA=tA, transB=tB, alpha=al, beta=be, lo=lo, emit_lo=el if el is not None else False)
        return
    for i0 in range(0, len(norm), 4):
        ch = norm[i0:i0 + 4]
        n = len(ch)
        Alo = [lo.lo_for(c[0], None) for c in ch]
        Blo = [lo.lo_for(c[1], None) for c in ch]
        IA, FA, PA = C_.c_int * n, C_.c_float * n, C_.c_void_p * n
        _lib.call("immtsf_gemm_group", n, IA(*[int(c[3]) for c in ch]), IA(*[int(c[4]) for c in ch]), IA(*[c[5] for c in ch]),
                  IA(*[c[6] for c in ch]), IA(*[c[7] for c in ch]), FA(*[c[8] for c in ch]),
                  PA(*[c[0].data_ptr() for c in ch]), PA(*[t.data_ptr() for t in Alo]), IA(*[c[0].stride(0) for c in ch]),
                  IA(*[t.stride(0) for t in Alo]), PA(*[c[1].data_ptr() for c in ch]), PA(*[t.data_ptr() for t in Blo]),
                  IA(*[c[1].stride(0) for c in ch]), IA(*[t.stride(0) for t in Blo]), FA(*[c[9] for c in ch]),
                  PA(*[c[2].data_ptr() for c in ch]), IA(*[c[2].stride(0) for c in ch]),
                  PA(*[_p(c[10]) for c in ch]), IA(*[c[10].stride(0) if c[10] is not None else 0 for c in ch]), _stream())
        for c in ch:
            if c[10] is not None:
                lo.put(c[2], c[10])


def linear_fwd(x, w, b, out=None, ragged=None, lo=None, emit_lo=False):
    """out[M,N] = x[M,K] w[N,K]^T + b."""
    if out is None:
        out = torch.empty(x.shape[0], w.shape[0], dtype=torch.float32, device=x.device)
    return gemm(x, w, out, transB=True, bias=b, ragged=ragged, ragged_dim=1 if ragged is not None else 0, lo=lo, emit_lo=emit_lo)


def linear_dgrad(dy, w, out=None, ragged=None, beta=0.0, lo=None, emit_lo=False):
    """dx[M,K] = dy[M,N] w[N,K].  On the tcgen05 path the weight is read TRANSPOSED (K-major, made once per step:
    LoCache.transposed) -- an MN-major fp32 operand costs the kernel 1.7x on these products."""
    if out is None:
        out = torch.empty(dy.shape[0], w.shape[1], dtype=torch.float32, device=dy.device)
    rd = 1 if ragged is not None else 0
    if (lo is not None and gemm_backend() != BACKEND_FFMA and min(w.shape) >= 64 and dy.shape[0] >= 256 and w.stride(1) == 1
            and os.environ.get("IMMTSF_DGRAD_T", "1") != "0"):
        return gemm(dy, lo.transposed(w), out, transB=True, beta=beta, ragged=ragged, ragged_dim=rd, lo=lo, emit_lo=emit_lo)
    return gemm(dy, w, out, beta=beta, ragged=ragged, ragged_dim=rd, lo=lo, emit_lo=emit_lo)


def linear_wgrad(dy, x, out=None, ragged=None, beta=0.0, lo=None, emit_lo=False):
    """dw[N,K] = dy[M,N]^T x[M,K]."""
    if out is None:
        out = torch.empty(dy.shape[1], x.shape[1], dtype=torch.float32, device=dy.device)
    return gemm(dy, x, out, transA=True, beta=beta, ragged=ragged, ragged_dim=2 if ragged is not None else 0, lo=lo, emit_lo=emit_lo)


def colsum(X, out=None, ragged=None, beta=0.0):
    _mat(X, "X")
    if out is None:
        out = torch.empty(X.shape[1], dtype=torch.float32, device=X.device)
    ws = _workspace(X.device, 16 << 20)
    _lib.call("immtsf_colsum", _p(X), X.shape[0], X.shape[1], X.stride(0), _p(out), float(beta), _p(ragged), _p(ws), ws.numel(),
              _stream())
    return out


def axpby(x, alpha, y, accumulate):
    _lib.call("immtsf_axpby", _p(x), float(alpha), _p(y), int(accumulate), x.numel(), _stream())
    return y


def group_sum_rows(x, R, T, d):
    out = torch.empty(R, d, dtype=torch.float32, device=x.device)
    _lib.call("immtsf_group_sum_rows", _p(x), d, R, T, d, _p(out), _stream())
    return out


# ------------------------------------------------------------------ RecAvg pooling
# Long segments x long windows: the pooling is dense matrix work -> tcgen05.  Selection measured on a B200 against the
# streaming kernels (profiles/r2_ab_recavg_fwd_bwd.txt): the forward wins at N 1024 x T 64 (1.4x), N 256 x T 256 (1.75x) and
# N 1024 x T 256 (2.8x) and ties at N 256 x T 64; the backward (its dV' product contracts over T only) wins at
# N 1024 x T 256 (1.43x) and loses at N 256 x T 256 and N 1024 x T 64.
RECAVG_TC_FWD_MIN = (256, 64, 65536)    # N_max >=, T >=, N_max * T >=
RECAVG_TC_BWD_MIN = (512, 128, 262144)


def recavg_tc_ok(r: RaggedNotes, Vp, T: int, d: int, backward: bool = False) -> bool:
    """The pooling of TTF_RecAvg.py:100 as batched tcgen05 products (csrc/recavg_tc.cu + immtsf_gemm_batched)
    (IMMTSF_RECAVG_TC=0 keeps the streaming kernels, =1 forces this path for any size)."""
    e = os.environ.get("IMMTSF_RECAVG_TC")
    if e == "0" or gemm_backend() == BACKEND_FFMA:
        return False
    ok = d % 4 == 0 and Vp.stride(0) % 4 == 0 and Vp.data_ptr() % 16 == 0 and r.B <= 65535 and r.N >= 1
    if e == "1":
        return ok
    n_min, t_min, nt_min = RECAVG_TC_BWD_MIN if backward else RECAVG_TC_FWD_MIN
    return ok and r.N >= n_min and T >= t_min and r.N * T >= nt_min


def _recavg_tc_operands(Vp, r: RaggedNotes, t_hat, log_sigma, T, d, with_c: bool):
    """Dense operands of the batched products and a LoCache that already holds their lo parts (written by the producers)."""
    B, Np = r.B, round_up(r.N, 4)
    dev = Vp.device
    f32 = torch.float32
    new = lambda *s: torch.empty(*s, dtype=f32, device=dev)
    Wn, Wn_lo = new(B, T, Np), new(B, T, Np)
    Cn, Cn_lo = (new(B, T, Np), new(B, T, Np)) if with_c else (None, None)
    wsum = new(B, T)
    csum = new(B, T) if with_c else None
    bstride = 0 if t_hat.dim() == 1 else t_hat.stride(0)
    _lib.call("immtsf_recavg_weights", _p(r.tau_flat), _p(r.offsets), _p(t_hat), bstride, _p(log_sigma), B, T, Np, _p(Wn), _p(Cn),
              _p(Wn_lo), _p(Cn_lo), _p(wsum), _p(csum), _stream())
    Vpad, Vpad_lo = new(B, Np, d), new(B, Np, d)
    _lib.call("immtsf_csr_to_padded", _p(Vp), Vp.stride(0), _p(r.offsets), B, Np, d, _p(Vpad), _p(Vpad_lo), _stream())
    lo = LoCache()
    for t, t_lo in ((Wn, Wn_lo), (Cn, Cn_lo), (Vpad, Vpad_lo)):
        if t is not None:
            lo[(t.data_ptr(), t.numel(), "flat")] = (t, t_lo)  # the key _flat_lo looks up (contiguous operands: extent = numel)
    return Np, Wn, Cn, wsum, csum, Vpad, lo


def _recavg_pool_fwd_tc(Vp, r: RaggedNotes, t_hat, log_sigma, gamma, beta, T, d, thr, seed, save):
    """E_raw[b] = Wn[b] [T x Np] . V'pad[b] [Np x d] on tcgen05 (Wn carries the 1 / clamp_min(denominator, 1e-6) of
    TTF_RecAvg.py:101-102), then dropout(LayerNorm(.)) with the streaming LayerNorm kernel (same dropout site and element
    indices as the fused kernels: identical masks)."""
    B = r.B
    Np, Wn, _, wsum, _, Vpad, lo = _recavg_tc_operands(Vp, r, t_hat, log_sigma, T, d, False)
    E_raw = torch.empty(B, T, d, dtype=torch.float32, device=Vp.device)
    gemm_batched(Wn, Vpad, E_raw, T, d, Np, (Np, T * Np, 0), (d, Np * d, 0), (d, T * d, 0), B, 1, lo=lo)
    E_drop, mean, rstd = ln_fwd(E_raw.view(B * T, d), None, None, T, gamma, beta, thr, seed, SITE_TTF_DROPOUT, save)
    if not save:
        return E_drop.view(B, T, d), None, None, None, None
    return E_drop.view(B, T, d), E_raw, mean.view(B, T), rstd.view(B, T), wsum


def _recavg_pool_bwd_tc(dE_drop, E_raw, mean, rstd, wsum, Vp, r: RaggedNotes, t_hat, log_sigma, gamma, T, d, thr, seed):
    """dE_raw = LayerNorm/dropout backward (streaming kernel); dV'pad[b] = Wn[b]^T dE_raw[b]; with Cn = Wn 2 (delta/sigma)^2 and
    R[b] = Cn[b] V'pad[b]:  dlog_sigma = sum dE_raw (R - csum E_raw)  (csrc/recavg_tc.cu).  Two batched tcgen05 products,
    nothing of size N x T x d is formed."""
    B = r.B
    dev = Vp.device
    f32 = torch.float32
    dE_raw, _, dgamma, dbeta = ln_bwd(dE_drop.reshape(B * T, d), E_raw.view(B * T, d), None, None, T, gamma, mean.reshape(B * T),
                                      rstd.reshape(B * T), thr, seed, SITE_TTF_DROPOUT)
    Np, Wn, Cn, _, csum, Vpad, lo = _recavg_tc_operands(Vp, r, t_hat, log_sigma, T, d, True)
    dVpad = torch.empty(B, Np, d, dtype=f32, device=dev)
    gemm_batched(Wn, dE_raw, dVpad, Np, d, T, (Np, T * Np, 0), (d, T * d, 0), (d, Np * d, 0), B, 1, transA=True, lo=lo)
    R = torch.empty(B, T, d, dtype=f32, device=dev)
    gemm_batched(Cn, Vpad, R, T, d, Np, (Np, T * Np, 0), (d, Np * d, 0), (d, T * d, 0), B, 1, lo=lo)
    dVp = torch.empty(r.M_alloc, d, dtype=f32, device=dev)
    _lib.call("immtsf_padded_to_csr", _p(dVpad), _p(r.offsets), B, Np, d, _p(dVp), d, _stream())
    zero_pad_rows(dVp, d, r.m_dev, r.M_alloc)
    dls = torch.zeros((), dtype=torch.float64, device=dev)
    _lib.call("immtsf_recavg_dls", _p(dE_raw), _p(R), _p(E_raw), _p(csum), B * T, d, _p(dls), _stream())
    return dVp, dgamma, dbeta, dls.float()


def recavg_pool_fwd(Vp, r: RaggedNotes, t_hat, log_sigma, gamma, beta, T, d, thr, seed, save):
    B = r.B
    _chk(t_hat, "t_hat")
    dev = Vp.device
    if recavg_tc_ok(r, Vp, T, d):
        return _recavg_pool_fwd_tc(Vp, r, t_hat, log_sigma, gamma, beta, T, d, thr, seed, save)
    E_drop = torch.empty(B, T, d, dtype=torch.float32, device=dev)
    E_raw = torch.empty(B, T, d, dtype=torch.float32, device=dev) if save else None
    mean = torch.empty(B, T, dtype=torch.float32, device=dev) if save else None
    rstd = torch.empty(B, T, dtype=torch.float32, device=dev) if save else None
    wsum = torch.empty(B, T, dtype=torch.float32, device=dev) if save else None
    bstride = 0 if t_hat.dim() == 1 else t_hat.stride(0)
    _lib.call("immtsf_recavg_pool_fwd", _p(Vp), Vp.stride(0), _p(r.tau_flat), _p(r.offsets), _p(t_hat), bstride,
              _p(log_sigma), _p(gamma), _p(beta), B, T, d, max(r.N, 1), LN_EPS, thr, seed, _p(E_drop), _p(E_raw), _p(mean), _p(rstd),
              _p(wsum), _stream())
    return E_drop, E_raw, mean, rstd, wsum


def recavg_pool_bwd(dE_drop, E_raw, mean, rstd, wsum, Vp, r: RaggedNotes, t_hat, log_sigma, gamma, T, d, thr, seed):
    dev = Vp.device
    if recavg_tc_ok(r, Vp, T, d, backward=True):
        return _recavg_pool_bwd_tc(dE_drop, E_raw, mean, rstd, wsum, Vp, r, t_hat, log_sigma, gamma, T, d, thr, seed)
    dVp = torch.empty(r.M_alloc, d, dtype=torch.float32, device=dev)
    dS = torch.empty(r.B * T * (d + 1), dtype=torch.float32, device=dev)  # rows [B*T, d] + one scalar per row
    dgamma = torch.zeros(d, dtype=torch.float32, device=dev)
    dbeta = torch.zeros(d, dtype=torch.float32, device=dev)
    dls = torch.zeros((), dtype=torch.float64, device=dev)  # summed in double by the kernel
    bstride = 0 if t_hat.dim() == 1 else t_hat.stride(0)
    _lib.call("immtsf_recavg_pool_bwd", _p(dE_drop), _p(E_raw), _p(mean), _p(rstd), _p(wsum), _p(Vp), Vp.stride(0),
              _p(r.tau_flat), _p(r.offsets), _p(t_hat), bstride, _p(log_sigma), _p(gamma), r.B, T, d, max(r.N, 1), thr, seed,
              _p(dS), _p(dVp), d, _p(dgamma), _p(dbeta), _p(dls), _stream())
    zero_pad_rows(dVp, d, r.m_dev, r.M_alloc)
    return dVp, dgamma, dbeta, dls.float()


# ------------------------------------------------------------------ Time2Vec / segment attention / LN
def time2vec_fwd(r: RaggedNotes, w_lin, b_lin, w_per, b_per, d_tau, out_view, lo_view=None):
    """out_view: [M_alloc, d_tau] column slice of the concat buffer; lo_view: the same slice of its lo operand."""
    _lib.call("immtsf_time2vec_fwd", _p(r.tau_flat), _p(w_lin), _p(b_lin), _p(w_per), _p(b_per), d_tau, _p(out_view),
              out_view.stride(0), _p(lo_view), lo_view.stride(0) if lo_view is not None else 0, _p(r.m_dev), r.M_alloc, _stream())


def time2vec_bwd(dphi_view, r: RaggedNotes, w_per, b_per, d_tau, buf=None):
    """buf: optional ZEROED [2 * d_tau] buffer the four gradients are accumulated into (views of it are returned)."""
    dev = dphi_view.device
    if buf is None:
        buf = torch.zeros(2 * d_tau, dtype=torch.float32, device=dev)
    dwl, dbl = buf[0:1].view(1, 1), buf[1:2]
    dwp, dbp = buf[2:d_tau + 1].view(d_tau - 1, 1), buf[d_tau + 1:2 * d_tau]
    _lib.call("immtsf_time2vec_bwd", _p(dphi_view), dphi_view.stride(0), _p(r.tau_flat), _p(w_per), _p(b_per), d_tau,
              _p(dwl), _p(dbl), _p(dwp), _p(dbp), _p(r.m_dev), r.M_alloc, _stream())
    return dwl, dbl, dwp, dbp


def segattn_fwd(q, KVp, r: RaggedNotes, T, H, d, per_query, thr, seed, save):
    R = r.B * T if per_query else r.B
    attn_cat = torch.empty(R, d, dtype=torch.float32, device=KVp.device)
    probs = torch.empty(r.M_alloc, H, dtype=torch.float32, device=KVp.device) if save else None
    _lib.call("immtsf_segattn_fwd", _p(q), _p(KVp), _p(r.offsets), r.B, T, H, d, max(r.N, 1), int(per_query), thr, seed,
              _p(attn_cat), _p(probs), _stream())
    return attn_cat, probs


def segattn_ln_ok(d, N_max) -> bool:
    return os.environ.get("IMMTSF_SEGATTN_FUSED", "1") != "0" and bool(_lib.load().immtsf_segattn_ln_ok(d, max(N_max, 1)))


def segattn_ln_fwd(q, KVp, r: RaggedNotes, T, d, thr, seed, xbias, res, gamma, beta, save):
    """One head, train mode: segment attention + bias + residual + LayerNorm + dropout in one launch (csrc/t2v_segattn.cu).
    Returns y [B*T, d] and what backward needs: attn_cat [B*T, d], probs [M_alloc, 1], mean, rstd [B*T]."""
    dev = KVp.device
    R = r.B * T
    attn_cat = torch.empty(R, d, dtype=torch.float32, device=dev)
    y = torch.empty(R, d, dtype=torch.float32, device=dev)
    probs = torch.empty(r.M_alloc, 1, dtype=torch.float32, device=dev) if save else None
    mean = torch.empty(R, dtype=torch.float32, device=dev) if save else None
    rstd = torch.empty(R, dtype=torch.float32, device=dev) if save else None
    _lib.call("immtsf_segattn_ln_fwd", _p(q), _p(KVp), _p(r.offsets), r.B, T, d, max(r.N, 1), thr, seed, _p(xbias), _p(res), _p(gamma),
              _p(beta), LN_EPS, _p(attn_cat), _p(probs), _p(y), _p(mean), _p(rstd), _stream())
    return y, attn_cat, probs, mean, rstd


def segattn_bwd(d_attn_cat, q, KVp, probs, r: RaggedNotes, T, H, d, per_query, thr, seed):
    dKVp = torch.empty(r.M_alloc, 2 * d, dtype=torch.float32, device=KVp.device)
    dq_partial = torch.empty(r.B, d, dtype=torch.float32, device=KVp.device)
    _lib.call("immtsf_segattn_bwd", _p(d_attn_cat), _p(q), _p(KVp), _p(probs), _p(r.offsets), r.B, T, H, d, max(r.N, 1),
              int(per_query), thr, seed, _p(dKVp), _p(dq_partial), _stream())
    zero_pad_rows(dKVp, 2 * d, r.m_dev, r.M_alloc)
    return dKVp, dq_partial


def t2vq_attn_fwd(A, a_sc, g, r: RaggedNotes, t_hat, t2v_params, T, H, d, d_tau, thr, seed, save):
    """Per-(note, query) Time2Vec attention over each ragged segment (csrc/t2v_perquery.cu).  Returns Z [B*T*H, d],
    Phi [B*T*H, d_tau], sp [B*T*H] and the saved softmax [H*T*M_alloc] (None unless save)."""
    R = r.B * T * H
    dev = A.device
    _chk(t_hat, "t_hat")
    Z = torch.empty(R, d, dtype=torch.float32, device=dev)
    Phi = torch.empty(R, d_tau, dtype=torch.float32, device=dev)
    sp = torch.empty(R, dtype=torch.float32, device=dev)
    probs = torch.empty(H * T * r.M_alloc, dtype=torch.float32, device=dev) if save else None
    w_lin, b_lin, w_per, b_per = t2v_params
    bstride = 0 if t_hat.dim() == 1 else t_hat.stride(0)
    _lib.call("immtsf_t2vq_attn_fwd", _p(A), A.stride(0), _p(a_sc), _p(g), _p(r.tau_flat), _p(r.offsets), _p(t_hat), bstride,
              _p(w_lin), _p(b_lin), _p(w_per), _p(b_per), r.B, T, H, d, d_tau, max(r.N, 1), r.M_alloc, thr, seed, _p(Z), _p(Phi),
              _p(sp), _p(probs), _stream())
    return Z, Phi, sp, probs


def t2vq_attn_bwd(dZ, dPhi, dsp, A, g, probs, r: RaggedNotes, t_hat, t2v_params, T, H, d, d_tau, thr, seed):
    """Returns dA [M_alloc, d] (pooling part), da [M_alloc, H] and the per-(sample, query tile) partials [B*tiles, (2+H)*d_tau]."""
    dev = A.device
    dA = torch.empty(r.M_alloc, d, dtype=torch.float32, device=dev)
    da = torch.empty(r.M_alloc, H, dtype=torch.float32, device=dev)
    lib = _lib.load()
    tiles = lib.immtsf_t2vq_bwd_tiles(T, H, max(r.N, 1))
    dpart = torch.empty(max(r.B, 1) * tiles, (2 + H) * d_tau, dtype=torch.float32, device=dev)
    ws_bytes = lib.immtsf_t2vq_bwd_workspace_bytes(T, H, r.M_alloc)
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)  # P~ and dS between the two launches
    w_lin, b_lin, w_per, b_per = t2v_params
    bstride = 0 if t_hat.dim() == 1 else t_hat.stride(0)
    _lib.call("immtsf_t2vq_attn_bwd", _p(dZ), _p(dPhi), _p(dsp), _p(A), A.stride(0), _p(g), _p(probs), _p(r.tau_flat), _p(r.offsets),
              _p(t_hat), bstride, _p(w_lin), _p(b_lin), _p(w_per), _p(b_per), r.B, T, H, d, d_tau, max(r.N, 1), r.M_alloc, thr, seed,
              _p(dA), dA.stride(0), _p(da), _p(dpart), _p(ws), ws_bytes, _stream())
    zero_pad_rows(dA, d, r.m_dev, r.M_alloc)
    zero_pad_rows(da, H, r.m_dev, r.M_alloc)
    return dA, da, dpart


def ln_fwd(x, res, valid, rows_per_sample, gamma, beta, thr, seed, site, save, xbias=None):
    R, d = x.shape
    y = torch.empty(R, d, dtype=torch.float32, device=x.device)
    mean = torch.empty(R, dtype=torch.float32, device=x.device) if save else None
    rstd = torch.empty(R, dtype=torch.float32, device=x.device) if save else None
    _lib.call("immtsf_ln_fwd", _p(x), x.stride(0), _p(xbias), _p(res), _p(valid), rows_per_sample, _p(gamma), _p(beta), R, d, LN_EPS,
              thr, seed, site, _p(y), _p(mean), _p(rstd), _stream())
    return y, mean, rstd


def ln_bwd(dy, x, res, valid, rows_per_sample, gamma, mean, rstd, thr, seed, site, xbias=None, acc=None):
    """acc: optional ZEROED [3 * d] buffer for (dres, dgamma, dbeta) -- one fill instead of three; views of it are returned."""
    R, d = x.shape
    dev = x.device
    dx = torch.empty(R, d, dtype=torch.float32, device=dev)
    if acc is None:
        acc = torch.zeros(3 * d, dtype=torch.float32, device=dev)
    dres = acc[0:d] if res is not None else None
    dgamma, dbeta = acc[d:2 * d], acc[2 * d:3 * d]
    _lib.call("immtsf_ln_bwd", _p(dy), _p(x), x.stride(0), _p(xbias), _p(res), _p(valid), rows_per_sample, _p(gamma), _p(mean), _p(rstd),
              R, d, thr, seed, site, _p(dx), _p(dres), _p(dgamma), _p(dbeta), _stream())
    return dx, dres, dgamma, dbeta


# ------------------------------------------------------------------ GR_Add
def gru_scan_fwd(G4, w_hh, b_hh, B, T, C):
    h_all = torch.empty(B * T, C, dtype=torch.float32, device=G4.device)
    h_prev = torch.empty(B * T, C, dtype=torch.float32, device=G4.device)
    # many channels: wide recurrence (one CTA per sample), which stores the gate activations for backward
    gates = torch.empty(B * T, 4 * C, dtype=torch.float32, device=G4.device) if C > 32 else None
    _lib.call("immtsf_gru_scan_fwd", _p(G4), _p(w_hh), _p(b_hh), B, T, C, _p(h_all), _p(h_prev), _p(gates), _stream())
    return h_all, h_prev, gates


def gr_tail_fwd(Y, G4, h_all, w_r, b_r, gamma, beta, m_txt, B, T, C, thr, seed, flags):
    Y_out = torch.empty(B, T, C, dtype=torch.float32, device=Y.device)
    _lib.call("immtsf_gr_tail_fwd", _p(Y), _p(G4), _p(h_all), _p(w_r), _p(b_r), _p(gamma), _p(beta), _p(m_txt), B, T, C,
              LN_EPS, thr, seed, _p(Y_out), _p(flags), _stream())
    return Y_out


def gr_tail_bwd(dY_out, G4, h_all, w_r, b_r, gamma, beta, m_txt, B, T, C, thr, seed, dG4):
    dev = G4.device
    d_delta = torch.empty(B * T, C, dtype=torch.float32, device=dev)
    dh_out = torch.empty(B * T, C, dtype=torch.float32, device=dev)
    dgamma = torch.zeros(C, dtype=torch.float32, device=dev)
    dbeta = torch.zeros(C, dtype=torch.float32, device=dev)
    _lib.call("immtsf_gr_tail_bwd", _p(dY_out), _p(G4), _p(h_all), _p(w_r), _p(b_r), _p(gamma), _p(beta), _p(m_txt), B, T, C,
              LN_EPS, thr, seed, _p(dG4), _p(d_delta), _p(dh_out), _p(dgamma), _p(dbeta), _stream())
    return d_delta, dh_out, dgamma, dbeta


def gru_scan_bwd(G4, h_prev, w_hh, b_hh, dh_out, B, T, C, dG4, gates=None):
    dGh = torch.empty(B * T, 3 * C, dtype=torch.float32, device=G4.device)
    _lib.call("immtsf_gru_scan_bwd", _p(G4), _p(h_prev), _p(w_hh), _p(b_hh), _p(dh_out), _p(gates), B, T, C, _p(dG4), _p(dGh),
              _stream())
    return dGh


# ------------------------------------------------------------------ XAttn_Add
XATTN_SMALL_T = 32  # up to here one CTA holds the whole T x T problem of a (sample, head) (csrc/xattn_small.cu)


def _flat_lo(lo: "LoCache", t: torch.Tensor, rows: int, cols: int, strides, batch1: int, batch2: int):
    """lo over the flat extent of a batched operand (same layout as the operand), cached per (pointer, extent)."""
    ld, s1, s2 = strides
    extent = (batch1 - 1) * s1 + (batch2 - 1) * s2 + (rows - 1) * ld + cols
    key = (t.data_ptr(), extent, "flat")
    e = lo.get(key)
    if e is None:
        buf = torch.empty(round_up(extent, 4), dtype=torch.float32, device=t.device)
        _lib.call("immtsf_split_lo", _p(t), round_up(extent, 4), 1, extent, _p(buf), round_up(extent, 4), None, _stream())
        e = (t, buf)
        lo[key] = e
    return e[1]


def gemm_batched(A, B, C, M, N, K, a_strides, b_strides, c_strides, batch1, batch2, transA=False, transB=False,
                 alpha=1.0, beta=0.0, lo: Optional["LoCache"] = None):
    """C(b1,b2)[M,N] = alpha*op(A(b1,b2)) op(B(b1,b2)) + beta*C(b1,b2); *_strides = (ld, s1, s2) in elements, A/B/C are
    the base tensors (only their data pointers are used).  tcgen05 3xTF32 (include/immtsf.h)."""
    lib = _lib.load()
    (lda, a1, a2), (ldb, b1, b2), (ldc, c1, c2) = a_strides, b_strides, c_strides
    need = lib.immtsf_gemm_batched_workspace_bytes(int(transA), int(transB), M, N, K, lda, a1, a2, ldb, b1, b2, batch1, batch2)
    ws = _workspace(C.device, need)
    A_lo = B_lo = None
    if lo is not None:
        ra, ca = (K, M) if transA else (M, K)
        rb, cb = (N, K) if transB else (K, N)
        A_lo = _flat_lo(lo, A, ra, ca, a_strides, batch1, batch2)
        B_lo = _flat_lo(lo, B, rb, cb, b_strides, batch1, batch2)
    _lib.call("immtsf_gemm_batched", int(transA), int(transB), M, N, K, float(alpha), _p(A), _p(A_lo), lda, a1, a2, _p(B),
              _p(B_lo), ldb, b1, b2, float(beta), _p(C), ldc, c1, c2, batch1, batch2, _p(ws), ws.numel(), _stream())
    return C


def _xattn_large_ok(q, k, v, T, H, d):
    hd = d // H
    al = lambda t: t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0
    return T > XATTN_SMALL_T and hd % 4 == 0 and al(q) and al(k) and al(v) and B_H_ok(q.shape[0] // T, H)


def B_H_ok(B, H):
    return B * H <= 65535


def xattn_core_fwd(q, k, v, m_txt, B, T, H, d, thr, seed, save, lo=None):
    """MHA core of MMF_XAttn_Add.  T <= 32: one fused kernel per direction.  T > 32: batched tcgen05 products around a
    row-softmax kernel.  Returns (o [B*T, d], saved) where saved is what xattn_core_bwd needs."""
    o = torch.empty(B * T, d, dtype=torch.float32, device=q.device)
    if _xattn_large_ok(q, k, v, T, H, d):
        hd, Tp = d // H, round_up(T, 4)
        scale = math.sqrt(1.0 / float(hd))
        P = torch.empty(B, H, T, Tp, dtype=torch.float32, device=q.device)
        Pt = torch.empty(B, H, T, Tp, dtype=torch.float32, device=q.device)
        sP = (Tp, H * T * Tp, T * Tp)
        sq, sk, sv, so = (q.stride(0), T * q.stride(0), hd), (k.stride(0), T * k.stride(0), hd), (v.stride(0), T * v.stride(0), hd), \
            (o.stride(0), T * o.stride(0), hd)
        gemm_batched(q, k, P, T, T, hd, sq, sk, sP, B, H, transB=True, lo=lo)  # S = Q K^T
        _lib.call("immtsf_softmax_rows_fwd", _p(P), _p(Pt), _p(m_txt), B, H, T, Tp, scale, thr, seed, _stream())
        gemm_batched(Pt, v, o, T, hd, T, sP, sv, so, B, H, lo=lo)  # O = P~ V
        return o, (P if save else None)
    probs = torch.empty(B, H, T, T, dtype=torch.float32, device=q.device) if save else None
    _lib.call("immtsf_xattn_core_fwd", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(m_txt), B, T, H, d,
              thr, seed, _p(o), o.stride(0), _p(probs), _stream())
    return o, probs


def xattn_core_bwd(d_o, q, k, v, probs, m_txt, B, T, H, d, thr, seed, dq, dk, dv, lo=None):
    if _xattn_large_ok(q, k, v, T, H, d) and probs.shape[-1] == round_up(T, 4) and d_o.stride(0) % 4 == 0:
        hd, Tp = d // H, round_up(T, 4)
        scale = math.sqrt(1.0 / float(hd))
        dS = torch.empty(B, H, T, Tp, dtype=torch.float32, device=q.device)
        Pt = torch.empty(B, H, T, Tp, dtype=torch.float32, device=q.device)
        sP = (Tp, H * T * Tp, T * Tp)
        st = lambda t: (t.stride(0), T * t.stride(0), hd)
        gemm_batched(d_o, v, dS, T, T, hd, st(d_o), st(v), sP, B, H, transB=True, lo=lo)  # dP~ = dO V^T
        _lib.call("immtsf_softmax_rows_bwd", _p(dS), _p(probs), _p(Pt), _p(m_txt), B, H, T, Tp, scale, thr, seed, _stream())
        # dS was rewritten in place by the softmax kernel after its use as an output: its lo is made now and shared
        gemm_batched(dS, k, dq, T, hd, T, sP, st(k), st(dq), B, H, lo=lo)  # dQ = dS K
        gemm_batched(dS, q, dk, T, hd, T, sP, st(q), st(dk), B, H, transA=True, lo=lo)  # dK = dS^T Q
        gemm_batched(Pt, d_o, dv, T, hd, T, sP, st(d_o), st(dv), B, H, transA=True, lo=lo)  # dV = P~^T dO
        return
    _lib.call("immtsf_xattn_core_bwd", _p(d_o), d_o.stride(0), _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0),
              _p(probs), _p(m_txt), B, T, H, d, thr, seed, _p(dq), dq.stride(0), _p(dk), dk.stride(0), _p(dv), dv.stride(0),
              _stream())


def xattn_lowrank_ok(T, H, d, C, *tensors) -> bool:
    """Rank-(C+1) query path applies (T <= 32, C + 1 <= 32) and the wide operands are 16B aligned."""
    al = all(t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0 for t in tensors)
    return al and bool(_lib.load().immtsf_xattn_lowrank_ok(T, H, d, C))


def xattn_lowrank_fwd(Y2, kq, v, m_txt, B, T, H, d, C, thr, seed, save):
    o = torch.empty(B * T, d, dtype=torch.float32, device=v.device)
    probs = torch.empty(B, H, T, T, dtype=torch.float32, device=v.device) if save else None
    _lib.call("immtsf_xattn_lowrank_fwd", _p(Y2), Y2.stride(0), _p(kq), kq.stride(0), _p(v), v.stride(0), _p(m_txt), B, T, H, d, C,
              thr, seed, _p(o), o.stride(0), _p(probs), _stream())
    return o, probs


def xattn_lowrank_bwd(d_o, Y2, kq, v, probs, m_txt, B, T, H, d, C, thr, seed, dv):
    z = torch.empty(B * T, H * (C + 1), dtype=torch.float32, device=v.device)
    dyh = torch.empty(H, B * T, C, dtype=torch.float32, device=v.device)
    _lib.call("immtsf_xattn_lowrank_bwd", _p(d_o), d_o.stride(0), _p(Y2), Y2.stride(0), _p(kq), kq.stride(0), _p(v), v.stride(0),
              _p(probs), _p(m_txt), B, T, H, d, C, thr, seed, _p(dv), dv.stride(0), _p(z), z.stride(0), _p(dyh), _stream())
    return z, dyh


def xattn_rank_ok(T, H, d, C) -> bool:
    if os.environ.get("IMMTSF_XATTN_RANK", "1") == "0":
        return False
    return bool(_lib.load().immtsf_xattn_rank_ok(T, H, d, C))


def xattn_rank_fwd(Y2, R, bo, m_txt, B, T, H, d, C, thr, seed, save):
    delta_y = torch.empty(B * T, C, dtype=torch.float32, device=R.device)
    probs = torch.empty(B, H, T, T, dtype=torch.float32, device=R.device) if save else None
    _lib.call("immtsf_xattn_rank_fwd", _p(Y2), Y2.stride(0), _p(R), R.stride(0), _p(bo), _p(m_txt), B, T, H, d, C, thr, seed,
              _p(delta_y), _p(probs), _stream())
    return delta_y, probs


def xattn_rank_bwd(d_delta, Y2, R, probs, m_txt, B, T, H, d, C, thr, seed):
    dR = torch.empty(R.shape[0], R.stride(0), dtype=torch.float32, device=R.device)[:, : R.shape[1]]  # same padding as R
    dY = torch.empty(B * T, C, dtype=torch.float32, device=R.device)
    _lib.call("immtsf_xattn_rank_bwd", _p(d_delta), _p(Y2), Y2.stride(0), _p(R), R.stride(0), _p(probs), _p(m_txt), B, T, H, d, C,
              thr, seed, _p(dR), dR.stride(0), _p(dY), _stream())
    return dR, dY


def xattn_rank_fused_ok(T, H, d, C, E2, Wr) -> bool:
    """The one-launch-per-direction data half of the rank form applies (csrc/xattn_rank_fused.cu)."""
    if os.environ.get("IMMTSF_XATTN_FUSED", "1") == "0":
        return False
    al = all(t.dim() == 2 and t.stride(1) == 1 and t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0 for t in (E2, Wr))
    return al and bool(_lib.load().immtsf_xattn_rank_fused_ok(T, H, d, C, E2.shape[1]))


def xattn_rank_fused_fwd(Y2, E2, Wr, br, bo, gamma, beta, m_txt, B, T, H, d, C, kappa, thr, seed, save, flags):
    """R = E Wr^T + br, the T x (2C+1) attention and the LayerNorm_C / dropout / blend tail in one launch.
    Returns Y_out [B, T, C] and what backward needs: R [B*T, nr], delta_y [B*T, C], probs [B, H, T, T] (None unless save)."""
    dev = E2.device
    nr = H * (2 * C + 1)
    R = torch.empty(B * T, round_up(nr, 4), dtype=torch.float32, device=dev)[:, :nr]
    delta_y = torch.empty(B * T, C, dtype=torch.float32, device=dev)
    probs = torch.empty(B, H, T, T, dtype=torch.float32, device=dev) if save else None
    Y_out = torch.empty(B, T, C, dtype=torch.float32, device=dev)
    _lib.call("immtsf_xattn_rank_fused_fwd", _p(E2), E2.stride(0), E2.shape[1], _p(Wr), Wr.stride(0), _p(br), _p(Y2), Y2.stride(0),
              _p(bo), _p(gamma), _p(beta), _p(m_txt), B, T, H, d, C, LN_EPS, float(kappa), thr, seed, _p(R), R.stride(0),
              _p(delta_y), _p(probs), _p(Y_out), _p(flags), _stream())
    return Y_out, R, delta_y, probs


def xattn_rank_fused_bwd(dY_out, delta_y, gamma, Y2, R, probs, m_txt, E2, Wr, B, T, H, d, C, kappa, thr, seed):
    """Returns dE [B*T, de], dY [B*T, C], dWr [nr, de], dbr [nr], d(bo_f) [C], dgamma [C], dbeta [C] and the buffer the last five
    are views of."""
    dev = E2.device
    nr, de = Wr.shape
    dE = torch.empty(B * T, de, dtype=torch.float32, device=dev)
    dY = torch.empty(B * T, C, dtype=torch.float32, device=dev)
    pack = dp_pack(nr * de + nr + 3 * C, dev)  # one buffer: the data-parallel all-reduce takes it whole
    dWr, small = pack[:nr * de].view(nr, de), pack[nr * de:]
    need = _lib.load().immtsf_xattn_rank_fused_bwd_workspace_bytes(B, H, C, de)
    ws = _workspace(dev, need + 256)
    off = (-ws.data_ptr()) % 16
    _lib.call("immtsf_xattn_rank_fused_bwd", _p(dY_out), _p(delta_y), _p(gamma), _p(Y2), Y2.stride(0), _p(R), R.stride(0), _p(probs),
              _p(m_txt), _p(E2), E2.stride(0), de, _p(Wr), Wr.stride(0), B, T, H, d, C, LN_EPS, float(kappa), thr, seed, _p(dE),
              dE.stride(0), _p(dY), _p(dWr), _p(small), ws.data_ptr() + off, ws.numel() - off, _stream())
    return dE, dY, dWr, small[:nr], small[nr:nr + C], small[nr + C:nr + 2 * C], small[nr + 2 * C:], pack


def xattn_tail_fwd(Y, delta_y, gamma, beta, m_txt, B, T, C, kappa, thr, seed, flags):
    Y_out = torch.empty(B, T, C, dtype=torch.float32, device=Y.device)
    _lib.call("immtsf_xattn_tail_fwd", _p(Y), _p(delta_y), _p(gamma), _p(beta), _p(m_txt), B, T, C, LN_EPS, float(kappa), thr,
              seed, _p(Y_out), _p(flags), _stream())
    return Y_out


def xattn_tail_bwd(dY_out, delta_y, gamma, m_txt, B, T, C, kappa, thr, seed):
    dev = delta_y.device
    d_delta = torch.empty(B * T, C, dtype=torch.float32, device=dev)
    dgamma = torch.zeros(C, dtype=torch.float32, device=dev)
    dbeta = torch.zeros(C, dtype=torch.float32, device=dev)
    _lib.call("immtsf_xattn_tail_bwd", _p(dY_out), _p(delta_y), _p(gamma), _p(m_txt), B, T, C, LN_EPS, float(kappa), thr, seed,
              _p(d_delta), _p(dgamma), _p(dbeta), _stream())
    return d_delta, dgamma, dbeta
