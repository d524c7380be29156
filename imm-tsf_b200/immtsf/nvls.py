"""Symmetric-memory arena + the hand-written in-switch all-reduce (csrc/nvls.cu) for the data-parallel gradient statistics.

torch.distributed._symmetric_memory is used for PLUMBING only: it allocates the arena on every rank, exchanges the
handles and maps the peers' replicas and the NVSwitch multicast address into this process.  The collective itself is
immtsf_nvls_allreduce_f32: a barrier over the ranks, multimem.ld_reduce + multimem.st of this rank's slice, a barrier.

    comm = NvlsComm.create(group)             # None when the fabric has no multicast (then NCCL does the all-reduce)
    comm.reset()                              # start of a step: the bump allocator starts over (same offsets every step)
    buf = comm.alloc(n_floats)                # a view of the arena; kernels write the statistics straight into it
    comm.all_reduce(buf)                      # in place, on the current stream; capturable in a CUDA graph
"""
from __future__ import annotations

import os

import torch

from . import _lib


class NvlsComm:
    CHANNELS = 16

    def __init__(self, group, arena, handle, chan_floats):
        self.group, self.arena, self.handle = group, arena, handle
        self.rank, self.world = handle.rank, handle.world_size
        # The flags sit in front of the data, one set per CHANNEL: the collectives of a step run on different streams and may
        # overlap in time, in a different order on different ranks, so the k-th collective of a step owns flag set k (every
        # rank issues the same sequence of collectives per step; the same collective of consecutive steps is ordered by the
        # streams).  NCCL serialises the collectives of a communicator instead.
        self.chan_floats = chan_floats
        self.flag_floats = chan_floats * self.CHANNELS
        self.cursor = self.flag_floats
        self.channel = 0
        self.err = torch.zeros(1, dtype=torch.int32, device=arena.device)
        self.base = arena.data_ptr()
        self.nbytes = arena.numel() * 4

    @staticmethod
    def create(group, arena_mb: float = 24.0):
        """Collective over `group`.  Returns None when symmetric memory / multicast is unavailable or IMMTSF_DP_NVLS=0."""
        if os.environ.get("IMMTSF_DP_NVLS", "1") == "0":
            return None
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            if dist.get_world_size(group) < 2:
                return None
            name = group.group_name
            try:
                symm_mem.enable_symm_mem_for_group(name)
            except Exception:
                pass
            dev = torch.device("cuda", torch.cuda.current_device())
            n = int(arena_mb * (1 << 20)) // 4
            arena = symm_mem.empty(n, dtype=torch.float32, device=dev)
            handle = symm_mem.rendezvous(arena, name)
            if not handle.multicast_ptr:
                return None
            flag_bytes = _lib.load().immtsf_nvls_flag_bytes(handle.world_size)
            chan_floats = (flag_bytes // 4 + 63) // 64 * 64
            arena[:chan_floats * NvlsComm.CHANNELS].zero_()
            torch.cuda.synchronize()
            handle.barrier()  # every rank's flags are zero before anyone signals
            torch.cuda.synchronize()
            return NvlsComm(group, arena, handle, chan_floats)
        except Exception as e:  # no symmetric memory on this system: NCCL takes over
            if os.environ.get("IMMTSF_DP_NVLS_DEBUG"):
                print(f"[immtsf.nvls] unavailable: {type(e).__name__}: {e}")
            return None

    def reset(self):
        self.cursor = self.flag_floats
        self.channel = 0

    def alloc(self, n_floats: int) -> torch.Tensor:
        n = (n_floats + 3) // 4 * 4
        if self.cursor + n > self.arena.numel():
            return None
        t = self.arena[self.cursor:self.cursor + n_floats]
        self.cursor += (n + 63) // 64 * 64
        return t

    def owns(self, t: torch.Tensor) -> bool:
        p = t.data_ptr()
        return self.base <= p < self.base + self.nbytes and t.is_contiguous() and (p - self.base) % 16 == 0

    def all_reduce(self, t: torch.Tensor):
        off = (t.data_ptr() - self.base) // 4
        n = (t.numel() + 3) // 4 * 4  # (allocations are padded to 4 floats; the pad is reduced along, harmlessly)
        h = self.handle
        flag_off = (self.channel % self.CHANNELS) * self.chan_floats
        self.channel += 1
        _lib.call("immtsf_nvls_allreduce_f32", h.multicast_ptr, h.buffer_ptrs_dev, off, n, flag_off, self.rank, self.world,
                  self.err.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def check(self):
        """Raise if a barrier of an earlier all-reduce timed out (one device->host read)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("immtsf NVLS all-reduce: a rank never arrived at a barrier (timeout)")
