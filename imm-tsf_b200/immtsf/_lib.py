"""ctypes binding of libimmtsf.so (the C ABI declared in include/immtsf.h).

There is deliberately no fallback: if the shared library is missing the import
of any compute path raises, and every entry point returns an error on a device
that is not sm_100.  Build the library with ``python __graft_entry__.py`` (or
``make -C imm-tsf_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libimmtsf.so")

P = C.c_void_p
I = C.c_int
F = C.c_float
U32 = C.c_uint32
U64 = C.c_uint64
SZ = C.c_size_t
L = C.c_long

# name -> argtypes, in the order of include/immtsf.h
SIGNATURES = {
    "immtsf_version": [],
    "immtsf_device_supported": [I],
    "immtsf_set_seed_offset_ptr": [P],
    "immtsf_seed_advance": [P, U64, P],
    "immtsf_profile_begin": [I],
    "immtsf_profile_end": [P, P, P, P, P, I],
    "immtsf_gemm_trace": [P],
    "immtsf_csr_build": [P, P, I, I, I, P, P, P, P, P, P, P, P, I, P],
    "immtsf_csr_build_ex": [P, P, I, I, I, P, P, P, P, P, I, P, I, P, P, P, I, P],
    "immtsf_nan_check": [P, SZ, P, I, P],
    "immtsf_zero_pad_rows": [P, I, I, P, I, P],
    "immtsf_gemm": [I, I, I, I, I, F, P, I, P, I, F, P, I, P, P, I, I, P, SZ, P],
    "immtsf_gemm_ex": [I, I, I, I, I, F, P, I, P, I, P, I, P, I, F, P, I, P, I, P, P, I, I, P, SZ, P],
    "immtsf_split_lo": [P, I, I, I, P, I, P, P],
    "immtsf_gemm_group": [I] + [P] * 19 + [P],
    "immtsf_transpose_split": [P, I, I, I, P, I, P, I, P],
    "immtsf_multi_split": [I, P, P, P, P, P, P, P, P, P],
    "immtsf_gemm_plan": [I, I, I, I, I, P, I, P, I, P, I, I],
    "immtsf_colsum": [P, I, I, I, P, F, P, P, SZ, P],
    "immtsf_recavg_pool_fwd": [P, I, P, P, P, I, P, P, P, I, I, I, I, F, U32, U64, P, P, P, P, P, P],
    "immtsf_recavg_pool_bwd": [P, P, P, P, P, P, I, P, P, P, I, P, P, I, I, I, I, U32, U64, P, P, I, P, P, P, P],
    "immtsf_recavg_weights": [P, P, P, I, P, I, I, I, P, P, P, P, P, P, P],
    "immtsf_csr_to_padded": [P, I, P, I, I, I, P, P, P],
    "immtsf_padded_to_csr": [P, P, I, I, I, P, I, P],
    "immtsf_recavg_dls": [P, P, P, P, L, I, P, P],
    "immtsf_time2vec_fwd": [P, P, P, P, P, I, P, I, P, I, P, I, P],
    "immtsf_time2vec_bwd": [P, I, P, P, P, I, P, P, P, P, P, I, P],
    "immtsf_segattn_fwd": [P, P, P, I, I, I, I, I, I, U32, U64, P, P, P],
    "immtsf_segattn_ln_ok": [I, I],
    "immtsf_segattn_ln_fwd": [P, P, P, I, I, I, I, U32, U64, P, P, P, P, F, P, P, P, P, P, P],
    "immtsf_segattn_bwd": [P, P, P, P, P, I, I, I, I, I, I, U32, U64, P, P, P],
    "immtsf_t2vq_attn_fwd": [P, I, P, P, P, P, P, I, P, P, P, P, I, I, I, I, I, I, I, U32, U64, P, P, P, P, P],
    "immtsf_t2vq_bwd_tiles": [I, I, I],
    "immtsf_t2vq_attn_bwd": [P, P, P, P, I, P, P, P, P, P, I, P, P, P, P, I, I, I, I, I, I, I, U32, U64, P, I, P, P, P, SZ, P],
    "immtsf_ln_fwd": [P, I, P, P, P, I, P, P, I, I, F, U32, U64, U32, P, P, P, P],
    "immtsf_ln_bwd": [P, P, I, P, P, P, I, P, P, P, I, I, U32, U64, U32, P, P, P, P, P],
    "immtsf_gru_scan_fwd": [P, P, P, I, I, I, P, P, P, P],
    "immtsf_gr_tail_fwd": [P, P, P, P, P, P, P, P, I, I, I, F, U32, U64, P, P, P],
    "immtsf_gr_tail_bwd": [P, P, P, P, P, P, P, P, I, I, I, F, U32, U64, P, P, P, P, P, P],
    "immtsf_gru_scan_bwd": [P, P, P, P, P, P, I, I, I, P, P, P],
    "immtsf_xattn_core_fwd": [P, I, P, I, P, I, P, I, I, I, I, U32, U64, P, I, P, P],
    "immtsf_xattn_core_bwd": [P, I, P, I, P, I, P, I, P, P, I, I, I, I, U32, U64, P, I, P, I, P, I, P],
    "immtsf_xattn_lowrank_ok": [I, I, I, I],
    "immtsf_xattn_lowrank_fwd": [P, I, P, I, P, I, P, I, I, I, I, I, U32, U64, P, I, P, P],
    "immtsf_xattn_lowrank_bwd": [P, I, P, I, P, I, P, I, P, P, I, I, I, I, I, U32, U64, P, I, P, I, P, P],
    "immtsf_xattn_rank_ok": [I, I, I, I],
    "immtsf_xattn_rank_fwd": [P, I, P, I, P, P, I, I, I, I, I, U32, U64, P, P, P],
    "immtsf_xattn_rank_bwd": [P, P, I, P, I, P, P, I, I, I, I, I, U32, U64, P, I, P, P],
    "immtsf_xattn_rank_fused_ok": [I, I, I, I, I],
    "immtsf_xattn_rank_fused_fwd": [P, I, I, P, I, P, P, I, P, P, P, P, I, I, I, I, I, F, F, U32, U64, P, I, P, P, P, P, P],
    "immtsf_xattn_rank_fused_bwd": [P, P, P, P, I, P, I, P, P, P, I, I, P, I, I, I, I, I, I, F, F, U32, U64, P, I, P, P, P, P, SZ, P],
    "immtsf_gemm_batched": [I, I, I, I, I, F, P, P, I, L, L, P, P, I, L, L, F, P, I, L, L, I, I, P, SZ, P],
    "immtsf_softmax_rows_fwd": [P, P, P, I, I, I, I, F, U32, U64, P],
    "immtsf_softmax_rows_bwd": [P, P, P, P, I, I, I, I, F, U32, U64, P],
    "immtsf_xattn_tail_fwd": [P, P, P, P, P, I, I, I, F, F, U32, U64, P, P, P],
    "immtsf_xattn_tail_bwd": [P, P, P, P, I, I, I, F, F, U32, U64, P, P, P, P],
    "immtsf_masked_mse_partial": [P, P, P, L, I, I, P, P, P, P, SZ, P],
    "immtsf_masked_mse_finalize": [P, P, I, P, P, P],
    "immtsf_masked_mse_bwd": [P, P, P, L, I, P, P, P, P],
    "immtsf_window_count": [P, P, P, P, P, I, P, P],
    "immtsf_exclusive_scan_i32": [P, I, P, P, P],
    "immtsf_window_fill": [P, P, P, P, P, I, P, P, P, P],
    "immtsf_batch_gather": [P, I, I, P, P, P, P, P, I, I, P, I, P, P],
    "immtsf_nvls_allreduce_f32": [P, P, SZ, SZ, SZ, I, I, P, P],
    "immtsf_axpby": [P, F, P, I, SZ, P],
    "immtsf_group_sum_rows": [P, I, I, I, I, P, P],
}

_lib = None


class ImmtsfError(RuntimeError):
    pass


def load():
    """Load libimmtsf.so once; raise loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImmtsfError(
            f"{LIB_PATH} not found: the immtsf CUDA library is not built. "
            "Run `python __graft_entry__.py` (nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = I
    lib.immtsf_last_error_string.argtypes = []
    lib.immtsf_last_error_string.restype = C.c_char_p
    lib.immtsf_launch_count.argtypes = []
    lib.immtsf_launch_count.restype = C.c_ulonglong
    lib.immtsf_gemm_workspace_bytes.argtypes = [I, I, I, I, I]
    lib.immtsf_gemm_workspace_bytes.restype = SZ
    lib.immtsf_t2vq_bwd_workspace_bytes.argtypes = [I, I, I]
    lib.immtsf_t2vq_bwd_workspace_bytes.restype = SZ
    lib.immtsf_xattn_rank_fused_bwd_workspace_bytes.argtypes = [I, I, I, I]
    lib.immtsf_xattn_rank_fused_bwd_workspace_bytes.restype = SZ
    lib.immtsf_nvls_flag_bytes.argtypes = [I]
    lib.immtsf_nvls_flag_bytes.restype = SZ
    lib.immtsf_masked_mse_workspace_bytes.argtypes = [I]
    lib.immtsf_masked_mse_workspace_bytes.restype = SZ
    lib.immtsf_gemm_batched_workspace_bytes.argtypes = [I, I, I, I, I, I, L, L, I, L, L, I, I]
    lib.immtsf_gemm_batched_workspace_bytes.restype = SZ
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load().immtsf_launch_count())


def profile_gemm_tc(max_records: int = 4096):
    """Context manager: device time of every gemm_tc_kernel launch inside the block -> list of (M, N, K, ragged_dim, ms)."""
    import contextlib

    @contextlib.contextmanager
    def cm():
        lib = load()
        out = []
        if lib.immtsf_profile_begin(max_records) != 0:
            raise ImmtsfError("immtsf_profile_begin failed")
        try:
            yield out
        finally:
            IA, FA = C.c_int * max_records, C.c_float * max_records
            m, n, k, rd, ms = IA(), IA(), IA(), IA(), FA()
            cnt = lib.immtsf_profile_end(m, n, k, rd, ms, max_records)
            out.extend((m[i], n[i], k[i], rd[i], ms[i]) for i in range(cnt))

    return cm()


# bench.py: {entry name: []} -> every call of a listed entry point is bracketed by a CUDA-event pair on the current
# (= launching) stream and (ev0, ev1) is appended to its list.  None: no instrumentation.
PROFILE = None


def call(name, *args):
    """Call an entry point; raise ImmtsfError with the library's message on failure."""
    lib = load()
    prof = PROFILE
    if prof is not None and name in prof:
        import torch

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = getattr(lib, name)(*args)
        ev1.record()
        prof[name].append((ev0, ev1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.immtsf_last_error_string()
        raise ImmtsfError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")
    return rc
