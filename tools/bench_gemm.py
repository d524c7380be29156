"""Micro-benchmark of immtsf_gemm backends at the cfg2 shapes (not a test)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops

def run(M, N, K, tA, tB, backend, iters=20):
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    for _ in range(3):
        ops.gemm(A, B, C, transA=tA, transB=tB, backend=backend)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(A, B, C, transA=tA, transB=tB, backend=backend)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, 2.0 * M * N * K / ms / 1e9

torch.backends.cuda.matmul.allow_tf32 = False
for (M, N, K, tA, tB, tag) in [(6144, 768, 768, False, True, "fwd"), (6144, 768, 768, False, False, "dgrad"),
                               (768, 768, 6144, True, False, "wgrad"), (2176, 768, 768, False, True, "fwd notes"),
                               (2176, 1536, 768, False, True, "kv inproj"), (768, 1152, 2176, True, False, "wgrad notes"),
                               (16384, 768, 4096, False, True, "cfg3 input_proj")]:
    f = run(M, N, K, tA, tB, ops.BACKEND_FFMA)
    t = run(M, N, K, tA, tB, ops.BACKEND_TC)
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda")
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    (A @ B.T); e0.record()
    for _ in range(20): (A @ B.T)
    e1.record(); torch.cuda.synchronize(); cb = e0.elapsed_time(e1) / 20
    print(f"{tag:16s} M{M} N{N} K{K}: ffma {f[0]*1e3:7.1f} us {f[1]:6.1f} TF/s | tc3x {t[0]*1e3:7.1f} us {t[1]:6.1f} TF/s | cublas fp32 {cb*1e3:7.1f} us {2.0*M*N*K/cb/1e9:6.1f} TF/s")
