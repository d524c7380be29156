"""Micro-benchmark of the tcgen05 3xTF32 GEMM at the cfg2 shapes (not a test).  The lo operands are split once
(ops.LoCache), so the time is the GEMM kernel (+ split-K reduce) alone.  IMMTSF_TC_BN=128|256|512 forces a variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops

def run(M, N, K, tA, tB, iters=30):
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.empty(M, N, device="cuda")
    lo = ops.LoCache()
    flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    for _ in range(3):
        ops.gemm(A, B, C, transA=tA, transB=tB, backend=ops.BACKEND_TC, lo=lo)
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(0.0)  # 256 MiB: cold L2, like the step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, B, C, transA=tA, transB=tB, backend=ops.BACKEND_TC, lo=lo)
        e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / iters
    ref = (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    return ms, 2.0 * M * N * K / ms / 1e9, err

tag_v = os.environ.get("IMMTSF_TC_BN", "auto")
for (M, N, K, tA, tB, tag) in [(6144, 768, 768, False, True, "fwd"), (6144, 1536, 768, False, True, "fwd 2d"), (6144, 768, 768, False, False, "dgrad"),
                               (768, 768, 6144, True, False, "wgrad"), (768, 768, 768, False, True, "fold"), (2176, 768, 768, False, True, "fwd notes"),
                               (2176, 1536, 768, False, True, "kv inproj"), (768, 1152, 2176, True, False, "wgrad notes"),
                               (16384, 768, 4096, False, True, "cfg3 input_proj"), (8192, 8192, 8192, False, True, "square 8k")]:
    ms, tf, err = run(M, N, K, tA, tB)
    print(f"[BN={tag_v}] {tag:16s} M{M} N{N} K{K} tA{int(tA)} tB{int(tB)}: {ms*1e3:8.1f} us {tf:7.1f} TF/s (fp32-exact)  err {err:.2e}", flush=True)
