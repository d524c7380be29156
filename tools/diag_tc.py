import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops
for (M, N, K) in [(128, 128, 32), (128, 128, 96), (128, 128, 128), (128, 128, 256), (256, 256, 768), (6144, 768, 768)]:
    for tA, tB in [(0, 1), (0, 0), (1, 0), (1, 1)]:
        g = torch.Generator().manual_seed(1)
        A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
        B = (torch.randn((N, K) if tB else (K, N), generator=g) * 0.3).cuda()
        C = torch.empty(M, N, device="cuda")
        ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=ops.BACKEND_TC)
        torch.cuda.synchronize()
        ref = (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())
        err = (C.double() - ref).abs()
        rel = err.max().item() / ref.abs().max().item()
        # where are the errors: per 32x32 block max
        bad = (err > 1e-4 * ref.abs().max()).float()
        print(f"M{M} N{N} K{K} tA{tA} tB{tB}: rel {rel:.3e} bad_frac {bad.mean().item():.3f}",
              "rows_bad", bad.sum(1).nonzero().numel(), "cols_bad", bad.sum(0).nonzero().numel())
