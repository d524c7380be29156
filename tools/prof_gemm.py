"""One GEMM shape for an ncu capture of the tcgen05 kernel (not a test):  python tools/prof_gemm.py M N K tA tB"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops

M, N, K, tA, tB = (int(x) for x in sys.argv[1:6])
A = torch.randn((K, M) if tA else (M, K), device="cuda")
B = torch.randn((N, K) if tB else (K, N), device="cuda")
C = torch.empty(M, N, device="cuda")
lo = ops.LoCache()
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for _ in range(3):
    flush.fill_(0.0)
    ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=ops.BACKEND_TC, lo=lo)
torch.cuda.synchronize()
