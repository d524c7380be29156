"""Does tcgen05.mma kind::tf32 truncate fp32 operands itself?  Run twice (IMMTSF_TC_RAWHI=0/1) and compare
the saved outputs bitwise (the env var is read once per process)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops

out = {}
for (M, N, K, tA, tB) in [(512, 256, 768, False, True), (512, 256, 768, False, False), (256, 384, 1024, True, False), (300, 200, 136, True, True)]:
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tB else (K, N), generator=g).cuda()
    C = torch.empty(M, N, device="cuda")
    ops.gemm(A, B, C, transA=tA, transB=tB, backend=ops.BACKEND_TC)
    ref = (A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())
    err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
    out[f"{M}x{N}x{K}_{int(tA)}{int(tB)}"] = C.cpu()
    print(f"RAWHI={os.environ.get('IMMTSF_TC_RAWHI','0')} {M}x{N}x{K} tA={tA} tB={tB}: rel err vs fp64 {err:.3e}")
path = sys.argv[1]
if os.path.exists(path):
    prev = torch.load(path)
    for k in out:
        same = torch.equal(prev[k], out[k])
        print(k, "bitwise equal to previous run:", same, "max abs diff", (prev[k] - out[k]).abs().max().item())
else:
    torch.save(out, path)
