"""Helper run as a subprocess by test_gpu_gemm_tc.py with IMMTSF_TC_BN=128|256|512 (512 = CTA pairs): the tile-width override is read
once per process.  Checks the forced variant against fp64 on shapes that exercise N tails, split-K, ragged
bounds, all transpositions and the epilogue (alpha/beta/bias)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from immtsf import ops  # noqa: E402

TOL = 4e-6
worst = 0.0
for (M, N, K) in [(128, 256, 64), (300, 200, 136), (384, 300, 40), (256, 520, 768), (768, 768, 6144), (2200, 1152, 768), (6144, 768, 768)]:
    for tA in (0, 1):
        for tB in (0, 1):
            g = torch.Generator().manual_seed(M + 3 * N + 7 * K + 2 * tA + tB)
            A = torch.randn((K, M) if tA else (M, K), generator=g).cuda()
            B = (torch.randn((N, K) if tB else (K, N), generator=g) * 0.3).cuda()
            bias = torch.randn(N, generator=g).cuda()
            C0 = torch.randn(M, N, generator=g).cuda()
            C = C0.clone()
            ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), bias=bias, alpha=0.5, beta=2.0, backend=ops.BACKEND_TC)
            ref = 0.5 * ((A.double().T if tA else A.double()) @ (B.double().T if tB else B.double())) + 2.0 * C0.double() + bias.double()
            err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
            worst = max(worst, err)
            assert err <= TOL, (M, N, K, tA, tB, err)
# ragged rows and ragged contraction
M, N, K, m = 640, 384, 128, 300
g = torch.Generator().manual_seed(10)
A = torch.randn(M, K, generator=g).cuda()
A[m:] = 0.0
W = torch.randn(N, K, generator=g).cuda()
m_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
out = torch.full((M, N), 7.0, device="cuda")
ops.gemm(A, W, out, transB=True, ragged=m_dev, ragged_dim=1, backend=ops.BACKEND_TC)
ref = A.double() @ W.double().T
assert ((out[:m].double() - ref[:m]).abs().max() / ref.abs().max()).item() <= TOL
tile = 256 if os.environ.get("IMMTSF_TC_BN") == "512" else 128  # pad rows are zeroed up to the end of the last touched tile
edge = (m + tile - 1) // tile * tile
assert (out[m:edge] == 0).all() and (out[edge:] == 7.0).all()
dy = torch.randn(M, N, generator=g).cuda()
dy[m:] = 0.0
dw = torch.empty(N, K, device="cuda")
ops.gemm(dy, A, dw, transA=True, ragged=m_dev, ragged_dim=2, backend=ops.BACKEND_TC)
refw = dy[:m].double().T @ A[:m].double()
assert ((dw.double() - refw).abs().max() / refw.abs().max()).item() <= TOL
print(f"OK IMMTSF_TC_BN={os.environ.get('IMMTSF_TC_BN')} worst rel err {worst:.3e}")
