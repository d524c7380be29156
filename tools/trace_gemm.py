"""Phase timeline of the CTA-pair tcgen05 GEMM (immtsf_gemm_trace): python tools/trace_gemm.py M N K tA tB [cold]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
os.environ.setdefault("IMMTSF_TC_BN", "512")
import torch
from immtsf import ops, _lib

M, N, K, tA, tB = (int(x) for x in sys.argv[1:6])
cold = len(sys.argv) > 6 and sys.argv[6] == "cold"
A = torch.randn((K, M) if tA else (M, K), device="cuda")
B = torch.randn((N, K) if tB else (K, N), device="cuda")
C = torch.empty(M, N, device="cuda")
lo = ops.LoCache()
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
ncta = 2 * ((M + 255) // 256) * ((N + 255) // 256) * 16
buf = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=ops.BACKEND_TC, lo=lo)
if cold:
    flush.fill_(0.0)
torch.cuda.synchronize()
_lib.call("immtsf_gemm_trace", buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=ops.BACKEND_TC, lo=lo)
e1.record()
torch.cuda.synchronize()
_lib.call("immtsf_gemm_trace", None)
t = buf.view(-1, 8).cpu()
t = t[t[:, 0] != 0]
d = lambda a, b, rows=t: (rows[:, a] - rows[:, b]).double()
lead = t[t[:, 2] != 0]
print(f"M{M} N{N} K{K} tA{tA} tB{tB} {'cold' if cold else 'warm'} L2: event {e0.elapsed_time(e1)*1e3:.1f} us, {t.shape[0]} CTAs traced")
def st(name, x):
    print(f"  {name:34s} mean {x.mean():9.0f}  min {x.min():9.0f}  max {x.max():9.0f} clk")
st("prologue (entry -> sync done)", d(1, 0))
st("first operands (sync -> full[0])", d(2, 1, lead))
st("mainloop issue (full[0] -> issued)", d(3, 2, lead))
st("accumulator complete - sync", d(4, 1))
st("epilogue tail (acc done -> stored)", d(5, 4))
st("exit sync (stored -> exit)", d(6, 5))
st("whole CTA", d(6, 0))
g = t[:, 7].double()
print(f"  CTA entry spread (globaltimer): {g.max() - g.min():.0f} ns")
