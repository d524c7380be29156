"""Diagnostic (not a test): where does fp32 error come from at width 768?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_common as G
from immtsf import ops
from oracle import immtsf_oracle as O

torch.backends.cuda.matmul.allow_tf32 = False
g = torch.Generator().manual_seed(1)
for (M, N, K) in [(768, 16, 772), (768, 768, 768), (6144, 768, 768)]:
    A = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) * 0.3
    ref = A.double() @ W.double().T
    C = torch.empty(M, N).cuda()
    ops.gemm(A.cuda(), W.cuda(), C, transB=True, backend=1)
    e1 = (C.cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    e2 = ((A.cuda() @ W.cuda().T).cpu().double() - ref).abs().max().item() / ref.abs().max().item()
    e3 = ((A @ W.T).double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"gemm {M}x{N}x{K}: ffma {e1:.2e}  cublas {e2:.2e}  cpu {e3:.2e}")

for ttf, mmf in [("TTF_RecAvg", "MMF_GR_Add"), ("TTF_T2V_XAttn", "MMF_GR_Add"), ("TTF_RecAvg", "MMF_XAttn_Add")]:
    cfg = dict(ttf=ttf, mmf=mmf, d_txt=768, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 768, dropout=0.0, seed=31)
    G.randomise_(fm, 32)
    notes, tau, t_hat, Y, Gw = G.synth_batch(32, 16, 24, 768, 4, 33)
    params = {k: v.detach().cpu() for k, v in fm.state_dict().items()}
    r64 = G.oracle_run(cfg, params, notes, tau, t_hat, Y, Gw, grads=False)
    r32 = G.oracle_run(cfg, params, notes, tau, t_hat, Y, Gw, dtype=torch.float32, grads=False)
    fm.eval()
    with torch.no_grad():
        E, M = fm.ttf(notes.cuda(), tau.cuda(), t_hat.cuda())
        Yo = fm(notes.cuda(), tau.cuda(), t_hat.cuda(), Y.cuda())
        Yo_from_ref_E = fm.mmf(Y.cuda(), r64["E_txt"].float().cuda(), M)
    rel = lambda a, b: (a.double().cpu() - b.double()).abs().max().item() / b.abs().max().item()
    print(ttf, mmf, f"E_txt: gpu {rel(E, r64['E_txt']):.2e} oracle32 {rel(r32['E_txt'], r64['E_txt']):.2e} | "
          f"Y_out: gpu {rel(Yo, r64['Y_out']):.2e} oracle32 {rel(r32['Y_out'], r64['Y_out']):.2e} | "
          f"mmf only (exact E): gpu {rel(Yo_from_ref_E, r64['Y_out']):.2e}")
