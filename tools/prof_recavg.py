"""Driver for ncu captures of the streaming kernels (RecAvg pooling fwd/bwd, CSR build) at B=2048 N<=16 T=24 d=768."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import torch
from immtsf import ops
from sweep_hbm import ragged

B, N, T, d = 2048, 16, 24, 768
notes, tau, sumN = ragged(B, N, d, False)
dev = notes.device
t_hat = (0.5 + 0.5 * torch.rand(B, T)).sort(dim=1)[0].cuda()
ls = torch.tensor(0.0, device=dev)
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
for _ in range(3):
    r = ops.csr_build(notes, tau)
    E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(r.emb_flat, r, t_hat, ls, gamma, beta, T, d, ops.drop_thr(0.1), 1, True)
    dE = torch.randn_like(E_drop)
    ops.recavg_pool_bwd(dE, E_raw, mean, rstd, wsum, r.emb_flat, r, t_hat, ls, gamma, T, d, ops.drop_thr(0.1), 1)
torch.cuda.synchronize()
