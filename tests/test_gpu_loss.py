"""Fused masked-MSE loss (immtsf/loss.py, csrc/loss.cu) against the golden vectors produced by the
reference's own compute_error (lib/evaluation.py:17-69) and against the oracle at the path's sizes."""
import os

import numpy as np
import pytest
import torch

import gpu_common as G
from oracle import immtsf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "loss_mse.npz")


@pytest.mark.parametrize("name", ["dense", "ragged", "novar", "one"])
def test_masked_mse_matches_reference_golden(name):
    from immtsf import loss as L

    g = np.load(GOLD)
    pred, truth, mask = (torch.from_numpy(g[f"{name}:{k}"]).cuda() for k in ("pred", "truth", "mask"))
    p = pred.clone().requires_grad_(True)
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    val = L.masked_mse(p, truth, mask, empty_flag=flag)
    val.backward()
    ref = float(g[f"{name}:loss64"])
    assert abs(float(val) - ref) <= 2e-6 * abs(ref)
    G.assert_close("dpred", p.grad.cpu(), g[f"{name}:dpred64"], 2e-6)
    assert int(flag.item()) == int((mask.reshape(mask.shape[0], -1).sum(1) == 0).any())


@pytest.mark.parametrize("B,T,C", [(256, 24, 4), (256, 192, 96), (3, 1, 128), (1000, 7, 5)])
def test_masked_mse_vs_oracle_sizes_determinism_and_flags(B, T, C):
    from immtsf import loss as L

    gen = torch.Generator().manual_seed(B + T + C)
    pred, truth = torch.randn(B, T, C, generator=gen), torch.randn(B, T, C, generator=gen)
    mask = (torch.rand(B, T, C, generator=gen) < 0.4).float()
    mask[B // 2] = 0.0  # a sample without any observed target: the reference raises (:128-132), here a flag
    if C > 2:
        mask[:, :, 1] = 0.0  # a variable without any observation
    p64 = pred.double().requires_grad_(True)
    ref = O.masked_mse(truth.double(), p64, mask.double())
    ref.backward()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    vals = []
    for _ in range(2):
        p = pred.cuda().requires_grad_(True)
        v = L.masked_mse(p, truth.cuda(), mask.cuda(), empty_flag=flag)
        (3.0 * v).backward()
        vals.append((float(v), p.grad.clone()))
    assert abs(vals[0][0] - float(ref)) <= 2e-6 * abs(float(ref))
    G.assert_close("dpred", vals[0][1].cpu() / 3.0, p64.grad, 3e-6)
    assert vals[0][0] == vals[1][0] and torch.equal(vals[0][1], vals[1][1])  # deterministic reduction order
    assert int(flag.item()) == 1
    # 4-D prediction with one trajectory sample, as models return it (lib/evaluation.py:21-23)
    p4 = pred.cuda().unsqueeze(0).requires_grad_(True)
    v4 = L.masked_mse(p4, truth.cuda(), mask.cuda())
    assert float(v4) == vals[0][0]
    (3.0 * v4).backward()  # the gradient comes back in the caller's 4-D layout
    assert tuple(p4.grad.shape) == (1,) + tuple(pred.shape) and torch.equal(p4.grad[0], vals[0][1])


def test_masked_mse_shares_add_up_under_sharding():
    """With the GLOBAL per-variable counts (what the all-reduce of C floats provides) each shard returns its share."""
    from immtsf import _lib, ops

    gen = torch.Generator().manual_seed(5)
    B, T, C = 64, 12, 6
    pred, truth = torch.randn(B, T, C, generator=gen).cuda(), torch.randn(B, T, C, generator=gen).cuda()
    mask = (torch.rand(B, T, C, generator=gen) < 0.5).float().cuda()
    whole = float(O.masked_mse(truth.double().cpu(), pred.double().cpu(), mask.double().cpu()))
    cnt = mask.reshape(-1, C).sum(0)
    lib = _lib.load()
    ws = ops._workspace(pred.device, lib.immtsf_masked_mse_workspace_bytes(C))
    ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
    total = 0.0
    for a, b in ((0, 20), (20, 64)):
        ec = torch.empty(2 * C, device="cuda")
        p, t, m = pred[a:b].contiguous(), truth[a:b].contiguous(), mask[a:b].contiguous()
        _lib.call("immtsf_masked_mse_partial", p.data_ptr(), t.data_ptr(), m.data_ptr(), (b - a) * T, T, C, ec.data_ptr(), None,
                  ticket.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream())
        loss, scale = torch.empty((), device="cuda"), torch.empty(C, device="cuda")
        _lib.call("immtsf_masked_mse_finalize", ec.data_ptr(), cnt.data_ptr(), C, loss.data_ptr(), scale.data_ptr(), ops._stream())
        total += float(loss)
    assert abs(total - whole) <= 3e-6 * abs(whole)
    assert int(ticket.item()) == 0
