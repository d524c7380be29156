"""Golden vectors for the masked-MSE training loss from the UNMODIFIED reference function
lib/evaluation.py:17-69 `compute_error(truth, pred, mask, "MSE", "mean" | "sum")`, run on CPU in the build container.
Test infrastructure; its output (tests/golden/loss_mse.npz) is committed.

    python oracle/make_golden_loss.py [--out tests/golden] [--ref /root/reference]

Cases: dense mask, ragged mask, a variable without any observation (count 0), a 4-D prediction with n_traj_samples = 1.
Stored per case: pred, truth, mask, the loss (fp32 and fp64), d loss / d pred (fp64), and the 'sum' reduction
(per-variable error sums and counts)."""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    sys.path.insert(0, a.ref)
    from lib.evaluation import compute_error  # the reference, unmodified

    g = torch.Generator().manual_seed(20260)
    out = {}
    cases = [("dense", 6, 9, 4, 1.0, None), ("ragged", 8, 11, 5, 0.45, None), ("novar", 5, 7, 6, 0.5, 2), ("one", 1, 1, 1, 1.0, None)]
    for name, B, T, C, keep, dead in cases:
        pred = torch.randn(B, T, C, generator=g)
        truth = torch.randn(B, T, C, generator=g)
        mask = (torch.rand(B, T, C, generator=g) < keep).float()
        mask[:, 0, :] = 1.0 if name != "novar" else mask[:, 0, :]
        if dead is not None:
            mask[:, :, dead] = 0.0
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            p = pred.to(dt).clone().requires_grad_(True)
            loss = compute_error(truth.to(dt), p, mask.to(dt), "MSE", "mean")
            loss.backward()
            out[f"{name}:loss{tag}"] = loss.detach().numpy()
            out[f"{name}:dpred{tag}"] = p.grad.numpy()
        s, c = compute_error(truth.double(), pred.double().unsqueeze(0), mask.double(), "MSE", "sum")
        out[f"{name}:sum"], out[f"{name}:count"] = s.numpy(), c.numpy()
        out[f"{name}:pred"], out[f"{name}:truth"], out[f"{name}:mask"] = pred.numpy(), truth.numpy(), mask.numpy()
    os.makedirs(a.out, exist_ok=True)
    np.savez_compressed(os.path.join(a.out, "loss_mse.npz"), **out)
    print("wrote", os.path.join(a.out, "loss_mse.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
