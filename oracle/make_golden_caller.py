"""Golden vectors from the reference's own CALLER of the fusion path: lib/evaluation.py:72-164 `compute_all_losses(model, fusion,
batch_dict)` -- backbone forecast -> fusion(notes_embeddings, tau, tp_to_predict, pred_y) -> masked MSE -- run UNMODIFIED on CPU
with the unmodified reference FusionModel and a small deterministic stand-in backbone (the 11 forecasters are out of scope;
the caller only needs `model.forecasting(tp_to_predict, observed_data, observed_tp, observed_mask) -> [B, Lp, C]`).

Test infrastructure; runs only in the build container (needs /root/reference).  Output: tests/golden/caller_*.npz with the
batch, the fusion state_dict, the backbone weight, the loss the reference returns and the gradients `loss.backward()`
(main.py:1097) leaves on the fusion parameters and on the backbone weight.

    python oracle/make_golden_caller.py [--out tests/golden] [--ref /root/reference]
"""
from __future__ import annotations

import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import D_MODEL_TABLE, import_reference  # noqa: E402

CASES = [("caller_t2v_xattn", "TTF_T2V_XAttn", "MMF_XAttn_Add"), ("caller_recavg_gr", "TTF_RecAvg", "MMF_GR_Add")]


class Backbone(torch.nn.Module):
    """Stand-in forecaster: pred[b, t, c] = sum_k W[c, k] * mean_l(observed_data[b, l, k] * observed_mask[b, l, k]) + tp[b, t]."""

    def __init__(self, C):
        super().__init__()
        self.W = torch.nn.Parameter(torch.eye(C) * 0.5 + 0.1)

    def forecasting(self, tp_to_predict, observed_data, observed_tp, observed_mask):
        feat = (observed_data * observed_mask).mean(dim=1)  # [B, C]
        return (feat @ self.W.T).unsqueeze(1) + tp_to_predict.unsqueeze(-1)


def make_batch(B, N, T, L, d_model, C, seed):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, N + 1, (B,), generator=g)
    counts[0] = N
    notes, tau = torch.zeros(B, N, d_model), torch.zeros(B, N)
    for b in range(B):
        n = int(counts[b])
        notes[b, :n] = torch.randn(n, d_model, generator=g)
        tau[b, :n] = torch.rand(n, generator=g) * 7.0
    tp = torch.sort(0.5 + 0.5 * torch.rand(B, T, generator=g), dim=1)[0]
    mask = (torch.rand(B, T, C, generator=g) > 0.3).float()
    mask[:, 0, :] = 1.0  # every sample has observations (lib/evaluation.py:128-132 raises otherwise)
    mask[:, :, C - 1] = 0.0  # a variable without any observation: excluded from the mean (:51-62)
    return {
        "notes_embeddings": notes, "tau": tau, "tp_to_predict": tp,
        "observed_data": torch.randn(B, L, C, generator=g), "observed_tp": torch.sort(torch.rand(B, L, generator=g), dim=1)[0],
        "observed_mask": (torch.rand(B, L, C, generator=g) > 0.2).float(),
        "data_to_predict": torch.randn(B, T, C, generator=g), "mask_predicted_data": mask,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    FusionModel = import_reference(a.ref)
    import lib.evaluation as E  # the reference's caller, unmodified

    for name, ttf, mmf in CASES:
        torch.manual_seed(7)
        C, d_txt, H = 4, 32, 1
        args = SimpleNamespace(TTF_module=ttf, MMF_module=mmf, llm_model_fusion="TINY", llm_layers_fusion=1, max_length=1024, device="cpu",
                               use_text_embeddings=True, recency_sigma=1.0, dropout=0.0, d_txt=d_txt, n_heads_fusion=H, C=C, kappa=0.5)
        fusion = FusionModel(args)
        model = Backbone(C)
        fusion.train(); model.train()
        batch = make_batch(6, 5, 8, 9, D_MODEL_TABLE["TINY"], C, seed=11)
        res = E.compute_all_losses(model, fusion, batch)  # lib/evaluation.py:72
        res["loss"].backward()  # main.py:1097
        out = {"meta": np.array([ttf, mmf, str(d_txt), str(C), str(H), "0.5"]), "loss": np.array(float(res["loss"])),
               "mse": np.array(res["mse"]), "backbone:W": model.W.detach().numpy(), "grad_backbone:W": model.W.grad.numpy()}
        for k, v in batch.items():
            out["batch:" + k] = v.numpy()
        for k, v in fusion.state_dict().items():
            out["param:" + k] = v.numpy()
        for k, p in fusion.named_parameters():
            out["grad:" + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        np.savez_compressed(os.path.join(a.out, name + ".npz"), **out)
        print(name, "loss", float(res["loss"]))


if __name__ == "__main__":
    main()
