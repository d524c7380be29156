"""Helpers shared by the parity tests (test infrastructure)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Fusion-path cases (oracle/make_golden.py); loss_mse.npz (oracle/make_golden_loss.py) and pq_*.npz
    (oracle/make_golden_perquery.py) and store_chunks.npz (oracle/make_golden_store.py) have their own tests."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("loss_", "pq_", "store_", "caller_"))]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = [str(x) for x in z["meta"]]
    cfg = dict(
        ttf=meta[0], mmf=meta[1], d_txt=None if meta[2] == "None" else int(meta[2]),
        C=int(meta[3]), H=int(meta[4]), kappa=float(meta[5]), no_note=bool(int(meta[6])),
    )
    params = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    inputs = {k[len("in:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
    rest = {k: z[k] for k in z.files if not (k.startswith("param:") or k.startswith("in:") or k == "meta")}
    return cfg, params, inputs, rest


def rel_max(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (SURVEY.md 8c metric)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.abs().max().item()
    num = (a - b).abs().max().item()
    return num / den if den > 0 else num
