"""Golden vectors for the per-(note, query) Time2Vec attention, produced by the reference's own
class ``fusions/TTF_T2V_XAttn_old.py: TTF_T2V_XAttn`` on CPU.

Test infrastructure; runs only in the build container (the reference tree does not exist on the GPU box), output is
committed under tests/golden/pq_*.npz.  Usage:  python oracle/make_golden_perquery.py

The file is dead code in the reference as shipped: it imports ``get_d_txt`` from ``fusions.load_llm``
(TTF_T2V_XAttn_old.py:4), a name load_llm.py no longer defines (it was renamed get_d_model and calls the HF hub).  The one
patch applied is therefore the same kind make_golden.py applies: ``load_llm.get_d_txt`` is set to a table lookup before
the module is imported.  Nothing else of the reference is touched.

Each case stores the state_dict (keys prefixed ``ttf.``), the inputs, eval-mode E_txt / M_txt in fp32 and from the fp64
copy of the module, and -- train mode, dropout 0 -- the gradient of every parameter for the loss sum(E_txt * G).
"""
from __future__ import annotations

import argparse
import copy
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import D_MODEL_TABLE, make_inputs, randomise_  # noqa: E402

CASES = [
    # name, H, B, N, T, t_hat_1d, no_note_sample
    ("pq_h1", 1, 5, 6, 7, False, False),
    ("pq_h4", 4, 4, 7, 5, False, False),
    ("pq_h2_t1d", 2, 3, 4, 9, True, False),
    ("pq_h1_nonote", 1, 4, 5, 6, False, True),
]


def import_reference(ref_root: str):
    sys.path.insert(0, ref_root)
    import fusions.load_llm as load_llm

    load_llm.get_d_txt = lambda alias: D_MODEL_TABLE[alias]
    return importlib.import_module("fusions.TTF_T2V_XAttn_old").TTF_T2V_XAttn


def run_case(cls, case, out_dir):
    name, H, B, N, T, t1d, no_note = case
    gen = torch.Generator().manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(name)))
    torch.manual_seed(4321)
    m = cls("TINY", 1, max_length=1024, device="cpu", use_text_embeddings=True, n_heads_fusion=H, dropout=0.0)
    randomise_(m, gen)
    d = D_MODEL_TABLE["TINY"]
    notes, tau, t_hat, _, _ = make_inputs(gen, B, N, T, d, 1, t1d, no_note)
    G = torch.randn(B, T, d, generator=gen)
    out = {"param:ttf." + k: v.detach().numpy().copy() for k, v in m.state_dict().items()}
    out["in:notes"], out["in:tau"], out["in:t_hat"], out["in:G"] = notes.numpy(), tau.numpy(), t_hat.numpy(), G.numpy()
    out["meta"] = np.array(["TTF_T2V_XAttn_old", "-", "None", "0", str(H), "0", str(int(no_note))])
    m.eval()
    with torch.no_grad():
        E, M = m(notes, tau, t_hat)
    out["eval:E_txt"], out["eval:M_txt"] = E.numpy(), M.numpy()
    m64 = copy.deepcopy(m).double()
    with torch.no_grad():
        E64, _ = m64(notes.double(), tau.double(), t_hat.double())
    out["eval64:E_txt"] = E64.numpy()
    if not no_note:
        for tag, model, cast in (("grad", m, torch.float32), ("grad64", m64, torch.float64)):
            model.train()
            model.zero_grad()
            E, _ = model(notes.to(cast), tau.to(cast), t_hat.to(cast))
            (E * G.to(cast)).sum().backward()
            for k, p_ in model.named_parameters():
                g = p_.grad if p_.grad is not None else torch.zeros_like(p_)
                out[f"{tag}:ttf.{k}"] = g.numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return name, float(np.abs(out["eval:E_txt"]).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    cls = import_reference(a.ref)
    torch.set_num_threads(1)
    for case in CASES:
        print(run_case(cls, case, a.out))


if __name__ == "__main__":
    main()
