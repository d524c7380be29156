"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED
reference modules (/root/reference/fusions) on CPU.

Test infrastructure.  Runs only in the build container (the reference tree does
not exist on the GPU box); its output is committed.  Usage:

    python oracle/make_golden.py [--out tests/golden] [--ref /root/reference]

The only patch applied to the reference is ``fusions.load_llm.get_d_model``
(fusions/load_llm.py:16-35 calls the HF hub; no network here) which is replaced
by a table lookup *before* the TTF modules bind the name at import
(fusions/TTF_RecAvg.py:4, fusions/TTF_T2V_XAttn.py:4).  The table maps the
alias ``"TINY"`` to 48 so fixtures stay small; widths do not change semantics.

Each case stores: the state_dict, the inputs, eval-mode outputs (Y_out, E_txt,
M_txt) in fp32 and from the fp64 copy of the module, and -- for train mode with
dropout 0.0 -- the gradients of every parameter and of Y_ts for the loss
``sum(Y_out * G)`` with a stored random ``G``.
"""
from __future__ import annotations

import argparse
import copy
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

D_MODEL_TABLE = {"TINY": 48, "GPT2": 768, "BERT": 768, "GPT2M": 1024, "GPT2L": 1280,
                 "GPT2XL": 1600, "Llama": 4096, "DeepSeek": 4096}

CASES = [
    # name, TTF, MMF, d_txt, C, H, kappa, B, N, T, t_hat_1d, no_note_sample
    ("recavg_gr", "TTF_RecAvg", "MMF_GR_Add", 32, 4, 1, 0.5, 5, 6, 7, False, False),
    ("recavg_xattn", "TTF_RecAvg", "MMF_XAttn_Add", None, 3, 2, 0.5, 4, 5, 6, True, False),
    ("t2v_xattn", "TTF_T2V_XAttn", "MMF_XAttn_Add", 32, 4, 1, 0.5, 5, 6, 7, False, False),
    ("t2v_gr", "TTF_T2V_XAttn", "MMF_GR_Add", None, 5, 4, 1.0, 4, 7, 5, False, False),
    ("t2v_xattn_h4", "TTF_T2V_XAttn", "MMF_XAttn_Add", 32, 6, 4, 2.0, 3, 4, 9, True, False),
    # a sample with zero notes: forward only (reference backward is NaN there)
    ("recavg_gr_nonote", "TTF_RecAvg", "MMF_GR_Add", 32, 4, 1, 0.5, 4, 5, 6, False, True),
    ("t2v_xattn_nonote", "TTF_T2V_XAttn", "MMF_XAttn_Add", 32, 4, 2, 0.5, 4, 5, 6, False, True),
    ("t2v_gr_nonote", "TTF_T2V_XAttn", "MMF_GR_Add", 32, 4, 1, 0.5, 4, 5, 6, False, True),
    ("recavg_xattn_nonote", "TTF_RecAvg", "MMF_XAttn_Add", 32, 4, 1, 0.5, 4, 5, 6, False, True),
]


def import_reference(ref_root: str):
    sys.path.insert(0, ref_root)
    import fusions.load_llm as load_llm  # noqa: E402

    def get_d_model(alias):
        return D_MODEL_TABLE[alias]

    load_llm.get_d_model = get_d_model
    from fusions.FusionModel import FusionModel  # noqa: E402

    return FusionModel


def make_inputs(gen, B, N, T, d_model, C, t_hat_1d, no_note, history=7.0, pred=7.0):
    """Time-IMM-shaped synthetic batch honouring the reference's invariants
    (lib/parse_datasets.py:209-213, 318, 345, 792, 811-819)."""
    counts = torch.randint(1, N + 1, (B,), generator=gen)
    counts[0] = N  # one full sample
    if B > 1:
        counts[1] = 1  # one single-note sample
    if no_note:
        counts[B - 1] = 0
    notes = torch.zeros(B, N, d_model)
    tau = torch.zeros(B, N)
    for b in range(B):
        n = int(counts[b])
        notes[b, :n] = torch.randn(n, d_model, generator=gen)
        tau[b, :n] = torch.rand(n, generator=gen) * history  # unsorted on purpose
    if B > 2 and N >= 3 and counts[0] == N:
        notes[0, 1] = 0.0  # an all-zero "real" row mid-sequence => masked by content
    if t_hat_1d:
        t_hat = torch.sort(history / (history + pred) + torch.rand(T, generator=gen) * pred / (history + pred))[0]
    else:
        t_hat = torch.zeros(B, T)
        for b in range(B):
            tl = int(torch.randint((T + 2) // 3, T + 1, (1,), generator=gen))
            if b == 0:
                tl = T
            v = history / (history + pred) + torch.rand(tl, generator=gen) * pred / (history + pred)
            t_hat[b, :tl] = torch.sort(v)[0]
    Y = torch.randn(B, T, C, generator=gen)
    G = torch.randn(B, T, C, generator=gen)
    return notes, tau, t_hat, Y, G


def randomise_(module, gen):
    """Move every parameter off its init so LayerNorm gains, biases, Q_param and
    log-sigma all carry signal."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("log_recency_sigma"):
                p.copy_(torch.tensor(-1.2))  # sigma ~ 0.3: weights are not all ~1
            elif p.dim() >= 2:
                p.add_(torch.randn(p.shape, generator=gen) * 0.05)
            else:
                p.add_(torch.randn(p.shape, generator=gen) * 0.1)
            if "time2vec.periodic.weight" in name:
                p.mul_(3.0)


def run_case(FusionModel, case, out_dir):
    (name, ttf, mmf, d_txt, C, H, kappa, B, N, T, t1d, no_note) = case
    gen = torch.Generator().manual_seed(sum(ord(c) * (i + 1) for i, c in enumerate(name)))
    torch.manual_seed(1234)
    args = SimpleNamespace(
        TTF_module=ttf, MMF_module=mmf, llm_model_fusion="TINY", llm_layers_fusion=1,
        max_length=1024, device="cpu", use_text_embeddings=True, recency_sigma=1.0,
        dropout=0.0, d_txt=d_txt, n_heads_fusion=H, C=C, kappa=kappa,
    )
    fm = FusionModel(args)
    randomise_(fm, gen)
    d_model = D_MODEL_TABLE["TINY"]
    notes, tau, t_hat, Y, G = make_inputs(gen, B, N, T, d_model, C, t1d, no_note)

    out = {}
    for k, v in fm.state_dict().items():
        out["param:" + k] = v.detach().numpy().copy()
    out["in:notes"], out["in:tau"], out["in:t_hat"] = notes.numpy(), tau.numpy(), t_hat.numpy()
    out["in:Y_ts"], out["in:G"] = Y.numpy(), G.numpy()
    out["meta"] = np.array([ttf, mmf, str(d_txt), str(C), str(H), str(kappa), str(int(no_note))])

    fm.eval()
    with torch.no_grad():
        E, M = fm.ttf(notes, tau, t_hat)
        Yo = fm(notes, tau, t_hat, Y)
    out["eval:E_txt"], out["eval:M_txt"], out["eval:Y_out"] = E.numpy(), M.numpy(), Yo.numpy()
    fm64 = copy.deepcopy(fm).double()
    with torch.no_grad():
        E64, _ = fm64.ttf(notes.double(), tau.double(), t_hat.double())
        Yo64 = fm64(notes.double(), tau.double(), t_hat.double(), Y.double())
    out["eval64:E_txt"], out["eval64:Y_out"] = E64.numpy(), Yo64.numpy()

    if not no_note:
        for tag, model, cast in (("grad", fm, torch.float32), ("grad64", fm64, torch.float64)):
            model.train()
            model.zero_grad()
            Yr = Y.to(cast).clone().requires_grad_(True)
            Yo = model(notes.to(cast), tau.to(cast), t_hat.to(cast), Yr)
            (Yo * G.to(cast)).sum().backward()
            out[f"{tag}:Y_out"] = Yo.detach().numpy()
            out[f"{tag}:Y_ts"] = Yr.grad.numpy()
            for k, p_ in model.named_parameters():
                g = p_.grad if p_.grad is not None else torch.zeros_like(p_)
                out[f"{tag}:{k}"] = g.numpy()
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **out)
    return name, float(np.abs(out["eval:Y_out"]).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    FusionModel = import_reference(a.ref)
    torch.set_num_threads(1)  # deterministic reduction order for the fixtures
    for case in CASES:
        print(run_case(FusionModel, case, a.out))


if __name__ == "__main__":
    main()
