"""CPU tests of row f3 (per-(note, query) Time2Vec attention, fusions/TTF_T2V_XAttn_old.py semantics):
the oracle restatement against golden vectors produced by the reference's own class, and the schedule model of the CUDA
path (oracle/perquery_schedule.py: kernel contracts + the GEMM composition around them) against the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, rel_max
from oracle import immtsf_oracle as O
from oracle import perquery_schedule as S

PQ = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "pq_*.npz")))


def load_pq(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = [str(x) for x in z["meta"]]
    params = {k[len("param:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param:")}
    inputs = {k[len("in:"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in:")}
    rest = {k: z[k] for k in z.files if not (k.startswith("param:") or k.startswith("in:") or k == "meta")}
    return dict(H=int(meta[4]), no_note=bool(int(meta[6]))), params, inputs, rest


def test_golden_present():
    assert len(PQ) >= 4


@pytest.mark.parametrize("name", PQ)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-6), (torch.float64, 1e-12)])
def test_oracle_matches_reference_golden(name, dtype, tol):
    cfg, params, inp, rest = load_pq(name)
    P = {k: v.to(dtype).clone().requires_grad_(True) for k, v in params.items()}
    E, M = O.ttf_t2v_xattn_perquery(P, inp["notes"].to(dtype), inp["tau"].to(dtype), inp["t_hat"].to(dtype), n_heads=cfg["H"])
    tag = "eval" if dtype == torch.float32 else "eval64"
    assert rel_max(E.detach(), rest[tag + ":E_txt"]) <= tol
    assert np.array_equal(M.numpy(), rest["eval:M_txt"])
    if cfg["no_note"]:
        return
    (E * inp["G"].to(dtype)).sum().backward()
    gtag = "grad" if dtype == torch.float32 else "grad64"
    gmax = max(np.abs(rest[f"{gtag}:{k}"]).max() for k in P)
    for k, v in P.items():
        ref = rest[f"{gtag}:{k}"]
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        err = (got.double() - torch.from_numpy(ref).double()).abs().max().item()
        assert err <= (5e-5 if dtype == torch.float32 else 1e-10) * max(np.abs(ref).max(), 1e-3 * gmax), (k, err)


def _rand_case(seed, B, N, T, d_model, d_txt, H, no_note=False, t1d=False):
    g = torch.Generator().manual_seed(seed)
    d = d_txt if d_txt is not None else d_model
    shapes = O.param_shapes("TTF_T2V_XAttn", "MMF_GR_Add", d_model, d_txt, 2)
    P = {k: torch.randn(s, generator=g, dtype=torch.float64) * (0.2 if len(s) >= 2 else 0.3) for k, s in shapes.items() if k.startswith("ttf.")}
    P["ttf.layer_norm.weight"] += 1.0
    P["ttf.time2vec.periodic.weight"] *= 8.0
    counts = torch.randint(1, N + 1, (B,), generator=g)
    counts[0] = N
    if no_note:
        counts[-1] = 0
    notes = torch.zeros(B, N, d_model, dtype=torch.float64)
    tau = torch.zeros(B, N, dtype=torch.float64)
    for b in range(B):
        n = int(counts[b])
        notes[b, :n] = torch.randn(n, d_model, generator=g, dtype=torch.float64)
        tau[b, :n] = torch.rand(n, generator=g, dtype=torch.float64) * 1.2  # lags of both signs: the clamp is exercised
    if N >= 3:
        notes[0, 1] = 0.0
    t_hat = torch.rand(T, generator=g, dtype=torch.float64) if t1d else torch.rand(B, T, generator=g, dtype=torch.float64)
    G = torch.randn(B, T, d, generator=g, dtype=torch.float64)
    return P, notes, tau, t_hat, G


@pytest.mark.parametrize("H,d_txt,p,t1d", [(1, None, 0.0, False), (4, 16, 0.0, True), (2, 16, 0.25, False), (1, None, 0.25, False)])
def test_schedule_equals_oracle(H, d_txt, p, t1d):
    """The restructured schedule (no per-pair vector of width d, K/V never projected per pair) reproduces the reference
    semantics: forward and every parameter gradient, with and without dropout masks."""
    B, N, T, d_model = 4, 6, 5, 24
    P, notes, tau, t_hat, G = _rand_case(11 + H, B, N, T, d_model, d_txt, H, t1d=t1d)
    d = d_txt if d_txt is not None else d_model
    masks = None
    if p > 0:
        g = torch.Generator().manual_seed(5)
        masks = {"ttf.attn_dropout": (torch.rand(B, T, H, N, generator=g) >= p).double(),
                 "ttf.dropout": (torch.rand(B, T, d, generator=g) >= p).double()}
    Pr = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    E_ref, M_ref = O.ttf_t2v_xattn_perquery(Pr, notes, tau, t_hat, n_heads=H, p=p, masks=masks)
    (E_ref * G).sum().backward()
    E, M, ctx = S.forward(P, notes, tau, t_hat, H, p=p, masks=masks)
    assert torch.equal(M, M_ref)
    assert rel_max(E, E_ref.detach()) <= 1e-12
    grads = S.backward(P, notes, tau, ctx, G)
    gmax = max(v.grad.abs().max().item() for v in Pr.values() if v.grad is not None)
    for k, v in Pr.items():
        ref = v.grad if v.grad is not None else torch.zeros_like(v)
        err = (grads[k] - ref).abs().max().item()
        assert err <= 1e-10 * max(ref.abs().max().item(), 1e-3 * gmax), (k, err)


def test_schedule_no_note_sample_forward_and_finite_grads():
    P, notes, tau, t_hat, G = _rand_case(3, 4, 5, 6, 24, None, 2, no_note=True)
    E_ref, M_ref = O.ttf_t2v_xattn_perquery(P, notes, tau, t_hat, n_heads=2)
    E, M, ctx = S.forward(P, notes, tau, t_hat, 2)
    assert torch.equal(M, M_ref) and not bool(M[-1])
    assert rel_max(E, E_ref) <= 1e-12
    grads = S.backward(P, notes, tau, ctx, G)
    assert all(torch.isfinite(v).all() for v in grads.values())  # the reference's backward is NaN here


@pytest.mark.parametrize("name", [n for n in PQ if "nonote" not in n])
def test_schedule_matches_reference_golden(name):
    cfg, params, inp, rest = load_pq(name)
    P = {k: v.double() for k, v in params.items()}
    E, M, ctx = S.forward(P, inp["notes"].double(), inp["tau"].double(), inp["t_hat"].double(), cfg["H"])
    assert rel_max(E, rest["eval64:E_txt"]) <= 1e-12
    grads = S.backward(P, inp["notes"].double(), inp["tau"].double(), ctx, inp["G"].double())
    gmax = max(np.abs(rest[f"grad64:{k}"]).max() for k in P)
    for k in P:
        ref = torch.from_numpy(rest[f"grad64:{k}"])
        assert (grads[k] - ref).abs().max().item() <= 1e-10 * max(ref.abs().max().item(), 1e-3 * gmax), k
