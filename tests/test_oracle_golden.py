"""Pin the oracle (oracle/immtsf_oracle.py) against outputs of the unmodified
reference modules, committed as tests/golden/*.npz by oracle/make_golden.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, golden_names, load_golden, rel_max
from oracle import immtsf_oracle as O

NAMES = golden_names()


def _run(cfg, params, inp, dtype, train=False):
    P = {k: v.to(dtype) if v.is_floating_point() else v for k, v in params.items()}
    if train:
        P = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    Y = inp["Y_ts"].to(dtype).clone().requires_grad_(train)
    out = O.fusion_forward(
        P, cfg["ttf"], cfg["mmf"], inp["notes"].to(dtype), inp["tau"].to(dtype), inp["t_hat"].to(dtype), Y,
        n_heads=cfg["H"], kappa=cfg["kappa"], return_intermediate=True,
    )
    return P, Y, out


def test_fixtures_present():
    assert len(NAMES) >= 9


@pytest.mark.parametrize("name", NAMES)
def test_forward_fp32_matches_reference(name):
    cfg, params, inp, ref = load_golden(name)
    _, _, (Yo, E, M) = _run(cfg, params, inp, torch.float32)
    assert np.array_equal(M.numpy(), ref["eval:M_txt"])  # bit-exact mask
    assert rel_max(E, ref["eval:E_txt"]) < 2e-6
    assert rel_max(Yo, ref["eval:Y_out"]) < 2e-6


@pytest.mark.parametrize("name", NAMES)
def test_forward_fp64_matches_reference(name):
    cfg, params, inp, ref = load_golden(name)
    _, _, (Yo, E, _) = _run(cfg, params, inp, torch.float64)
    assert rel_max(E, ref["eval64:E_txt"]) < 1e-12
    assert rel_max(Yo, ref["eval64:Y_out"]) < 1e-12


@pytest.mark.parametrize("name", [n for n in NAMES if not n.endswith("nonote")])
@pytest.mark.parametrize("dtype,tag,tol", [(torch.float32, "grad", 5e-5), (torch.float64, "grad64", 1e-10)])
def test_gradients_match_reference(name, dtype, tag, tol):
    cfg, params, inp, ref = load_golden(name)
    P, Y, (Yo, _, _) = _run(cfg, params, inp, dtype, train=True)
    (Yo * inp["G"].to(dtype)).sum().backward()
    assert rel_max(Yo.detach(), ref[f"{tag}:Y_out"]) < tol
    assert rel_max(Y.grad, ref[f"{tag}:Y_ts"]) < tol
    gmax = max(float(np.abs(ref[f"{tag}:{k}"]).max()) for k in P)
    for k, v in P.items():
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        r = ref[f"{tag}:{k}"]
        # structurally-zero grads (e.g. mmf.proj_q after T2V_XAttn) need an absolute floor
        err = float((g.double() - torch.from_numpy(r).double()).abs().max())
        assert err <= tol * max(float(np.abs(r).max()), 1e-3 * gmax), (k, err)


@pytest.mark.parametrize("name", [n for n in NAMES if load_golden(n)[0]["ttf"] == "TTF_T2V_XAttn"])
def test_expand_and_shared_kv_agree(name):
    """The T_f-fold K/V expansion (reference :151-159) is numerically a no-op (TTF_RecAvg has no expansion)."""
    cfg, params, inp, _ = load_golden(name)
    a, _ = O.ttf_t2v_xattn(params, inp["notes"], inp["tau"], inp["t_hat"], cfg["H"], faithful_expand=True)
    b, _ = O.ttf_t2v_xattn(params, inp["notes"], inp["tau"], inp["t_hat"], cfg["H"], faithful_expand=False)
    assert rel_max(a, b) < 1e-6


def test_csr_oracle_definition():
    cfg, params, inp, _ = load_golden("recavg_gr")
    off, rows, seg, mask = O.csr_from_padded(inp["notes"])
    B, N = mask.shape
    assert off[0] == 0 and off[-1] == mask.sum()
    for b in range(B):
        r = rows[off[b]:off[b + 1]]
        assert torch.equal(r, (torch.nonzero(mask[b]).reshape(-1) + b * N).to(torch.int32))
        assert (seg[off[b]:off[b + 1]] == b).all()
    assert not mask[0, 1]  # the all-zero "real" row mid-sequence is masked by content


def test_shapes_table_matches_reference_state_dict():
    for name in NAMES:
        cfg, params, inp, _ = load_golden(name)
        shapes = O.param_shapes(cfg["ttf"], cfg["mmf"], inp["notes"].shape[2], cfg["d_txt"], cfg["C"])
        assert {k: tuple(v.shape) for k, v in params.items()} == shapes, name


def test_nan_raises_like_reference():
    cfg, params, inp, _ = load_golden("recavg_gr")
    bad = inp["notes"].clone()
    bad[0, 0, 0] = float("nan")
    with pytest.raises(ValueError):
        O.fusion_forward(params, cfg["ttf"], cfg["mmf"], bad, inp["tau"], inp["t_hat"], inp["Y_ts"])
    badY = inp["Y_ts"].clone()
    badY[0, 0, 0] = float("nan")
    with pytest.raises(ValueError):
        O.fusion_forward(params, cfg["ttf"], cfg["mmf"], inp["notes"], inp["tau"], inp["t_hat"], badY)


@pytest.mark.skipif(not os.path.isdir("/root/reference/fusions"), reason="reference tree only exists in the build container")
def test_fixtures_regenerate_identically(tmp_path):
    """Re-run the generator against the live reference and compare with the
    committed fixtures (catches a stale tests/golden/)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "oracle", "make_golden.py"), "--out", str(tmp_path)],
                   check=True, capture_output=True, timeout=600)
    for name in ("recavg_gr", "t2v_xattn"):
        a = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        b = np.load(os.path.join(str(tmp_path), name + ".npz"))
        for k in a.files:
            if k == "meta":
                continue
            assert np.allclose(a[k], b[k], rtol=1e-6, atol=1e-7), (name, k)


def test_masked_mse_oracle_matches_reference_golden():
    """oracle.masked_mse (lib/evaluation.py:17-69 restated) against vectors produced by the reference's compute_error
    (oracle/make_golden_loss.py): loss, d loss / d pred, the 'sum' reduction, and a variable without observations."""
    import os

    import numpy as np
    import torch

    from oracle import immtsf_oracle as O

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "loss_mse.npz"))
    for name in ("dense", "ragged", "novar", "one"):
        pred, truth, mask = (torch.from_numpy(g[f"{name}:{k}"]) for k in ("pred", "truth", "mask"))
        for dt, tag, tol in ((torch.float32, "32", 2e-6), (torch.float64, "64", 1e-12)):
            p = pred.to(dt).clone().requires_grad_(True)
            loss = O.masked_mse(truth.to(dt), p, mask.to(dt))
            loss.backward()
            ref = torch.from_numpy(g[f"{name}:loss{tag}"])
            assert abs(float(loss) - float(ref)) <= tol * max(abs(float(ref)), 1e-30), (name, tag)
            rg = torch.from_numpy(g[f"{name}:dpred{tag}"])
            assert (p.grad - rg).abs().max() <= tol * max(float(rg.abs().max()), 1e-30), (name, tag)
        s, c = O.masked_mse(truth.double(), pred.double(), mask.double(), reduce="sum")
        assert torch.allclose(s, torch.from_numpy(g[f"{name}:sum"]), rtol=1e-12) and torch.equal(c, torch.from_numpy(g[f"{name}:count"]))
    # shares of a sharded batch add up to the single-process loss
    pred, truth, mask = (torch.from_numpy(g[f"ragged:{k}"]).double() for k in ("pred", "truth", "mask"))
    _, cnt = O.masked_mse(truth, pred, mask, reduce="sum")
    shares = sum(O.masked_mse(truth[a:b], pred[a:b], mask[a:b], count=cnt) for a, b in ((0, 3), (3, 8)))
    assert abs(float(shares) - float(g["ragged:loss64"])) < 1e-12
