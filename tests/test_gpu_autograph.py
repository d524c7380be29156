"""Transparent CUDA-graph replay behind FusionModel.forward / loss.backward() (immtsf/autograph.py): the unmodified caller's
loop -- forward, its own loss, backward, with padded shapes that change from batch to batch (lib/evaluation.py:95-100,
main.py:1097) -- must give the eager path's results, capture one graph pair per shape bucket, and keep raising the
reference's ValueErrors."""
import pytest
import torch

import gpu_common as G

pytestmark = pytest.mark.gpu


def _pad_notes(notes, tau, Nb):
    B, N, dm = notes.shape
    pn, pt = torch.zeros(B, Nb, dm), torch.zeros(B, Nb)
    pn[:, :N], pt[:, :N] = notes, tau
    return pn, pt


@pytest.mark.parametrize("ttf,mmf,p", [("TTF_T2V_XAttn", "MMF_XAttn_Add", 0.1), ("TTF_RecAvg", "MMF_GR_Add", 0.1),
                                       ("TTF_T2V_XAttn", "MMF_GR_Add", 0.0), ("TTF_RecAvg", "MMF_XAttn_Add", 0.0)])
def test_autograph_equals_eager_over_varying_shapes(ttf, mmf, p):
    from immtsf import autograph, runtime

    cfg = dict(ttf=ttf, mmf=mmf, d_txt=64, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, 96, dropout=p, seed=1)
    G.randomise_(fm, 2)
    fm.train()
    fm.enable_graphs()
    runtime.SEEDS.fixed = 0xABCDEF
    try:
        # shapes as a loader produces them: N_max varies inside and across buckets of 8, T and B vary; some repeat
        shapes = [(12, 6, 10), (12, 7, 10), (12, 13, 10), (12, 6, 10), (5, 6, 10), (12, 6, 9), (12, 7, 10)]
        for i, (B, N, T) in enumerate(shapes):
            notes, tau, t_hat, Y, Gw = G.synth_batch(B, N, T, 96, 4, 100 + i)
            fm.zero_grad(set_to_none=True)
            Yc = Y.cuda().requires_grad_(True)
            out = fm(notes.cuda(), tau.cuda(), t_hat.cuda(), Yc)
            (out * Gw.cuda()).sum().backward()
            got = {"Y_out": out.detach().clone(), "dY": Yc.grad.clone(), "grads": {k: v.grad.clone() for k, v in fm.named_parameters()}}
            # eager reference on the padded batch the graph saw, with the replay's effective seed
            Nb = (N + autograph.N_BUCKET - 1) // autograph.N_BUCKET * autograph.N_BUCKET
            pn, pt = _pad_notes(notes, tau, Nb)
            off = int(autograph._seed_offset(torch.device("cuda", torch.cuda.current_device())).item())
            fm.enable_graphs(False)
            runtime.SEEDS.fixed = 0xABCDEF + off
            ref = G.gpu_run(fm, pn, pt, t_hat, Y, Gw, train=True)
            runtime.SEEDS.fixed = 0xABCDEF
            fm.enable_graphs(True)
            G.assert_close("Y_out", got["Y_out"].cpu(), ref["Y_out"], 1e-6)
            G.assert_close("dY", got["dY"].cpu(), ref["dY"], 1e-5)
            gmax = max(float(v.abs().max()) for v in ref["grads"].values())
            for k, g in ref["grads"].items():
                G.assert_close(k, got["grads"][k].cpu(), g, 2e-5, floor=1e-3 * gmax)
        # (12,6,10) and (12,7,10) share a bucket; repeats hit the cache: 4 distinct keys
        assert fm._autograph.captures == 4, fm._autograph.captures
    finally:
        runtime.SEEDS.fixed = None


def test_autograph_eval_and_error_conventions():
    cfg = dict(ttf="TTF_RecAvg", mmf="MMF_GR_Add", d_txt=32, C=3, H=1, kappa=0.5)
    fm = G.build_model(cfg, 48, dropout=0.1, seed=3)
    fm.eval()
    notes, tau, t_hat, Y, _ = G.synth_batch(6, 5, 7, 48, 3, 9)
    n, ta, th, Yc = notes.cuda(), tau.cuda(), t_hat.cuda(), Y.cuda()
    with torch.no_grad():
        ref = fm(n, ta, th, Yc).clone()
        fm.enable_graphs()
        a = fm(n, ta, th, Yc).clone()
        b = fm(n, ta, th, Yc).clone()
    assert torch.equal(a, b)
    G.assert_close("eval", a.cpu(), ref.cpu(), 1e-6)
    bad = n.clone()
    bad[1, 0, 3] = float("nan")
    with torch.no_grad(), pytest.raises(ValueError, match="V contain NaN"):
        fm(bad, ta, th, Yc)
    badY = Yc.clone()
    badY[0, 0, 0] = float("nan")
    with torch.no_grad(), pytest.raises(ValueError, match="Y_ts contains NaN"):
        fm(n, ta, th, badY)
    with torch.no_grad():
        G.assert_close("after the errors", fm(n, ta, th, Yc).cpu(), ref.cpu(), 1e-6)


def test_autograph_under_anomaly_mode_like_the_reference_loop():
    """main.py:1079 wraps the training step in torch.autograd.set_detect_anomaly(True)."""
    cfg = dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", d_txt=32, C=3, H=1, kappa=0.5)
    fm = G.build_model(cfg, 48, dropout=0.1, seed=3)
    fm.train()
    fm.enable_graphs()
    opt = torch.optim.Adam(fm.parameters(), lr=1e-3)
    losses = []
    for i in range(4):
        notes, tau, t_hat, Y, _ = G.synth_batch(6, 5, 7, 48, 3, 20 + i)
        opt.zero_grad()
        with torch.autograd.set_detect_anomaly(True):
            out = fm(notes.cuda(), tau.cuda(), t_hat.cuda(), Y.cuda().requires_grad_(True))
            loss = out.square().mean()
            loss.backward()
        torch.nn.utils.clip_grad_norm_(fm.parameters(), 1.0)
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses) and fm._autograph.captures == 1
