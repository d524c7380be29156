"""GPU parity tests of SURVEY.md 8f row f3: the per-(note, query) Time2Vec attention (fusions/TTF_T2V_XAttn_old.py
semantics) through the C ABI -- the two fused kernels against their contract (oracle/perquery_schedule.py), the module
against golden vectors produced by the reference's own class, FusionModel compositions against the oracle with Philox
dropout masks rebuilt on the host, and size-independent properties at the cfg2 shape.
Tolerances as in test_gpu_parity.py: outputs 1e-5 max-norm relative, gradients 5e-5 (floor for structural zeros)."""
import numpy as np
import pytest
import torch

import gpu_common as G
import philox_ref
from test_gpu_parity import grad_check, OUT_TOL
from test_perquery_cpu import PQ, load_pq
from oracle import immtsf_oracle as O
from oracle import perquery_schedule as S

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fixed_seed():
    from immtsf import runtime

    runtime.SEEDS.fixed = 0x5EED1234ABCD
    yield
    runtime.SEEDS.fixed = None


def _ragged(B, N, d, seed, no_note=False):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, N + 1, (B,), generator=g)
    counts[0] = N
    if B > 1:
        counts[1] = 1
    if no_note:
        counts[-1] = 0
    mask = torch.arange(N)[None, :] < counts[:, None]
    return counts, mask, g


# ------------------------------------------------------------------ the two kernels against their contract
@pytest.mark.parametrize("B,N,T,H,d,dt,p,t1d,no_note", [
    (5, 6, 7, 1, 48, 24, 0.0, False, False),
    (3, 40, 19, 4, 64, 32, 0.0, False, False),   # T > tile, N > 32, several heads
    (4, 9, 11, 2, 32, 16, 0.25, True, True),     # dropout, shared 1-D t_hat, a no-note sample
    (2, 70, 3, 8, 64, 33, 0.1, False, False),    # 8 heads, odd d_tau
    (6, 16, 24, 1, 768, 384, 0.1, False, False),  # cfg2 widths
])
def test_kernels_match_contract(B, N, T, H, d, dt, p, t1d, no_note):
    from immtsf import ops, runtime

    counts, mask, g = _ragged(B, N, d, B * 100 + N + T, no_note)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    A = rn(B, N, d) * mask[..., None]
    a_sc = rn(B, N, H) * mask[..., None]
    gg = rn(H, dt) * 0.5
    tau = torch.rand(B, N, generator=g, dtype=torch.float64) * 1.2 * mask
    t_hat = torch.rand(T, generator=g, dtype=torch.float64) if t1d else torch.rand(B, T, generator=g, dtype=torch.float64)
    w_lin, b_lin = rn(1, 1), rn(1)
    w_per, b_per = rn(dt - 1, 1) * 4.0, rn(dt - 1)
    dZ, dPhi, dsp = rn(B, T, H, d), rn(B, T, H, dt), rn(B, T, H)
    # fp32 copies are the ground truth the kernel sees
    f32 = lambda t: t.float()
    A, a_sc, gg, tau, t_hat, w_lin, b_lin, w_per, b_per, dZ, dPhi, dsp = map(
        lambda t: f32(t).double(), (A, a_sc, gg, tau, t_hat, w_lin, b_lin, w_per, b_per, dZ, dPhi, dsp))
    seed = runtime.SEEDS.fixed
    thr = ops.drop_thr(p)
    ks = None
    if p > 0:
        rank = (torch.cumsum(mask.to(torch.int64), 1) - 1).clamp_min(0).numpy().astype(np.uint64)
        b_ = np.arange(B, dtype=np.uint64)[:, None, None, None]
        t_ = np.arange(T, dtype=np.uint64)[None, :, None, None]
        h_ = np.arange(H, dtype=np.uint64)[None, None, :, None]
        idx = ((b_ * np.uint64(T) + t_) * np.uint64(H) + h_) * np.uint64(N) + rank[:, None, None, :]
        ks = torch.from_numpy(philox_ref.keep_mask(seed, G.SITE_TTF_ATTN, idx, p)).double() / (1.0 - philox_ref.realised_p(p))
    t2 = t_hat if t_hat.dim() == 2 else t_hat[None].repeat(B, 1)
    Z_r, Phi_r, sp_r, P_r = S.pool_fwd(A, a_sc, gg, tau, t2, mask, w_lin, b_lin, w_per, b_per, keep_scale=ks)
    dA_r, da_r, dw_r, db_r, dg_r = S.pool_bwd(dZ, dPhi, dsp, A, gg, P_r, tau, t2, w_lin, b_lin, w_per, b_per, keep_scale=ks)

    # CSR layout on the device (valid rows are a prefix of every sample)
    notes = torch.zeros(B, N, 4)
    notes[mask] = 1.0
    r = ops.csr_build(notes.cuda(), tau.float().cuda())
    rows = r.rows.cpu().long()
    total = int(r.offsets[-1])
    flat = lambda x, w: x.reshape(B * N, w)[rows[:total]].float()
    A_c = torch.zeros(r.M_alloc, d)
    A_c[:total] = flat(A, d)
    a_c = torch.zeros(r.M_alloc, H)
    a_c[:total] = flat(a_sc, H)
    cu = lambda t: t.float().contiguous().cuda()
    t2v = (cu(w_lin), cu(b_lin), cu(w_per), cu(b_per))
    th = cu(t_hat)
    Z, Phi, sp, probs = ops.t2vq_attn_fwd(A_c.cuda(), a_c.cuda(), cu(gg), r, th, t2v, T, H, d, dt, thr, seed, True)
    G.assert_close("Z", Z.cpu().view(B, T, H, d), Z_r, OUT_TOL)
    G.assert_close("Phi", Phi.cpu().view(B, T, H, dt), Phi_r, OUT_TOL)
    G.assert_close("sp", sp.cpu().view(B, T, H), sp_r, OUT_TOL)
    pg = probs.cpu().view(H, T, r.M_alloc)[:, :, :total]  # [H,T,total]
    P_flat = P_r.permute(2, 1, 0, 3).reshape(H, T, B * N)[:, :, rows[:total]]
    G.assert_close("probs", pg, P_flat, OUT_TOL)
    dA, da, dpart = ops.t2vq_attn_bwd(cu(dZ.view(-1, d)), cu(dPhi.view(-1, dt)), cu(dsp.view(-1)), A_c.cuda(), cu(gg), probs, r, th,
                                      t2v, T, H, d, dt, thr, seed)
    torch.cuda.synchronize()
    G.assert_close("dA", dA.cpu()[:total], flat(dA_r, d).double(), 2e-5)
    G.assert_close("da", da.cpu()[:total], flat(da_r, H).double(), 2e-5)
    tg = dpart.cpu().double().view(-1, 2 + H, dt).sum(0)  # rows = (sample, query tile)
    gmax = max(dw_r.abs().max().item(), db_r.abs().max().item(), dg_r.abs().max().item())
    G.assert_close("dw", tg[0], dw_r, 2e-5, floor=1e-2 * gmax)
    G.assert_close("db", tg[1], db_r, 2e-5, floor=1e-2 * gmax)
    G.assert_close("dg", tg[2:], dg_r, 2e-5, floor=1e-2 * gmax)
    pad_end = min((total + 127) // 128 * 128, r.M_alloc)
    assert (dA[total:pad_end] == 0).all() and (da[total:pad_end] == 0).all()


# ------------------------------------------------------------------ the module against the reference's golden vectors
def _module(params, H, dropout=0.0):
    import fusions.load_llm as L
    from fusions.TTF_T2V_XAttn_old import TTF_T2V_XAttn

    d = params["ttf.proj_out.weight"].shape[0]
    d_model = params["ttf.input_proj.weight"].shape[1] if "ttf.input_proj.weight" in params else d
    L.register_d_model(f"SYN{d_model}", d_model)
    m = TTF_T2V_XAttn(f"SYN{d_model}", 1, device="cuda", n_heads_fusion=H, dropout=dropout,
                      d_txt=d if "ttf.input_proj.weight" in params else None)
    m.load_state_dict({k[len("ttf."):]: v for k, v in params.items()}, strict=True)
    return m.cuda()


@pytest.mark.parametrize("name", PQ)
def test_module_matches_reference_golden(name):
    cfg, params, inp, rest = load_pq(name)
    m = _module(params, cfg["H"])
    m.eval()
    with torch.no_grad():
        E, M = m(inp["notes"].cuda(), inp["tau"].cuda(), inp["t_hat"].cuda())
    assert np.array_equal(M.cpu().numpy(), rest["eval:M_txt"])
    G.assert_close("E_txt", E.cpu(), rest["eval64:E_txt"], OUT_TOL)
    m.train()
    E, _ = m(inp["notes"].cuda(), inp["tau"].cuda(), inp["t_hat"].cuda())
    (E * inp["G"].cuda()).sum().backward()
    torch.cuda.synchronize()
    grads = {"ttf." + k: p.grad.cpu() for k, p in m.named_parameters()}
    assert all(torch.isfinite(g).all() for g in grads.values())
    if cfg["no_note"]:
        return  # the reference's backward is NaN for a no-note sample; ours is finite
    gmax = max(np.abs(rest[f"grad64:{k}"]).max() for k in grads)
    for k, g in grads.items():
        grad_check(k, g, rest[f"grad64:{k}"], rest[f"grad:{k}"], gmax)


# ------------------------------------------------------------------ FusionModel compositions against the oracle
@pytest.mark.parametrize("mmf,d_model,d_txt,C,H,B,N,T,p", [
    ("MMF_GR_Add", 96, 64, 5, 2, 9, 7, 11, 0.0),
    ("MMF_XAttn_Add", 96, 64, 4, 1, 8, 12, 10, 0.1),
    ("MMF_GR_Add", 64, None, 3, 4, 6, 5, 40, 0.1),
    ("MMF_XAttn_Add", 768, 768, 4, 1, 16, 16, 24, 0.1),  # cfg2 widths (rank form of MMF_XAttn_Add)
])
def test_fusion_model_matches_oracle(mmf, d_model, d_txt, C, H, B, N, T, p):
    from immtsf import runtime

    cfg = dict(ttf="TTF_T2V_XAttn_old", mmf=mmf, d_txt=d_txt, C=C, H=H, kappa=0.5)
    fm = G.build_model(cfg, d_model, dropout=p, seed=3)
    G.randomise_(fm, 4)
    with torch.no_grad():
        fm.ttf.time2vec.periodic.weight.mul_(4.0)
    d = d_txt if d_txt is not None else d_model
    notes, tau, t_hat, Y, Gw = G.synth_batch(B, N, T, d_model, C, seed=77 + B)
    tau = tau / 7.0 * 1.1  # lags of both signs around the query times: the clamp is exercised
    params = {k: v.detach().cpu() for k, v in fm.state_dict().items()}
    masks = G.oracle_masks(cfg, notes, T, C, d, p, runtime.SEEDS.fixed)
    ref = G.oracle_run(cfg, params, notes, tau, t_hat, Y, Gw, p=p, masks=masks)
    ref32 = G.oracle_run(cfg, params, notes, tau, t_hat, Y, Gw, dtype=torch.float32, p=p, masks=masks)
    out = G.gpu_run(fm, notes, tau, t_hat, Y, Gw, train=True)
    G.assert_close("Y_out", out["Y_out"], ref["Y_out"], OUT_TOL)
    gmax = max(v.abs().max().item() for v in ref["grads"].values())
    grad_check("dY_ts", out["dY"], ref["dY"], ref32["dY"], ref["dY"].abs().max().item())
    for k, g in out["grads"].items():
        grad_check(k, g, ref["grads"][k], ref32["grads"][k], gmax)
    # eval mode: E_txt itself
    fm.eval()
    with torch.no_grad():
        E, M = fm.ttf(notes.cuda(), tau.cuda(), t_hat.cuda())
    refe = G.oracle_run(cfg, params, notes, tau, t_hat, Y, Gw, grads=False)
    G.assert_close("E_txt", E.cpu(), refe["E_txt"], OUT_TOL)
    assert torch.equal(M.cpu(), refe["M_txt"])


# ------------------------------------------------------------------ properties at the cfg2 shape
def test_properties_cfg2_shape():
    """Permutation invariance over notes, independence of the padding width, 1-D t_hat == repeated 2-D t_hat,
    dependence on the query time (the point of this variant)."""
    B, N, T, d = 256, 16, 24, 768
    cfg = dict(ttf="TTF_T2V_XAttn_old", mmf="MMF_GR_Add", d_txt=d, C=4, H=1, kappa=0.5)
    fm = G.build_model(cfg, d, dropout=0.0, seed=5)
    G.randomise_(fm, 6)
    fm.eval()
    notes, tau, t_hat, Y, _ = G.synth_batch(B, N, T, d, 4, seed=9, full=True)
    tau = tau / 7.0
    with torch.no_grad():
        E0, _ = fm.ttf(notes.cuda(), tau.cuda(), t_hat.cuda())
        perm = torch.randperm(N, generator=torch.Generator().manual_seed(1))
        E1, _ = fm.ttf(notes[:, perm].cuda(), tau[:, perm].cuda(), t_hat.cuda())
        pad_n = torch.cat([notes, torch.zeros(B, 5, d)], 1)
        pad_t = torch.cat([tau, torch.zeros(B, 5)], 1)
        E2, _ = fm.ttf(pad_n.cuda(), pad_t.cuda(), t_hat.cuda())
        E3, _ = fm.ttf(notes.cuda(), tau.cuda(), t_hat[0].cuda())
        E4, _ = fm.ttf(notes.cuda(), tau.cuda(), t_hat[0][None].repeat(B, 1).cuda())
    G.assert_close("permutation", E1.cpu(), E0.cpu(), 2e-6)
    G.assert_close("padding width", E2.cpu(), E0.cpu(), 2e-6)  # (M_alloc changes the GEMM tiling of the ragged rows)
    assert torch.equal(E3, E4)
    assert (E0[:, 0] - E0[:, -1]).abs().max().item() > 1e-3 * E0.abs().max().item()
