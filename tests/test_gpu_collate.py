"""CSR-emitting collate (immtsf/collate.py): FusionModel.forward_csr on the ragged layout must give
the results of FusionModel.forward on the zero-padded batch the reference collate produces (lib/parse_datasets.py:
764-824) -- outputs and every gradient -- including a sample without notes and N_i = N_max."""
import pytest
import torch

import gpu_common as G

pytestmark = pytest.mark.gpu


def _samples(B, N, d_model, seed):
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, N + 1, (B,), generator=g).tolist()
    counts[0], counts[-1] = N, 0
    return [(torch.rand(n, generator=g) * 7.0, torch.randn(n, d_model, generator=g)) for n in counts]


@pytest.mark.parametrize("ttf,mmf", [("TTF_RecAvg", "MMF_GR_Add"), ("TTF_T2V_XAttn", "MMF_XAttn_Add"),
                                     ("TTF_RecAvg", "MMF_XAttn_Add"), ("TTF_T2V_XAttn", "MMF_GR_Add")])
@pytest.mark.parametrize("train", [False, True])
def test_forward_csr_equals_padded_forward(ttf, mmf, train):
    from immtsf import collate, runtime

    B, N, T, d_model, C = 9, 7, 10, 96, 4
    cfg = dict(ttf=ttf, mmf=mmf, d_txt=64, C=C, H=2, kappa=0.5)
    fm = G.build_model(cfg, d_model, dropout=0.1, seed=1)
    G.randomise_(fm, 2)
    fm.train(train)
    samples = _samples(B, N, d_model, 5)
    g = torch.Generator().manual_seed(6)
    t_hat = torch.sort(0.5 + 0.5 * torch.rand(B, T, generator=g), dim=1)[0].cuda()
    Y = torch.randn(B, T, C, generator=g).cuda()
    Gw = torch.randn(B, T, C, generator=g).cuda()
    notes, tau = collate.pad_from_ragged(samples, "cuda")
    runtime.SEEDS.fixed = 4242
    try:
        outs = []
        for use_csr in (False, True):
            fm.zero_grad(set_to_none=True)
            Yr = Y.clone().requires_grad_(train)
            with torch.set_grad_enabled(train):
                if use_csr:
                    r = collate.ragged_collate(samples, "cuda")
                    assert r.offsets.tolist() == [0] + torch.tensor([e.shape[0] for _, e in samples]).cumsum(0).tolist()
                    out = fm.forward_csr(r, t_hat, Yr)
                else:
                    out = fm(notes, tau, t_hat, Yr)
            grads = None
            if train:
                (out * Gw).sum().backward()
                grads = {k: p.grad.clone() for k, p in fm.named_parameters()}
                grads["dY"] = Yr.grad.clone()
            outs.append((out.detach().clone(), grads))
    finally:
        runtime.SEEDS.fixed = None
    G.assert_close("Y_out", outs[1][0].cpu(), outs[0][0].cpu(), 1e-6)
    if train:
        for k, gref in outs[0][1].items():
            G.assert_close(k, outs[1][1][k].cpu(), gref.cpu(), 2e-5, floor=1e-3)


def test_ragged_collate_flags_nan_like_the_reference():
    from immtsf import collate

    cfg = dict(ttf="TTF_RecAvg", mmf="MMF_GR_Add", d_txt=32, C=3, H=1, kappa=0.5)
    fm = G.build_model(cfg, 48, dropout=0.0, seed=1).eval()
    samples = _samples(4, 5, 48, 9)
    samples[1][1][0, 3] = float("nan")
    r = collate.ragged_collate(samples, "cuda")
    with pytest.raises(ValueError, match="Input embeddings V contain NaN"):
        fm.forward_csr(r, torch.rand(4, 6).cuda(), torch.randn(4, 6, 3).cuda())
