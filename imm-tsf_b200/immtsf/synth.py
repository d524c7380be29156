"""Synthetic Time-IMM-shaped workloads and model construction for the benchmark, the tools and the
tests: seeded generators that honour the reference data path's invariants (valid rows first then zero padding,
every sample has >= 1 note unless asked otherwise, tau in [0, history) unsorted, t_hat in [h/(h+p), 1) sorted then zero
padded; lib/parse_datasets.py:209-221, 318, 345, 786-819).  No oracle and no reference code is involved."""
from types import SimpleNamespace

import torch


def make_args(ttf, mmf, alias, d_txt, C, H, kappa, dropout):
    return SimpleNamespace(TTF_module=ttf, MMF_module=mmf, llm_model_fusion=alias, llm_layers_fusion=1, max_length=1024,
                           device="cuda", use_text_embeddings=True, recency_sigma=1.0, dropout=dropout, d_txt=d_txt,
                           n_heads_fusion=H, C=C, kappa=kappa)


def build_model(cfg, d_model, params=None, dropout=0.0, seed=0):
    import fusions.load_llm as L
    from fusions.FusionModel import FusionModel

    alias = f"SYN{d_model}"
    L.register_d_model(alias, d_model)
    torch.manual_seed(seed)
    fm = FusionModel(make_args(cfg["ttf"], cfg["mmf"], alias, cfg["d_txt"], cfg["C"], cfg["H"], cfg["kappa"], dropout))
    if params is not None:
        fm.load_state_dict(params, strict=True)
    return fm.cuda()


def synth_batch(B, N, T, d_model, C, seed, history=7.0, pred=7.0, no_note=False, t1d=False, full=False):
    """Time-IMM-shaped synthetic batch: ragged N_i ~ U{1..N}, one sample full, zero tail
    padding, tau in [0,history) unsorted, t_hat in [h/(h+p),1) sorted then zero padded."""
    g = torch.Generator().manual_seed(seed)
    counts = torch.randint(1, N + 1, (B,), generator=g)
    if full:
        counts[:] = N
    counts[0] = N
    if B > 1 and not full:
        counts[1] = 1
    if no_note:
        counts[B - 1] = 0
    notes = torch.zeros(B, N, d_model)
    tau = torch.zeros(B, N)
    for b in range(B):
        n = int(counts[b])
        notes[b, :n] = torch.randn(n, d_model, generator=g)
        tau[b, :n] = torch.rand(n, generator=g) * history
    lo = history / (history + pred)
    if t1d:
        t_hat = torch.sort(lo + torch.rand(T, generator=g) * (1 - lo))[0]
    else:
        t_hat = torch.zeros(B, T)
        for b in range(B):
            tl = T if b == 0 else int(torch.randint((T + 2) // 3, T + 1, (1,), generator=g))
            t_hat[b, :tl] = torch.sort(lo + torch.rand(tl, generator=g) * (1 - lo))[0]
    Y = torch.randn(B, T, C, generator=g)
    G = torch.randn(B, T, C, generator=g)
    return notes, tau, t_hat, Y, G


def randomise_(fm, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in fm.named_parameters():
            if name.endswith("log_recency_sigma"):
                p.copy_(torch.tensor(-1.2))
            elif p.dim() >= 2:
                p.add_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
            else:
                p.add_((torch.randn(p.shape, generator=g) * 0.1).to(p.device))
