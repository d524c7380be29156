"""GPU tests of SURVEY.md 8f row f4 through the C ABI: the window index (count / scan / order-preserving fill) bit-exact
against the fixture produced by the reference's ChunkedTimeSeriesDataset and against the oracle on a larger random
case; the batch gather against the reference's collate (restated) + the pad->CSR definition; the fusion path fed from
the store equals the padded path."""
import numpy as np
import pytest
import torch

import gpu_common as G
from test_store_cpu import load_store_golden
from oracle import immtsf_oracle as O

pytestmark = pytest.mark.gpu


def _check_index(ix, recs, ent, st, history):
    off = ix.chunk_offsets_host
    rows_all, tau_all = ix.chunk_rows.cpu().numpy(), ix.chunk_tau.cpu().numpy()
    eo = ix.store.entity_offsets_host
    for i, (e, s) in enumerate(zip(ent, st)):
        rel = recs[e][0]
        sel = O.select_window_notes([(t.item(), j) for j, t in enumerate(rel)], float(s), history)
        rows = np.asarray([j for (_, j) in sel], dtype=np.int64) + eo[e]
        tau = torch.tensor([t for (t, _) in sel], dtype=torch.float32).numpy()
        assert off[i + 1] - off[i] == len(sel), i
        assert np.array_equal(rows_all[off[i]:off[i + 1]], rows), i
        assert np.array_equal(tau_all[off[i]:off[i + 1]].view(np.uint32), tau.view(np.uint32)), i


def test_window_index_matches_reference_dataset():
    from immtsf.store import EmbeddingStore, WindowIndex

    z, names, recs = load_store_golden()
    history = float(z["history"][0])
    store = EmbeddingStore.from_records(names, recs, "cuda")
    ix = WindowIndex(store, z["ent"], z["st"], z["st"] + history)
    eo = store.entity_offsets_host
    assert np.array_equal(ix.chunk_offsets_host, z["sel_offsets"])
    assert np.array_equal(ix.chunk_rows.cpu().numpy()[:199], z["sel_rows"] + eo[z["ent"]].repeat(np.diff(z["sel_offsets"])))
    assert np.array_equal(ix.chunk_tau.cpu().numpy()[:199].view(np.uint32), z["sel_tau"].view(np.uint32))
    assert len(ix.nonempty()) == 37


def test_window_index_random_large_bit_exact():
    from immtsf.store import EmbeddingStore, WindowIndex

    g = torch.Generator().manual_seed(3)
    recs, names = [], []
    for e in range(24):
        n = int(torch.randint(0, 300, (1,), generator=g)) if e != 5 else 0  # one record without notes
        rel = (torch.rand(n, generator=g) * 50).float()
        if n > 4:
            rel[:4] = torch.tensor([0.0, 7.0, 14.0, 7.0])  # exact window edges, duplicates
        recs.append((rel, torch.randn(n, 12, generator=g)))
        names.append(f"r{e}")
    n_chunks = 2500  # > 1024: the scan carries across chunks of 1024
    ent = torch.randint(0, 24, (n_chunks,), generator=g).numpy().astype(np.int32)
    st = (torch.randint(0, 60, (n_chunks,), generator=g).double() * 0.5).numpy()  # some windows past every note
    store = EmbeddingStore.from_records(names, recs, "cuda")
    ix = WindowIndex(store, ent, st, st + 7.0)
    _check_index(ix, recs, ent, st, 7.0)
    ne = ix.nonempty()
    assert 0 < len(ne) < n_chunks and (ix.counts_host[ne] > 0).all()


def test_batch_gather_equals_reference_collate_and_feeds_the_fusion_path(tmp_path):
    from immtsf import ops
    from immtsf.store import EmbeddingStore, WindowIndex
    import immtsf.store as S

    z, names, recs = load_store_golden()
    history = float(z["history"][0])
    path = str(tmp_path / "s.bin")
    EmbeddingStore.from_records(names, recs, "cuda").save(path)
    S._SLAB_ROWS_BYTES = 8 * 4 * 16  # 16 rows per slab: the double-buffered upload runs 7 slabs
    store = EmbeddingStore.open(path, device="cuda")
    assert torch.equal(store.emb_all.cpu(), torch.cat([e for _, e in recs])) and torch.equal(store.rel_all.cpu(), torch.cat([r for r, _ in recs]))
    # two extra windows that select nothing (kept in a batch on purpose: M_txt = False rows)
    ent = np.concatenate([z["ent"], [1, 2]]).astype(np.int32)
    st = np.concatenate([z["st"], [1000.0, 60.0]])
    ix = WindowIndex(store, ent, st, st + history)
    assert len(ix.nonempty()) == 37 and ix.n == 39
    ids = [3, 38, 20, 0, 36, 17, 37, 9]
    r = ix.batch(ids)
    raws = []
    for c in ids:
        rel, emb = recs[ent[c]]
        raws.append(O.select_window_notes([(t.item(), emb[j]) for j, t in enumerate(rel)], float(st[c]), history))
    tau_p, emb_p = O.collate_text(raws)  # what the reference's multimodal_collate hands to the TTF modules
    off, rows, seg, mask = O.csr_from_padded(emb_p)
    total = int(off[-1])
    assert torch.equal(r.offsets.cpu(), off) and r.N == emb_p.shape[1]
    assert torch.equal(r.m_txt[: len(ids)].cpu().bool(), mask.any(1))
    assert torch.equal(r.emb_flat[:total].cpu(), emb_p.reshape(-1, 8)[rows.long()])
    assert torch.equal(r.tau_flat[:total].cpu(), tau_p.reshape(-1)[rows.long()])
    pad_end = min((total + 127) // 128 * 128, r.M_alloc)
    assert (r.emb_flat[total:pad_end] == 0).all() and (r.tau_flat[total:pad_end] == 0).all()
    # fusion path from the store == fusion path from the reference-collated padded batch
    for ttf in ("TTF_RecAvg", "TTF_T2V_XAttn", "TTF_T2V_XAttn_old"):
        cfg = dict(ttf=ttf, mmf="MMF_GR_Add", d_txt=16, C=3, H=2, kappa=0.5)
        fm = G.build_model(cfg, 8, dropout=0.0, seed=1)
        G.randomise_(fm, 2)
        fm.eval()
        gen = torch.Generator().manual_seed(4)
        t_hat = torch.sort(torch.rand(len(ids), 6, generator=gen))[0].cuda()
        Y = torch.randn(len(ids), 6, 3, generator=gen).cuda()
        with torch.no_grad():
            a = fm.forward_csr(ix.batch(ids), t_hat, Y)
            b = fm(emb_p.cuda(), tau_p.cuda(), t_hat, Y)
        G.assert_close(ttf, a.cpu(), b.cpu(), 1e-6)
