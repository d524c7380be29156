"""CPU tests of SURVEY.md 8f row f4: the oracle's restatement of the reference's chunk-window note selection against
the fixture produced by the reference's own ChunkedTimeSeriesDataset (oracle/make_golden_store.py), and the host
side of the embedding store (packed file format, per-record .pt reader, error conventions, no CPU compute)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import immtsf_oracle as O


def load_store_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "store_chunks.npz"))
    names = [str(n) for n in z["names"]]
    recs = [(torch.from_numpy(z[f"rel:{n}"]), torch.from_numpy(z[f"emb:{n}"])) for n in names]
    return z, names, recs


def test_oracle_selection_matches_reference_dataset():
    z, names, recs = load_store_golden()
    history = float(z["history"][0])
    off = z["sel_offsets"]
    assert len(z["ent"]) == 37 and off[-1] == 199
    for i, (e, st) in enumerate(zip(z["ent"], z["st"])):
        rel, emb = recs[e]
        texts = [(t.item(), j) for j, t in enumerate(rel)]  # payload = row index
        sel = O.select_window_notes(texts, float(st), history)
        rows = np.asarray([j for (_, j) in sel], dtype=np.int32)
        tau = torch.tensor([t for (t, _) in sel], dtype=torch.float32).numpy()
        assert np.array_equal(rows, z["sel_rows"][off[i]:off[i + 1]])
        assert np.array_equal(tau.view(np.uint32), z["sel_tau"][off[i]:off[i + 1]].view(np.uint32))  # bit-exact
        assert len(rows) > 0  # the reference drops chunks without notes (lib/parse_datasets.py:217-221)


def test_collate_text_pads_like_the_reference():
    e = torch.arange(12.0).reshape(4, 3)
    tau, emb = O.collate_text([[(0.5, e[0]), (1.5, e[1])], [], [(2.0, e[3])]])
    assert tau.tolist() == [[0.5, 1.5], [0.0, 0.0], [2.0, 0.0]]
    assert emb.shape == (3, 2, 3) and torch.equal(emb[0, 1], e[1]) and emb[1].abs().sum() == 0 and torch.equal(emb[2, 0], e[3])


def test_packed_store_roundtrip(tmp_path):
    from immtsf.store import EmbeddingStore

    _, names, recs = load_store_golden()
    s = EmbeddingStore.from_records(names, recs, "cpu")
    assert s.num_notes == sum(r.shape[0] for r, _ in recs) and s.d_model == 8
    assert s.entity_offsets_host.tolist() == [0, 37, 46, 110]
    path = str(tmp_path / "store.bin")
    s.save(path)
    t = EmbeddingStore.open(path, device="cpu")
    assert t.names == names and np.array_equal(t.entity_offsets_host, s.entity_offsets_host)
    assert torch.equal(t.rel_all, s.rel_all) and torch.equal(t.emb_all, s.emb_all)
    with open(path, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError, match="not an immtsf embedding store"):
        EmbeddingStore.open(path, device="cpu")


def test_pt_dir_reader_and_error_conventions(tmp_path):
    from immtsf.store import EmbeddingStore, WindowIndex, pt_file_name

    _, names, recs = load_store_golden()
    fname = pt_file_name("GPT2", 6, 1024)
    assert fname == "text_embeddings_model=GPT2_layers=6_maxlen=1024.pt"
    assert pt_file_name("BERT", None, 512) == "text_embeddings_model=BERT_layers=full_maxlen=512.pt"
    for n, (rel, emb) in zip(names, recs):
        os.makedirs(tmp_path / n)
        torch.save({"embeddings": emb, "rel_times": rel}, tmp_path / n / fname)
    s = EmbeddingStore.from_pt_dir(str(tmp_path), "GPT2", 6, 1024, device="cpu")
    assert s.names == sorted(names) and torch.equal(s.emb_all[37:46], recs[1][1])
    os.remove(tmp_path / names[1] / fname)
    with pytest.raises(FileNotFoundError, match="Missing text embeddings file"):
        EmbeddingStore.from_pt_dir(str(tmp_path), "GPT2", 6, 1024, device="cpu")
    bad = recs[0][1].clone()
    bad[3, 2] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        EmbeddingStore.from_records(names[:1], [(recs[0][0], bad)], "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        WindowIndex(s, [0], [0.0], [5.0])
