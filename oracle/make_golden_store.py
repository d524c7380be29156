"""Golden fixture for the text store / window index, produced by the reference's own
``lib.parse_datasets.ChunkedTimeSeriesDataset`` on a small synthetic dataset tree.

Test infrastructure; runs only in the build container.  Usage:  python oracle/make_golden_store.py
Writes tests/golden/store_chunks.npz.  The reference module is imported unmodified; ``prettytable`` and
``reformer_pytorch`` (absent here, unused on this path) are stubbed so that ``lib.parse_datasets`` imports.

The tree: processed/<rec>/time_series.csv (irregular timestamps, NaN gaps) + the per-record text-embedding file in the
format of compute_text_embeddings.py:92-98, with rel_times deliberately NOT sorted (the precomputed path keeps file
order).  Stored per chunk the reference emits: record index, window start (recomputed with the oracle's restatement of
the chunk loop and cross-checked against the reference's sub_tt), the selected taus (fp32, as multimodal_collate makes
them) and the selected row indices (found by exact row match).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import immtsf_oracle as O  # noqa: E402

HISTORY, PRED, STRIDE = 5, 3, 2
D_MODEL = 8
LLM, LAYERS, MAXLEN = "GPT2", 6, 1024


def build_tree(root, seed=7):
    rng = np.random.default_rng(seed)
    proc = os.path.join(root, "processed")
    recs = {}
    for r, (days, n_notes) in enumerate([(40, 37), (25, 9), (31, 64)]):
        name = f"rec{r:02d}"
        os.makedirs(os.path.join(proc, name))
        # irregular series: 1-3 observations a day at random hours, some all-NaN rows
        stamps = []
        for dday in range(days):
            for _ in range(int(rng.integers(1, 4))):
                stamps.append(pd.Timestamp("2021-01-01") + pd.Timedelta(days=dday, hours=float(rng.uniform(0, 24))))
        stamps = sorted(set(stamps))
        vals = rng.normal(size=(len(stamps), 3))
        vals[rng.uniform(size=vals.shape) < 0.3] = np.nan
        if r == 1:
            vals[5:12] = np.nan  # a stretch without any observation: some windows fail the hist/pred mask test
        pd.DataFrame({"date_time": stamps, "a": vals[:, 0], "b": vals[:, 1], "c": vals[:, 2]}).to_csv(
            os.path.join(proc, name, "time_series.csv"), index=False)
        rel = rng.uniform(0, days, size=n_notes).astype(np.float32)  # file order, unsorted
        rel[:3] = [0.0, float(HISTORY), 2.0]  # window-edge cases: st <= t and t < st + history
        if r == 2:
            rel[10:20] = np.float32(100.0)  # notes after the series' end: never selected
        emb = torch.from_numpy(rng.normal(size=(n_notes, D_MODEL)).astype(np.float32))
        fname = f"text_embeddings_model={LLM}_layers={LAYERS}_maxlen={MAXLEN}.pt"
        torch.save({"embeddings": emb, "rel_times": torch.from_numpy(rel)}, os.path.join(proc, name, fname))
        recs[name] = (torch.from_numpy(rel), emb)
    return recs


def main():
    for m in ("prettytable", "reformer_pytorch"):
        mod = types.ModuleType(m)
        mod.PrettyTable = object
        mod.LSHSelfAttention = object
        sys.modules.setdefault(m, mod)
    sys.path.insert(0, "/root/reference")
    import lib.parse_datasets as PD

    out = {}
    with tempfile.TemporaryDirectory() as root:
        recs = build_tree(root)
        ds = PD.ChunkedTimeSeriesDataset(root=root, history=HISTORY, pred_window=PRED, stride=STRIDE, device=torch.device("cpu"),
                                         time_unit="days", normalize=True, enable_text=True, use_text_embeddings=True,
                                         llm_model_fusion=LLM, llm_layers_fusion=LAYERS, max_length=MAXLEN)
        names = sorted(recs)
        # numeric side, as the reference parses it (lib/parse_datasets.py:93-123), to restate the chunk loop
        series = {}
        for name in names:
            df = pd.read_csv(os.path.join(root, "processed", name, "time_series.csv"))
            df["_ts_raw"] = pd.to_datetime(df["date_time"])
            df = df.sort_values("_ts_raw")
            secs = (df["_ts_raw"] - df["_ts_raw"].min()).dt.total_seconds()
            tt = torch.tensor((secs / 86400.0).values, dtype=torch.float32)
            mask = torch.tensor((~pd.isna(df[["a", "b", "c"]].values.astype("float32"))).astype("float32"))
            series[name] = (tt, mask)
    ent, sts, taus, rows, ids = [], [], [], [], []
    it = iter(ds.chunks)
    for e, name in enumerate(names):
        tt, mask = series[name]
        rel, emb = recs[name]
        texts = [(t.item(), emb[i]) for i, t in enumerate(rel)]  # :146-147
        k = 0
        for st in O.chunk_windows(tt, mask, HISTORY, PRED, STRIDE):
            sel = O.select_window_notes(texts, st, HISTORY)
            cid = f"{name}_chunk{k}"
            k += 1
            if not sel:
                continue  # the reference drops it (:217-221) but has consumed the chunk number
            chunk_id, sub_tt, _, _, selected = next(it)
            assert chunk_id == cid, (chunk_id, cid)
            idx = ((tt >= st) & (tt < st + HISTORY + PRED)).nonzero().squeeze(1)
            assert torch.equal(sub_tt, tt[idx] - st)
            assert len(selected) == len(sel) and all(a[0] == b[0] and torch.equal(a[1], b[1]) for a, b in zip(selected, sel))
            r_idx = [int((emb == p).all(dim=1).nonzero()[0, 0]) for (_, p) in selected]
            ent.append(e)
            sts.append(st)
            ids.append(cid)
            taus.append(torch.tensor([t for (t, _) in selected], dtype=torch.float32).numpy())  # :786-790
            rows.append(np.asarray(r_idx, dtype=np.int32))
    assert next(it, None) is None
    out["names"] = np.array(names)
    for name in names:
        out[f"rel:{name}"], out[f"emb:{name}"] = recs[name][0].numpy(), recs[name][1].numpy()
    out["ent"], out["st"], out["chunk_id"] = np.asarray(ent, dtype=np.int32), np.asarray(sts, dtype=np.float64), np.array(ids)
    out["sel_offsets"] = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    out["sel_tau"], out["sel_rows"] = np.concatenate(taus), np.concatenate(rows)
    out["history"] = np.array([HISTORY, PRED, STRIDE], dtype=np.float64)
    path = os.path.join(HERE, "..", "tests", "golden", "store_chunks.npz")
    np.savez_compressed(path, **out)
    print(len(ids), "chunks,", int(out["sel_offsets"][-1]), "selected notes ->", os.path.normpath(path))


if __name__ == "__main__":
    main()
