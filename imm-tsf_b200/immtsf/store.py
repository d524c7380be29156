"""GPU-resident text-embedding store and window index.

The reference loads one ``.pt`` file per record -- ``{"embeddings": [N, d_model] fp32, "rel_times": [N] fp32}``, written
by compute_text_embeddings.py:92-98 -- into a Python list of ``(rel_time, embedding row)`` tuples
(lib/parse_datasets.py:132-147), filters that list with a comprehension for every chunk window (:204-209) and, for
every batch, stacks and pads the selected rows again (multimodal_collate, :786-819).  Here

* :class:`EmbeddingStore` holds every embedding row of every record ONCE, in HBM: ``emb_all [sumN, d_model]``,
  ``rel_all [sumN]``, ``entity_offsets [E+1]``.  It is built from the reference's per-record files
  (:meth:`from_pt_dir`) or from a packed single-file form that is memory-mapped and uploaded in slabs through pinned
  staging (:meth:`save` / :meth:`open`), so a store larger than host RAM's comfort zone never exists as Python objects;
* :class:`WindowIndex` runs the window filter for ALL chunks of a dataset as two kernels around an exclusive scan and
  keeps the result as a CSR over chunks (selected rows in file order, ``tau = t - st``);
* :meth:`WindowIndex.batch` gathers a batch of chunks from the resident store straight into the ragged layout
  (``ops.RaggedNotes``) that ``FusionModel.forward_csr`` consumes: one launch, no padded tensor, no host copy of data.

What stays with the caller: which windows exist (the reference's chunk loop over the numeric series, :178-227, is
dataset parsing and out of scope) -- the caller passes ``(record index, st, st + history)`` per chunk, in double, as
the reference computes them.  Chunks that select no note are dropped by the reference (:217-221); :meth:`nonempty`
returns the surviving chunk ids.

Packed file format (little endian): 64-byte header ``b"IMMTSFES" u32 version=1, u32 E, u64 sumN, u32 d_model,
u32 names_bytes``, then ``names_bytes`` of UTF-8 record names joined by ``\\n`` (padded to 8 bytes), ``entity_offsets``
(E+1 int64), ``rel_all`` (sumN fp32, padded to 8 bytes), ``emb_all`` (sumN * d_model fp32).
"""
from __future__ import annotations

import os
import struct
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops

MAGIC = b"IMMTSFES"
VERSION = 1
_HDR = struct.Struct("<8sIIQII")  # magic, version, E, sumN, d_model, names_bytes  (+ zero padding to 64 bytes)
_SLAB_ROWS_BYTES = 64 << 20


def pt_file_name(llm_model_fusion: str, llm_layers_fusion, max_length: int) -> str:
    """The file name of compute_text_embeddings.py:56-60 / lib/parse_datasets.py:134-138."""
    return f"text_embeddings_model={llm_model_fusion}_layers={llm_layers_fusion or 'full'}_maxlen={max_length}.pt"


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


class EmbeddingStore:
    """Every text embedding of every record, resident on one device.  ``names[e]`` owns rows
    ``entity_offsets[e] : entity_offsets[e+1]`` (file order)."""

    def __init__(self, names: Sequence[str], entity_offsets: np.ndarray, rel_all: torch.Tensor, emb_all: torch.Tensor):
        self.names = list(names)
        self.index = {n: i for i, n in enumerate(self.names)}
        self.entity_offsets_host = np.asarray(entity_offsets, dtype=np.int64)
        assert self.entity_offsets_host.shape == (len(self.names) + 1,) and int(self.entity_offsets_host[-1]) == emb_all.shape[0]
        if int(self.entity_offsets_host[-1]) >= 2**31:
            raise ValueError("EmbeddingStore: more than 2^31 - 1 notes do not fit the int32 row index of the kernels")
        self.rel_all, self.emb_all = rel_all, emb_all
        self.device = emb_all.device
        self.entity_offsets = torch.from_numpy(self.entity_offsets_host.astype(np.int32)).to(self.device)

    @property
    def d_model(self) -> int:
        return int(self.emb_all.shape[1])

    @property
    def num_notes(self) -> int:
        return int(self.emb_all.shape[0])

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_records(cls, names: Sequence[str], records, device) -> "EmbeddingStore":
        """records[e] = (rel_times [N_e], embeddings [N_e, d_model]) CPU tensors, in file order."""
        counts = [int(e.shape[0]) for _, e in records]
        d_m = max((int(e.shape[1]) for _, e in records if e.dim() == 2 and e.shape[0]), default=0)
        off = np.zeros(len(counts) + 1, dtype=np.int64)
        off[1:] = np.cumsum(counts)
        total = int(off[-1])
        dev = torch.device(device)
        emb_all = torch.empty(total, d_m, dtype=torch.float32, device=dev)
        rel_all = torch.empty(total, dtype=torch.float32, device=dev)
        for (rel, emb), o, c in zip(records, off[:-1], counts):
            if c:
                emb_all[o:o + c].copy_(emb.reshape(c, d_m).float(), non_blocking=True)
                rel_all[o:o + c].copy_(rel.reshape(c).float(), non_blocking=True)
        store = cls(names, off, rel_all, emb_all)
        store._raise_on_nan()
        return store

    @classmethod
    def from_pt_dir(cls, proc_dir: str, llm_model_fusion: str, llm_layers_fusion, max_length: int = 1024,
                    rec_ids: Optional[Sequence[str]] = None, device="cuda") -> "EmbeddingStore":
        """Reads the reference's per-record files under ``<proc_dir>/<record>/`` (lib/parse_datasets.py:80-83, 132-149);
        records in sorted directory order unless ``rec_ids`` is given (:86-88)."""
        if rec_ids is None:
            rec_ids = sorted(d for d in os.listdir(proc_dir) if os.path.isdir(os.path.join(proc_dir, d)))
        fname = pt_file_name(llm_model_fusion, llm_layers_fusion, max_length)
        records = []
        for rec in rec_ids:
            path = os.path.join(proc_dir, rec, fname)
            if not os.path.isfile(path):
                raise FileNotFoundError(f"Missing text embeddings file: {path}")  # :148
            data = torch.load(path, map_location="cpu")
            records.append((data["rel_times"], data["embeddings"]))
        return cls.from_records(list(rec_ids), records, device)

    def _raise_on_nan(self):
        """The reference's load-time guard (lib/parse_datasets.py:142-143), once for the whole store."""
        if self.emb_all.is_cuda and self.num_notes:
            flags = torch.zeros(4, dtype=torch.int32, device=self.device)
            ops.nan_check(self.emb_all, flags, ops.FLAG_V)
            bad = bool(flags[ops.FLAG_V].item())
        else:
            bad = bool(torch.isnan(self.emb_all).any())
        if bad:
            raise ValueError("text embeddings contains NaN values.")

    # ------------------------------------------------------------------ packed file
    def save(self, path: str):
        names_blob = "\n".join(self.names).encode("utf-8")
        with open(path, "wb") as f:
            f.write(_HDR.pack(MAGIC, VERSION, len(self.names), self.num_notes, self.d_model, len(names_blob)).ljust(64, b"\0"))
            f.write(names_blob.ljust(_pad8(len(names_blob)), b"\0"))
            f.write(self.entity_offsets_host.astype("<i8").tobytes())
            rel = self.rel_all.detach().cpu().numpy().astype("<f4")
            f.write(rel.tobytes().ljust(_pad8(rel.nbytes), b"\0"))
            rows = max(1, _SLAB_ROWS_BYTES // max(4 * self.d_model, 1))
            for r0 in range(0, self.num_notes, rows):
                f.write(self.emb_all[r0:r0 + rows].detach().cpu().numpy().astype("<f4").tobytes())

    @classmethod
    def open(cls, path: str, device="cuda") -> "EmbeddingStore":
        """Memory-maps the packed file and uploads the embedding matrix slab by slab through a pinned staging buffer."""
        with open(path, "rb") as f:
            magic, ver, E, total, d_m, nb = _HDR.unpack(f.read(64)[: _HDR.size])
        if magic != MAGIC or ver != VERSION:
            raise ValueError(f"{path}: not an immtsf embedding store (magic {magic!r}, version {ver})")
        pos = 64
        mm = np.memmap(path, mode="r", dtype=np.uint8)
        names = bytes(mm[pos:pos + nb]).decode("utf-8").split("\n") if nb else []
        pos += _pad8(nb)
        off = np.frombuffer(mm, dtype="<i8", count=E + 1, offset=pos).astype(np.int64)
        pos += 8 * (E + 1)
        rel = np.frombuffer(mm, dtype="<f4", count=total, offset=pos)
        pos += _pad8(4 * total)
        emb = np.frombuffer(mm, dtype="<f4", count=total * d_m, offset=pos).reshape(total, d_m)
        if len(names) != E or int(off[-1]) != total or pos + 4 * total * d_m != mm.shape[0]:
            raise ValueError(f"{path}: inconsistent embedding store header")
        dev = torch.device(device)
        rel_all = torch.from_numpy(np.array(rel, dtype=np.float32)).to(dev)
        emb_all = torch.empty(total, d_m, dtype=torch.float32, device=dev)
        rows = max(1, _SLAB_ROWS_BYTES // max(4 * d_m, 1))
        stage = [torch.empty(min(rows, max(total, 1)), d_m, dtype=torch.float32).pin_memory() if dev.type == "cuda" else None
                 for _ in range(2)]
        events = [None, None]
        for k, r0 in enumerate(range(0, total, rows)):
            r1 = min(r0 + rows, total)
            if dev.type != "cuda":
                emb_all[r0:r1] = torch.from_numpy(np.array(emb[r0:r1], dtype=np.float32))
                continue
            s = stage[k & 1]
            if events[k & 1] is not None:
                events[k & 1].synchronize()  # the slab staged two iterations ago has left the pinned buffer
            np.copyto(s[: r1 - r0].numpy(), emb[r0:r1])  # page cache -> pinned buffer, one host copy
            emb_all[r0:r1].copy_(s[: r1 - r0], non_blocking=True)
            events[k & 1] = torch.cuda.Event()
            events[k & 1].record()
        store = cls(names, off, rel_all, emb_all)
        store._raise_on_nan()
        return store


class WindowIndex:
    """CSR over the chunks of a dataset: for chunk i of record ``ent[i]`` with window ``[st[i], hist_end[i])`` the rows
    of the store whose ``rel_time`` falls in the window, in file order, and ``tau = rel_time - st`` (fp32 of the double
    difference) -- lib/parse_datasets.py:203-209 for every chunk at once."""

    def __init__(self, store: EmbeddingStore, ent, st, hist_end):
        if not store.emb_all.is_cuda:
            raise RuntimeError("WindowIndex: the store must live on a CUDA device -- the immtsf path has no CPU fallback")
        self.store = store
        dev = store.device
        ent = np.asarray(ent, dtype=np.int32)
        st = np.asarray(st, dtype=np.float64)
        he = np.asarray(hist_end, dtype=np.float64)
        if not (ent.ndim == st.ndim == he.ndim == 1 and len(ent) == len(st) == len(he)):
            raise ValueError("WindowIndex: ent, st, hist_end must be 1-D arrays of one length")
        if len(ent) and (ent.min() < 0 or ent.max() >= len(store.names)):
            raise ValueError("WindowIndex: record index out of range")
        n = self.n = len(ent)
        self.ent = torch.from_numpy(ent).to(dev)
        self.st = torch.from_numpy(st).to(dev)
        self.hist_end = torch.from_numpy(he).to(dev)
        counts = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.chunk_offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
        overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        s = ops._stream()
        p = ops._p
        _lib.call("immtsf_window_count", p(store.rel_all), p(store.entity_offsets), p(self.ent), p(self.st), p(self.hist_end), n,
                  p(counts), s)
        _lib.call("immtsf_exclusive_scan_i32", p(counts), n, p(self.chunk_offsets), p(overflow), s)
        self.chunk_offsets_host = self.chunk_offsets.cpu().numpy().astype(np.int64)  # one sync per dataset
        if int(overflow.item()):
            raise ValueError("WindowIndex: more than 2^31 - 1 selected notes")
        total = int(self.chunk_offsets_host[-1])
        self.chunk_rows = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
        self.chunk_tau = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        _lib.call("immtsf_window_fill", p(store.rel_all), p(store.entity_offsets), p(self.ent), p(self.st), p(self.hist_end), n,
                  p(self.chunk_offsets), p(self.chunk_rows), p(self.chunk_tau), s)
        self.counts_host = np.diff(self.chunk_offsets_host)

    def nonempty(self) -> np.ndarray:
        """Chunk ids that select at least one note -- the reference drops the others (lib/parse_datasets.py:217-221)."""
        return np.nonzero(self.counts_host > 0)[0]

    def batch(self, chunk_ids, flags: Optional[torch.Tensor] = None) -> ops.RaggedNotes:
        """The batch ``[chunk_ids[0], chunk_ids[1], ...]`` in the kernels' ragged layout: what the reference's
        multimodal_collate (:786-819) followed by the pad -> CSR adapter would produce, with one gather launch."""
        ids = np.asarray(chunk_ids, dtype=np.int64)
        if ids.ndim != 1 or (len(ids) and (ids.min() < 0 or ids.max() >= self.n)):
            raise ValueError("WindowIndex.batch: chunk ids out of range")
        B = len(ids)
        dev = self.store.device
        d_m = self.store.d_model
        cnt = self.counts_host[ids] if B else np.zeros(0, dtype=np.int64)
        off = np.zeros(B + 1, dtype=np.int32)
        off[1:] = np.cumsum(cnt)
        total, N_max = int(off[-1]), int(cnt.max()) if B else 0
        M_alloc = max(ops.round_up(max(total, 1), 128), 128)
        host = torch.empty(2 * B + 1 + max(B, 1), dtype=torch.int32).pin_memory()
        host[:B] = torch.from_numpy(ids.astype(np.int32))
        host[B:2 * B + 1] = torch.from_numpy(off)
        host[2 * B + 1:2 * B + 1 + B] = torch.from_numpy((cnt > 0).astype(np.int32))
        devbuf = host.to(dev, non_blocking=True)
        ids_d, offsets = devbuf[:B], devbuf[B:2 * B + 1]
        m_txt = devbuf[2 * B + 1:2 * B + 1 + max(B, 1)].to(torch.uint8)
        emb_flat = torch.empty(M_alloc, max(d_m, 1), dtype=torch.float32, device=dev)
        tau_flat = torch.empty(M_alloc, dtype=torch.float32, device=dev)
        p = ops._p
        _lib.call("immtsf_batch_gather", p(self.store.emb_all), self.store.emb_all.stride(0), d_m, p(self.chunk_offsets),
                  p(self.chunk_rows), p(self.chunk_tau), p(ids_d), p(offsets), B, N_max, p(emb_flat), emb_flat.stride(0),
                  p(tau_flat), ops._stream())
        m_dev = offsets[B:]
        ops.zero_pad_rows(emb_flat, max(d_m, 1), m_dev, M_alloc)
        ops.zero_pad_rows(tau_flat.view(M_alloc, 1), 1, m_dev, M_alloc)
        if flags is None:
            flags = torch.zeros(4, dtype=torch.int32, device=dev)
        return ops.RaggedNotes(B, max(N_max, 1), d_m, M_alloc, None, offsets, None, None, emb_flat, tau_flat, m_txt, flags)

    def rows_of(self, chunk_id: int) -> List[int]:
        """Selected store rows of one chunk (host copy; for inspection and tests)."""
        a, b = int(self.chunk_offsets_host[chunk_id]), int(self.chunk_offsets_host[chunk_id + 1])
        return self.chunk_rows[a:b].cpu().tolist()
