"""CSR-emitting collate for the text side of a batch.

The reference's `multimodal_collate` (lib/parse_datasets.py:764-824) pads every sample's list of (time, embedding)
pairs to the batch maximum with `pad_sequence` (:792, :817-819), and the TTF modules then recover the ragged
structure from the zero rows (TTF_RecAvg.py:69).  The dataset already holds ragged per-sample lists (:209-213), so
this collate emits the ragged layout the kernels consume directly -- `emb_flat [sum N_i, d_model]`, `tau_flat`,
`offsets` -- and the padded tensor, the content mask pass and the compaction copy never exist:
HBM traffic drops from B*N_max rows (read twice, written once) to sum N_i rows (written once).

    r = ragged_collate([(tau_i, emb_i), ...], device)          # tau_i [N_i], emb_i [N_i, d_model]
    Y_out = fusion.forward_csr(r, t_hat, Y_ts)                  # same results as fusion(padded notes, padded tau, ...)

Difference to the padded path, by construction: validity comes from the list lengths, not from the content, so an
all-zero embedding row that a caller passes explicitly counts as a note here (the reference would mask it).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch

from . import ops


def ragged_collate(samples: Sequence[Tuple[torch.Tensor, torch.Tensor]], device, flags: torch.Tensor = None) -> ops.RaggedNotes:
    """samples[i] = (tau_i [N_i], emb_i [N_i, d_model]); N_i may be 0.  Returns the kernels' ragged layout."""
    B = len(samples)
    counts = [int(e.shape[0]) for _, e in samples]
    d_m = int(samples[0][1].shape[1]) if B else 0
    N_max = max(counts) if counts else 0
    total = sum(counts)
    M_alloc = max(ops.round_up(max(total, 1), 128), 128)
    dev = torch.device(device)
    emb_flat = torch.zeros(M_alloc, max(d_m, 1), dtype=torch.float32, device=dev)  # pad rows stay zero
    tau_flat = torch.zeros(M_alloc, dtype=torch.float32, device=dev)
    if total:
        emb_flat[:total].copy_(torch.cat([e.reshape(-1, d_m).float() for _, e in samples if e.shape[0]]), non_blocking=True)
        tau_flat[:total].copy_(torch.cat([t.reshape(-1).float() for t, e in samples if e.shape[0]]), non_blocking=True)
    off = [0]
    for c in counts:
        off.append(off[-1] + c)
    offsets = torch.tensor(off, dtype=torch.int32).to(dev, non_blocking=True)
    m_txt = torch.tensor([1 if c > 0 else 0 for c in counts] or [0], dtype=torch.uint8).to(dev, non_blocking=True)
    if flags is None:
        flags = torch.zeros(4, dtype=torch.int32, device=dev)
    r = ops.RaggedNotes(B, max(N_max, 1), d_m, M_alloc, None, offsets, None, None, emb_flat, tau_flat, m_txt, flags)
    if total and d_m:
        ops.nan_check(emb_flat[:total], flags, ops.FLAG_V)  # the reference's guard on V (TTF_RecAvg.py:75)
    return r


def pad_from_ragged(samples: Sequence[Tuple[torch.Tensor, torch.Tensor]], device):
    """What the reference collate would have produced (zero tail padding): for tests and for feeding the padded API."""
    B = len(samples)
    N_max = max(max((int(e.shape[0]) for _, e in samples), default=0), 1)
    d_m = int(samples[0][1].shape[1])
    notes = torch.zeros(B, N_max, d_m)
    tau = torch.zeros(B, N_max)
    for i, (t, e) in enumerate(samples):
        n = int(e.shape[0])
        notes[i, :n] = e
        tau[i, :n] = t
    return notes.to(device), tau.to(device)
