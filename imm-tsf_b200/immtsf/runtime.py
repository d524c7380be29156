"""Small runtime pieces shared by the fusion modules: the deferred NaN guard
and dropout seeding.

The reference raises ValueError on NaN at five places (fusions/FusionModel.py:
103-112, fusions/TTF_*.py:75/116, fusions/MMF_XAttn_Add.py:84-91), each one a
device->host sync.  Here every kernel that already touches the data sets a
device flag and the flags are read ONCE per forward (one sync), keeping the
ValueError convention.  IMMTSF_NAN_CHECK=0 (or CUDA-graph capture) skips the
read entirely."""
from __future__ import annotations

import os

import torch

from . import ops

_MESSAGES = {
    ops.FLAG_V: "Input embeddings V contain NaN values.",
    ops.FLAG_Y: "Y_ts contains NaN values.",
    ops.FLAG_E: "E_txt contains NaN values.",
    ops.FLAG_OUT: "Y_out contains NaN values.",
}


def nan_check_enabled() -> bool:
    if os.environ.get("IMMTSF_NAN_CHECK", "1") == "0":
        return False
    return not torch.cuda.is_current_stream_capturing()


def new_flags(device) -> torch.Tensor:
    return torch.zeros(4, dtype=torch.int32, device=device)


def raise_on_flags(flags: torch.Tensor, slots=(ops.FLAG_Y, ops.FLAG_V, ops.FLAG_E, ops.FLAG_OUT)):
    """One device->host read; raises the reference's ValueError for the first set flag."""
    if not nan_check_enabled():
        return
    host = flags.tolist()
    for s in slots:
        if host[s]:
            raise ValueError(_MESSAGES[s])


class SeedSource:
    """Dropout seeds.  Default: drawn from torch's CPU generator per call.  A fixed
    seed can be pinned (tests, CUDA graphs replaying one captured step)."""

    def __init__(self):
        self.fixed = None

    def next(self) -> int:
        return self.fixed if self.fixed is not None else ops.new_seed()


SEEDS = SeedSource()
