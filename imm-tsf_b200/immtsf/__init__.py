"""immtsf: host side of the B200-native IMM-TSF fusion kernels.

Layout:
  _lib.py        ctypes binding of libimmtsf.so (C ABI, include/immtsf.h)
  ops.py         tensor-level wrappers (pointers, streams, workspace)
  functional.py  autograd Functions, one per reference module
  runtime.py     NaN-flag handling, seeds, CUDA-graph capture of a whole step
  dp.py          batch-sharded data parallelism (NCCL gradient all-reduce)
The drop-in `fusions` package next to this one mirrors the reference's
fusions/ plugin surface on top of these.
"""
from ._lib import ImmtsfError, LIB_PATH, load  # noqa: F401

__all__ = ["ImmtsfError", "LIB_PATH", "load"]
