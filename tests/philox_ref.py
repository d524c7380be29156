"""Host mirror of the kernels' dropout RNG (csrc/common.cuh): Philox4x32-10 keyed
by the seed, counter = (idx>>3 lo, idx>>3 hi, site, 0); element idx uses the 16-bit
field (idx & 1) of word (idx>>1) & 3; keep iff field >= floor(p * 2^16).  Test
infrastructure."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint32).copy() for x in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def keep_mask(seed: int, site: int, idx: np.ndarray, p: float) -> np.ndarray:
    """0/1 keep mask for flat element indices `idx` (any shape)."""
    idx = np.asarray(idx, dtype=np.uint64)
    if p <= 0:
        return np.ones(idx.shape, dtype=np.float32)
    thr = min(int(p * 65536.0), 65535)
    c = idx >> np.uint64(3)
    c0 = (c & MASK32).astype(np.uint32)
    c1 = (c >> np.uint64(32)).astype(np.uint32)
    z = np.zeros_like(c0)
    w = philox4x32_10(c0, c1, np.full_like(c0, site), z, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    sel = ((idx >> np.uint64(1)) & np.uint64(3)).astype(np.int64)
    word = np.choose(sel, w)
    field = np.where((idx & np.uint64(1)) == 1, word >> np.uint32(16), word & np.uint32(0xFFFF))
    return (field >= np.uint32(thr)).astype(np.float32)


def realised_p(p: float) -> float:
    """The drop rate the kernels realise for a nominal p: floor(p * 2^16) / 2^16 (what 1/(1-p) must use)."""
    return min(int(p * 65536.0), 65535) / 65536.0 if p > 0 else 0.0
