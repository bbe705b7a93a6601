"""Philox4x32-10 counter-based uniforms, numpy restatement (TEST INFRASTRUCTURE ONLY, see ``oracle/__init__.py``).

Algorithm: Salmon, Moraes, Dror, Shaw, "Parallel Random Numbers: As Easy as 1, 2, 3" (SC'11), Philox-4x32 with 10 rounds;
multipliers 0xD2511F53 / 0xCD9E8D57 and Weyl key increments 0x9E3779B9 / 0xBB67AE85 (the Random123 / cuRAND constants).
Pinned on the Random123 known-answer vectors (``tests/test_philox.py``).

The reference itself draws scrambled-Sobol base samples (``discrete_mixed_bo/input.py:655-673, 696-713``); Philox is the
build's device-side replacement stream (BASELINE.json north_star), laid out so that -- like the reference's base samples --
the uniform depends only on (MC index i, discrete parameter column c), never on the restart:

    counter = (i, c, offset_lo, offset_hi), key = (seed_lo, seed_hi)
    u = ((r0 << 32 | r1) >> 11) * 2^-53            in [0, 1)
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32 array [..., 4]; key: uint32 array [..., 2] (broadcastable) -> uint32 array [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64)
    c0, c1, c2, c3 = c[..., 0].copy(), c[..., 1].copy(), c[..., 2].copy(), c[..., 3].copy()
    k = np.asarray(key, dtype=np.uint64)
    k0, k1 = k[..., 0].copy(), k[..., 1].copy()
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def philox_uniform(seed: int, offset: int, mc: int, n_cols: int) -> np.ndarray:
    """[mc, n_cols] float64 uniforms of the library's stream (``prb_philox_uniform``, include/prb200.h)."""
    i, c = np.meshgrid(np.arange(mc, dtype=np.uint64), np.arange(n_cols, dtype=np.uint64), indexing="ij")
    ctr = np.stack([i, c, np.full_like(i, offset & 0xFFFFFFFF), np.full_like(i, (offset >> 32) & 0xFFFFFFFF)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    r = philox4x32_10(ctr, key).astype(np.uint64)
    bits = ((r[..., 0] << np.uint64(32)) | r[..., 1]) >> np.uint64(11)
    return bits.astype(np.float64) * (1.0 / 9007199254740992.0)
