"""Philox4x32-10 exactly as libmmg's sampler keys it (TEST INFRASTRUCTURE): counter = (v, step, row_lo, row_hi),
key = (seed_lo, seed_hi); the first output word >> 8, times 2^-24, is the uniform for vocabulary index v of global row `row`."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def uniform(seed, step, row, vocab):
    c0 = np.arange(vocab, dtype=np.uint64)
    c1 = np.full(vocab, step & 0xFFFFFFFF, dtype=np.uint64)
    c2 = np.full(vocab, row & 0xFFFFFFFF, dtype=np.uint64)
    c3 = np.full(vocab, (row >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & mask, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & mask, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return ((c0 >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / (1 << 24)))
