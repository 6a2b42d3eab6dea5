"""Philox4x32-10 exactly as libmmg's sampler keys it (TEST INFRASTRUCTURE): counter = (v, step, row_lo, row_hi),
key = (seed_lo, seed_hi); the first output word >> 8, times 2^-24, is the uniform for vocabulary index v of global row `row`."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def uniform_at(seed, step, row, c0):
    """single draw at an explicit first counter word (the critic-noise stream uses c0 = 0xFFFFFFFF)."""
    return _uniform(seed, step, row, np.array([c0], dtype=np.uint64))[0]


def uniform(seed, step, row, vocab):
    return _uniform(seed, step, row, np.arange(vocab, dtype=np.uint64))


def _uniform(seed, step, row, c0):
    vocab = c0.shape[0]
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


# ---------------------------------------------------------------------------------------------------------------------------
# ATen's CUDA `tensor.uniform_(0, 1)` stream (fp32), restated so the in-kernel "aten" noise mode can be checked without a
# 4 GiB injected-noise tensor.  Restates, from the published PyTorch sources (aten/src/ATen/native/cuda/
# DistributionTemplates.h: calc_execution_policy, distribution_elementwise_grid_stride_kernel, uniform_kernel) and cuRAND's
# curand_init / curand_uniform4 for Philox4_32_10:
#   block = 256, grid = min(SMs * (maxThreadsPerSM / 256), ceil(numel / 256)), stride S = 256 * grid, unroll 4
#   thread t handles elements li = t + S * (4 * j + ii): word ii of Philox(counter = (offset/4 + j, subsequence = t), key = seed)
#   u = word * 2^-32 + 2^-33 in fp32 (so u in (0, 1]);  u == 1 -> 0
#   each call advances the generator offset by ((numel - 1) / (S * 4) + 1) * 4
# tests/test_gpu_aten_rng.py pins this model against torch.cuda on the GPU box.
# ---------------------------------------------------------------------------------------------------------------------------
def philox4(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on uint64 numpy arrays holding 32-bit words; returns the four output words."""
    mask = np.uint64(0xFFFFFFFF)
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & mask, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & mask, lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def aten_stride(numel, sm_count, max_threads_per_sm):
    grid = min(sm_count * (max_threads_per_sm // 256), (numel + 255) // 256)
    return 256 * grid


def aten_offset_increment(numel, stride):
    return ((numel - 1) // (stride * 4) + 1) * 4


def aten_uniform(seed, offset, numel, stride, index=None):
    """Values of `torch.empty(numel, device='cuda').uniform_(0, 1)` at flat positions `index` (default: all) for a generator
    with (seed, offset) on a device whose launch stride is `stride`."""
    li = np.arange(numel, dtype=np.uint64) if index is None else np.asarray(index, dtype=np.uint64)
    t, k = li % np.uint64(stride), li // np.uint64(stride)
    j, ii = k // np.uint64(4), (k % np.uint64(4)).astype(np.int64)
    ctr = np.uint64(offset // 4) + j
    words = philox4(ctr & np.uint64(0xFFFFFFFF), ctr >> np.uint64(32), t & np.uint64(0xFFFFFFFF), t >> np.uint64(32),
                    seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    w = np.choose(ii, [x.astype(np.float32) for x in words])          # uint32 -> fp32, round to nearest even
    u = w * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)
    return np.where(u == np.float32(1.0), np.float32(0.0), u).astype(np.float32)
