"""TEST INFRASTRUCTURE (imported only by tests/): numpy restatement of the fp32-on-tensor-cores operand split of libmmg (`mmg_split3`,
include/mmg.h) and of the product it feeds.

A fp32 value x is written as three bf16 terms  hi = bf16(x), mid = bf16(x - hi), lo = bf16(x - hi - mid)  (round to nearest even; both
subtractions are exact in fp32), so hi + mid + lo == x to 24 mantissa bits.  A left operand row is laid out along K as
[lo | hi | mid | mid | hi | hi], a right operand row as [hi | lo | mid | hi | mid | hi]: one bf16 x bf16 -> fp32 product over the 6K columns
then accumulates  lo*hi + hi*lo + mid*mid + mid*hi + hi*mid + hi*hi,  i.e. the fp32 product without the three terms below 2^-24 of it
(mid*lo, lo*mid, lo*lo).  There is no reference counterpart: the reference computes these products with fp32 torch ops
(muse_maskgit_pytorch.py:85-89, 118-124, 225; vqgan_vae.py:224-277), which is what the six-term product reproduces to fp32 accuracy."""
import numpy as np

LEFT = ("lo", "hi", "mid", "mid", "hi", "hi")
RIGHT = ("hi", "lo", "mid", "hi", "mid", "hi")


def bf16_round(x):
    """fp32 array -> nearest bf16 value (ties to even), returned as fp32."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return rounded.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def terms(x):
    x = np.asarray(x, dtype=np.float32)
    hi = bf16_round(x)
    r1 = (x - hi).astype(np.float32)
    mid = bf16_round(r1)
    lo = bf16_round((r1 - mid).astype(np.float32))
    return {"hi": hi, "mid": mid, "lo": lo}


def split3(x2d, side):
    """[rows, K] fp32 -> [rows, 6K] array of bf16 VALUES (fp32 dtype) in the column order libmmg writes for `side` (0 left, 1 right)."""
    t = terms(x2d)
    return np.concatenate([t[name] for name in (LEFT if side == 0 else RIGHT)], axis=1)


def product(a, w):
    """a [M, K] . w [N, K]^T through the split operands, accumulated in float64 (the tensor core accumulates in fp32): the value the six cross
    terms carry before any accumulation rounding."""
    return split3(a, 0).astype(np.float64) @ split3(w, 1).astype(np.float64).T
