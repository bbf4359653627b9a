"""CPU restatement of the emulated-FP64 number format and digit GEMM of the product
(scfgp_b200/csrc/umma_i8.cuh, emu_kernels.cuh) -- TEST INFRASTRUCTURE ONLY: imported by
tests/ (and nothing in the product).  Integer arithmetic in NumPy int64, so it is exact; the
GPU kernels (tcgen05.mma kind::i8, int32 accumulators) must reproduce it bit for bit.

The reference computes these products in float64 (SCFGP/SCFGP.py:103,112,129 via Theano's
Dot22); this file states the *replacement arithmetic*, whose distance from float64 the parity
tests bound.
"""
from __future__ import annotations

import numpy as np

DIGIT_BITS = 8
PROD_SCALE = 2.0 ** -14


def pow2_scale(bound):
    """Digit scale of quantities bounded by `bound`: the smallest power of two >= 2*bound
    (2 for bound == 0), so that |x| / scale <= 1/2 and the leading digit stays within +-64."""
    bound = np.asarray(bound, dtype=np.float64)
    m, e = np.frexp(bound)                      # bound = m 2^e, m in [0.5, 1)
    e = np.where(m == 0.5, e - 1, e) + 1
    return np.where(bound > 0, np.ldexp(1.0, e), 2.0)


def digits_of(x_over_scale, ks):
    """Balanced base-256 digits of q = rint(x/scale * 2^(8 ks - 1)), clamped to |q| <= 2^(8 ks - 2).
    Returns int64 array [ks, ...]: digit 0 most significant (|.| <= 64), the others in [-128, 127]."""
    lim = 1 << (DIGIT_BITS * ks - 2)
    q = np.rint(np.asarray(x_over_scale, dtype=np.float64) * 2.0 ** (DIGIT_BITS * ks - 1)).astype(np.int64)
    q = np.clip(q, -lim, lim)
    out = np.zeros((ks,) + q.shape, dtype=np.int64)
    for i in range(ks - 1, 0, -1):
        lo = ((q + 128) & 255) - 128
        out[i] = lo
        q = (q - lo) >> DIGIT_BITS
    out[0] = q
    return out


def value_of(digits, scale):
    """Inverse map: scale * 2^-(8 ks - 1) * sum_i d_i 256^(ks-1-i)  (exact in float64 for ks <= 6)."""
    ks = digits.shape[0]
    q = np.zeros(digits.shape[1:], dtype=np.int64)
    for i in range(ks):
        q = q * 256 + digits[i]
    return q.astype(np.float64) * 2.0 ** (-(DIGIT_BITS * ks - 1)) * scale


def emu_gemm(A, B, ks):
    """C = A . B^T through per-row power-of-two scales, `ks` digits and the truncated digit
    products (classes i + j < ks), combined exactly like the GPU epilogue (umma_i8.cuh):
    ks <= 5: S = sum_q acc_q 256^(ks-1-q) exactly in int64, v = float64(S) 2^(-8 (ks-1)) (one rounding);
    ks >= 6: v = fma(acc_q, 2^(-8q), v) for q = ks-1 .. 0;  then C = v * (scaleA 2^-14) * scaleB."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    sa = pow2_scale(np.abs(A).max(axis=1))
    sb = pow2_scale(np.abs(B).max(axis=1))
    dA = digits_of(A * (1.0 / sa)[:, None], ks)          # scaling by a power of two is exact
    dB = digits_of(B * (1.0 / sb)[:, None], ks)
    v = np.zeros((A.shape[0], B.shape[0]))
    total = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
    for q in range(ks - 1, -1, -1):
        acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.int64)
        for i in range(q + 1):
            acc += dA[i] @ dB[q - i].T
        assert np.abs(acc).max(initial=0) < 2 ** 31, "int32 accumulator would overflow"
        v = v + acc.astype(np.float64) * 2.0 ** (-DIGIT_BITS * q)   # product exact -> one rounding, like fma
        total += acc << (DIGIT_BITS * (ks - 1 - q))
    if ks <= 5:
        v = total.astype(np.float64) * 2.0 ** (-DIGIT_BITS * (ks - 1))   # int64 -> float64 rounds to nearest even
    return v * (sa * PROD_SCALE)[:, None] * sb[None, :]
