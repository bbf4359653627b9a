"""The emulated-FP64 number format on the CPU (oracle/emu_digits.py): representation error,
the leading-digit-planes property the mixed digit counts rely on, and the int32 headroom."""
import numpy as np
import pytest

from oracle import emu_digits as ED


@pytest.mark.parametrize("ks", [4, 5, 6])
def test_digits_reconstruct_within_half_ulp(ks):
    rng = np.random.default_rng(ks)
    x = rng.standard_normal(20000) * np.exp(rng.standard_normal(20000) * 4)
    scale = ED.pow2_scale(np.abs(x).max())
    assert scale >= 2 * np.abs(x).max() and scale < 4 * np.abs(x).max()
    d = ED.digits_of(x / scale, ks)
    assert np.abs(d[0]).max() <= 64 and d[1:].min() >= -128 and d[1:].max() <= 127
    err = np.abs(ED.value_of(d, scale) - x)
    assert err.max() <= 0.5 * scale * 2.0 ** (-(8 * ks - 1))


def test_extremes_stay_inside_int8():
    for ks in (4, 5, 6, 7, 8):
        x = np.array([0.5, -0.5, 0.5 - 2.0 ** -40, -0.5 + 2.0 ** -40, 0.0, 2.0 ** -60, 0.499999999, -0.25])
        d = ED.digits_of(x, ks)
        assert np.abs(d[0]).max() <= 64 and d[1:].min() >= -128 and d[1:].max() <= 127
        assert np.all(d[:, 4] == 0)
        # out-of-range input is clamped, never wrapped
        assert ED.digits_of(np.array([0.75]), ks)[0, 0] == 64


@pytest.mark.parametrize("ks,n", [(5, 4), (6, 4), (6, 5), (8, 5)])
def test_leading_planes_are_a_valid_shorter_representation(ks, n):
    """Balanced digits: dropping the tail of a ks-digit expansion rounds to nearest, so the
    first n planes represent x to within (1/2 + 2^-8) units of digit n (scfgp_set_precision2)."""
    rng = np.random.default_rng(ks * 10 + n)
    x = rng.uniform(-0.5, 0.5, 50000)
    d = ED.digits_of(x, ks)
    short = ED.value_of(d[:n], 1.0)
    unit = 2.0 ** (-(8 * n - 1))
    assert np.abs(short - x).max() <= (0.5 + 2.0 ** -8 + 2.0 ** -16) * unit


def class_bound(ks, depth):
    """Largest possible |class sum|: digit 0 is within +-64, the others within +-128."""
    m = [64] + [128] * (ks - 1)
    return depth * max(sum(m[i] * m[q - i] for i in range(q + 1)) for q in range(ks))


def test_int32_headroom_of_the_depths_in_use():
    """The library refuses reduction depths beyond (2^31 - 1) / (16384 (ks - 1)) (scfgp_set_chunk_rows,
    scfgp_set_precision): check that bound, and that the defaults sit far inside it."""
    for ks in range(4, 9):
        limit = (2 ** 31 - 1) // (16384 * (ks - 1))
        assert class_bound(ks, limit) < 2 ** 31 <= class_bound(ks, limit + 130)
        assert limit >= 16384                       # 8192-row chunks, 2K = 8320 (cfg5) fit for every digit count
    # adversarial digits reach the bound exactly
    ks, depth = 5, 1000
    a = np.full((1, depth), 0.5); b = np.full((1, depth), 0.5)
    dA, dB = ED.digits_of(a, ks), ED.digits_of(b, ks)
    assert int((dA[0] @ dB[0].T)[0, 0]) == 64 * 64 * depth


@pytest.mark.parametrize("ks", [4, 5, 6])
def test_emu_gemm_error_bound(ks):
    rng = np.random.default_rng(ks)
    A = rng.standard_normal((70, 300)) * np.exp(3 * rng.standard_normal((70, 1)))
    B = rng.standard_normal((50, 300)) * np.exp(3 * rng.standard_normal((50, 1)))
    C = ED.emu_gemm(A, B, ks)
    bound = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * A.shape[1]
    assert (np.abs(C - A @ B.T) / bound).max() <= 4 * 2.0 ** (-(8 * ks - 2))


def test_emu_gemm_exact_on_small_integers():
    rng = np.random.default_rng(0)
    A = rng.integers(-100, 100, (40, 64)).astype(float)
    B = rng.integers(-100, 100, (30, 64)).astype(float)
    np.testing.assert_array_equal(ED.emu_gemm(A, B, 4), A @ B.T)
