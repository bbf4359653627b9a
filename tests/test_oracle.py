"""The oracle against itself: the literal restatement of SCFGP/SCFGP.py:92-148 (Hermite
quadrature + autograd) vs the streaming three-pass form the kernels implement."""
import numpy as np
import pytest

from oracle import scfgp_oracle as O
from scfgp_b200 import synth

SHAPES = [(400, 13, 20, 50), (300, 8, 4, 100), (333, 1, 100, 50), (257, 10, 5, 50), (150, 64, 16, 40)]


def _problem(N, D, S, M, seed=0):
    return synth.make_problem(N, D, S, M, seed=seed)


@pytest.mark.parametrize("shape", SHAPES)
def test_hermite_equals_closed_form(shape):
    N, D, S, M = shape
    X, y, h = _problem(*shape)
    c_h, a_h, L_h = O.literal_forward(h, X, y, S, M, hermite=True)
    c_c, a_c, L_c = O.literal_forward(h, X, y, S, M, hermite=False)
    assert abs(c_h - c_c) <= 1e-13 * abs(c_h)
    np.testing.assert_array_equal(a_h, a_c)


@pytest.mark.parametrize("shape", SHAPES)
def test_streaming_matches_literal_autograd(shape):
    N, D, S, M = shape
    X, y, h = _problem(*shape)
    c0, g0, a0, L0 = O.literal_cost_grad(h, X, y, S, M, hermite=True)
    c1, g1, a1, L1 = O.streaming_eval(h, X, y, S, M, chunk=97)
    assert abs(c1 - c0) <= 1e-12 * abs(c0)
    assert np.linalg.norm(g1 - g0) <= 1e-10 * np.linalg.norm(g0)
    assert np.abs(a1 - a0).max() <= 1e-9 * max(1.0, np.abs(a0).max())
    assert np.abs(L1 - L0).max() <= 1e-9 * max(1.0, np.abs(L0).max())


def test_finite_difference_gradient():
    N, D, S, M = 120, 3, 2, 6
    X, y, h = _problem(N, D, S, M, seed=3)
    _, g, _, _ = O.streaming_eval(h, X, y, S, M)
    eps = 1e-6
    for i in list(range(3)) + [3, 5, 9, 11, 15, 20]:
        hp, hm = h.copy(), h.copy()
        hp[i] += eps; hm[i] -= eps
        fd = (O.literal_forward(hp, X, y, S, M)[0] - O.literal_forward(hm, X, y, S, M)[0]) / (2 * eps)
        assert abs(fd - g[i]) <= 1e-6 * max(1.0, abs(g[i])), (i, fd, g[i])


def test_phase_invariance():
    """cost and predictions do not depend on l_p, p (finding 2)."""
    N, D, S, M = 200, 5, 3, 20
    X, y, h = _problem(N, D, S, M, seed=4)
    rng = np.random.default_rng(9)
    h2 = h.copy()
    h2[-(S + M):] = 2 * np.pi * rng.random(S + M)
    c1, g1, a1, L1 = O.streaming_eval(h, X, y, S, M)
    c2, g2, a2, L2 = O.streaming_eval(h2, X, y, S, M)
    assert abs(c1 - c2) <= 1e-12 * abs(c1)
    assert np.abs(g1[-(S + M):]).max() <= 1e-12
    Xs = rng.random((50, D))
    m1, s1 = O.literal_predict(h, Xs, a1, L1, S, M)
    m2, s2 = O.literal_predict(h2, Xs, a2, L2, S, M)
    np.testing.assert_allclose(m1, m2, atol=1e-10)
    np.testing.assert_allclose(s1, s2, atol=1e-10)


def test_phase_oracle_matches_streaming():
    N, D, S, M = 300, 7, 3, 30
    X, y, h = _problem(N, D, S, M, seed=5)
    po = O.PhaseOracle(X, y, S, M, chunk=64)
    po.begin(h); po.pass1(); po.factor(); po.pass2(); po.backward(); po.pass3()
    c, g, a, L = po.finish()
    c0, g0, a0, L0 = O.streaming_eval(h, X, y, S, M, chunk=97)
    assert abs(c - c0) <= 1e-13 * abs(c0)
    np.testing.assert_allclose(g, g0, rtol=1e-9, atol=1e-10)


def test_oracle_times_in_seconds():
    import time
    X, y, h = _problem(2000, 8, 4, 100)
    t0 = time.time()
    O.streaming_eval(h, X, y, 4, 100)
    assert time.time() - t0 < 20
