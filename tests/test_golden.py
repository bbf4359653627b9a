"""The one reference-produced numerical pin: arrays recovered from
experiments/boston_housing/boston_scfgp.pkl (tests/golden/make_golden.py).  X and y of that
run were not stored, so the fixture pins structure (column order, scale, jitter, sigma_n^2
parametrisation, lower-triangular inverse factor), not a cost value; SURVEY.md section 4.3."""
import numpy as np

from oracle import scfgp_oracle as O


def _unpacked(golden):
    D, S, M = int(golden["D"]), int(golden["S"]), int(golden["M"])
    return D, S, M, O.unpack_params(golden["params"], D, S, M)


def test_shapes_and_layout(golden):
    D, S, M, _ = _unpacked(golden)
    assert (D, S, M) == (13, 20, 50)
    assert golden["params"].shape == (O.num_params(D, S, M),) == (1333,)
    assert golden["Li"].shape == (140, 140) and golden["alpha"].shape == (140, 1)
    np.testing.assert_allclose(golden["params"][:3],
                               [-1.575304860841505, -0.6655536119352673, -1.7756344623098925], rtol=0, atol=0)


def test_Li_is_inverse_of_lower_cholesky(golden):
    Li = golden["Li"]
    assert np.abs(np.triu(Li, 1)).max() < 1e-13          # lower triangular: Theano's cholesky is lower
    np.testing.assert_allclose(np.diag(Li)[:3],
                               [0.7400923412458766, 0.8280543906237307, 1.2999690852555987], rtol=1e-15)
    L = np.linalg.inv(Li)
    A = L @ L.T
    # the oracle's factor step (SCFGP.py:106-107 restated) reproduces the stored Li from A
    Li2 = np.linalg.inv(np.linalg.cholesky(A))
    np.testing.assert_allclose(Li2, Li, rtol=0, atol=1e-9)


def test_gram_pair_sum_identity(golden):
    """G[j,j] + G[j+J,j+J] = N * sig_f^2 * 2/M  (cos^2+sin^2 = 1): pins the column order
    [cos lat | cos main | sin lat | sin main], the scale with M (not M+S), sig2_n = exp(2a),
    the 1e-6 jitter and N = 400."""
    D, S, M, (a, b, c, *_rest) = _unpacked(golden)
    J = S + M
    L = np.linalg.inv(golden["Li"])
    A = L @ L.T
    G = A - (np.exp(2 * a) + O.EPSILON) * np.eye(2 * J)
    pair = np.diag(G)[:J] + np.diag(G)[J:]
    expected = 400 * np.exp(2 * b) * 2.0 / M
    assert abs(expected - 4.226953405125291) < 1e-12
    np.testing.assert_allclose(pair, expected, rtol=2e-14)
    assert np.linalg.eigvalsh(G).min() > -1e-9            # a Gram matrix


def test_alpha_solves_normal_equations(golden):
    """alpha = LiT Li p  <=>  A alpha = p with p = PhiT y; check consistency of the stored pair and
    the scalar summaries quoted in SURVEY 4.3."""
    Li, alpha = golden["Li"], golden["alpha"][:, 0]
    np.testing.assert_allclose(alpha[:3], [1.3999818658608723, 0.3782002510646413, -0.10093712344293192],
                               rtol=1e-14)
    L = np.linalg.inv(Li)
    A = L @ L.T
    p = A @ alpha
    beta = Li @ p
    np.testing.assert_allclose(Li.T @ beta, alpha, rtol=1e-8)
    assert abs(-2 * np.log(np.diag(Li)).sum() - (-251.45269728665568)) < 1e-9
    assert abs(beta @ beta - 380.5099665620399) < 1e-6


def test_penalty_at_trained_params(golden):
    D, S, M, (a, b, c, l_F, r_F, F, l_FC, FC) = _unpacked(golden)
    pen, _, _ = O.penalty_and_grad(l_F, F, S, M)
    assert abs(pen - 18.014982476271832) < 1e-10


def test_noise_scalars(golden):
    a, b, c = golden["params"][:3]
    assert abs(np.exp(2 * a) - 0.0428260) < 1e-6
    assert abs(np.exp(b) - 0.5139889) < 1e-6
    assert abs(np.log(1 + np.exp(c)) - 0.1564702) < 1e-6


def test_history(golden):
    assert golden["evals_COST"].shape == (33,)
    assert abs(golden["evals_COST"][-1] - (-1.061668727315276)) < 1e-12
    assert abs(golden["evals_NMSE"][-1] - 0.0732) < 1e-3
