"""The oracle pinned to the reference's own compiled Theano functions.

``tests/golden/reference_graph_fixtures.npz`` holds outputs of the reference's pickled
``train_func`` / ``pred_func`` graphs (SCFGP/SCFGP.py:132-135, 145-148), evaluated by the NumPy
interpreter ``oracle/theano_graph.py`` on seeded inputs (``tests/golden/make_reference_graph_fixtures.py``).
Here the CPU oracle (literal graph, streaming form, autograd gradient) is checked against those
reference outputs; ``tests/test_gpu_reference_graph.py`` checks the CUDA kernels against the same
file.  When ``/root/reference`` is present (build container) the interpreter is also run live.
"""
import os

import numpy as np
import pytest

from oracle import scfgp_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = os.path.join(HERE, "golden", "reference_graph_fixtures.npz")
REFERENCE_PKL = "/root/reference/experiments/boston_housing/boston_scfgp.pkl"
CASES = ["trained_n400", "initlaw_n257", "illcond_n1000", "tiny_n7"]
# Round-off amplification of each case: (y'y - b'b)/sigma_n^2 cancels and cond(A) ~ N sigma_f^2/sigma_n^2.
# illcond_n1000 has sigma_n^2 = e^-8: two float64 evaluation orders of the same graph (Solve vs
# inv(L) products, Gemm vs dot + add) already differ by ~1e-12 relative in the cost there.
LOOSEN = {"trained_n400": 1.0, "initlaw_n257": 1.0, "illcond_n1000": 300.0, "tiny_n7": 1.0}


@pytest.fixture(scope="module")
def fx():
    with np.load(FIXTURES, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def case(fx, name):
    keys = ("X", "y", "Xs", "params", "cost", "alpha", "Li", "mu", "std")
    return {k: fx[f"{name}_{k}"] for k in keys}


def test_fixture_inventory(fx):
    assert list(fx["cases"]) == CASES
    assert (int(fx["D"]), int(fx["S"]), int(fx["M"])) == (13, 20, 50)
    # the compiled graph the vectors came from (Theano fast_run output): BLAS, fused loops, LAPACK
    census = dict(s.split(":") for s in fx["op_census_train"])
    assert int(census["Cholesky"]) == 1 and int(census["MatrixInverse"]) == 1 and int(census["Solve"]) == 3
    assert int(census["Dot22"]) == 5 and int(census["Gemm"]) == 1 and sum(map(int, census.values())) == 85
    assert sum(int(s.split(":")[1]) for s in fx["op_census_pred"]) == 37


@pytest.mark.parametrize("name", CASES)
def test_literal_forward_matches_reference_train_func(fx, name):
    """a3-a9, a11: cost, alpha, Li of the restated graph == the reference's compiled train_func."""
    c = case(fx, name)
    S, M = int(fx["S"]), int(fx["M"])
    cost, alpha, Li = O.literal_forward(c["params"], c["X"], c["y"], S, M)
    k = LOOSEN[name]
    assert abs(cost - float(c["cost"])) <= k * 1e-12 * abs(float(c["cost"]))
    scale = max(1.0, np.abs(c["Li"]).max())
    # alpha: the reference graph solves L^T alpha = beta, the oracle multiplies by inv(L)^T; the two
    # differ by cond(A)*eps
    assert np.abs(alpha - c["alpha"]).max() <= k * 1e-9 * max(1.0, np.abs(c["alpha"]).max())
    assert np.abs(Li - c["Li"]).max() <= k * 1e-10 * scale
    assert alpha.shape == c["alpha"].shape == (140, 1) and Li.shape == (140, 140)


@pytest.mark.parametrize("name", CASES)
def test_closed_form_and_streaming_match_reference(fx, name):
    """The closed form of the 30-node Hermite term (a7) and the three-pass form the kernels
    implement reproduce the reference cost."""
    c = case(fx, name)
    S, M = int(fx["S"]), int(fx["M"])
    cost_cf, _, _ = O.literal_forward(c["params"], c["X"], c["y"], S, M, hermite=False)
    cost_st, _, alpha, Li = O.streaming_eval(c["params"], c["X"], c["y"], S, M, chunk=97)
    ref, k = float(c["cost"]), LOOSEN[name]
    assert abs(cost_cf - ref) <= k * 1e-12 * abs(ref)
    assert abs(cost_st - ref) <= k * 1e-11 * abs(ref)
    assert np.abs(alpha - c["alpha"]).max() <= k * 1e-9 * max(1.0, np.abs(c["alpha"]).max())
    assert np.abs(Li - c["Li"]).max() <= k * 1e-10 * max(1.0, np.abs(c["Li"]).max())


@pytest.mark.parametrize("name", CASES)
def test_literal_predict_matches_reference_pred_func(fx, name):
    """a12: mu_pred (T,1), std_pred (T,) at the reference's alpha, Li."""
    c = case(fx, name)
    mu, std = O.literal_predict(c["params"], c["Xs"], c["alpha"], c["Li"], int(fx["S"]), int(fx["M"]))
    assert mu.shape == c["mu"].shape and std.shape == c["std"].shape and std.ndim == 1
    assert np.abs(mu - c["mu"]).max() <= 1e-11 * max(1.0, np.abs(c["mu"]).max())
    assert np.abs(std - c["std"]).max() <= 1e-12 * max(1.0, np.abs(c["std"]).max())


def test_checkpoint_arrays_through_reference_pred_func(fx):
    """The alpha / Li the checkpoint stores, pushed through the reference pred_func."""
    mu, std = O.literal_predict(fx["ckpt_params"], fx["ckpt_Xs"], fx["ckpt_alpha"], fx["ckpt_Li"], 20, 50)
    assert np.abs(mu - fx["ckpt_mu"]).max() <= 1e-12 and np.abs(std - fx["ckpt_std"]).max() <= 1e-12


@pytest.mark.parametrize("name", CASES)
def test_gradient_matches_finite_differences_of_reference_cost(fx, name):
    """a10: ``TT.grad(cost, params)`` (SCFGP.py:129) is the derivative of the reference's cost graph;
    its 4th-order central differences pin both the autograd gradient of the literal restatement and
    the hand adjoints of the streaming form."""
    c = case(fx, name)
    S, M = int(fx["S"]), int(fx["M"])
    _, g_auto, _, _ = O.literal_cost_grad(c["params"], c["X"], c["y"], S, M, threads=1)
    _, g_hand, _, _ = O.streaming_eval(c["params"], c["X"], c["y"], S, M, chunk=97)
    idx, fd = fx[f"{name}_grad_idx"], fx[f"{name}_grad_fd"]
    tol = LOOSEN[name] * 2e-7 * max(1.0, np.abs(fd).max())
    for g in (g_auto, g_hand):
        assert np.abs(g[idx] - fd).max() <= tol, np.abs(g[idx] - fd).max()
        proj = fx[f"{name}_dirs"] @ g
        assert np.abs(proj - fx[f"{name}_dir_fd"]).max() <= tol
    if name == "trained_n400":                              # every coordinate of the P-vector
        assert idx.size == g_hand.size == 1333
        assert np.linalg.norm(g_hand - fd) <= 1e-8 * np.linalg.norm(fd)


# --------------------------------------------------------------------------------------
# live: run the interpreter itself (build container only)
# --------------------------------------------------------------------------------------
needs_reference = pytest.mark.skipif(not os.path.exists(REFERENCE_PKL), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    from oracle.theano_graph import ReferenceFunctions
    return ReferenceFunctions()


@needs_reference
def test_interpreter_reproduces_committed_fixtures(fx, ref):
    for name in CASES:
        c = case(fx, name)
        cost, alpha, Li = ref.train_func(c["X"], c["y"], c["params"])
        assert abs(cost - float(c["cost"])) <= 1e-13 * abs(float(c["cost"]))
        assert np.abs(alpha - c["alpha"]).max() <= 1e-9 and np.abs(Li - c["Li"]).max() <= 1e-10
        mu, std = ref.pred_func(c["Xs"], c["alpha"], c["Li"], c["params"])
        np.testing.assert_allclose(mu, c["mu"], rtol=0, atol=1e-13)
        np.testing.assert_allclose(std, c["std"], rtol=0, atol=1e-13)


@needs_reference
def test_interpreter_vs_oracle_on_fresh_inputs(ref):
    """Inputs that are in no fixture: the restatement tracks the reference graph everywhere."""
    rng = np.random.default_rng(987)
    for N in (5, 140, 613):
        X = rng.random((N, 13)) * rng.uniform(0.5, 2.0)
        y = rng.standard_normal((N, 1))
        h = ref.trained_params + 0.1 * rng.standard_normal(1333)
        cost, alpha, Li = ref.train_func(X, y, h)
        c0, a0, L0 = O.literal_forward(h, X, y, 20, 50)
        assert abs(cost - c0) <= 1e-12 * abs(c0)
        assert np.abs(alpha - a0).max() <= 1e-9 * max(1, np.abs(a0).max())
        assert np.abs(Li - L0).max() <= 1e-10 * max(1, np.abs(L0).max())
        Xs = rng.random((17, 13))
        mu, std = ref.pred_func(Xs, alpha, Li, h)
        mu0, std0 = O.literal_predict(h, Xs, alpha, Li, 20, 50)
        assert np.abs(mu - mu0).max() <= 1e-11 and np.abs(std - std0).max() <= 1e-12


@needs_reference
def test_reference_graph_is_phase_invariant(ref):
    """SURVEY section 0 finding 2, now on the reference's own graph: the cost does not move with the
    phase parameters (last S+M entries)."""
    rng = np.random.default_rng(5)
    X, y = rng.random((200, 13)), rng.standard_normal((200, 1))
    h = ref.trained_params.copy()
    c0 = ref.train_func(X, y, h)[0]
    h[-70:] = 2 * np.pi * rng.random(70)
    assert abs(ref.train_func(X, y, h)[0] - c0) <= 1e-12 * abs(c0)
