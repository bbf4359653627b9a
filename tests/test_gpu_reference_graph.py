"""The CUDA path (through the C ABI) against outputs of the reference's own compiled Theano
functions: ``tests/golden/reference_graph_fixtures.npz`` (see ``tests/test_reference_graph.py`` and
``oracle/theano_graph.py``).  No oracle in the loop here: kernels vs reference-produced numbers.

Tolerances are north_star's: cost <= 1e-6 relative; gradient, alpha / Li and predictions <= 1e-5."""
import os

import numpy as np
import pytest

from scfgp_b200 import Engine

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = os.path.join(HERE, "golden", "reference_graph_fixtures.npz")
CASES = ["trained_n400", "initlaw_n257", "illcond_n1000", "tiny_n7"]
D, S, M = 13, 20, 50


@pytest.fixture(scope="module")
def fx():
    with np.load(FIXTURES, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("precision", ["fp64", "emu5", "emu6"])
@pytest.mark.parametrize("name", CASES)
def test_eval_matches_reference_train_func(fx, name, precision):
    X, y, h = fx[f"{name}_X"], fx[f"{name}_y"].ravel(), fx[f"{name}_params"]
    eng = Engine(D, S, M, device=0, precision=precision)
    try:
        eng.set_data(X, y)
        cost, grad, alpha, Li = eng.eval(h)
    finally:
        eng.close()
    ref_cost, ref_alpha, ref_Li = float(fx[f"{name}_cost"]), fx[f"{name}_alpha"], fx[f"{name}_Li"]
    assert abs(cost - ref_cost) <= 1e-6 * abs(ref_cost), (cost, ref_cost)
    assert alpha.shape == ref_alpha.shape and Li.shape == ref_Li.shape
    assert np.abs(alpha - ref_alpha).max() <= 1e-5 * max(1.0, np.abs(ref_alpha).max())
    assert np.abs(Li - ref_Li).max() <= 1e-5 * max(1.0, np.abs(ref_Li).max())
    # gradient vs central differences of the reference's cost graph (what TT.grad differentiates)
    idx, fd = fx[f"{name}_grad_idx"], fx[f"{name}_grad_fd"]
    scale = max(1.0, np.abs(fd).max())
    assert np.abs(grad[idx] - fd).max() <= 1e-5 * scale, np.abs(grad[idx] - fd).max()
    proj = fx[f"{name}_dirs"] @ grad
    assert np.abs(proj - fx[f"{name}_dir_fd"]).max() <= 1e-5 * scale
    if name == "trained_n400":
        assert np.linalg.norm(grad - fd) <= 1e-5 * np.linalg.norm(fd)


@pytest.mark.parametrize("precision", ["fp64", "emu5"])
@pytest.mark.parametrize("name", CASES + ["ckpt"])
def test_predict_matches_reference_pred_func(fx, name, precision):
    h, Xs = fx[f"{name}_params"], fx[f"{name}_Xs"]
    alpha, Li = fx[f"{name}_alpha"], fx[f"{name}_Li"]
    eng = Engine(D, S, M, device=0, precision=precision)
    try:
        mu, std = eng.predict(h, Xs, alpha, Li)
    finally:
        eng.close()
    ref_mu, ref_std = fx[f"{name}_mu"], fx[f"{name}_std"]
    assert mu.shape == ref_mu.shape and std.shape == ref_std.shape
    assert np.abs(mu - ref_mu).max() <= 1e-5 * max(1.0, np.abs(ref_mu).max())
    assert np.abs(std - ref_std).max() <= 1e-5 * max(1.0, np.abs(ref_std).max())
