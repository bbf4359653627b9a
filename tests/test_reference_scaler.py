"""Row f3 pinned to the reference: ``scfgp_b200/scaler.py`` against outputs of the reference's own
``SCFGP/Scaler.py`` (fit / forward_transform / backward_transform, Scaler.py:39-135), which runs in
the build container (NumPy/SciPy only).  Fixtures: ``tests/golden/reference_scaler_fixtures.npz``
made by ``tests/golden/make_reference_scaler_fixtures.py``."""
import os

import numpy as np
import pytest

from scfgp_b200 import Scaler

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = os.path.join(HERE, "golden", "reference_scaler_fixtures.npz")
REFERENCE_SRC = "/root/reference/SCFGP/Scaler.py"
ALGOS = ["min-max", "normal", "inv-normal", "auto-normal", "auto-inv-normal"]
TABLES = ["mixed", "target", "wide"]
FIELDS = ("min", "max", "mu", "std", "boxcox")


@pytest.fixture(scope="module")
def fx():
    with np.load(FIXTURES, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("table", TABLES)
def test_fit_and_transforms_equal_reference(fx, table, algo):
    """Same NumPy/SciPy calls in the same order => bit-identical state and transforms."""
    key = f"{table}_{algo}"
    sc = Scaler(algo)
    sc.fit(fx[f"{table}_Xfit"])
    assert list(sc.data["cols"]) == list(fx[key + "_cols"])
    for f in FIELDS:
        if f"{key}_{f}" in fx:
            np.testing.assert_array_equal(np.asarray(sc.data[f]), fx[f"{key}_{f}"], err_msg=f)
    fwd = sc.forward_transform(fx[f"{table}_Xfit"])
    np.testing.assert_array_equal(fwd, fx[key + "_fwd_fit"])
    np.testing.assert_array_equal(sc.forward_transform(fx[f"{table}_Xnew"]), fx[key + "_fwd_new"])
    np.testing.assert_array_equal(sc.backward_transform(fwd), fx[key + "_bwd_fit"])


def test_fixture_covers_the_special_columns(fx):
    assert list(fx["mixed_auto-normal_cols"]) == [0, 1, 3, 4, 5, 6]      # constant column 2 dropped
    assert fx["mixed_auto-normal_boxcox"][2] == 1.0                      # < 10 distinct values


@pytest.mark.skipif(not os.path.exists(REFERENCE_SRC), reason="/root/reference not present")
def test_live_reference_scaler_equals_ours_on_fresh_table():
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_scfgp_scaler", REFERENCE_SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(99)
    X = np.column_stack([rng.lognormal(size=300), rng.random(300), rng.gamma(2.0, size=300)])
    for algo in ALGOS:
        ref, ours = mod.Scaler(algo), Scaler(algo)
        ref.fit(X)
        ours.fit(X)
        np.testing.assert_array_equal(ref.forward_transform(X), ours.forward_transform(X))
        np.testing.assert_array_equal(ref.backward_transform(ref.forward_transform(X)),
                                      ours.backward_transform(ours.forward_transform(X)))
