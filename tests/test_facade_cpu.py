"""Facade logic on CPU: the SCFGP training loop / predict / metrics / save-load driven through
an oracle-backed stand-in engine (test double injected in place of the CUDA Engine; the
product itself never falls back to it)."""
import numpy as np
import pytest

from oracle import scfgp_oracle as O
from scfgp_b200 import SCFGP


class OracleEngine:
    def __init__(self, D, S, M):
        self.D, self.S, self.M = D, S, M
        self.n_eval = 0

    def set_data(self, X, y, n_global=None):
        self.X, self.y = X, np.asarray(y).reshape(-1)
        self.fullX, self.fully = self.X, self.y

    def select_rows(self, idx=None, n_global=None):
        self.X, self.y = (self.fullX, self.fully) if idx is None else (self.fullX[idx], self.fully[idx])

    def eval(self, h, want_grad=True, **_):
        self.n_eval += 1
        return O.streaming_eval(h, self.X, self.y, self.S, self.M, want_grad=want_grad)

    def predict(self, h, Xs, a, L):
        return O.literal_predict(h, Xs, a, L, self.S, self.M)


def _data(seed=3, N=160, D=5):
    rng = np.random.default_rng(seed)
    X = rng.lognormal(size=(N, D))
    y = np.tanh(X @ rng.standard_normal(D) / 4) * 5 + 20 + 0.3 * rng.standard_normal(N)
    return X[:120], y[:120, None], X[120:], y[120:, None]


def test_optimize_loop_and_metrics():
    np.random.seed(1)
    Xtr, ytr, Xte, yte = _data()
    m = SCFGP(sparsity=3, nfeats=12)
    m.set_data(Xtr, ytr)
    assert m.X.min() > 0 and m.X.max() < 1 and abs((m.y ** 2).sum() - len(ytr)) < 1e-8
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, None, None, max_iter=25, cvrg_tol=1e-5, max_cvrg=8)
    ev = m.evals
    n = len(ev["COST"][1])
    assert n == 26 and all(len(ev[k][1]) == n for k in ev)      # one entry per iteration + final train_func
    assert ev["COST"][1][-1] < ev["COST"][1][0]
    assert ev["SCORE"][1][-1] == pytest.approx(ev["NMSE"][1][-1] / (1 + np.exp(-ev["MNLP"][1][-1])))
    assert ev["NMAE"][1][-1] == pytest.approx(ev["MAE"][1][-1] / np.std(yte))
    assert m.alpha.shape == (2 * (3 + 12), 1) and m.Li.shape == (30, 30)
    mu, sd = m.predict(Xte)
    assert mu.shape == (40, 1) and sd.shape == (40, 1) and np.all(sd > 0)
    f = m.get_compiled_funcs()
    assert len(f) == 3 and callable(f[0]) and f.engine is m._engine


def test_obj_without_validation_defaults_to_cost():
    np.random.seed(2)
    Xtr, ytr, _, _ = _data()
    m = SCFGP(sparsity=2, nfeats=6)
    m.set_data(Xtr, ytr)
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(obj="nmse", max_iter=5)
    assert len(m.evals["COST"][1]) == 6 and m.evals["NMSE"][1] == [0]


def test_fit_api_and_save_load(tmp_path):
    np.random.seed(3)
    Xtr, ytr, Xte, yte = _data()
    m = SCFGP(-1, 8, "phase", False)                    # rank=-1 -> S = D (documented, unpinned)
    m.set_data(Xtr, ytr)
    assert m.S == m.D == 5
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, max_iter=6)
    for attr in ("TsMAE", "TsMSE", "TsRMSE", "TsNMSE", "TsMNLP", "SCORE"):
        assert np.isfinite(getattr(m, attr))
    path = str(tmp_path / "m.npz")
    m.save(path)
    m2 = SCFGP().load(path)
    assert (m2.S, m2.M, m2.D) == (m.S, m.M, m.D)
    np.testing.assert_array_equal(m2.params, m.params)
    m2._engine = OracleEngine(m2.D, m2.S, m2.M)
    m2.build_models()
    mu1, sd1 = m.predict(Xte)
    mu2, sd2 = m2.predict(Xte)
    np.testing.assert_allclose(mu2, mu1, rtol=1e-12)
    np.testing.assert_allclose(sd2, sd1, rtol=1e-12)


def test_minibatch_mode_runs():
    np.random.seed(4)
    Xtr, ytr, _, _ = _data()
    m = SCFGP(sparsity=2, nfeats=6)
    m.set_data(Xtr, ytr)
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(max_iter=3, nbatches=2, batchsize=40)
    assert len(m.evals["COST"][1]) == 4


def test_load_reference_golden(golden, tmp_path):
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "boston_scfgp_golden.npz")
    m = SCFGP().load_reference_golden(path)
    assert (m.S, m.M, m.D) == (20, 50, 13) and m.Li.shape == (140, 140)
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.build_models()
    rng = np.random.default_rng(0)
    lo, hi = golden["xs_min"], golden["xs_max"]
    Xs = lo + (hi - lo) * rng.random((30, 13))          # raw-unit inputs inside the fitted range
    mu, sd = m.predict(Xs)
    assert mu.shape == (30, 1) and np.all(np.isfinite(mu)) and np.all(sd > 0)
    assert 0 < mu.mean() < 60                            # Boston house prices (k$)


class CountingEngine(OracleEngine):
    def __init__(self, *a):
        super().__init__(*a)
        self.uploads = 0

    def set_data(self, X, y, n_global=None):
        self.uploads += 1
        super().set_data(np.array(X, copy=True), np.array(y, copy=True), n_global)


def test_upload_cache_sees_in_place_updates_and_new_arrays():
    """The compiled functions evaluate the arrays they are passed (SCFGP.py:132-137): the device copy is
    reused only for the same, unchanged array objects."""
    np.random.seed(5)
    Xtr, ytr, _, _ = _data()
    m = SCFGP(sparsity=2, nfeats=6)
    m.set_data(Xtr, ytr)
    m._engine = eng = CountingEngine(m.D, m.S, m.M)
    m.build_models()
    c0 = m.train_func(m.X, m.y)[0]
    c1 = m.train_func(m.X, m.y)[0]
    assert eng.uploads == 1 and c0 == c1                      # unchanged arrays: one upload
    m.y[3, 0] += 0.5                                          # in-place edit of a training target
    c2 = m.train_func(m.X, m.y)[0]
    assert eng.uploads == 2 and c2 != c0
    X2, y2 = m.X.copy(), m.y.copy()                           # equal content, new objects
    m.train_func(X2, y2)
    assert eng.uploads == 3
    del X2, y2
    X3, y3 = m.X.copy() * 0.5, m.y.copy()                     # may recycle the ids of X2 / y2
    c3 = m.train_func(X3, y3)[0]
    assert eng.uploads == 4 and c3 != c2
    m.invalidate_data()
    m.train_func(X3, y3)
    assert eng.uploads == 5


@pytest.mark.parametrize("algo", ["adam", "adamax", "adadelta", "sgd"])
def test_checkpoint_resumes_the_update_rule(tmp_path, algo):
    """save()/load() carry the optimiser slots and step counter: 4 + 4 iterations across a checkpoint
    equal 8 uninterrupted iterations."""
    Xtr, ytr, _, _ = _data()
    opt = dict(algo={"algo": algo, "algo_params": {"learning_rate": 0.02}})

    def fresh():
        np.random.seed(6)
        m = SCFGP(sparsity=2, nfeats=6)
        m.set_data(Xtr, ytr)
        m._engine = OracleEngine(m.D, m.S, m.M)
        return m

    ref = fresh()
    ref.optimize(max_iter=8, **opt)
    a = fresh()
    a.optimize(max_iter=4, **opt)
    path = str(tmp_path / "ckpt.npz")
    a.save(path)
    b = SCFGP().load(path)
    b.set_data(Xtr, ytr)                                      # same table -> same scalers; params kept
    np.testing.assert_array_equal(b.params, a.params)
    b._engine = OracleEngine(b.D, b.S, b.M)
    b.optimize(max_iter=4, **opt)
    np.testing.assert_allclose(b.params, ref.params, rtol=0, atol=1e-12)
    np.testing.assert_allclose(b.evals["COST"][1][-1], ref.evals["COST"][1][-1], rtol=1e-12)


def test_save_before_training(tmp_path):
    np.random.seed(8)
    Xtr, ytr, _, _ = _data()
    m = SCFGP(sparsity=2, nfeats=6)
    m.set_data(Xtr, ytr)
    path = str(tmp_path / "early.npz")
    m.save(path)                                              # alpha / Li do not exist yet
    m2 = SCFGP().load(path)
    assert m2.alpha is None and m2.Li is None
    np.testing.assert_array_equal(m2.params, m.params)
