"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle.

Tolerances are north_star's: NLML (cost) <= 1e-6 relative; gradients, alpha/Li and
predictions <= 1e-5 (block-wise norm-relative for the gradient, absolute for the phase
block whose true value is 0).  In practice the FP64 DMMA path lands near 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import scfgp_oracle as O
from scfgp_b200 import Engine, SCFGP, ScfgpError, _lib, engine as E, synth, shard_bounds

pytestmark = pytest.mark.gpu

COST_RTOL = 1e-6
GRAD_RTOL = 1e-5
PRED_TOL = 1e-5


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def grad_blocks(D, S, M):
    P = 3 + D * S + M * S + S + M
    return {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S),
            "phase": slice(3 + D * S + M * S, P)}


def check_eval(eng, h, X, y, S, M, chunk=1024):
    D = X.shape[1]
    c, g, al, Li = eng.eval(h)
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M, chunk=chunk)
    assert abs(c - c0) <= COST_RTOL * abs(c0), (c, c0)
    for name, sl in grad_blocks(D, S, M).items():
        if name == "phase":
            assert np.abs(g[sl] - g0[sl]).max() <= GRAD_RTOL, name
        else:
            assert rel(g[sl], g0[sl]) <= GRAD_RTOL, (name, rel(g[sl], g0[sl]))
    assert np.abs(al - al0).max() <= PRED_TOL * max(1.0, np.abs(al0).max())
    assert np.abs(Li - Li0).max() <= PRED_TOL * max(1.0, np.abs(Li0).max())
    assert al.shape == (2 * (S + M), 1) and Li.shape == (2 * (S + M),) * 2
    assert np.abs(np.triu(Li, 1)).max() == 0.0           # Li = inv(L) is lower triangular
    return c, g, al, Li


# ------------------------------------------------------------------ kernel families
@pytest.mark.parametrize("mnk", [(128, 128, 16), (130, 70, 33), (300, 260, 200), (64, 512, 1000), (1, 2, 3)])
@pytest.mark.parametrize("layout", [0, 1, 2])
def test_dmma_gemm_layouts(layout, mnk):
    m, n, k = mnk
    rng = np.random.default_rng(m * 7 + n + k)
    if layout == 0:
        A, B = rng.standard_normal((m, k)), rng.standard_normal((n, k)); ref = A @ B.T
    elif layout == 1:
        A, B = rng.standard_normal((k, m)), rng.standard_normal((k, n)); ref = A.T @ B
    else:
        A, B = rng.standard_normal((m, k)), rng.standard_normal((k, n)); ref = A @ B
    np.testing.assert_allclose(E.dbg_gemm(layout, A, B), ref, rtol=0, atol=1e-12 * k)


@pytest.mark.parametrize("mnk", [(128, 128, 32), (128, 128, 2080), (130, 70, 34), (300, 260, 200), (256, 384, 1000),
                                 (2, 2, 2), (200, 160, 100), (140, 192, 70), (128, 33, 64), (64, 64, 64)])
@pytest.mark.parametrize("layout", [10, 11])
def test_tma_dmma_gemm(layout, mnk):
    """The TMA/mbarrier pipeline (128-byte swizzled 16x16 boxes, producer warpgroup): NT and TN."""
    m, n, k = mnk
    rng = np.random.default_rng(m * 7 + n + k)
    if layout == 10:
        A, B = rng.standard_normal((m, k)), rng.standard_normal((n, k)); ref = A @ B.T
    else:
        A, B = rng.standard_normal((k, m)), rng.standard_normal((k, n)); ref = A.T @ B
    np.testing.assert_allclose(E.dbg_gemm(layout, A, B), ref, rtol=0, atol=1e-12 * k)


@pytest.mark.parametrize("K", [4, 8, 64, 110, 140, 208, 600])
def test_blocked_cholesky_and_inverse(K):
    rng = np.random.default_rng(K)
    Q = rng.standard_normal((K + 50, K))
    A = Q.T @ Q + 0.1 * np.eye(K)
    L, Li, info = E.dbg_factor(A)
    assert info == 0
    L0 = np.linalg.cholesky(A)
    assert rel(np.tril(L), L0) < 1e-12
    assert rel(Li, np.linalg.inv(L0)) < 1e-11


def test_not_positive_definite_reports_pivot():
    A = np.eye(64); A[10, 10] = -1.0
    _, _, info = E.dbg_factor(A)
    assert info == 11                                    # pivot index + 1


@pytest.mark.parametrize("shape", [(400, 13, 20, 50), (300, 1, 100, 50), (1000, 64, 16, 120)])
def test_feature_kernel(shape):
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=2)
    eng = Engine(D, S, M)
    Phi = eng.features(h, X)
    np.testing.assert_allclose(Phi, synth.features_np(h, X, S, M), rtol=0, atol=1e-13)
    eng.close()


@pytest.mark.parametrize("scale", [3.0e4, 1.0e7, 1.0e12])
def test_feature_kernel_large_arguments(scale):
    """The feature map's sincos is straight-line code up to |z| = 1e6 and the library function beyond (per warp):
    arguments on both sides of the switch and far beyond it (Payne-Hanek range) against NumPy.  D = S = 1 makes
    z = x w + phi one product and one sum, i.e. bit-identical on both sides -- at |z| ~ 1e7 an ulp of z is already 2e-9,
    so any other geometry would compare summation orders of X.W, not the trigonometric functions."""
    N, D, S, M = 600, 1, 1, 29
    X, y, h = synth.make_problem(N, D, S, M, seed=4)
    Xb = X * scale
    Xb[::7] = X[::7]                                     # mixed warps: small and large arguments side by side
    eng = Engine(D, S, M)
    Phi = eng.features(h, Xb)
    ref = synth.features_np(h, Xb, S, M)
    assert np.isfinite(Phi).all()
    np.testing.assert_allclose(Phi, ref, rtol=0, atol=1e-15 * max(1.0, np.abs(ref).max()))
    eng.close()


def test_emulated_feature_kernel_large_arguments():
    """The fused feature+digit kernel of the emulated pass 1 takes the same switch: one evaluation on a table with
    arguments up to ~1e8 against the FP64 path on the same data (the two feature kernels must see the same Phi)."""
    N, D, S, M = 3000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=6)
    Xb = X.copy()
    Xb[::5] *= 1.0e8
    res = {}
    for prec in ("fp64", "emu5"):
        eng = Engine(D, S, M, precision=prec)
        eng.set_data(Xb, y)
        res[prec] = eng.eval(h)
        eng.close()
    a, b = res["emu5"], res["fp64"]
    assert abs(a[0] - b[0]) <= 1e-7 * abs(b[0])
    assert np.abs(a[2] - b[2]).max() <= 1e-6 * max(1.0, np.abs(b[2]).max())
    assert np.linalg.norm(a[1] - b[1]) <= 1e-5 * np.linalg.norm(b[1])


# ------------------------------------------------------------------ full evaluations (BASELINE configs)
@pytest.mark.parametrize("name,shape", [
    ("boston cfg1 (train split)", (400, 13, 5, 50)),
    ("boston shipped-pickle shape", (400, 13, 20, 50)),
    ("kin8nm cfg2", (5000, 8, 4, 100)),
    ("abalone cfg3", (3133, 10, 5, 50)),
    ("co2 cfg3", (2225, 1, 100, 50)),
    ("ragged: odd D, odd J, N not a tile multiple", (777, 7, 3, 30)),
    ("tiny", (5, 2, 2, 2)),
])
def test_eval_matches_oracle(name, shape):
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=N)
    eng = Engine(D, S, M)
    eng.set_data(X, y)
    c, g, al, Li = check_eval(eng, h, X, y, S, M)
    # forward-only call (train_func) returns the same cost / factors
    c2, g2, al2, Li2 = eng.eval(h, want_grad=False)
    assert g2 is None and c2 == c
    np.testing.assert_array_equal(al2, al)
    eng.close()


def test_eval_cfg4_geometry_reduced_rows():
    """cfg4's D, S, M (K = 2080) at a row count the oracle finishes in seconds."""
    N, D, S, M = 6000, 64, 16, 1024
    X, y, h = synth.make_problem(N, D, S, M, seed=4)
    eng = Engine(D, S, M)
    eng.set_data(X, y)
    check_eval(eng, h, X, y, S, M, chunk=2048)
    eng.close()


def test_literal_reference_graph_agrees():
    """Against the literal restatement of the Theano graph (Hermite quadrature + autograd)."""
    N, D, S, M = 400, 13, 20, 50
    X, y, h = synth.make_problem(N, D, S, M, seed=9)
    eng = Engine(D, S, M)
    eng.set_data(X, y)
    c, g, al, Li = eng.eval(h)
    c0, g0, al0, Li0 = O.literal_cost_grad(h, X, y, S, M, hermite=True)
    assert abs(c - c0) <= COST_RTOL * abs(c0)
    assert rel(g[:-(S + M)], g0[:-(S + M)]) <= GRAD_RTOL
    assert np.abs(al - al0).max() <= PRED_TOL
    eng.close()


@pytest.mark.parametrize("chunk", [128, 256, 1024])
def test_chunk_size_does_not_change_results(chunk):
    N, D, S, M = 1500, 8, 4, 60
    X, y, h = synth.make_problem(N, D, S, M, seed=6)
    eng = Engine(D, S, M, chunk_rows=chunk)
    eng.set_data(X, y)
    check_eval(eng, h, X, y, S, M)
    eng.close()


def test_repeatable_bitwise():
    N, D, S, M = 3000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=8)
    eng = Engine(D, S, M)
    eng.set_data(X, y)
    r1 = eng.eval(h)
    r2 = eng.eval(h)
    assert r1[0] == r2[0]
    np.testing.assert_array_equal(r1[1], r2[1])
    np.testing.assert_array_equal(r1[3], r2[3])
    eng.close()


@pytest.mark.parametrize("precision", ["fp64", "emu5"])
def test_nophase_fftype(precision):
    """fftype='non-phase' (Z = X.W): same cost as the oracle evaluated with the phase
    parameters chosen so that the centred phase vanishes (both arithmetic modes)."""
    N, D, S, M = 600, 6, 3, 40
    X, y, h = synth.make_problem(N, D, S, M, seed=10)
    a, b, c, l_F, r_F, F, l_FC, FC = O.unpack_params(h, D, S, M)
    h0 = h.copy()
    h0[-(S + M):-M] = l_F.mean(0)           # l_p = mean(l_F)  -> l_FC = 0
    h0[-M:] = F.mean(0)                     # p = mean(F)      -> FC = 0
    eng = Engine(D, S, M, fftype=_lib.FF_NOPHASE, precision=precision)
    eng.set_data(X, y)
    c1, g1, al1, Li1 = eng.eval(h)
    c0, g0, al0, Li0 = O.streaming_eval(h0, X, y, S, M)
    assert abs(c1 - c0) <= COST_RTOL * abs(c0)
    assert np.abs(al1 - al0).max() <= PRED_TOL
    assert np.all(g1[-(S + M):] == 0)
    # and, by phase invariance, the same cost as the phase variant at h
    eng2 = Engine(D, S, M, fftype=_lib.FF_PHASE, precision=precision)
    eng2.set_data(X, y)
    assert abs(eng2.eval(h, want_grad=False)[0] - c1) <= (1e-9 if precision == "fp64" else COST_RTOL) * abs(c1)
    eng.close(); eng2.close()


# ------------------------------------------------------------------ prediction
@pytest.mark.parametrize("shape,T", [((400, 13, 20, 50), 106), ((2000, 8, 4, 100), 3192), ((300, 3, 2, 10), 1)])
def test_predict_matches_oracle(shape, T):
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=12)
    Xs = np.random.default_rng(1).random((T, D))
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M)
    eng = Engine(D, S, M)
    mu, sd = eng.predict(h, Xs, al0, Li0)                     # pred_func(Xs, alpha, Li)
    mu0, sd0 = O.literal_predict(h, Xs, al0, Li0, S, M)
    assert mu.shape == (T, 1) and sd.shape == (T,)
    np.testing.assert_allclose(mu, mu0, rtol=0, atol=PRED_TOL)
    np.testing.assert_allclose(sd, sd0, rtol=0, atol=PRED_TOL)
    # with the factors left on the device by an evaluation
    eng.set_data(X, y)
    eng.eval(h, want_grad=False)
    mu2, sd2 = eng.predict(h, Xs)
    np.testing.assert_allclose(mu2, mu0, rtol=0, atol=PRED_TOL)
    np.testing.assert_allclose(sd2, sd0, rtol=0, atol=PRED_TOL)
    eng.close()


def test_shipped_boston_model_predicts(golden):
    """The reference's trained Boston model (params, alpha, Li from its pickle) through
    pred_func: kernel vs literal graph, and the reference's own scalers around it."""
    D, S, M = int(golden["D"]), int(golden["S"]), int(golden["M"])
    h, al, Li = golden["params"], golden["alpha"], golden["Li"]
    rng = np.random.default_rng(0)
    Xs = rng.random((106, D))
    eng = Engine(D, S, M)
    mu, sd = eng.predict(h, Xs, al, Li)
    mu0, sd0 = O.literal_predict(h, Xs, al, Li, S, M)
    np.testing.assert_allclose(mu, mu0, rtol=0, atol=PRED_TOL)
    np.testing.assert_allclose(sd, sd0, rtol=0, atol=PRED_TOL)
    assert np.all(sd > np.sqrt(np.log(1 + np.exp(h[2]))) - 1e-12)    # std^2 >= noise
    eng.close()


# ------------------------------------------------------------------ size-independent properties
def test_shard_additivity():
    """Row-sharded partial sums add up: evaluating two shards through the phase API and
    summing the reduce buffers on the host equals the single evaluation."""
    N, D, S, M = 4100, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=14)
    lib = _lib.load()
    full = Engine(D, S, M); full.set_data(X, y)
    c_full, g_full, al_full, Li_full = full.eval(h)
    import torch
    engs = []
    for r in range(2):
        lo, hi = shard_bounds(N, r, 2)
        e = Engine(D, S, M); e.set_data(X[lo:hi], y[lo:hi], n_global=N)
        engs.append(e)
    hp = _lib.dptr(_lib.as_f64(h))
    for e in engs:
        e._check(lib.scfgp_eval_begin(e._h, hp, 1))

    def allreduce(which):
        bufs = [e._reduce_tensor(which) for e in engs]
        tot = bufs[0] + bufs[1]
        for b in bufs:
            b.copy_(tot)
        torch.cuda.synchronize()
    for e in engs: e._check(lib.scfgp_eval_pass1(e._h))
    allreduce(1)
    for e in engs: e._check(lib.scfgp_eval_factor(e._h))
    for e in engs: e._check(lib.scfgp_eval_pass2(e._h))
    allreduce(2)
    for e in engs: e._check(lib.scfgp_eval_backward(e._h))
    for e in engs: e._check(lib.scfgp_eval_pass3(e._h))
    allreduce(3)
    outs = []
    for e in engs:
        cost = np.zeros(1); g = np.zeros(e.P); al = np.zeros(e.K); Li = np.zeros((e.K, e.K))
        e._check(lib.scfgp_eval_finish(e._h, _lib.dptr(cost), _lib.dptr(g), _lib.dptr(al), _lib.dptr(Li)))
        outs.append((cost[0], g, al, Li))
    assert outs[0][0] == outs[1][0]                       # replicated phases: identical on both "ranks"
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    assert abs(outs[0][0] - c_full) <= 1e-11 * abs(c_full)
    assert rel(outs[0][1], g_full) <= 1e-8
    for e in engs + [full]:
        e.close()


def test_row_scaling_linearity():
    """Duplicating every row doubles G, p, yy: cost*N of the data terms follows; checked via
    the oracle-free identity G[j,j] + G[j+J,j+J] = N s^2 on the device Gram."""
    import torch
    N, D, S, M = 1000, 5, 3, 33
    X, y, h = synth.make_problem(N, D, S, M, seed=15)
    lib = _lib.load()
    e = Engine(D, S, M); e.set_data(X, y)
    e._check(lib.scfgp_eval_begin(e._h, _lib.dptr(_lib.as_f64(h)), 0))
    e._check(lib.scfgp_eval_pass1(e._h))
    torch.cuda.synchronize()
    K, J = e.K, e.J
    red = e._reduce_tensor(1).cpu().numpy()
    G = red[:K * K].reshape(K, K)
    pair = np.diag(G)[:J] + np.diag(G)[J:]
    np.testing.assert_allclose(pair, N * np.exp(2 * h[1]) * 2.0 / M, rtol=1e-12)
    assert abs(red[-1] - y @ y) <= 1e-10 * (y @ y)
    e.close()


# ------------------------------------------------------------------ errors
def test_errors():
    eng = Engine(4, 2, 8)
    with pytest.raises(ValueError):
        eng.eval(np.zeros(eng.P))                                 # no data
    X = np.random.default_rng(0).random((50, 4)); y = np.zeros(50)
    with pytest.raises(ValueError):
        eng.set_data(X[:, :3], y)                                 # wrong D
    eng.set_data(X, y)
    with pytest.raises(ValueError):
        eng.eval(np.zeros(3))                                     # wrong P
    h = synth.draw_params(4, 2, 8, np.random.default_rng(1))
    h[0] = -400.0; h[1] = 200.0                                   # sig2_n -> 0, sig_f huge: A not PD / not finite
    with pytest.raises((np.linalg.LinAlgError, FloatingPointError)):
        eng.eval(h)
    eng.close()


# ------------------------------------------------------------------ facade end to end (cfg1: Boston-shaped fit+predict)
def test_fit_predict_boston_shaped():
    rng = np.random.default_rng(3)
    N, D = 506, 13
    X = rng.lognormal(size=(N, D))
    w = rng.standard_normal(D)
    yraw = np.tanh(X @ w / 5) * 10 + 22 + rng.standard_normal(N)
    tr = rng.choice(N, 400, replace=False); te = np.setdiff1d(np.arange(N), tr)
    np.random.seed(7)                       # init_params draws from the global NumPy state (SCFGP.py:65-71)
    model = SCFGP(rank=5, feature_size=50, fftype="phase", msg=False)
    model.fit(X[tr], yraw[tr, None], X[te], yraw[te, None], max_iter=60)
    costs = model.evals["COST"][1]
    assert costs[-1] < costs[0]
    mu, sd = model.predict(X[te], yraw[te, None])
    assert mu.shape == (106, 1) and sd.shape == (106, 1) and np.all(sd > 0)
    assert model.TsNMSE < 0.9, model.TsNMSE
    f = model.get_compiled_funcs()
    assert len(f) == 3 and f.engine is not None
    # second model reuses the compiled handle (optimize(funcs=...), SCFGP.py:209-210)
    m2 = SCFGP(sparsity=5, nfeats=50)
    m2.set_data(X[tr], yraw[tr, None])
    m2.optimize(X[te], yraw[te, None], f, None, max_iter=12)
    assert len(m2.evals["COST"][1]) == 13


def test_save_load_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    X = rng.random((200, 4)); y = np.sin(X.sum(1))[:, None] + 0.05 * rng.standard_normal((200, 1))
    m = SCFGP(sparsity=2, nfeats=10)
    m.set_data(X, y); m.optimize(max_iter=12)
    mu, sd = m.predict(X[:20])
    path = str(tmp_path / "model.npz")
    m.save(path)
    m2 = SCFGP().load(path)
    mu2, sd2 = m2.predict(X[:20])
    np.testing.assert_allclose(mu2, mu, rtol=1e-12)
    np.testing.assert_allclose(sd2, sd, rtol=1e-12)


class _OracleEngine:
    """Oracle-backed stand-in for the CUDA engine (checker only): drives the same facade loop on the CPU."""

    def __init__(self, D, S, M):
        self.D, self.S, self.M = D, S, M

    def set_data(self, X, y, n_global=None):
        self.X, self.y = X, np.asarray(y).reshape(-1)

    def eval(self, h, want_grad=True, **_):
        return O.streaming_eval(h, self.X, self.y, self.S, self.M, want_grad=want_grad)

    def predict(self, h, Xs, a, L):
        return O.literal_predict(h, Xs, a, L, self.S, self.M)


@pytest.mark.parametrize("precision,ptol,ctol", [("fp64", 1e-8, 1e-10), ("emu5", None, 1e-6)])
def test_optimizer_trajectory_matches_oracle_driven_fit(precision, ptol, ctol):
    """Row f1 of SURVEY section 8: the whole training loop (SCFGP.py:184-276; Adam + the Nesterov wrapper,
    Optimizer.py) driven by the CUDA evaluation follows the loop driven by the oracle: same costs per
    iteration, same final parameters (FP64 path) and the same predictions."""
    rng = np.random.default_rng(11)
    X = rng.lognormal(size=(300, 6))
    y = (np.tanh(X @ rng.standard_normal(6) / 4) * 5 + 20 + 0.3 * rng.standard_normal(300))[:, None]
    Xtr, ytr, Xte, yte = X[:240], y[:240], X[240:], y[240:]
    runs = []
    for which in ("oracle", "cuda"):
        np.random.seed(21)
        m = SCFGP(rank=3, feature_size=20, fftype="phase", msg=False,
                  **({"precision": precision} if which == "cuda" else {}))
        m.set_data(Xtr, ytr)
        if which == "oracle":
            m._engine = _OracleEngine(m.D, m.S, m.M)
        m.optimize(Xte, yte, None, None, max_iter=20, restore_best=False)
        mu, sd = m.predict(Xte)
        runs.append((np.array(m.evals["COST"][1]), m.params.copy(), mu, sd))
    (c0, p0, mu0, sd0), (c1, p1, mu1, sd1) = runs
    assert len(c0) == len(c1) == 21
    np.testing.assert_allclose(c1, c0, rtol=ctol)
    if ptol is not None:
        np.testing.assert_allclose(p1, p0, atol=ptol)
    scale = float(np.std(yte))
    np.testing.assert_allclose(mu1, mu0, atol=1e-5 * scale if ptol else 1e-3 * scale)
    np.testing.assert_allclose(sd1, sd0, atol=1e-5 * scale if ptol else 1e-3 * scale)


# ------------------------------------------------------------------ minibatch mode on the device (row f5)
@pytest.mark.parametrize("precision", ["fp64", "emu5"])
def test_select_rows_matches_oracle_on_the_batch(precision):
    """scfgp_select_rows: per-batch cost / gradient / alpha / Li equal the oracle on (X[batch], y[batch])
    (SCFGP.py:226-231: train_iter_func(X[batch], y[batch])); the resident set is not uploaded again."""
    N, D, S, M = 3000, 8, 4, 40
    X, y, h = synth.make_problem(N, D, S, M, seed=21)
    eng = Engine(D, S, M, precision=precision)
    eng.set_data(X, y)
    rng = np.random.default_rng(5)
    for nb in (150, 1, 1000):
        idx = rng.permutation(N)[:nb]
        eng.select_rows(idx)
        c, g, al, Li = eng.eval(h)
        c0, g0, al0, Li0 = O.streaming_eval(h, X[idx], y[idx], S, M)
        assert abs(c - c0) <= COST_RTOL * abs(c0)
        assert rel(g[:3 + D * S + M * S], g0[:3 + D * S + M * S]) <= GRAD_RTOL
        assert np.abs(al - al0).max() <= PRED_TOL * max(1.0, np.abs(al0).max())
        assert np.abs(Li - Li0).max() <= PRED_TOL * max(1.0, np.abs(Li0).max())
    eng.select_rows(None)                                   # back to the whole set
    c, _, _, _ = eng.eval(h)
    c0 = O.streaming_eval(h, X, y, S, M, want_grad=False)[0]
    assert abs(c - c0) <= COST_RTOL * abs(c0)
    with pytest.raises(ValueError):
        eng.select_rows(np.array([N]))                      # out of range
    eng.close()


def test_empty_local_batch_contributes_nothing():
    """A rank whose shard holds no row of the current batch (n = 0): its three partial reductions are zero, so
    the sum over shards is the other shard's -- checked through the phase API like test_shard_additivity."""
    N, D, S, M = 600, 5, 3, 20
    X, y, h = synth.make_problem(N, D, S, M, seed=23)
    eng = Engine(D, S, M)
    eng.set_data(X, y)
    eng.select_rows(np.zeros(0, dtype=np.int64), n_global=100)
    lib, hd = eng.lib, eng._h
    eng._check(lib.scfgp_eval_begin(hd, _lib.dptr(_lib.as_f64(h)), 1))
    eng._check(lib.scfgp_eval_pass1(hd))
    red1 = eng._reduce_tensor(1).cpu().numpy()
    assert np.all(red1 == 0.0)
    eng.close()


def test_facade_minibatch_with_device_resident_training_set():
    """nbatches > 1 with CUDA tensors (scaler_device='gpu'): batches are gathered on the device."""
    import torch
    rng = np.random.default_rng(14)
    N, D = 2000, 6
    X = rng.lognormal(size=(N, D))
    yraw = np.sin(X.sum(1) / 3) + 0.1 * rng.standard_normal(N)
    np.random.seed(3)
    m = SCFGP(rank=3, feature_size=20, scaler_device="gpu")
    m.set_data(torch.as_tensor(X, device="cuda"), torch.as_tensor(yraw, device="cuda"))
    m.optimize(max_iter=12, nbatches=3, batchsize=200)
    costs = m.evals["COST"][1]
    assert len(costs) == 13 and np.all(np.isfinite(costs)) and costs[-2] < costs[0]
    assert m._engine.N == N                                  # the final train_func ran on the whole set
