"""Parity of the emulated-FP64 path (int8 tensor cores, tcgen05.mma kind::i8 through the C
ABI) against the CPU oracle and against the FP64 DMMA path.

north_star's tolerances are cost <= 1e-6 relative, gradient / alpha / Li / predictions
<= 1e-5.  The default of the emulated mode (5 base-256 digits = 38 fractional bits of every
operand row's power-of-two scale) is held to bounds 10x tighter than that here; 6 digits
to 1000x tighter."""
import numpy as np
import pytest

from oracle import scfgp_oracle as O
from scfgp_b200 import Engine, SCFGP, _lib, engine as E, synth, shard_bounds

pytestmark = pytest.mark.gpu

TOL = {  # precision -> (cost rtol, gradient block rtol / phase atol, alpha & Li & prediction atol)
    "emu5": (1e-7, 1e-6, 1e-6),
    "emu6": (1e-9, 1e-8, 1e-8),
}


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def grad_blocks(D, S, M):
    P = 3 + D * S + M * S + S + M
    return {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S),
            "phase": slice(3 + D * S + M * S, P)}


def check_eval(eng, prec, h, X, y, S, M, chunk=1024):
    ctol, gtol, atol = TOL[prec]
    D = X.shape[1]
    c, g, al, Li = eng.eval(h)
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M, chunk=chunk)
    assert abs(c - c0) <= ctol * abs(c0), (c, c0)
    for name, sl in grad_blocks(D, S, M).items():
        if name == "phase":
            assert np.abs(g[sl] - g0[sl]).max() <= gtol, name
        else:
            assert rel(g[sl], g0[sl]) <= gtol, (name, rel(g[sl], g0[sl]))
    assert np.abs(al - al0).max() <= atol * max(1.0, np.abs(al0).max())
    assert np.abs(Li - Li0).max() <= atol * max(1.0, np.abs(Li0).max())
    assert np.abs(np.triu(Li, 1)).max() == 0.0
    return c, g, al, Li


# ------------------------------------------------------------------ digit GEMM (extraction + tcgen05 mainloop)
@pytest.mark.parametrize("digits", [4, 5, 6, 7, 8])
@pytest.mark.parametrize("mnk", [(128, 64, 64), (130, 70, 33), (300, 260, 200), (257, 129, 1000), (1, 1, 1),
                                 (64, 512, 4160)])
def test_digit_gemm_error_bound(digits, mnk):
    """C = A.BT with per-row power-of-two scales.  Each operand entry is rounded to
    2^-(8 digits) of its row scale (scale < 4 rowmax), i.e. an absolute error below
    2^-(8 digits - 2) rowmax; both operands and the dropped digit products (i + j >= digits)
    together stay below 4 * 2^-(8 digits - 2) * k * rowmax(A) * rowmax(B)."""
    m, n, k = mnk
    rng = np.random.default_rng(m + 3 * n + 7 * k + digits)
    A = rng.standard_normal((m, k)) * np.exp(3 * rng.standard_normal((m, 1)))   # rows of very different size
    B = rng.standard_normal((n, k)) * np.exp(3 * rng.standard_normal((n, 1)))
    Cm = E.dbg_emu_gemm(A, B, digits)
    bound = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * k
    err = np.abs(Cm - A @ B.T) / bound
    assert err.max() <= max(4 * 2.0 ** (-(8 * digits - 2)), 4e-16)


def test_digit_gemm_exact_on_integers():
    """Small integers are representable exactly in 4 digits: the product must be exact."""
    rng = np.random.default_rng(0)
    A = rng.integers(-100, 100, (200, 300)).astype(float)
    B = rng.integers(-100, 100, (100, 300)).astype(float)
    np.testing.assert_array_equal(E.dbg_emu_gemm(A, B, 4), A @ B.T)


def test_digit_gemm_zero_rows_and_columns():
    A = np.zeros((130, 70)); B = np.ones((65, 70))
    A[3] = 2.5
    Cm = E.dbg_emu_gemm(A, B, 5)
    ref = A @ B.T
    np.testing.assert_array_equal(Cm, ref)


# ------------------------------------------------------------------ full evaluations (BASELINE configs)
@pytest.mark.parametrize("prec", ["emu5", "emu6"])
@pytest.mark.parametrize("name,shape", [
    ("boston cfg1 (train split)", (400, 13, 5, 50)),
    ("boston shipped-pickle shape", (400, 13, 20, 50)),
    ("kin8nm cfg2", (5000, 8, 4, 100)),
    ("abalone cfg3", (3133, 10, 5, 50)),
    ("co2 cfg3", (2225, 1, 100, 50)),
    ("ragged: odd D, odd J, N not a tile multiple", (777, 7, 3, 30)),
    ("more than one 16384-row chunk", (17000, 8, 4, 100)),
    ("tiny", (5, 2, 2, 2)),
])
def test_eval_matches_oracle(name, shape, prec):
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=N)
    eng = Engine(D, S, M, precision=prec)
    assert eng.precision == prec
    eng.set_data(X, y)
    c, g, al, Li = check_eval(eng, prec, h, X, y, S, M)
    c2, g2, al2, Li2 = eng.eval(h, want_grad=False)          # train_func: same cost / factors
    assert g2 is None and c2 == c
    np.testing.assert_array_equal(al2, al)
    eng.close()


def test_eval_cfg4_geometry_reduced_rows():
    N, D, S, M = 6000, 64, 16, 1024
    X, y, h = synth.make_problem(N, D, S, M, seed=4)
    eng = Engine(D, S, M, precision="emu5")
    eng.set_data(X, y)
    check_eval(eng, "emu5", h, X, y, S, M, chunk=2048)
    eng.close()


def test_modes_agree_on_one_engine():
    """Switching the same handle between the FP64 tensor path and the emulated one."""
    N, D, S, M = 3000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=21)
    eng = Engine(D, S, M)
    assert eng.precision == "fp64"
    eng.set_data(X, y)
    c0, g0, al0, Li0 = eng.eval(h)
    eng.set_precision("emu")
    assert eng.precision == "emu5"
    c1, g1, al1, Li1 = eng.eval(h)
    eng.set_precision("fp64")
    c2, g2, al2, Li2 = eng.eval(h)
    assert c2 == c0
    np.testing.assert_array_equal(g2, g0)
    assert abs(c1 - c0) <= 1e-8 * abs(c0)
    assert rel(g1, g0) <= 1e-6
    with pytest.raises(ValueError):
        eng.set_precision("emu3")
    with pytest.raises(ValueError):
        eng.set_precision("bf16")
    eng.close()


@pytest.mark.parametrize("chunk", [128, 384, 1024])
def test_chunk_size_does_not_change_results(chunk):
    N, D, S, M = 1500, 8, 4, 60
    X, y, h = synth.make_problem(N, D, S, M, seed=6)
    eng = Engine(D, S, M, chunk_rows=chunk, precision="emu5")
    eng.set_data(X, y)
    check_eval(eng, "emu5", h, X, y, S, M)
    eng.close()


def test_repeatable_bitwise():
    """Integer accumulation + fixed-order float64 sums: evaluations repeat bit for bit."""
    N, D, S, M = 20000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=8)
    eng = Engine(D, S, M, precision="emu5")
    eng.set_data(X, y)
    r1 = eng.eval(h)
    r2 = eng.eval(h)
    assert r1[0] == r2[0]
    np.testing.assert_array_equal(r1[1], r2[1])
    np.testing.assert_array_equal(r1[3], r2[3])
    eng.close()


# ------------------------------------------------------------------ prediction
@pytest.mark.parametrize("shape,T", [((400, 13, 20, 50), 106), ((2000, 8, 4, 100), 3192), ((300, 3, 2, 10), 1)])
def test_predict_matches_oracle(shape, T):
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=12)
    Xs = np.random.default_rng(1).random((T, D))
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M)
    eng = Engine(D, S, M, precision="emu5")
    mu, sd = eng.predict(h, Xs, al0, Li0)                     # pred_func(Xs, alpha, Li)
    mu0, sd0 = O.literal_predict(h, Xs, al0, Li0, S, M)
    assert mu.shape == (T, 1) and sd.shape == (T,)
    np.testing.assert_allclose(mu, mu0, rtol=0, atol=1e-6)
    np.testing.assert_allclose(sd, sd0, rtol=0, atol=1e-6)
    eng.set_data(X, y)
    eng.eval(h, want_grad=False)
    mu2, sd2 = eng.predict(h, Xs)
    np.testing.assert_allclose(mu2, mu0, rtol=0, atol=1e-6)
    np.testing.assert_allclose(sd2, sd0, rtol=0, atol=1e-6)
    eng.close()


def test_shipped_boston_model_predicts(golden):
    D, S, M = int(golden["D"]), int(golden["S"]), int(golden["M"])
    h, al, Li = golden["params"], golden["alpha"], golden["Li"]
    Xs = np.random.default_rng(0).random((106, D))
    eng = Engine(D, S, M, precision="emu5")
    mu, sd = eng.predict(h, Xs, al, Li)
    mu0, sd0 = O.literal_predict(h, Xs, al, Li, S, M)
    np.testing.assert_allclose(mu, mu0, rtol=0, atol=1e-6)
    np.testing.assert_allclose(sd, sd0, rtol=0, atol=1e-6)
    eng.close()


# ------------------------------------------------------------------ sharding
def test_shard_additivity():
    """Two row shards through the phase API, reduce buffers summed on the host (what the
    NCCL all-reduce does), against the single evaluation."""
    import torch
    N, D, S, M = 4100, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=14)
    lib = _lib.load()
    full = Engine(D, S, M, precision="emu5"); full.set_data(X, y)
    c_full, g_full, al_full, Li_full = full.eval(h)
    engs = []
    for r in range(2):
        lo, hi = shard_bounds(N, r, 2)
        e = Engine(D, S, M, precision="emu5"); e.set_data(X[lo:hi], y[lo:hi], n_global=N)
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
    # the reduce buffers are identical after the all-reduce, so every rank finishes with the
    # same bits; against the unsharded evaluation only the digit rounding differs (each rank
    # scales its Psi digits by its own bound of |gv|)
    assert outs[0][0] == outs[1][0]
    np.testing.assert_array_equal(outs[0][1], outs[1][1])
    np.testing.assert_array_equal(outs[0][3], outs[1][3])
    assert abs(outs[0][0] - c_full) <= 1e-9 * abs(c_full)
    assert rel(outs[0][1], g_full) <= 1e-7
    for e in engs + [full]:
        e.close()


# ------------------------------------------------------------------ facade
def test_facade_fit_in_emulated_mode():
    rng = np.random.default_rng(3)
    N, D = 506, 13
    X = rng.lognormal(size=(N, D))
    yraw = np.tanh(X @ rng.standard_normal(D) / 5) * 10 + 22 + rng.standard_normal(N)
    tr = rng.choice(N, 400, replace=False); te = np.setdiff1d(np.arange(N), tr)
    np.random.seed(7)
    model = SCFGP(rank=5, feature_size=50, fftype="phase", msg=False, precision="emu")
    model.fit(X[tr], yraw[tr, None], X[te], yraw[te, None], max_iter=40)
    costs = model.evals["COST"][1]
    assert costs[-1] < costs[0]
    assert model.get_compiled_funcs().engine.precision == "emu5"
    mu, sd = model.predict(X[te], yraw[te, None])
    assert np.all(np.isfinite(mu)) and np.all(sd > 0)


def test_cfg5_geometry_reduced_rows_both_modes():
    """cfg5's D = 256, rank 32, feature_size 2048 (K = 4160, P = 75 811) at a row count the
    oracle finishes in seconds: FP64 DMMA path and emulated path against the oracle."""
    N, D, S, M = 2500, 256, 32, 2048
    X, y, h = synth.make_problem(N, D, S, M, seed=5)
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M, chunk=1250)
    for prec, (ctol, gtol, atol) in (("fp64", (1e-10, 1e-8, 1e-8)), ("emu5", TOL["emu5"])):
        eng = Engine(D, S, M, precision=prec)
        eng.set_data(X, y)
        c, g, al, Li = eng.eval(h)
        assert abs(c - c0) <= ctol * abs(c0), (prec, c, c0)
        for name, sl in grad_blocks(D, S, M).items():
            if name == "phase":
                assert np.abs(g[sl] - g0[sl]).max() <= gtol, (prec, name)
            else:
                assert rel(g[sl], g0[sl]) <= gtol, (prec, name, rel(g[sl], g0[sl]))
        assert np.abs(al - al0).max() <= atol * max(1.0, np.abs(al0).max()), prec
        assert np.abs(Li - Li0).max() <= atol * max(1.0, np.abs(Li0).max()), prec
        eng.close()


def test_errors_surface_in_emulated_mode():
    """The digit extraction would turn NaN / Inf into finite integers; the float64 side paths
    (Phi^T y, Phi alpha) still carry them into the cost, and the factorisation reports a
    non-positive-definite A exactly as in FP64 mode."""
    eng = Engine(4, 2, 8, precision="emu5")
    X = np.random.default_rng(0).random((300, 4)); y = np.random.default_rng(1).standard_normal(300)
    h = synth.draw_params(4, 2, 8, np.random.default_rng(1))
    eng.set_data(X, y)
    c, g, _, _ = eng.eval(h)
    assert np.isfinite(c) and np.all(np.isfinite(g))
    Xbad = X.copy(); Xbad[17, 2] = np.nan
    eng.set_data(Xbad, y)
    with pytest.raises((FloatingPointError, np.linalg.LinAlgError)):
        eng.eval(h)
    Xbad[17, 2] = np.inf
    eng.set_data(Xbad, y)
    with pytest.raises((FloatingPointError, np.linalg.LinAlgError)):
        eng.eval(h)
    eng.set_data(X, y)
    hb = h.copy(); hb[0] = -400.0; hb[1] = 200.0          # sig2_n -> 0, sig_f huge
    with pytest.raises((np.linalg.LinAlgError, FloatingPointError)):
        eng.eval(hb)
    c2, _, _, _ = eng.eval(h)                              # the handle stays usable
    assert c2 == c
    eng.close()


def test_async_upload_matches_synchronous():
    """scfgp_set_data_host_async: row blocks land while pass 1 runs; same bits as the blocking copy."""
    N, D, S, M = 70000, 8, 4, 60
    X, y, h = synth.make_problem(N, D, S, M, seed=31)
    eng = Engine(D, S, M, precision="emu5", chunk_rows=1024)
    eng.set_data(X, y)
    r1 = eng.eval(h)
    for _ in range(2):
        eng.set_data(X, y, overlap=True)
        r2 = eng.eval(h)
        assert r1[0] == r2[0]
        np.testing.assert_array_equal(r1[1], r2[1])
    X2 = X[::-1].copy(); y2 = y[::-1].copy()               # new data through the same buffers
    eng.set_data(X2, y2, overlap=True)
    r3 = eng.eval(h)
    assert abs(r3[0] - r1[0]) <= 1e-9 * abs(r1[0])         # a row permutation only reorders the sums
    eng.close()


def test_depth_limit_is_enforced():
    """Reduction depths that could overflow an int32 class sum are refused, not silently wrong."""
    eng = Engine(4, 2, 8, precision="emu5")
    with pytest.raises(ValueError):
        eng._check(eng.lib.scfgp_set_chunk_rows(eng._h, 32768))       # limit at 5 digits: 32767
    eng._check(eng.lib.scfgp_set_chunk_rows(eng._h, 16384))
    with pytest.raises(ValueError):
        eng.set_precision("emu8")                                     # 16384 rows < 18724: fine ...
        eng._check(eng.lib.scfgp_set_chunk_rows(eng._h, 18816))       # ... but not 18816
    eng.close()


@pytest.mark.parametrize("digits", [4, 5, 6])
@pytest.mark.parametrize("mnk", [(130, 70, 33), (257, 129, 500), (128, 64, 2080)])
def test_digit_gemm_bit_exact_against_cpu_restatement(digits, mnk):
    """Scales, digit extraction, int32 class sums (tcgen05.mma kind::i8) and the float64
    recombination reproduce oracle/emu_digits.py (NumPy int64) bit for bit."""
    from oracle import emu_digits as ED
    m, n, k = mnk
    rng = np.random.default_rng(11 * m + n + k + digits)
    A = rng.standard_normal((m, k)) * np.exp(2 * rng.standard_normal((m, 1)))
    B = rng.standard_normal((n, k)) * np.exp(2 * rng.standard_normal((n, 1)))
    A[5] = 0.0                                              # an all-zero row (scale 2)
    np.testing.assert_array_equal(E.dbg_emu_gemm(A, B, digits), ED.emu_gemm(A, B, digits))


def test_digit_cache_on_and_off_agree(monkeypatch):
    """With the cache, passes 2 and 3 read Phi back from its digits instead of recomputing the
    feature map: same results up to the digit rounding of Phi in the by-products."""
    N, D, S, M = 20000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=41)
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SCFGP_PHI_CACHE", flag)
        eng = Engine(D, S, M, precision="emu5")
        eng.set_data(X, y)
        res[flag] = eng.eval(h)
        r2 = eng.eval(h)                                   # and each mode repeats bit for bit
        assert r2[0] == res[flag][0]
        np.testing.assert_array_equal(r2[1], res[flag][1])
        eng.close()
    a, b = res["1"], res["0"]
    np.testing.assert_array_equal(a[3], b[3])              # pass 1 is identical: same Li, alpha
    assert abs(a[0] - b[0]) <= 1e-9 * abs(b[0])
    assert rel(a[1], b[1]) <= 1e-7


@pytest.mark.parametrize("shape", [(3000, 13, 5, 50), (2500, 8, 4, 100), (1500, 64, 16, 120), (900, 1, 100, 50)])
def test_tile_widths_agree(monkeypatch, shape):
    """The two K-major products run on 128 x 96 tiles by default (at most 5 digits) and on 128 x 64 tiles with
    SCFGP_TILE_N=64: the same exact integer sums; only float64 partial sums over tile halves are grouped differently."""
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=N + 1)
    Xs = np.random.default_rng(3).random((200, D))
    res = {}
    for width in ("96", "64"):
        monkeypatch.setenv("SCFGP_TILE_N", width)
        eng = Engine(D, S, M, precision="emu5")
        eng.set_data(X, y)
        c, g, al, Li = eng.eval(h)
        mu, sd = eng.predict(h, Xs)
        res[width] = (c, g, al, Li, mu, sd)
        eng.close()
    a, b = res["96"], res["64"]
    np.testing.assert_array_equal(a[2], b[2])              # pass 1 does not depend on the tile width: same alpha, Li
    np.testing.assert_array_equal(a[3], b[3])
    np.testing.assert_array_equal(a[4], b[4])              # mu = Phi.alpha: one exact digit sum per row
    assert abs(a[0] - b[0]) <= 1e-14 * abs(b[0])           # row sums of T^2 are grouped by tile half
    np.testing.assert_allclose(a[5], b[5], rtol=1e-13)
    assert rel(a[1], b[1]) <= 1e-12
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M)
    assert abs(a[0] - c0) <= 1e-7 * abs(c0) and rel(a[1], g0) <= 1e-6


@pytest.mark.parametrize("shape", [(2500, 8, 4, 124), (20000, 8, 4, 100), (1500, 64, 16, 1024), (700, 5, 3, 33)])
@pytest.mark.parametrize("knob,alt", [("SCFGP_FUSED_FEATURES", "0"), ("SCFGP_EPILOGUE", "1"), ("SCFGP_DIGIT_KERNEL", "generic")])
def test_kernel_variants_agree(monkeypatch, shape, knob, alt):
    """Every fast kernel of the 5-digit path has a plainer sibling behind an environment knob; they must compute the same
    thing.  SCFGP_FUSED_FEATURES=0: k_features + k_digits_slab instead of the fused feature+digit kernel (the same
    sincos and the same rounding: bit-identical digits, hence bit-identical Gram and Li; only p = PhiT.y is summed in
    another order).  SCFGP_EPILOGUE=1: the one-phase Phi.LiT / backward kernels (dphi and sbar summed in the kernel
    instead of by the dW contraction / the K x K traces).  SCFGP_DIGIT_KERNEL=generic: k_digits_slab instead of the
    16-byte-wide digit kernel (bit-identical digits; q = PhiT.gmu summed in another order).
    Shapes: J = 128 (16-byte plane stores), J = 104 (4-byte words in the sin half), cfg4 geometry, J = 36 (J % 16 != 0)."""
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=N + 7)
    res = {}
    for val in (None, alt):
        if val is None:
            monkeypatch.delenv(knob, raising=False)
        else:
            monkeypatch.setenv(knob, val)
        eng = Engine(D, S, M, precision="emu5")
        eng.set_data(X, y)
        res[val] = eng.eval(h)
        eng.close()
    a, b = res[None], res[alt]
    if knob != "SCFGP_FUSED_FEATURES":
        np.testing.assert_array_equal(a[3], b[3])          # pass 1 untouched: the same Li and alpha
        np.testing.assert_array_equal(a[2], b[2])
    else:
        assert np.abs(a[3] - b[3]).max() == 0.0            # the Gram is bit-identical; alpha sees p's summation order
        assert np.abs(a[2] - b[2]).max() <= 1e-12 * max(1.0, np.abs(b[2]).max())
    assert abs(a[0] - b[0]) <= 1e-12 * abs(b[0])
    # (one-phase: sbar = sum <Phibar, Phi> carries the digit rounding of the backward product, the traces that of the
    #  Gram products -- they agree to the accuracy of the arithmetic, not to round-off)
    assert rel(a[1], b[1]) <= (1e-7 if knob == "SCFGP_EPILOGUE" else 1e-9)


@pytest.mark.parametrize("prec,amp,gtol", [("emu5", 30.0, 1e-5), ("emu6", 1000.0, 1e-6)])
def test_outliers_in_y(prec, amp, gtol):
    """Outlier rows stretch the range of the per-row seeds gv (one power-of-two scale per chunk /
    per rank for Psi = gv Phi): 30-sigma outliers stay inside north_star's gradient tolerance at 5
    digits; 1000-sigma outliers need 6 digits (tools/outlier_check.py prints the whole table)."""
    N, D, S, M = 5000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=N)
    y = y.copy()
    y[::500] += amp
    c0, g0, al0, Li0 = O.streaming_eval(h, X, y, S, M, chunk=1024)
    eng = Engine(D, S, M, precision=prec)
    eng.set_data(X, y)
    c, g, al, Li = eng.eval(h)
    assert abs(c - c0) <= 1e-6 * abs(c0)
    for name, sl in grad_blocks(D, S, M).items():
        if name != "phase":
            assert rel(g[sl], g0[sl]) <= gtol, (name, rel(g[sl], g0[sl]))
    assert np.abs(al - al0).max() <= 1e-5 * max(1.0, np.abs(al0).max())
    eng.close()


def test_pinned_result_buffers():
    """pinned_results=True: same numbers, returned as views of page-locked buffers the engine reuses."""
    N, D, S, M = 3000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=8)
    ref = Engine(D, S, M, precision="emu5"); ref.set_data(X, y)
    c0, g0, a0, L0 = ref.eval(h)
    eng = Engine(D, S, M, precision="emu5", pinned_results=True); eng.set_data(X, y)
    c1, g1, a1, L1 = eng.eval(h)
    assert c1 == c0
    np.testing.assert_array_equal(g1, g0); np.testing.assert_array_equal(L1, L0); np.testing.assert_array_equal(a1, a0)
    h2 = h.copy(); h2[1] += 0.1
    c2, g2, a2, L2 = eng.eval(h2)
    assert g2 is g1 and L2 is L1 and not np.array_equal(g2, g0)     # reused storage, new contents
    ref.close(); eng.close()
