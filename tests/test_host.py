"""Host-side logic (no GPU): C-ABI exports, optimiser rules, scalers, facade plumbing."""
import ctypes
import os
import re

import numpy as np
import pytest

import scfgp_b200
from scfgp_b200 import Optimizer, Scaler, SCFGP, _lib, shard_bounds, weighted_bounds
from scfgp_b200 import build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ C ABI
def test_library_builds_and_exports_every_declared_symbol():
    libbuild.build()
    header = open(os.path.join(ROOT, "include", "scfgp_b200.h")).read()
    declared = set(re.findall(r"\b(scfgp_[a-z0-9_]+)\s*\(", header))
    declared.discard("scfgp_handle")
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/scfgp_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), declared.symmetric_difference(_lib.PROTOTYPES)
    lib.scfgp_abi_version.restype = ctypes.c_int
    assert lib.scfgp_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    """create() must fail loudly (not fall back) when there is no CUDA device."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(scfgp_b200.ScfgpError):
        scfgp_b200.Engine(4, 2, 8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "scfgp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


# ------------------------------------------------------------------ sharding
def test_shard_bounds_cover_rows():
    for n in (0, 1, 7, 8, 1000, 10_000_001):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_weighted_bounds_cover_rows_and_follow_the_weights():
    for n in (0, 5, 1000, 10_000_001):
        for w in ([1.0], [1.0, 1.0], [1.0, 0.97, 1.02, 1.0], [3.0, 1.0]):
            for granule in (1, 128):
                spans = [weighted_bounds(n, r, w, granule) for r in range(len(w))]
                assert spans[0][0] == 0 and spans[-1][1] == n
                assert all(spans[i][1] == spans[i + 1][0] and spans[i][0] <= spans[i][1] for i in range(len(w) - 1))
                for (a, b), wr in zip(spans, w):
                    assert abs((b - a) - n * wr / sum(w)) <= granule + 1
    # equal weights: the plain balanced split (up to the rounding of the cut points)
    assert [weighted_bounds(1000, r, [1, 1, 1, 1]) for r in range(4)] == [shard_bounds(1000, r, 4) for r in range(4)]
    with pytest.raises(ValueError):
        weighted_bounds(10, 0, [1.0, 0.0])


# ------------------------------------------------------------------ optimiser (Optimizer.py)
def test_sgd_nesterov_is_real_nesterov():
    rule = Optimizer.build("sgd", {"learning_rate": 0.1}, nesterov=0.9)
    p = np.array([1.0, -2.0]); g = np.array([0.5, 0.25])
    vel = np.zeros(2)
    for _ in range(3):
        upd = p - 0.1 * g                   # Optimizer.py:118
        x = 0.9 * vel + upd - p             # Optimizer.py:94
        vel = x
        expect = 0.9 * x + upd              # Optimizer.py:96
        p = rule.step(p, g)
        np.testing.assert_allclose(p, expect, rtol=1e-15)


def test_adam_matches_reference_formulas_with_quirk():
    kw = dict(learning_rate=0.01, beta1=0.9, beta2=0.999, epsilon=1e-8)
    rule = Optimizer.build("adam", kw, nesterov=0.9, reference_quirk=True)
    rng = np.random.default_rng(0)
    p = rng.standard_normal(5)
    m = np.zeros(5); v = np.zeros(5); vel = np.zeros(5); t = 0
    for _ in range(4):
        g = rng.standard_normal(5)
        t += 1
        a_t = 0.01 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        m_t = 0.9 * m + 0.1 * g
        v_t = 0.999 * v + 0.001 * g ** 2
        p_ref = p - a_t * m_t / (np.sqrt(v_t) + 1e-8)          # Optimizer.py:322-329
        x = 0.9 * vel + m_t - m                                 # wrapper acts on m_prev (Optimizer.py:89-96)
        vel = x
        m = 0.9 * x + m_t
        v = v_t
        p = rule.step(p, g)
        np.testing.assert_allclose(p, p_ref, rtol=1e-14)
    np.testing.assert_allclose(rule.slots["m"], m, rtol=1e-14)


@pytest.mark.parametrize("algo", ["sgd", "adagrad", "rmsprop", "adadelta", "adam", "adamax"])
def test_all_rules_descend_a_quadratic(algo):
    rule = Optimizer.build(algo, {"learning_rate": 0.05}, nesterov=0.9)
    p = np.array([3.0, -4.0])
    f0 = (p ** 2).sum()
    for _ in range(300):
        p = rule.step(p, 2 * p)
    assert (p ** 2).sum() < f0


def test_optimizer_state_roundtrip():
    r1 = Optimizer.build("adam", {"learning_rate": 0.01})
    p = np.ones(3)
    for _ in range(3):
        p = r1.step(p, p)
    r2 = Optimizer.build("adam", {"learning_rate": 0.01})
    r2.load_state_dict(r1.state_dict())
    np.testing.assert_array_equal(r1.step(p, p), r2.step(p, p))


# ------------------------------------------------------------------ scalers (Scaler.py)
@pytest.mark.parametrize("algo", ["min-max", "normal", "auto-normal", "auto-inv-normal"])
def test_scaler_roundtrip(algo):
    rng = np.random.default_rng(1)
    X = np.column_stack([rng.lognormal(size=300), rng.random(300) * 5 + 1, np.ones(300),
                         rng.integers(0, 3, 300).astype(float)])
    sc = Scaler(algo)
    sc.fit(X)
    assert sc.data["cols"] == [0, 1, 3]                      # constant column dropped (Scaler.py:40-41)
    Z = sc.forward_transform(X)
    assert Z.shape == (300, 3)
    back = sc.backward_transform(Z)
    np.testing.assert_allclose(back, X[:, [0, 1, 3]], rtol=1e-6, atol=1e-8)
    if algo == "auto-inv-normal":
        assert Z.min() > 0 and Z.max() < 1
    if algo == "auto-normal":
        np.testing.assert_allclose(Z.mean(0), 0, atol=1e-9)
        np.testing.assert_allclose((Z ** 2).sum(0), 300, rtol=1e-9)   # sum y^2 = N


def test_scaler_boxcox_reduces_skew():
    from scipy.stats import skew
    rng = np.random.default_rng(2)
    X = rng.lognormal(sigma=1.0, size=(500, 1))
    sc = Scaler("auto-normal"); sc.fit(X)
    assert abs(skew(sc.forward_transform(X)[:, 0])) < 0.1 < abs(skew(X[:, 0]))


def test_scaler_from_golden_state(golden):
    st = {k[3:]: v for k, v in golden.items() if k.startswith("ys_")}
    sc = Scaler.from_state_dict(st)
    assert sc.algo == "auto-normal"
    y = np.linspace(float(golden["ys_min"][0]), float(golden["ys_max"][0]), 20)[:, None]
    np.testing.assert_allclose(sc.backward_transform(sc.forward_transform(y)), y, rtol=1e-8, atol=1e-8)
    xs = Scaler.from_state_dict({k[3:]: v for k, v in golden.items() if k.startswith("xs_")})
    assert xs.algo == "auto-inv-normal" and len(xs.data["cols"]) == 13


# ------------------------------------------------------------------ facade plumbing
def test_constructor_spellings():
    m = SCFGP(sparsity=20, nfeats=18)                        # SCFGP.py:33
    assert (m.S, m.M) == (20, 18) and m.fftype == _lib.FF_PHASE and not m.verbose
    m = SCFGP(-1, 20, False)                                 # experiments/kin8nm/test_regression.py
    assert m.S_requested == -1 and m.M == 20
    m = SCFGP(rank=4, feature_size=100, fftype="non-phase", msg=True)   # README.md:47-50
    assert (m.S, m.M, m.fftype, m.verbose) == (4, 100, _lib.FF_NOPHASE, True)
    assert set(m.evals) == {"SCORE", "COST", "MAE", "NMAE", "MSE", "NMSE", "MNLP", "TIME(s)"}
    with pytest.raises(ValueError):
        SCFGP(4, 10, "bogus")


def test_init_params_layout():
    m = SCFGP(sparsity=3, nfeats=7)
    m.D = 5
    m.init_params(np.random.default_rng(0))
    D, S, M = 5, 3, 7
    assert m.params.shape == (3 + D * S + M * S + S + M,)
    r_f = m.params[3 + D * S:3 + D * S + M * S]
    ph = m.params[3 + D * S + M * S:]
    assert r_f.min() >= 0 and r_f.max() < 1 and ph.min() >= 0 and ph.max() < 2 * np.pi
