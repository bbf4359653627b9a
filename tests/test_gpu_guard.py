"""The a-posteriori guard of the emulated arithmetic (scfgp_set_guard / scfgp_guard_last, csrc/host_guard.cuh):
quiet in the benign regime, escalating 5 -> 6 digits -> FP64 where the digit Gram is not accurate enough for the
conditioning of A -- so that results stay inside north_star's tolerances whatever the
hyper-parameters.  Geometry and grid points are those of tools/envelope.py (profiles/r2_envelope_*.txt)."""
import numpy as np
import pytest

from scfgp_b200 import Engine, synth

pytestmark = pytest.mark.gpu

N, D, S, M = 200_000, 8, 4, 100


def blocks():
    P = 3 + D * S + M * S + S + M
    return {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S),
            "phase": slice(3 + D * S + M * S, P)}


def within_tolerance(res, ref):
    c, g, al, Li = res
    c0, g0, al0, Li0 = ref
    ok = abs(c - c0) <= 1e-6 * abs(c0)
    ok &= np.abs(al - al0).max() <= 1e-5 * max(1.0, np.abs(al0).max())
    ok &= np.abs(Li - Li0).max() <= 1e-5 * max(1.0, np.abs(Li0).max())
    for name, sl in blocks().items():
        if name == "phase":
            ok &= np.abs(g[sl] - g0[sl]).max() <= 1e-5
        else:
            ok &= np.linalg.norm(g[sl] - g0[sl]) <= 1e-5 * np.linalg.norm(g0[sl])
    return bool(ok)


@pytest.fixture(scope="module")
def problem():
    X, y, h = synth.make_problem(N, D, S, M, seed=5, lf_scale=0.3)
    e64 = Engine(D, S, M, precision="fp64")
    e64.set_data(X, y)
    yield X, y, h, e64
    e64.close()


def evaluate(problem, abc, precision, guard):
    X, y, h0, e64 = problem
    h = h0.copy()
    h[:3] = abc
    ref = tuple(np.array(v, copy=True) if hasattr(v, "shape") else v for v in e64.eval(h))
    eng = Engine(D, S, M, precision=precision)
    eng.set_guard(guard)
    eng.set_data(X, y)
    res = eng.eval(h)
    info = eng.guard_info()
    eng.close()
    return res, ref, info


def test_quiet_at_benign_hyperparameters(problem):
    res, ref, info = evaluate(problem, synth.TRAINED_ABC, "emu5", True)
    assert within_tolerance(res, ref)
    assert info["escalations"] == 0 and info["precision"] == "emu5" and info["sample_rows"] == 2048
    assert 0 < info["est_alpha"] < 1e-7 and info["est_cost"] < 1e-8


def test_five_digits_leave_the_tolerance_when_ill_conditioned(problem):
    """Why the guard exists: sigma_n^2 = e^-8, sigma_f = e^2 (cond(A) ~ 3e8) -- 5 digits miss the gradient by 5e-4."""
    res, ref, info = evaluate(problem, (-4.0, 2.0, -1.8), "emu5", False)
    assert not within_tolerance(res, ref)
    assert info["escalations"] == 0 and info["precision"] == "emu5"


@pytest.mark.parametrize("abc", [(-4.0, 2.0, -1.8), (-3.0, 2.0, -1.8), (-3.0, 0.0, 1.0)])
def test_guard_escalates_and_results_stay_inside_the_tolerance(problem, abc):
    res, ref, info = evaluate(problem, abc, "emu5", True)
    assert within_tolerance(res, ref), info
    assert info["escalations"] >= 1 and info["precision"] in ("emu6", "fp64")
    assert not info["exceeded"]                       # the estimates of the precision finally used are below the thresholds


def test_escalation_is_sticky_and_reported(problem):
    X, y, h0, _ = problem
    h = h0.copy()
    h[:3] = (-4.0, 2.0, -1.8)
    eng = Engine(D, S, M, precision="emu5")
    eng.set_data(X, y)
    eng.eval(h)
    first = eng.guard_info()
    eng.eval(h0)                                       # benign parameters again: the handle stays where it went
    assert eng.precision == first["precision"] != "emu5"
    assert eng.guard_info()["escalations"] == first["escalations"]
    eng.set_precision("emu5")                          # the caller may go back explicitly
    eng.eval(h0)
    assert eng.precision == "emu5"
    eng.close()
