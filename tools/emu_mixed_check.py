"""Parity of mixed digit counts ('emu5/4' etc.) against the oracle on the BASELINE shapes and
against the FP64 DMMA path at the cfg4 geometry.  python tools/emu_mixed_check.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scfgp_b200 import engine as E, synth
from oracle import scfgp_oracle as O

def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))

shapes = [(400, 13, 20, 50), (400, 13, 5, 50), (5000, 8, 4, 100), (3133, 10, 5, 50), (2225, 1, 100, 50), (777, 7, 3, 30),
          (9000, 8, 4, 100), (5000, 64, 16, 120)]
precs = sys.argv[1:] or ["emu5/4", "emu6/4", "emu5"]
for (N, D, S, M) in shapes:
    X, y, h0 = synth.make_problem(N, D, S, M, seed=N)
    c0, g0, al0, Li0 = O.streaming_eval(h0, X, y, S, M, chunk=1024)
    Xs = np.random.default_rng(5).random((257, D))
    mu0, sd0 = O.literal_predict(h0, Xs, al0, Li0, S, M)
    for prec in precs:
        eng = E.Engine(D, S, M, precision=prec)
        eng.set_data(X, y)
        c, g, al, Li = eng.eval(h0)
        b = {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S)}
        mu, sd = eng.predict(h0, Xs, al, Li)
        print(prec, (N, D, S, M), "cost", f"{abs(c - c0) / abs(c0):.1e}", "grad", {k: f"{rel(g[v], g0[v]):.1e}" for k, v in b.items()},
              "phase", f"{np.abs(g[3 + D * S + M * S:] - g0[3 + D * S + M * S:]).max():.1e}",
              "alpha", f"{np.abs(al - al0).max():.1e}", "Li", f"{np.abs(Li - Li0).max():.1e}",
              "pred", f"{np.abs(mu - mu0).max():.1e}", f"{np.abs(sd - sd0).max():.1e}", flush=True)
        eng.close()
N, D, S, M = 400_000, 64, 16, 1024
X, y, h0 = synth.make_problem(N, D, S, M, seed=3)
res = {}
for prec in ["fp64"] + precs:
    eng = E.Engine(D, S, M, precision=prec)
    eng.set_data(X, y)
    eng.eval(h0)
    t0 = time.time(); r = eng.eval(h0); t1 = time.time()
    print("cfg4 geometry N=400k", prec, "sec", round(t1 - t0, 4), "pass ms", [round(x, 1) for x in eng.pass_ms()], flush=True)
    res[prec] = r
    eng.close()
c0, g0, al0, Li0 = res["fp64"]
for prec in precs:
    c, g, al, Li = res[prec]
    print("  ", prec, "vs fp64: cost", f"{abs(c - c0) / abs(c0):.1e}", "grad", f"{rel(g, g0):.1e}",
          "grad max/max", f"{np.abs(g - g0).max() / np.abs(g0).max():.1e}", "alpha", f"{np.abs(al - al0).max():.1e}",
          "Li", f"{np.abs(Li - Li0).max():.1e}")
