import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np
from scfgp_b200 import engine as E, synth
from oracle import scfgp_oracle as O
def rel(a, b): return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
for amp in (0.0, 30.0, 1e3, 1e5):
    N, D, S, M = 5000, 8, 4, 100
    X, y, h0 = synth.make_problem(N, D, S, M, seed=N)
    y = y.copy(); y[::500] += amp
    c0, g0, al0, Li0 = O.streaming_eval(h0, X, y, S, M, chunk=1024)
    for prec in ("emu5", "emu6", "fp64"):
        eng = E.Engine(D, S, M, precision=prec); eng.set_data(X, y)
        c, g, al, Li = eng.eval(h0)
        b = {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S)}
        print("outlier amp", amp, prec, "cost", f"{abs(c-c0)/abs(c0):.1e}", "grad", {k: f"{rel(g[v], g0[v]):.1e}" for k, v in b.items()},
              "alpha", f"{np.abs(al-al0).max():.1e}", flush=True)
        eng.close()
