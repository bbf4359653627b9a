"""Self-test of the emulated-FP64 (int8 tensor core) path on a GPU box (developer tool;
tests/test_gpu_emu.py covers the same ground under pytest).

    python tools/emu_selftest.py [--big]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from scfgp_b200 import engine as E, synth  # noqa: E402
from oracle import scfgp_oracle as O  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


def main():
    rng = np.random.default_rng(1)
    # ---- digit GEMM: error against the normwise bound rowmax(A) rowmax(B) k
    for digits in (4, 5, 6, 7, 8):
        for (m, n, k) in [(128, 64, 64), (130, 70, 33), (300, 260, 200), (257, 129, 1000)]:
            A = rng.standard_normal((m, k)) * np.exp(rng.standard_normal((m, 1)) * 3)
            B = rng.standard_normal((n, k)) * np.exp(rng.standard_normal((n, 1)) * 3)
            Cm = E.dbg_emu_gemm(A, B, digits)
            ref = A @ B.T
            bound = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * k
            print("emu gemm digits", digits, (m, n, k), "err/bound", float((np.abs(Cm - ref) / bound).max()),
                  "2^-(8d-2)", 2.0 ** (-(8 * digits - 2)), flush=True)

    shapes = [(400, 13, 20, 50), (506, 13, 5, 50), (1000, 8, 4, 100), (333, 1, 100, 50), (2500, 10, 5, 50),
              (5000, 64, 16, 120), (9000, 8, 4, 100)]
    for (N, D, S, M) in shapes:
        X, y, h0 = synth.make_problem(N, D, S, M, seed=N)
        c0, g0, al0, Li0 = O.streaming_eval(h0, X, y, S, M, chunk=1024)
        Xs = np.random.default_rng(5).random((257, D))
        mu0, sd0 = O.literal_predict(h0, Xs, al0, Li0, S, M)
        for prec in ("emu4", "emu5", "emu6"):
            eng = E.Engine(D, S, M, precision=prec)
            eng.set_data(X, y)
            c, g, al, Li = eng.eval(h0)
            P = len(h0)
            blocks = {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S)}
            msg = {k: rel(g[v], g0[v]) for k, v in blocks.items()}
            ph = float(np.abs(g[3 + D * S + M * S:] - g0[3 + D * S + M * S:]).max())
            cf, _, _, _ = eng.eval(h0, want_grad=False)
            mu, sd = eng.predict(h0, Xs, al, Li)
            print(prec, (N, D, S, M), "cost rel", abs(c - c0) / abs(c0), "grad", msg, "phase abs", ph,
                  "alpha", float(np.abs(al - al0).max()), "Li", float(np.abs(Li - Li0).max()),
                  "fwd-only", abs(cf - c), "pred", float(np.abs(mu - mu0).max()), float(np.abs(sd - sd0).max()),
                  flush=True)
            eng.close()

    if "--big" in sys.argv:
        N, D, S, M = 200_000, 64, 16, 1024
        X, y, h0 = synth.make_problem(N, D, S, M, seed=3)
        res = {}
        for prec in ("fp64", "emu5", "emu4", "emu6"):
            eng = E.Engine(D, S, M, precision=prec)
            eng.set_data(X, y)
            for rep in range(2):
                t0 = time.time()
                c, g, al, Li = eng.eval(h0)
                t1 = time.time()
                print("big eval", prec, N, "sec", round(t1 - t0, 4), "pass ms", eng.pass_ms(), "cost", c, flush=True)
            res[prec] = (c, g, al, Li)
            eng.close()
        c0, g0, al0, Li0 = res["fp64"]
        for prec in ("emu4", "emu5", "emu6"):
            c, g, al, Li = res[prec]
            print("cfg4-geometry N=200000", prec, "vs fp64: cost rel", abs(c - c0) / abs(c0), "grad rel", rel(g, g0),
                  "grad max rel-to-max", float(np.abs(g - g0).max() / np.abs(g0).max()),
                  "alpha", float(np.abs(al - al0).max()), "Li", float(np.abs(Li - Li0).max()), flush=True)


if __name__ == "__main__":
    main()
