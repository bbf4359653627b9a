"""Kernel-by-kernel numerical self-test on a GPU box (developer tool; the pytest -m gpu
suite covers the same ground).  Prints one line per check so that a single gpurun tells
which kernel family is off.

    python tools/gpu_selftest.py [--big]
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
    # ---- GEMM layouts, ragged sizes
    for (m, n, k) in [(128, 128, 16), (128, 128, 64), (130, 70, 33), (300, 260, 200), (64, 512, 1000)]:
        A = rng.standard_normal((m, k)); B = rng.standard_normal((n, k))
        print("gemm NT", m, n, k, rel(E.dbg_gemm(0, A, B), A @ B.T), flush=True)
        A = rng.standard_normal((k, m)); B = rng.standard_normal((k, n))
        print("gemm TN", m, n, k, rel(E.dbg_gemm(1, A, B), A.T @ B), flush=True)
        A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
        print("gemm NN", m, n, k, rel(E.dbg_gemm(2, A, B), A @ B), flush=True)
    # ---- factor
    for K in [8, 64, 110, 140, 208, 600]:
        Q = rng.standard_normal((K + 50, K))
        A = Q.T @ Q + 0.1 * np.eye(K)
        L, Li, info = E.dbg_factor(A)
        L0 = np.linalg.cholesky(A)
        print("factor", K, "info", info, "L", rel(L, L0), "Li", rel(Li, np.linalg.inv(L0)), flush=True)
    A = -np.eye(64)
    print("factor notpd info", E.dbg_factor(A)[2], flush=True)

    # ---- features / eval / predict vs oracle
    shapes = [(400, 13, 20, 50), (506, 13, 5, 50), (1000, 8, 4, 100), (333, 1, 100, 50), (2500, 10, 5, 50),
              (5000, 64, 16, 120)]
    for (N, D, S, M) in shapes:
        X, y, h0 = synth.make_problem(N, D, S, M, seed=N)
        for fft in (0,):
            eng = E.Engine(D, S, M, fftype=fft)
            Phi = eng.features(h0, X)
            Phi0 = synth.features_np(h0, X, S, M)
            print("features", (N, D, S, M), float(np.abs(Phi - Phi0).max()), flush=True)
            eng.set_data(X, y)
            t0 = time.time()
            c, g, al, Li = eng.eval(h0)
            t1 = time.time()
            c0, g0, al0, Li0 = O.streaming_eval(h0, X, y, S, M, chunk=1024)
            P = len(h0)
            blocks = {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S),
                      "phase": slice(3 + D * S + M * S, P)}
            msg = {k: (rel(g[v], g0[v]) if k != "phase" else float(np.abs(g[v] - g0[v]).max())) for k, v in blocks.items()}
            print("eval", (N, D, S, M), "cost", c, c0, abs(c - c0) / abs(c0), "grad", msg,
                  "alpha", float(np.abs(al - al0).max()), "Li", float(np.abs(Li - Li0).max()),
                  "sec", round(t1 - t0, 4), "launches", eng.launches, flush=True)
            cf, _, _, _ = eng.eval(h0, want_grad=False)
            print("  fwd-only cost diff", abs(cf - c), flush=True)
            Xs = np.random.default_rng(5).random((257, D))
            mu, sd = eng.predict(h0, Xs, al, Li)
            mu0, sd0 = O.literal_predict(h0, Xs, al0, Li0, S, M)
            print("  predict", float(np.abs(mu - mu0).max()), float(np.abs(sd - sd0).max()), flush=True)
            eng.close()

    if "--big" in sys.argv:
        N, D, S, M = 200_000, 64, 16, 1024
        X, y, h0 = synth.make_problem(N, D, S, M, seed=3)
        eng = E.Engine(D, S, M)
        eng.set_data(X, y)
        for rep in range(2):
            t0 = time.time()
            c, g, al, Li = eng.eval(h0)
            t1 = time.time()
            print("big eval", N, "sec", t1 - t0, "pass ms", eng.pass_ms(), "cost", c, flush=True)
        t0 = time.time()
        c0, g0, al0, Li0 = O.streaming_eval(h0, X[:20000], y[:20000], S, M, chunk=4096)
        print("oracle 20000 rows sec", time.time() - t0, flush=True)
        eng.set_data(X[:20000], y[:20000])
        c, g, al, Li = eng.eval(h0)
        print("cfg4-shape N=20000: cost rel", abs(c - c0) / abs(c0), "grad rel", rel(g, g0), "alpha", float(np.abs(al - al0).max()),
              "Li", float(np.abs(Li - Li0).max()), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
