"""Validity envelope of the emulated (int8 digit) arithmetic: errors of emu5 / emu6 against the FP64 DMMA path
over a grid of hyper-parameters, next to the a-posteriori guard's estimates (scfgp_guard_last).

    python tools/envelope.py [small|cfg4|all] [out.json]

small: D=8, S=4, M=100 (K=208), N=200k -- the full grid a x b x c x lf_scale (108 points)
cfg4 : D=64, S=16, M=1024 (K=2080), N=1M -- a subset, incl. parameters after 200 optimiser steps on a learnable
       teacher problem (small geometry)
Tolerances (north_star): cost 1e-6 relative; gradient blocks, alpha, Li 1e-5."""
import itertools, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scfgp_b200 import Engine, synth

A_GRID = [-4.0, -3.0, -1.5, 0.0]
B_GRID = [-2.0, 0.0, 2.0]
C_GRID = [-4.0, -1.8, 1.0]
LF_GRID = [0.05, 0.3, 1.0]


def blocks(D, S, M):
    P = 3 + D * S + M * S + S + M
    return {"abc": slice(0, 3), "l_f": slice(3, 3 + D * S), "r_f": slice(3 + D * S, 3 + D * S + M * S),
            "phase": slice(3 + D * S + M * S, P)}


def errors(res, ref, D, S, M):
    c, g, al, Li = res
    c0, g0, al0, Li0 = ref
    out = {"cost_rel": abs(c - c0) / max(abs(c0), 1e-300),
           "alpha": float(np.abs(al - al0).max() / max(1.0, np.abs(al0).max())),
           "Li": float(np.abs(Li - Li0).max() / max(1.0, np.abs(Li0).max()))}
    gb = 0.0
    for name, sl in blocks(D, S, M).items():
        if name == "phase":
            gb = max(gb, float(np.abs(g[sl] - g0[sl]).max()))
        else:
            gb = max(gb, float(np.linalg.norm(g[sl] - g0[sl]) / max(np.linalg.norm(g0[sl]), 1e-300)))
    out["grad"] = gb
    out["ok"] = bool(out["cost_rel"] <= 1e-6 and out["grad"] <= 1e-5 and out["alpha"] <= 1e-5 and out["Li"] <= 1e-5)
    return out


def run_point(engs, X, y, h, D, S, M):
    rec = {}
    ref = None
    for prec in ("fp64", "emu5", "emu6"):
        eng = engs[prec]
        try:
            res = eng.eval(h)
        except Exception as ex:                 # not positive definite in this arithmetic, ...
            rec[prec] = {"error": type(ex).__name__}
            if prec == "fp64":
                return rec
            continue
        if prec == "fp64":
            ref = tuple(np.array(v, copy=True) if hasattr(v, "shape") else v for v in res)
            rec["cost"] = ref[0]
            rec["cond_proxy"] = float((np.diag(ref[3]).max() / np.diag(ref[3]).min()) ** 2)
        else:
            rec[prec] = errors(res, ref, D, S, M)
            g = eng.guard_info()
            rec[prec].update(est_cost=g["est_cost"], est_alpha=g["est_alpha"],
                             guard_fires=bool(g["est_cost"] > 1e-7 or g["est_alpha"] > 1e-7))
    return rec


def make_engines(D, S, M, X, y):
    engs = {}
    for prec in ("fp64", "emu5", "emu6"):
        e = Engine(D, S, M, precision=prec)
        e.set_guard(True, 1e300, 1e300)         # measure, never escalate: this tool maps the raw arithmetic
        e.set_data(X, y)
        engs[prec] = e
    return engs


def grid(N, D, S, M, points, tag, out):
    for lf in sorted({p[3] for p in points}):
        X, y, h0 = synth.make_problem(N, D, S, M, seed=5, lf_scale=lf)
        engs = make_engines(D, S, M, X, y)
        for (a, b, c, lf2) in points:
            if lf2 != lf:
                continue
            h = h0.copy()
            h[:3] = [a, b, c]
            rec = run_point(engs, X, y, h, D, S, M)
            rec.update(geometry=tag, N=N, a=a, b=b, c=c, lf_scale=lf)
            out.append(rec)
            e5 = rec.get("emu5", {})
            print(tag, f"a={a:5.1f} b={b:4.1f} c={c:4.1f} lf={lf:4.2f}", "cond~%.1e" % rec.get("cond_proxy", float("nan")),
                  "emu5:", {k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in e5.items()}, flush=True)
        for e in engs.values():
            e.close()


def trained_point(out):
    """Parameters after 200 Adam steps on a learnable teacher problem (small geometry, FP64 path)."""
    from scfgp_b200.optimizer import Optimizer as OPT
    N, D, S, M = 200_000, 8, 4, 100
    X, y, h = synth.make_problem(N, D, S, M, seed=9, lf_scale=0.3, noise=0.02)
    engs = make_engines(D, S, M, X, y)
    rule = OPT.build("adam", {"learning_rate": 0.01, "beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8}, nesterov=0.9)
    for it in range(200):
        c, g, _, _ = engs["fp64"].eval(h)
        h = rule.step(h, g)
    rec = run_point(engs, X, y, h, D, S, M)
    rec.update(geometry="small, after 200 adam steps", N=N, a=float(h[0]), b=float(h[1]), c=float(h[2]), lf_scale=0.3)
    out.append(rec)
    print("trained:", rec, flush=True)
    for e in engs.values():
        e.close()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    dst = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/envelope.json"
    out = []
    t0 = time.time()
    if which in ("small", "all"):
        pts = list(itertools.product(A_GRID, B_GRID, C_GRID, LF_GRID))
        grid(200_000, 8, 4, 100, pts, "small", out)
        trained_point(out)
    if which in ("cfg4", "all"):
        pts = [(-1.575, -0.666, -1.776, 0.3), (-3.0, 0.0, -1.8, 0.3), (-4.0, 2.0, -4.0, 0.3), (0.0, -2.0, 1.0, 0.3),
               (-3.0, 0.0, -1.8, 0.05), (-4.0, 0.0, -1.8, 1.0), (-1.5, 2.0, -1.8, 0.05)]
        grid(1_000_000, 64, 16, 1024, pts, "cfg4", out)
    json.dump(out, open(dst, "w"), indent=1)
    # summary
    for prec in ("emu5", "emu6"):
        rows = [r for r in out if prec in r and "error" not in r[prec]]
        bad = [r for r in rows if not r[prec]["ok"]]
        missed = [r for r in bad if not r[prec]["guard_fires"]]
        false_alarm = [r for r in rows if r[prec]["ok"] and r[prec]["guard_fires"]]
        print(f"{prec}: {len(rows)} points, {len(bad)} outside tolerance, guard misses {len(missed)}, "
              f"guard fires although within tolerance {len(false_alarm)}")
    print("seconds", round(time.time() - t0, 1))


if __name__ == "__main__":
    main()
