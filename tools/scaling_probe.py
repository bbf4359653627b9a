"""What does not scale with the rows: one GPU, the cfg4 geometry at N = 10M / {1, 2, 4, 8} rows (the per-rank work of a
1 / 2 / 4 / 8-GPU run without the all-reduces), wall time per evaluation and the per-pass device times, and the
least-squares fit  t(N) = fixed + per_row * N.   python tools/scaling_probe.py [precision=emu5] [Nmax=10000000]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scfgp_b200 import Engine, synth

prec = sys.argv[1] if len(sys.argv) > 1 else "emu5"
Nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
D, S, M = 64, 16, 1024
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.rand((Nmax, D), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn(Nmax, dtype=torch.float64, device="cuda", generator=g)
h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
eng = Engine(D, S, M, precision=prec, pinned_results=True)
pts = []
for div in (1, 2, 4, 8):
    n = Nmax // div
    eng.set_data_device(X[:n], y[:n], n_global=Nmax)
    for _ in range(2):
        eng.eval(h)
    torch.cuda.synchronize()
    reps = 3 if div < 4 else 6
    t0 = time.perf_counter()
    pm = []
    for _ in range(reps):
        eng.eval(h)
        pm.append(eng.pass_ms())
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    pmean = [sum(p[i] for p in pm) / reps for i in range(4)]
    pts.append((n, ms, pmean))
    print(f"N={n:9d}  {ms:9.2f} ms/eval   pass1 {pmean[0]:8.2f}  pass2 {pmean[1]:8.2f}  pass3 {pmean[2]:8.2f}  dense {pmean[3]:6.2f}"
          f"   outside the passes {ms - sum(pmean):6.2f}", flush=True)
A = np.array([[1.0, p[0]] for p in pts]); 
for name, vals in (("eval", [p[1] for p in pts]), ("pass1", [p[2][0] for p in pts]), ("pass2", [p[2][1] for p in pts]),
                   ("pass3", [p[2][2] for p in pts])):
    f, w = np.linalg.lstsq(A, np.array(vals), rcond=None)[0]
    print(f"fit {name:6s}: fixed {f:7.2f} ms + {w * 1e6:8.3f} ms per million rows")
f, w = np.linalg.lstsq(A, np.array([p[1] for p in pts]), rcond=None)[0]
for gpus in (2, 4, 8):
    print(f"  predicted efficiency on {gpus} GPUs without collectives: {(f + w * Nmax) / gpus / (f + w * Nmax / gpus):.4f}")
