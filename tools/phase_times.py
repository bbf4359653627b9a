"""Wall time of each phase of one evaluation (phase API, device synchronised between phases):
shows the replicated, N-independent part that limits multi-GPU scaling.
python tools/phase_times.py [N] [precision]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scfgp_b200 import Engine, synth, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prec = sys.argv[2] if len(sys.argv) > 2 else "emu5"
D, S, M = 64, 16, 1024
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
eng = Engine(D, S, M, precision=prec)
eng.set_data_device(X, y)
eng.eval(h)
lib, hd = eng.lib, eng._h
hp = _lib.dptr(_lib.as_f64(h))
cost = np.zeros(1); grad = np.zeros(eng.P); al = np.zeros(eng.K); Li = np.zeros((eng.K, eng.K))
for rep in range(2):
    ts = []
    def run(name, fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        ts.append((name, (time.perf_counter() - t0) * 1e3))
    run("begin", lambda: eng._check(lib.scfgp_eval_begin(hd, hp, 1)))
    run("pass1", lambda: eng._check(lib.scfgp_eval_pass1(hd)))
    run("factor", lambda: eng._check(lib.scfgp_eval_factor(hd)))
    run("pass2", lambda: eng._check(lib.scfgp_eval_pass2(hd)))
    run("backward", lambda: eng._check(lib.scfgp_eval_backward(hd)))
    run("pass3", lambda: eng._check(lib.scfgp_eval_pass3(hd)))
    run("finish", lambda: eng._check(lib.scfgp_eval_finish(hd, _lib.dptr(cost), _lib.dptr(grad), _lib.dptr(al), _lib.dptr(Li))))
print(prec, "N", N, " ".join(f"{n} {t:.2f}" for n, t in ts), "total", round(sum(t for _, t in ts), 2), "ms")
