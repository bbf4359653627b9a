"""In-kernel timing of the tcgen05 kernels' roles (scfgp_dbg_tile_timing): where one tile's time goes.
Runs one evaluation of the cfg4 geometry at reduced N through the phase API and records each streaming
pass separately.   python tools/tile_timing.py [N] [streams]"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scfgp_b200 import Engine, synth, _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 327680
streams = int(sys.argv[2]) if len(sys.argv) > 2 else 0
D, S, M = 64, 16, 1024
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
eng = Engine(D, S, M, precision="emu5")
if streams:
    eng.set_streams(streams)
eng.set_data_device(X, y)
eng.eval(h); eng.eval(h)
lib, hd = eng.lib, eng._h
hp = _lib.dptr(_lib.as_f64(h))
cost = np.zeros(1); grad = np.zeros(eng.P); al = np.zeros(eng.K); Li = np.zeros((eng.K, eng.K))
NC = 148
names = ["tiles", "mma:wait acc_empty", "mma:wait operands", "mma:tile span", "epi:wait acc_full",
         "epi:acc_full->release", "epi:release->end", "producer:wait free stage"]

def timed(label, fn):
    out = (C.c_ulonglong * (NC * 8))()
    torch.cuda.synchronize()
    eng._check(lib.scfgp_dbg_tile_timing(0, 1, None, NC))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    eng._check(lib.scfgp_dbg_tile_timing(0, 0, out, NC))
    a = np.array(out[:], dtype=np.float64).reshape(NC, 8)
    tiles = a[:, 0].sum()
    print(f"== {label}: {e0.elapsed_time(e1):.2f} ms, {int(tiles)} tiles, per tile (clk, mean over CTAs):")
    for i in range(1, 8):
        print(f"   {names[i]:28s} {a[:, i].sum() / max(tiles, 1):10.0f}")

eng._check(lib.scfgp_eval_begin(hd, hp, 1))
timed("pass1 (gram)", lambda: eng._check(lib.scfgp_eval_pass1(hd)))
eng._check(lib.scfgp_eval_factor(hd))
timed("pass2 (tgemm + gram)", lambda: eng._check(lib.scfgp_eval_pass2(hd)))
eng._check(lib.scfgp_eval_backward(hd))
timed("pass3 (backward)", lambda: eng._check(lib.scfgp_eval_pass3(hd)))
eng._check(lib.scfgp_eval_finish(hd, _lib.dptr(cost), _lib.dptr(grad), _lib.dptr(al), _lib.dptr(Li)))
print("cost", cost[0], "pass_ms", eng.pass_ms())
