"""Short target for `ncu --set full`: two evaluations (the first warms up) of the cfg4 geometry at N rows on
ONE chunk stream with the a-posteriori guard off, so that the launch sequence per evaluation is fixed:
pass 1 [features+digits, gram] x chunks, pass 2 [tgemm, digits, gram] x chunks, pass 3 [digits, backward, dw] x chunks.
    python tools/ncu_target.py [N=16384] [precision=emu5]"""
import os, sys
os.environ.setdefault("SCFGP_GUARD", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scfgp_b200 import Engine, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
prec = sys.argv[2] if len(sys.argv) > 2 else "emu5"
D, S, M = 64, 16, 1024
g = torch.Generator(device="cuda"); g.manual_seed(1)
X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
eng = Engine(D, S, M, precision=prec)
eng.set_streams(1)
eng.set_data_device(X, y)
l0 = eng.launches
c0 = eng.eval(h)[0]
l1 = eng.launches
c1 = eng.eval(h)[0]
torch.cuda.synchronize()
print("cost", c0, c1, "launches per evaluation", l1 - l0, eng.launches - l1, "cache bytes", eng.phi_cache_bytes)
