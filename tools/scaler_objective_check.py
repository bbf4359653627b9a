"""How closely the GPU Box-Cox objective (two deterministic reduction passes, scfgp_scaler_column_moments) reproduces
the host objective of the reference (skew(boxcox(x, lm), bias=False) ** 2, Scaler.py:58-66), and what that does to
the fitted exponents.  Uses the reference-produced fixtures tests/golden/reference_scaler_fixtures.npz."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scipy.stats import skew
from scfgp_b200 import GpuScaler
from scfgp_b200.scaler import _boxcox, _softplus

fx = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "reference_scaler_fixtures.npz")))
for table in ("mixed", "target", "wide"):
    X = fx[f"{table}_Xfit"]
    ref_lm = fx[f"{table}_auto-normal_boxcox"]
    cols = fx[f"{table}_auto-normal_cols"]
    mn, mx = fx[f"{table}_auto-normal_min"], fx[f"{table}_auto-normal_max"]
    g = GpuScaler("auto-normal")
    g.fit(X)
    print(table, "lambda gpu - reference:", np.array2string(g.data["boxcox"] - ref_lm, precision=3),
          "mu diff", np.abs(g.data["mu"] - fx[f"{table}_auto-normal_mu"]).max(), "evals", g.objective_evals)
    # objective values at the same t: GPU moments vs host formula
    n = X.shape[0]
    worst = 0.0
    for j, d in enumerate(cols):
        col = (X[:, d] - mn[j]) / (mx[j] - mn[j])
        ct = torch.as_tensor(np.ascontiguousarray(col)).cuda()
        for t in (-1.3, 0.0, 0.7, float(np.log(np.expm1(max(ref_lm[j], 1e-3))))):
            lm = _softplus(t)
            host = skew(_boxcox(col, lm), bias=False) ** 2
            _, _, g1 = g._moments(ct, n, lm)
            dev = g._unbiased_skew(g1, n) ** 2
            worst = max(worst, abs(dev - host) / max(abs(host), 1e-300))
    print(table, "max relative difference of the objective (GPU vs host, same lambda):", worst)
