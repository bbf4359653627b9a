"""Golden vectors from the reference's own ``Scaler`` (row f3).

``/root/reference/SCFGP/Scaler.py`` imports only NumPy/SciPy, so -- unlike the Theano model -- it
runs in the build container.  This script loads that file *by path* (the package ``__init__`` would
pull in Theano/matplotlib), runs ``fit`` / ``forward_transform`` / ``backward_transform``
(Scaler.py:39-135) of every algorithm on seeded tables and stores inputs and outputs in
``tests/golden/reference_scaler_fixtures.npz`` (committed; the GPU box has no ``/root/reference``).

    python tests/golden/make_reference_scaler_fixtures.py        # build container only

``tests/test_reference_scaler.py`` pins ``scfgp_b200/scaler.py`` (CPU) and
``tests/test_gpu_scaler.py`` pins ``GpuScaler`` against these reference outputs.
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/SCFGP/Scaler.py"
DST = os.path.join(HERE, "reference_scaler_fixtures.npz")
ALGOS = ["min-max", "normal", "inv-normal", "auto-normal", "auto-inv-normal"]


def load_reference_scaler():
    spec = importlib.util.spec_from_file_location("reference_scfgp_scaler", SRC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.Scaler


def tables():
    """name -> (fit table, table to transform).  Covers skewed columns (Box-Cox search), a constant
    column (dropped, Scaler.py:40-41), a column with fewer than 10 distinct values (lambda = 1,
    Scaler.py:61-63), a single-column target and out-of-range transform inputs."""
    rng = np.random.default_rng(7)
    N = 400
    mixed = np.column_stack([
        rng.lognormal(0.0, 0.8, N),            # right-skewed
        rng.standard_normal(N) * 3 + 10,       # symmetric
        np.full(N, 2.5),                       # constant -> dropped
        rng.integers(0, 4, N).astype(float),   # 4 distinct values -> lambda 1
        rng.beta(5.0, 1.5, N),                 # left-skewed
        rng.gamma(2.0, 2.0, N),                # right-skewed
        rng.exponential(1.0, N) ** 2,          # heavy right tail
    ])
    mixed_new = np.column_stack([
        rng.lognormal(0.0, 0.8, 57), rng.standard_normal(57) * 3 + 10, np.full(57, 2.5),
        rng.integers(0, 4, 57).astype(float), rng.beta(5.0, 1.5, 57), rng.gamma(2.0, 2.0, 57),
        rng.exponential(1.0, 57) ** 2])
    target = rng.gamma(3.0, 7.0, (N, 1))       # a y column (auto-normal, SCFGP.py:35)
    target_new = rng.gamma(3.0, 7.0, (31, 1))
    wide = rng.lognormal(0.0, 0.5, (500, 13)) * np.linspace(0.5, 20, 13)
    return {"mixed": (mixed, mixed_new), "target": (target, target_new), "wide": (wide, wide[::17] * 1.01)}


def main():
    Scaler = load_reference_scaler()
    out = {"algos": np.array(ALGOS), "tables": []}
    for tname, (Xfit, Xnew) in tables().items():
        out["tables"].append(tname)
        out[f"{tname}_Xfit"], out[f"{tname}_Xnew"] = Xfit, Xnew
        for algo in ALGOS:
            sc = Scaler(algo)
            sc.fit(Xfit)
            key = f"{tname}_{algo}"
            out[key + "_cols"] = np.asarray(sc.data["cols"], dtype=np.int64)
            for field in ("min", "max", "mu", "std", "boxcox"):
                if field in sc.data:
                    out[f"{key}_{field}"] = np.asarray(sc.data[field], dtype=np.float64)
            fwd_fit = sc.forward_transform(Xfit)
            fwd_new = sc.forward_transform(Xnew)
            out[key + "_fwd_fit"], out[key + "_fwd_new"] = fwd_fit, fwd_new
            out[key + "_bwd_fit"] = sc.backward_transform(fwd_fit)
            print(f"{key}: cols {sc.data['cols']} lambda {np.round(sc.data.get('boxcox', np.nan), 6)}")
    out["tables"] = np.array(out["tables"])
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
