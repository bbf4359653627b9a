"""Extract the only reference-produced numerical artefact into a small fixture.

Source: ``/root/reference/experiments/boston_housing/boston_scfgp.pkl`` -- written by the
reference's ``SCFGP.save`` (SCFGP/SCFGP.py:296-302).  It pickles compiled Theano
functions, so it cannot be ``pickle.load``-ed without Theano.  This script walks it with
a *restricted* unpickler: only NumPy array reconstruction and a few builtins are allowed,
every other class is replaced by an inert stub (nothing from the pickle is executed).

Output (committed): ``tests/golden/boston_scfgp_golden.npz`` with
  params  (1333,)  trained hyper-parameter vector (the shared variable bound inside the
                   pickled ``train_func``; the top-level ``params`` key is an untrained
                   vector)
  alpha   (140,1), Li (140,140)   outputs of the final ``train_func`` call (SCFGP.py:265)
  M, S, D, N_train
  evals_<METRIC>   the 33-step history
  xs_*/ys_*        scaler states (Scaler.py:24-36)

Run from the repo root in the build container (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.restricted_pickle import BOSTON_PKL as SRC, load_stub_tree, walk_arrays  # noqa: E402

DST = os.path.join(HERE, "boston_scfgp_golden.npz")


def main():
    top = load_stub_tree(SRC)      # restricted unpickler: NumPy arrays + inert stubs, nothing is executed
    assert isinstance(top, dict), type(top)
    print("keys:", sorted(top.keys()))
    Li = np.asarray(top["Li"], dtype=np.float64)
    alpha = np.asarray(top["alpha"], dtype=np.float64)
    M = int(top["M"])
    K = Li.shape[0]
    S = K // 2 - M
    arrs = walk_arrays(top["train_func"])
    cands = [a for a in arrs if a.dtype == np.float64 and a.ndim == 1 and a.size > 100]
    sizes = sorted({a.size for a in cands})
    print("1-D float64 arrays under train_func:", sizes)
    # P = 3 + D*S + M*S + S + M  ->  D
    params = None
    for a in cands:
        rem = a.size - 3 - M * S - S - M
        if rem > 0 and rem % S == 0:
            # the trained vector has r_f outside [0,1); the optimiser slots (m, v) are same-sized
            if params is None or np.abs(a[:3] - [-1.5753, -0.6656, -1.7756]).max() < 1e-3:
                params = a
    assert params is not None
    D = (params.size - 3 - M * S - S - M) // S
    print("M,S,D,P =", M, S, D, params.size, "params[:3] =", params[:3])

    out = dict(params=params, alpha=alpha, Li=Li, M=M, S=S, D=D, N_train=400)
    for metric, val in top["evals"].items():
        out["evals_" + metric.replace("(s)", "_s")] = np.asarray(val[1], dtype=np.float64)
    for tag, key in (("xs", "X_scaler"), ("ys", "y_scaler")):
        sc = top[key]
        st = sc._state if sc._state is not None else vars(sc)
        if isinstance(st, dict):
            out[tag + "_algo"] = str(st.get("algo"))
            data = st.get("data", {})
            for k, v in data.items():
                if v is None:
                    continue
                out[f"{tag}_{k}"] = np.asarray(v)
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")
    for k, v in out.items():
        print(" ", k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
