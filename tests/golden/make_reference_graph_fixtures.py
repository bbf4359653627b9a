"""Reference-produced golden vectors for the hot path: outputs of the reference's OWN compiled
Theano functions, evaluated by ``oracle/theano_graph.py`` (a NumPy interpreter of the
``FunctionGraph``s pickled inside ``/root/reference/experiments/boston_housing/boston_scfgp.pkl``).

    python tests/golden/make_reference_graph_fixtures.py        # build container only

Writes ``tests/golden/reference_graph_fixtures.npz`` (committed; the GPU box has no
``/root/reference``).  Per case ``<c>`` the file holds the inputs (``<c>_X``, ``<c>_y``, ``<c>_Xs``,
``<c>_params``) and what the reference functions return on them:

  ``<c>_cost, <c>_alpha, <c>_Li``  = ``train_func(X, y)``            (SCFGP/SCFGP.py:132-135)
  ``<c>_mu,   <c>_std``            = ``pred_func(Xs, alpha, Li)``    (SCFGP/SCFGP.py:145-148)
  ``<c>_grad_idx, <c>_grad_fd``    = d cost / d params[idx] of the reference's cost graph by 4th-order
                                     central differences (what ``TT.grad``, SCFGP.py:129, differentiates)
  ``<c>_dirs, <c>_dir_fd``         = directional derivatives of the same cost along random directions
                                     (each pins the whole P-vector gradient in one projection)

Geometry is the checkpoint's (D=13, S=20, M=50, K=140, P=1333): those shapes are constants of the
pickled graph.  Cases vary N, the parameter vector (trained / init law / ill-conditioned) and the
input range.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.theano_graph import ReferenceFunctions  # noqa: E402
from scfgp_b200 import synth  # noqa: E402  (NumPy-only data generator)

DST = os.path.join(HERE, "reference_graph_fixtures.npz")
D, S, M = 13, 20, 50
P = synth.num_params(D, S, M)


def teacher_problem(N, seed):
    X, y, params0 = synth.make_problem(N, D, S, M, seed=seed)
    return X, y.reshape(-1, 1), params0


def main():
    R = ReferenceFunctions()
    assert R.trained_params.shape == (P,)
    out = {"D": D, "S": S, "M": M, "cases": []}
    rng = np.random.default_rng(20260)

    def add(name, X, y, params, Xs, n_coord, n_dirs):
        cost, alpha, Li = R.train_func(X, y, params)
        mu, std = R.pred_func(Xs, alpha, Li, params)
        out["cases"].append(name)
        out[name + "_X"], out[name + "_y"], out[name + "_Xs"] = X, y, Xs
        out[name + "_params"] = params
        out[name + "_cost"], out[name + "_alpha"], out[name + "_Li"] = cost, alpha, Li
        out[name + "_mu"], out[name + "_std"] = mu, std
        if n_coord:
            # always the three scalars a, b, c; the rest spread over l_f, r_f and the phase blocks
            idx = np.unique(np.concatenate([[0, 1, 2], rng.choice(np.arange(3, P), n_coord - 3, replace=False)]))
            out[name + "_grad_idx"] = idx
            out[name + "_grad_fd"] = R.cost_gradient_fd(X, y, params, index=list(idx))
        if n_dirs:
            dirs = rng.standard_normal((n_dirs, P))
            dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
            from threadpoolctl import threadpool_limits
            with threadpool_limits(1):
                out[name + "_dirs"] = dirs
                out[name + "_dir_fd"] = np.array([R.cost_directional_fd(X, y, params, d) for d in dirs])
        print(f"{name}: N={X.shape[0]} cost={cost:.15g} |alpha|max={np.abs(alpha).max():.3g} "
              f"cond-ish diag(Li) range [{np.diag(Li).min():.3g}, {np.diag(Li).max():.3g}]")

    # 1. the trained Boston hyper-parameters on a Boston-shaped teacher problem (N = 400)
    X, y, _ = teacher_problem(400, seed=11)
    Xs = np.random.default_rng(12).random((106, D))
    add("trained_n400", X, y, R.trained_params, Xs, n_coord=P, n_dirs=4)

    # 2. parameters from the init_params law (a, b, c random too), ragged N
    X, y, _ = teacher_problem(257, seed=21)
    h = synth.draw_params(D, S, M, np.random.default_rng(22), lf_scale=1.0, abc=None)
    Xs = np.random.default_rng(23).random((33, D))
    add("initlaw_n257", X, y, h, Xs, n_coord=120, n_dirs=8)

    # 3. ill-conditioned: small noise variance (a lowered by 2.5), larger signal, N = 1000,
    #    test inputs outside the training range (|Z| large)
    X, y, h = teacher_problem(1000, seed=31)
    h = h.copy()
    h[:3] = [-4.0, 0.5, -3.0]
    Xs = 3.0 * np.random.default_rng(33).standard_normal((64, D))
    add("illcond_n1000", X, y, h, Xs, n_coord=60, n_dirs=8)

    # 4. fewer rows than features (N < M: the 2(N-M)a term changes sign), single test row
    X, y, h = teacher_problem(7, seed=41)
    Xs = np.random.default_rng(43).random((1, D))
    add("tiny_n7", X, y, h, Xs, n_coord=40, n_dirs=4)

    # the arrays the checkpoint itself stores (final train_func outputs, SCFGP.py:265)
    out["ckpt_alpha"], out["ckpt_Li"], out["ckpt_params"] = R.alpha, R.Li, R.trained_params
    Xs = np.random.default_rng(51).random((50, D))
    mu, std = R.pred_func(Xs, R.alpha, R.Li)
    out["ckpt_Xs"], out["ckpt_mu"], out["ckpt_std"] = Xs, mu, std
    out["cases"] = np.array(out["cases"])
    out["op_census_train"] = np.array(sorted(f"{k}:{v}" for k, v in R._train.op_census.items()))
    out["op_census_pred"] = np.array(sorted(f"{k}:{v}" for k, v in R._pred.op_census.items()))
    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
