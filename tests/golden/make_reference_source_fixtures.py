"""Golden vectors produced by EXECUTING the reference's own, unmodified Python sources
(``/root/reference/SCFGP/SCFGP.py``, ``Optimizer.py``, ``Scaler.py``) in the build container.

Theano is not installable here, so the sources run on ``oracle/theano_shim.py``: a lazy torch-CPU
float64 implementation of the ~25 Theano entry points they call (``shared``, ``function`` with
``givens``/``updates``, ``tensor.*``, ``grad``, ``sandbox.linalg``).  Model code, optimiser rules,
training loop, minibatch loop, scalers, metrics and the global-NumPy-RNG draw order are the
reference's; only the tensor backend is substituted.  That the backend computes what Theano computes
is checked against Theano's own compiled graph (``oracle/theano_graph.py``) in
``tests/test_reference_source.py``.

    python tests/golden/make_reference_source_fixtures.py        # build container only

Writes ``tests/golden/reference_source_fixtures.npz`` (committed):

  ``eval_<shape>_*``   one ``train_func`` call + the gradient ``TT.grad(cost, params)`` (SCFGP.py:129) at the
                       BASELINE shapes kin8nm / abalone / co2 / boston (cfg1-3), extracted from one
                       ``train_iter_func`` step with ``algo='sgd'`` (p1 = p0 - 1.9 lr g on the first step,
                       Optimizer.py:62-119), and ``pred_func`` outputs
  ``algo_<name>_*``    parameter trajectories of six ``train_iter_func`` steps for every update rule
                       (SCFGP.py:130-137, Optimizer.py:27-382) -- including the Nesterov wrapper's
                       ``list(updates.keys())[0]`` behaviour
  ``fit_boston_*``     a whole ``set_data`` + ``optimize`` + ``predict`` run on a Boston-shaped table:
                       scaled data, initial parameters from ``np.random.seed``, COST / metric
                       histories, final parameters, alpha, Li, predictions (SCFGP.py:153-294)
  ``minibatch_*``      the same with ``nbatches=2, batchsize=150`` (SCFGP.py:172-182,226-235)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import theano_shim  # noqa: E402
from scfgp_b200 import synth  # noqa: E402

DST = os.path.join(HERE, "reference_source_fixtures.npz")
METRICS = ["SCORE", "COST", "MAE", "NMAE", "MSE", "NMSE", "MNLP"]


def bound_params(model):
    """The parameter vector the compiled functions act on (``givens``, SCFGP.py:135)."""
    return model.train_func.givens[0][1].get_value()


def identity_scaled_model(pkg, X, y, S, M):
    """A reference model whose X, y are already in scaler-output form (``set_data`` would rescale them):
    mirrors what ``set_data`` leaves behind (SCFGP.py:164-169) without the scalers."""
    m = pkg.SCFGP(sparsity=S, nfeats=M)
    m.X, m.y = X, y
    m.N, m.D = X.shape
    m.init_params()
    return m


def boston_table(seed, n_train=400, n_valid=106, D=13):
    rng = np.random.default_rng(seed)
    w = rng.standard_normal(D)

    def draw(n):
        X = rng.lognormal(size=(n, D)) * np.linspace(0.5, 8.0, D)
        X[:, 3] = rng.integers(0, 2, n)                      # a binary column, like CHAS
        y = np.tanh(np.log1p(X) @ w / 3) * 9 + 22 + rng.standard_normal(n)
        return X, y[:, None]
    return draw(n_train) + draw(n_valid)


def main():
    pkg = theano_shim.import_reference()
    out = {}

    # ---------------------------------------------------------------- eval + gradient at the cfg shapes
    shapes = {"boston": (400, 13, 5, 50), "kin8nm": (2048, 8, 4, 100), "abalone": (1044, 10, 5, 50),
              "co2": (555, 1, 100, 50)}
    for name, (N, D, S, M) in shapes.items():
        X, y, h = synth.make_problem(N, D, S, M, seed=len(name))
        y = y[:, None]
        np.random.seed(1)
        m = identity_scaled_model(pkg, X, y, S, M)
        m.params.set_value(h)
        lr = 1e-3
        m.build_theano_models("sgd", {"learning_rate": lr})
        cost0, alpha0, Li0 = m.train_func(X, y)
        cost1, alpha1, Li1 = m.train_iter_func(X, y)          # outputs at the pre-update parameters
        assert float(cost0) == float(cost1)
        grad = (h - bound_params(m)) / (1.9 * lr)
        m.params.set_value(h)
        Xs = np.random.default_rng(5).random((64, D))
        mu, std = m.pred_func(Xs, alpha0, Li0)
        k = f"eval_{name}_"
        out[k + "X"], out[k + "y"], out[k + "params"], out[k + "Xs"] = X, y, h, Xs
        out[k + "SM"] = np.array([S, M])
        out[k + "cost"], out[k + "alpha"], out[k + "Li"], out[k + "grad"] = float(cost0), alpha0, Li0, grad
        out[k + "mu"], out[k + "std"] = mu, std
        print(f"{k}: cost {float(cost0):.12g} |grad| {np.linalg.norm(grad):.6g}")

    # ---------------------------------------------------------------- update rules
    N, D, S, M = 150, 4, 3, 8
    X, y, h = synth.make_problem(N, D, S, M, seed=77)
    y = y[:, None]
    out["algo_X"], out["algo_y"], out["algo_params0"], out["algo_SM"] = X, y, h, np.array([S, M])
    settings = {
        "sgd": {"learning_rate": 0.02},
        "adagrad": {"learning_rate": 0.05, "epsilon": 1e-6},
        "adadelta": {"learning_rate": 1.0, "rho": 0.95, "epsilon": 1e-6},
        "adam": {"learning_rate": 0.01, "beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8},
        "adamax": {"learning_rate": 0.02, "beta1": 0.9, "beta2": 0.999, "epsilon": 1e-8},
    }
    for algo, kw in settings.items():
        np.random.seed(2)
        m = identity_scaled_model(pkg, X, y, S, M)
        m.params.set_value(h)
        m.build_theano_models(algo, kw)
        traj, costs = [], []
        for _ in range(6):
            c, _, _ = m.train_iter_func(X, y)
            costs.append(float(c))
            traj.append(bound_params(m))
        out[f"algo_{algo}_traj"], out[f"algo_{algo}_cost"] = np.array(traj), np.array(costs)
        out[f"algo_{algo}_kw"] = np.array([f"{k}={v}" for k, v in kw.items()])
        print(f"algo {algo}: cost {costs[0]:.6g} -> {costs[-1]:.6g}")
    try:
        m = identity_scaled_model(pkg, X, y, S, M)
        m.build_theano_models("rmsprop", {"learning_rate": 0.01})
        out["algo_rmsprop_raises"] = np.array("")
    except NameError as e:                                    # Optimizer.py:210 reads an undefined name
        out["algo_rmsprop_raises"] = np.array(f"NameError: {e}")
    print("rmsprop:", out["algo_rmsprop_raises"])

    # ---------------------------------------------------------------- whole fits through the public API
    Xtr, ytr, Xva, yva = boston_table(seed=2026)
    for tag, opt in (("fit_boston", dict(max_iter=25)),
                     ("minibatch", dict(max_iter=10, nbatches=2, batchsize=150))):
        np.random.seed(7)
        m = pkg.SCFGP(sparsity=5, nfeats=50)
        m.set_data(Xtr, ytr)
        p0 = m.params.get_value()
        with theano_shim.numpy1_names():
            m.optimize(Xva, yva, **opt)
        mu_y, std_y = m.predict(Xva)
        k = tag + "_"
        out[k + "Xtr"], out[k + "ytr"], out[k + "Xva"], out[k + "yva"] = Xtr, ytr, Xva, yva
        out[k + "seed"], out[k + "SM"] = 7, np.array([5, 50])
        out[k + "opt"] = np.array([f"{a}={b}" for a, b in opt.items()])
        out[k + "ID"] = np.array(m.ID)
        out[k + "X_scaled"], out[k + "y_scaled"] = m.X, m.y
        out[k + "params0"], out[k + "params_final"] = p0, bound_params(m)
        out[k + "alpha"], out[k + "Li"] = m.alpha, m.Li
        out[k + "mu_y"], out[k + "std_y"] = mu_y, std_y
        for metric in METRICS:
            out[k + "evals_" + metric] = np.array([float(v) for v in m.evals[metric][1]])
        print(f"{tag}: COST {out[k + 'evals_COST'][0]:.6g} -> {out[k + 'evals_COST'][-1]:.6g}, "
              f"NMSE {out[k + 'evals_NMSE'][-1]:.4g}, {len(out[k + 'evals_COST'])} entries")

    np.savez_compressed(DST, **out)
    print("wrote", DST, os.path.getsize(DST), "bytes")


if __name__ == "__main__":
    main()
