"""Column-wise data preprocessing on the host (restates SCFGP/Scaler.py:11-135).

Five algorithms (Scaler.py:15-21): ``min-max``, ``normal``, ``inv-normal``, ``auto-normal``,
``auto-inv-normal``.  The two ``auto-*`` variants min-max scale each column, pick a
per-column Box-Cox exponent lambda = softplus(t), t in [-5, 5], that minimises the squared
sample skewness (SLSQP, Scaler.py:58-70 / 81-93), then standardise (``auto-normal``) or map
through the fitted normal CDF (``auto-inv-normal``, output in (0,1)).  Constant columns
are dropped at fit time (Scaler.py:40-41).
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize
from scipy.stats import norm, skew


def _boxcox(x, lm):
    return (np.sign(x) * np.abs(x) ** lm - 1) / lm


def _inv_boxcox(x, lm):
    return np.sign(x * lm + 1) * np.abs(x * lm + 1) ** (1. / lm)


def _softplus(t):
    return np.log(np.exp(t) + 1)


class Scaler(object):
    " Scaler (Data Preprocessing) "

    algos = ["min-max", "normal", "inv-normal", "auto-normal", "auto-inv-normal"]

    _FIELDS = {
        "min-max": ("min", "max"),
        "normal": ("std", "mu"),
        "inv-normal": ("std", "mu"),
        "auto-normal": ("min", "max", "std", "mu", "boxcox"),
        "auto-inv-normal": ("min", "max", "std", "mu", "boxcox"),
    }

    def __init__(self, algo):
        assert algo.lower() in self.algos, "Invalid Scaling Algorithm!"
        self.algo = algo.lower()
        self.data = {"cols": None}
        self.data.update({k: 0 for k in self._FIELDS[self.algo]})

    # ------------------------------------------------------------------ fit
    def _fit_boxcox(self, tX):
        """Per-column lambda (Scaler.py:58-70): columns with < 10 distinct values keep lambda = 1."""
        lms = np.zeros(tX.shape[1])
        for d in range(tX.shape[1]):
            col = tX[:, d]
            if np.unique(col).shape[0] < 10:
                lms[d] = 1
                continue
            obj = lambda t: skew(_boxcox(col, _softplus(t[0])), bias=False) ** 2
            t = minimize(obj, [0.], method='SLSQP', bounds=[(-5, 5)],
                         options={'ftol': 1e-8, 'maxiter': 100, 'disp': False})['x']
            lms[d] = _softplus(t[0])
        return lms

    def fit(self, X):
        X = np.asarray(X)
        const = np.where(np.all(X == X[0, :], axis=0))[0]
        self.data["cols"] = sorted(set(range(X.shape[1])).difference(const))
        tX = X[:, self.data["cols"]]
        if self.algo == "min-max":
            self.data["min"], self.data["max"] = np.min(tX, axis=0), np.max(tX, axis=0)
        elif self.algo in ("normal", "inv-normal"):
            self.data["mu"], self.data["std"] = np.mean(tX, axis=0), np.std(tX, axis=0)
        else:
            self.data["min"], self.data["max"] = np.min(tX, axis=0), np.max(tX, axis=0)
            tX = (tX - self.data["min"]) / (self.data["max"] - self.data["min"])
            self.data["boxcox"] = self._fit_boxcox(tX)
            tX = _boxcox(tX, self.data["boxcox"][None, :])
            self.data["mu"], self.data["std"] = np.mean(tX, axis=0), np.std(tX, axis=0)

    # ------------------------------------------------------------------ transforms
    def forward_transform(self, X):
        d = self.data
        tX = np.asarray(X)[:, d["cols"]]
        if self.algo == "min-max":
            return (tX - d["min"]) / (d["max"] - d["min"])
        if self.algo == "normal":
            return (tX - d["mu"]) / d["std"]
        if self.algo == "inv-normal":
            return norm.cdf((tX - d["mu"]) / d["std"])
        tX = _boxcox((tX - d["min"]) / (d["max"] - d["min"]), d["boxcox"][None, :])
        if self.algo == "auto-normal":
            return (tX - d["mu"]) / d["std"]
        return norm.cdf(tX, d["mu"], d["std"])

    def backward_transform(self, X):
        d = self.data
        X = np.asarray(X)
        assert len(d["cols"]) == X.shape[1], "Backward Transform Error"
        if self.algo == "min-max":
            return X * (d["max"] - d["min"]) + d["min"]
        if self.algo == "normal":
            return X * d["std"] + d["mu"]
        if self.algo == "inv-normal":
            # as in the reference (Scaler.py:124-125): not the exact inverse of forward_transform
            return (norm.ppf(X) - d["mu"]) / d["std"]
        lm = d["boxcox"][None, :]
        if self.algo == "auto-normal":
            tX = X * d["std"] + d["mu"]
        else:
            tX = norm.ppf(X, d["mu"], d["std"])
        return _inv_boxcox(tX, lm) * (d["max"] - d["min"]) + d["min"]

    # ------------------------------------------------------------------ checkpoint helpers
    def state_dict(self):
        out = {"algo": self.algo}
        for k, v in self.data.items():
            out[k] = np.asarray(v)
        return out

    @classmethod
    def from_state_dict(cls, st):
        sc = cls(str(st["algo"]))
        for k, v in st.items():
            if k == "algo":
                continue
            v = np.asarray(v)
            sc.data[k] = [int(c) for c in v.ravel()] if k == "cols" else v
        return sc
