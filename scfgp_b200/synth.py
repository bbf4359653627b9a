"""Synthetic problems of the reference's shapes.

The reference's datasets (Boston, kin8nm, abalone, CO2) are not available offline, so the
parity tests and the bench use seeded stand-ins: X ~ U(0,1)^(N x D) (what the
``auto-inv-normal`` X scaler outputs, Scaler.py:112-116), a teacher drawn from the
``init_params`` law (SCFGP/SCFGP.py:64-72), and y standardised so that sum(y^2) = N as the
``auto-normal`` y scaler guarantees.
"""
from __future__ import annotations

import numpy as np

# (N, D, S, M) of BASELINE.json's configs
CONFIGS = {
    "boston": (506, 13, 5, 50),
    "boston_pkl": (400, 13, 20, 50),
    "kin8nm": (8192, 8, 4, 100),
    "abalone": (4177, 10, 5, 50),
    "co2": (2225, 1, 100, 50),
    "cfg4": (10_000_000, 64, 16, 1024),
    "cfg5": (100_000_000, 256, 32, 2048),
}

TRAINED_ABC = (-1.575304860841505, -0.6655536119352673, -1.7756344623098925)  # shipped Boston model


def num_params(D, S, M):
    return 3 + D * S + M * S + S + M


def draw_params(D, S, M, rng, lf_scale=1.0, abc=TRAINED_ABC):
    """Parameter vector with the layout/law of SCFGP.init_params (SCFGP/SCFGP.py:64-72)."""
    a = rng.standard_normal(1)
    b = rng.standard_normal(1)
    c = rng.standard_normal(1)
    l_f = lf_scale * rng.standard_normal(D * S)
    r_f = rng.random(M * S)
    l_p = 2 * np.pi * rng.random(S)
    p = 2 * np.pi * rng.random(M)
    h = np.concatenate([a, b, c, l_f, r_f, l_p, p])
    if abc is not None:
        h[:3] = abc
    return h


def features_np(h, X, S, M):
    """Phi of SCFGP/SCFGP.py:98-102 (NumPy; data generation only)."""
    N, D = X.shape
    t = 3
    b = h[1]
    l_F = h[t:t + D * S].reshape(D, S); t += D * S
    r_F = h[t:t + M * S].reshape(M, S); t += M * S
    F = l_F @ r_F.T
    l_p = h[t:t + S]; t += S
    p = h[t:t + M]
    W = np.concatenate([l_F, F], 1)
    phi = np.concatenate([l_p - l_F.mean(0), p - F.mean(0)])
    Z = X @ W + phi
    return np.exp(b) * np.sqrt(2.0 / M) * np.concatenate([np.cos(Z), np.sin(Z)], 1)


def make_problem(N, D, S, M, seed=0, chunk=65536, lf_scale=0.3, noise=0.05):
    """Returns X (N,D), y (N,), params0 (P,).  Teacher features are generated chunk-wise so
    N = 1e7 fits in host memory."""
    rng = np.random.default_rng(seed)
    X = rng.random((N, D))
    teacher = draw_params(D, S, M, rng, lf_scale=lf_scale)
    w = rng.standard_normal(2 * (S + M))
    y = np.empty(N)
    for i in range(0, N, chunk):
        y[i:i + chunk] = features_np(teacher, X[i:i + chunk], S, M) @ w
    y += noise * y.std() * rng.standard_normal(N)
    y = (y - y.mean()) / y.std()
    params0 = draw_params(D, S, M, rng, lf_scale=lf_scale)
    return X, y, params0
