"""CPU float64 oracle for SCFGP's marginal-likelihood hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``scfgp_b200``) never routes through this
module and fails loudly when its CUDA library is missing.

PARITY STATUS: pinned to reference-produced outputs (round 2).  The reference's arithmetic lives in the
third-party dependency **Theano** (``requirements.txt:7``, unpinned; the shipped pickle's module
paths indicate the 0.8/0.9 line), which is not installed and not installable offline, and the
reference ships no tests (``tests/__init__.py`` is empty).  The pins are therefore outputs of the
reference itself, produced in the build container and committed as fixtures with their scripts:

* ``oracle/theano_graph.py`` executes the reference's own *compiled Theano functions* -- the
  ``FunctionGraph``s of ``train_func`` / ``pred_func`` pickled inside
  ``experiments/boston_housing/boston_scfgp.pkl`` -- node by node in NumPy: cost, alpha, Li,
  predictions and finite differences of that cost on seeded inputs
  (``tests/golden/reference_graph_fixtures.npz``, ``tests/test_reference_graph.py``);
* ``oracle/theano_shim.py`` runs the reference's unmodified sources (SCFGP.py, Optimizer.py, Scaler.py)
  on a torch-backed stand-in for the Theano API: evaluations + ``TT.grad`` at the cfg1-3 shapes, every
  update rule's trajectory, whole ``optimize`` runs incl. minibatch mode
  (``tests/golden/reference_source_fixtures.npz``, ``tests/test_reference_source.py``);
* the arrays the checkpoint stores and the identities they satisfy (``tests/test_golden.py``).

Two independent restatements live here and are checked against each other:

* ``literal_*``  -- follows ``SCFGP/SCFGP.py:74-148`` line by line (materialised
  Phi, 30-node Gauss-Hermite expected NLL, explicit ``inv(cholesky(A))``), with
  the gradient taken by torch-CPU reverse-mode autograd exactly where the
  reference calls ``TT.grad`` (``SCFGP.py:129``).
* ``streaming_eval`` -- the three-pass, Phi-never-materialised formulation with
  hand-derived adjoints that the CUDA kernels implement,
  in NumPy.
"""
from __future__ import annotations

import math

import numpy as np

EPSILON = 1e-6  # SCFGP.py:93


# --------------------------------------------------------------------------------------
# parameter layout  (SCFGP.py:64-90)
# --------------------------------------------------------------------------------------
def num_params(D: int, S: int, M: int) -> int:
    """Length of the flat vector built by ``init_params`` (SCFGP.py:64-72)."""
    return 3 + D * S + M * S + S + M


def init_params(D: int, S: int, M: int, rng) -> np.ndarray:
    """Same law and order as ``SCFGP.init_params`` (SCFGP.py:64-72), but from an
    explicit generator instead of the global NumPy state."""
    a = rng.standard_normal(1)
    b = rng.standard_normal(1)
    c = rng.standard_normal(1)
    l_f = rng.standard_normal(D * S)
    r_f = rng.random(M * S)
    l_p = 2 * np.pi * rng.random(S)
    p = 2 * np.pi * rng.random(M)
    return np.concatenate([a, b, c, l_f, r_f, l_p, p])


def unpack_params(h: np.ndarray, D: int, S: int, M: int):
    """``SCFGP.unpack_params`` (SCFGP.py:74-90) in NumPy."""
    h = np.asarray(h, dtype=np.float64)
    assert h.shape == (num_params(D, S, M),)
    t = 3
    a, b, c = h[0], h[1], h[2]
    l_F = h[t:t + D * S].reshape(D, S); t += D * S
    r_F = h[t:t + M * S].reshape(M, S); t += M * S
    F = l_F.dot(r_F.T)
    l_P = h[t:t + S].reshape(1, S); t += S
    P = h[t:t + M].reshape(1, M); t += M
    l_FC = l_P - np.mean(l_F, 0)[None, :]
    FC = P - np.mean(F, 0)[None, :]
    return a, b, c, l_F, r_F, F, l_FC, FC


# --------------------------------------------------------------------------------------
# literal restatement of the Theano graph (SCFGP.py:92-148)
# --------------------------------------------------------------------------------------
def literal_forward(h, X, y, S: int, M: int, hermite: bool = True):
    """``train_func(X, y) -> [cost, alpha, Li]`` exactly as the graph builds it
    (SCFGP.py:92-128, 132-135), in NumPy/LAPACK float64.

    ``hermite=False`` replaces the 30-node quadrature (SCFGP.py:118-124) by its closed
    form; the two agree to round-off (tests/test_oracle.py)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    N, D = X.shape
    a, b, c, l_F, r_F, F, l_FC, FC = unpack_params(h, D, S, M)
    kl = lambda mu, sig: sig + mu ** 2 - np.log(sig)
    sig2_n, sig_f = np.exp(2 * a), np.exp(b)
    l_FF = X.dot(l_F) + l_FC
    FF = np.concatenate((l_FF, X.dot(F) + FC), 1)
    Phi = np.concatenate((np.cos(FF), np.sin(FF)), 1)
    Phi = sig_f * np.sqrt(2. / M) * Phi
    noise = np.log(1 + np.exp(c))
    PhiTPhi = Phi.T.dot(Phi)
    A = PhiTPhi + (sig2_n + EPSILON) * np.eye(PhiTPhi.shape[0])
    L = np.linalg.cholesky(A)
    Li = np.linalg.inv(L)
    PhiTy = Phi.T.dot(y)
    beta = Li.dot(PhiTy)
    alpha = Li.T.dot(beta)
    mu_f = Phi.dot(alpha)
    var_f = (Phi.dot(Li.T) ** 2).sum(1)[:, None]
    dsp = noise * (var_f + 1)
    mu_l = np.sum(np.mean(l_F, axis=1))
    sig_l = np.sum(np.std(l_F, axis=1))
    mu_w = np.sum(np.mean(F, axis=1))
    sig_w = np.sum(np.std(F, axis=1))
    if hermite:
        hx, hw = np.polynomial.hermite.hermgauss(30)
        herm_x = hx[None, None, :]
        herm_w = (hw / np.sqrt(np.pi))[None, None, :]
        herm_f = np.sqrt(2 * var_f[:, :, None]) * herm_x + mu_f[:, :, None]
        nlk = (0.5 * herm_f ** 2. - y[:, :, None] * herm_f) / dsp[:, :, None] + 0.5 * (
            np.log(2 * np.pi * dsp[:, :, None]) + y[:, :, None] ** 2 / dsp[:, :, None])
        enll_sum = (herm_w * nlk).sum()
    else:
        enll_sum = 0.5 * (((mu_f - y) ** 2 + var_f) / dsp + np.log(2 * np.pi * dsp)).sum()
    nlml = 2 * np.log(np.diagonal(L)).sum() + 2 * enll_sum + 1. / sig2_n * (
        (y ** 2).sum() - (beta ** 2).sum()) + 2 * (N - M) * a
    penalty = (kl(mu_w, sig_w) * M + kl(mu_l, sig_l) * S) / (S + M)
    cost = (nlml + penalty) / N
    return float(cost), alpha, Li


def literal_cost_grad(h, X, y, S: int, M: int, hermite: bool = True, threads: int | None = None):
    """``cost`` and ``TT.grad(cost, params)`` (SCFGP.py:129) by reverse-mode autograd of
    the literal graph, torch-CPU float64.  Returns (cost, grad, alpha, Li)."""
    import torch

    if threads is not None:
        torch.set_num_threads(int(threads))
    X = torch.as_tensor(np.asarray(X, dtype=np.float64))
    y = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1, 1))
    N, D = X.shape
    hp = torch.tensor(np.asarray(h, dtype=np.float64), requires_grad=True)
    t = 3
    a, b, c = hp[0], hp[1], hp[2]
    l_F = hp[t:t + D * S].reshape(D, S); t += D * S
    r_F = hp[t:t + M * S].reshape(M, S); t += M * S
    F = l_F @ r_F.T
    l_P = hp[t:t + S].reshape(1, S); t += S
    P = hp[t:t + M].reshape(1, M); t += M
    l_FC = l_P - l_F.mean(0)[None, :]
    FC = P - F.mean(0)[None, :]
    kl = lambda mu, sig: sig + mu ** 2 - torch.log(sig)
    sig2_n, sig_f = torch.exp(2 * a), torch.exp(b)
    l_FF = X @ l_F + l_FC
    FF = torch.cat((l_FF, X @ F + FC), 1)
    Phi = torch.cat((torch.cos(FF), torch.sin(FF)), 1)
    Phi = sig_f * math.sqrt(2. / M) * Phi
    noise = torch.log(1 + torch.exp(c))
    PhiTPhi = Phi.T @ Phi
    K = PhiTPhi.shape[0]
    A = PhiTPhi + (sig2_n + EPSILON) * torch.eye(K, dtype=torch.float64)
    L = torch.linalg.cholesky(A)
    Li = torch.linalg.inv(L)
    PhiTy = Phi.T @ y
    beta = Li @ PhiTy
    alpha = Li.T @ beta
    mu_f = Phi @ alpha
    var_f = ((Phi @ Li.T) ** 2).sum(1)[:, None]
    dsp = noise * (var_f + 1)
    mu_l = l_F.mean(1).sum()
    sig_l = l_F.std(1, unbiased=False).sum()
    mu_w = F.mean(1).sum()
    sig_w = F.std(1, unbiased=False).sum()
    if hermite:
        hx, hw = np.polynomial.hermite.hermgauss(30)
        herm_x = torch.as_tensor(hx)[None, None, :]
        herm_w = torch.as_tensor(hw / np.sqrt(np.pi))[None, None, :]
        herm_f = torch.sqrt(2 * var_f[:, :, None]) * herm_x + mu_f[:, :, None]
        nlk = (0.5 * herm_f ** 2. - y[:, :, None] * herm_f) / dsp[:, :, None] + 0.5 * (
            torch.log(2 * np.pi * dsp[:, :, None]) + y[:, :, None] ** 2 / dsp[:, :, None])
        enll_sum = (herm_w * nlk).sum()
    else:
        enll_sum = 0.5 * (((mu_f - y) ** 2 + var_f) / dsp + torch.log(2 * np.pi * dsp)).sum()
    nlml = 2 * torch.log(torch.diagonal(L)).sum() + 2 * enll_sum + 1. / sig2_n * (
        (y ** 2).sum() - (beta ** 2).sum()) + 2 * (N - M) * a
    penalty = (kl(mu_w, sig_w) * M + kl(mu_l, sig_l) * S) / (S + M)
    cost = (nlml + penalty) / N
    cost.backward()
    return (float(cost.detach()), hp.grad.detach().numpy().copy(),
            alpha.detach().numpy().copy(), Li.detach().numpy().copy())


def literal_predict(h, Xs, alpha, Li, S: int, M: int):
    """``pred_func(Xs, alpha, Li) -> [mu_pred (T,1), std_pred (T,)]`` (SCFGP.py:138-148)."""
    Xs = np.asarray(Xs, dtype=np.float64)
    D = Xs.shape[1]
    a, b, c, l_F, r_F, F, l_FC, FC = unpack_params(h, D, S, M)
    sig_f = np.exp(b)
    noise = np.log(1 + np.exp(c))
    l_FFs = Xs.dot(l_F) + l_FC
    FFs = np.concatenate((l_FFs, Xs.dot(F) + FC), 1)
    Phis = np.concatenate((np.cos(FFs), np.sin(FFs)), 1)
    Phis = sig_f * np.sqrt(2. / M) * Phis
    mu_pred = Phis.dot(alpha)
    std_pred = (noise * (1 + (Phis.dot(Li.T) ** 2).sum(1))) ** 0.5
    return mu_pred, std_pred


# --------------------------------------------------------------------------------------
# streaming three-pass restatement -- what the kernels implement
# --------------------------------------------------------------------------------------
def _features(Xc, W, phi, s):
    Z = Xc.dot(W) + phi[None, :]
    return np.concatenate((s * np.cos(Z), s * np.sin(Z)), 1)


def penalty_and_grad(l_F, F, S, M):
    """Penalty of SCFGP.py:94,114-117,127 and its adjoints w.r.t. l_F and F."""
    def stats(Q):
        m = Q.mean(1)
        sd = Q.std(1)
        return m.sum(), sd.sum(), m, sd
    mu_l, sig_l, m_l, sd_l = stats(l_F)
    mu_w, sig_w, m_w, sd_w = stats(F)
    kl = lambda mu, sig: sig + mu ** 2 - np.log(sig)
    pen = (kl(mu_w, sig_w) * M + kl(mu_l, sig_l) * S) / (S + M)

    def adj(Q, mu, sig, m, sd, weight):
        n = Q.shape[1]
        return weight * (2 * mu / n + (1 - 1 / sig) * (Q - m[:, None]) / (n * sd[:, None]))
    dl_F = adj(l_F, mu_l, sig_l, m_l, sd_l, S / (S + M))
    dF = adj(F, mu_w, sig_w, m_w, sd_w, M / (S + M))
    return pen, dl_F, dF


def streaming_eval(h, X, y, S: int, M: int, chunk: int = 4096, want_grad: bool = True):
    """cost, grad, alpha, Li with Phi generated chunk-by-chunk (three passes over X).

    Follows SURVEY.md Appendix A.2-A.6, which restates SCFGP.py:92-129 with the
    Gauss-Hermite term in closed form and the reverse pass derived by hand."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    N, D = X.shape
    a, b, c, l_F, r_F, F, l_FC, FC = unpack_params(h, D, S, M)
    J = S + M
    K = 2 * J
    W = np.concatenate((l_F, F), 1)
    phi = np.concatenate((l_FC.ravel(), FC.ravel()))
    sig2_n = np.exp(2 * a)
    sig_f = np.exp(b)
    tau = np.log(1 + np.exp(c))
    s = sig_f * np.sqrt(2. / M)

    # ---- pass 1
    G = np.zeros((K, K)); p = np.zeros(K); yy = 0.0
    for i in range(0, N, chunk):
        Ph = _features(X[i:i + chunk], W, phi, s)
        yc = y[i:i + chunk]
        G += Ph.T.dot(Ph); p += Ph.T.dot(yc); yy += yc.dot(yc)
    A = G + (sig2_n + EPSILON) * np.eye(K)
    L = np.linalg.cholesky(A)
    Li = np.linalg.inv(L)
    beta = Li.dot(p)
    alpha = Li.T.dot(beta)
    Bm = Li.T.dot(Li)

    # ---- pass 2
    E = 0.0; g_tau = 0.0
    q = np.zeros(K); H = np.zeros((K, K))
    gmu = np.empty(N); gv = np.empty(N)
    for i in range(0, N, chunk):
        Ph = _features(X[i:i + chunk], W, phi, s)
        yc = y[i:i + chunk]
        T = Ph.dot(Li.T)
        v = (T ** 2).sum(1)
        mu = Ph.dot(alpha)
        r = mu - yc
        d = tau * (v + 1)
        E += ((r ** 2 + v) / d + np.log(2 * np.pi * d)).sum()
        k = -(r ** 2 + v) / d ** 2 + 1 / d
        gm = 2 * r / d
        gvv = 1 / d + tau * k
        g_tau += (k * (v + 1)).sum()
        gmu[i:i + chunk] = gm; gv[i:i + chunk] = gvv
        if want_grad:
            q += Ph.T.dot(gm)
            H += Ph.T.dot(gvv[:, None] * Ph)
    nlml = 2 * np.log(np.diagonal(L)).sum() + E + (yy - beta.dot(beta)) / sig2_n + 2 * (N - M) * a
    pen, pen_dlF, pen_dF = penalty_and_grad(l_F, F, S, M)
    cost = (nlml + pen) / N
    if not want_grad:
        return float(cost), None, alpha[:, None], Li

    # ---- small replicated backward (A.4)
    t = Bm.dot(q)
    Abar = Bm - Bm.dot(H).dot(Bm) - 0.5 * (np.outer(t, alpha) + np.outer(alpha, t)) \
        + np.outer(alpha, alpha) / sig2_n
    pbar = t - 2 * alpha / sig2_n
    d_a = 2 * sig2_n * (np.trace(Abar) - (yy - p.dot(alpha)) / sig2_n ** 2) + 2 * (N - M)
    d_c = g_tau / (1 + np.exp(-c))

    # ---- pass 3
    sbar = 0.0
    Wbar = np.zeros((D, J)); phibar = np.zeros(J)
    for i in range(0, N, chunk):
        Xc = X[i:i + chunk]
        Ph = _features(Xc, W, phi, s)
        yc = y[i:i + chunk]
        Phbar = Ph.dot(2 * Abar) + (2 * gv[i:i + chunk])[:, None] * Ph.dot(Bm) \
            + np.outer(yc, pbar) + np.outer(gmu[i:i + chunk], alpha)
        sbar += (Phbar * Ph).sum() / s
        # s*Sn = Ph[:, J:], s*C = Ph[:, :J]
        Zbar = -Phbar[:, :J] * Ph[:, J:] + Phbar[:, J:] * Ph[:, :J]
        Wbar += Xc.T.dot(Zbar)
        phibar += Zbar.sum(0)
    d_b = sbar * s

    # ---- chain to the parameter vector (A.6)
    Wl, WF = Wbar[:, :S].copy(), Wbar[:, S:].copy()
    phl, phF = phibar[:S], phibar[S:]
    Wl -= phl[None, :] / D
    WF -= phF[None, :] / D
    Wl += pen_dlF
    WF += pen_dF
    d_lF = Wl + WF.dot(r_F)
    d_rF = WF.T.dot(l_F)
    grad = np.concatenate(([d_a, d_b, d_c], d_lF.ravel(), d_rF.ravel(), phl, phF)) / N
    return float(cost), grad, alpha[:, None], Li


# --------------------------------------------------------------------------------------
# bounded CPU baseline used by bench.py (reference-shaped: materialised Phi + autograd)
# --------------------------------------------------------------------------------------
def cpu_eval_reference_shaped(h, X, y, S, M, threads=None):
    """One NLML+gradient evaluation the way the reference executes it (dense float64
    ops on a materialised Phi, reverse-mode gradient): used for CPU timing only."""
    return literal_cost_grad(h, X, y, S, M, hermite=True, threads=threads)


# --------------------------------------------------------------------------------------
# phase-split streaming oracle (same phase API as include/scfgp_b200.h) -- lets the CPU
# tests drive the multi-rank reduction protocol with gloo, one row shard per rank
# --------------------------------------------------------------------------------------
class PhaseOracle:
    """NumPy counterpart of the scfgp_eval_* phase API over one row shard.  Buffers returned
    by ``reduce_buffer`` are flat float64 arrays laid out like the CUDA library's:
    RED1 = [G | p | yy], RED2 = [H | q | E | g_tau], RED3 = [dW | dphi | dnlml/db]."""

    def __init__(self, X, y, S, M, n_global=None, chunk=1024):
        self.X = np.asarray(X, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64).reshape(-1)
        self.N, self.D = self.X.shape
        self.S, self.M = S, M
        self.J = S + M
        self.K = 2 * self.J
        self.Ng = n_global or self.N
        self.chunk = chunk

    def _chunks(self):
        for i in range(0, self.N, self.chunk):
            yield i, self.X[i:i + self.chunk], self.y[i:i + self.chunk]

    def begin(self, h):
        a, b, c, l_F, r_F, F, l_FC, FC = unpack_params(h, self.D, self.S, self.M)
        self.a, self.b, self.c, self.l_F, self.r_F, self.F = a, b, c, l_F, r_F, F
        self.W = np.concatenate((l_F, F), 1)
        self.phi = np.concatenate((l_FC.ravel(), FC.ravel()))
        self.sig2n = np.exp(2 * a)
        self.tau = np.log(1 + np.exp(c))
        self.s = np.exp(b) * np.sqrt(2. / self.M)
        K = self.K
        self.red = {1: np.zeros(K * K + K + 1), 2: np.zeros(K * K + K + 2),
                    3: np.zeros(self.D * self.J + self.J + 1)}

    def reduce_buffer(self, which):
        return self.red[which]

    def pass1(self):
        K = self.K
        G = np.zeros((K, K)); p = np.zeros(K); yy = 0.0
        for _, Xc, yc in self._chunks():
            Ph = _features(Xc, self.W, self.phi, self.s)
            G += Ph.T.dot(Ph); p += Ph.T.dot(yc); yy += yc.dot(yc)
        self.red[1][:] = np.concatenate((G.ravel(), p, [yy]))

    def factor(self):
        K = self.K
        G = self.red[1][:K * K].reshape(K, K)
        self.p = self.red[1][K * K:K * K + K].copy()
        self.yy = self.red[1][-1]
        A = G + (self.sig2n + EPSILON) * np.eye(K)
        self.L = np.linalg.cholesky(A)
        self.Li = np.linalg.inv(self.L)
        self.beta = self.Li.dot(self.p)
        self.alpha = self.Li.T.dot(self.beta)
        self.Bm = self.Li.T.dot(self.Li)

    def pass2(self):
        K = self.K
        E = 0.0; gt = 0.0; q = np.zeros(K); H = np.zeros((K, K))
        self.gmu = np.empty(self.N); self.gv = np.empty(self.N)
        for i, Xc, yc in self._chunks():
            Ph = _features(Xc, self.W, self.phi, self.s)
            v = (Ph.dot(self.Li.T) ** 2).sum(1)
            r = Ph.dot(self.alpha) - yc
            d = self.tau * (v + 1)
            E += ((r ** 2 + v) / d + np.log(2 * np.pi * d)).sum()
            k = -(r ** 2 + v) / d ** 2 + 1 / d
            gm = 2 * r / d; gvv = 1 / d + self.tau * k
            gt += (k * (v + 1)).sum()
            self.gmu[i:i + len(yc)] = gm; self.gv[i:i + len(yc)] = gvv
            q += Ph.T.dot(gm); H += Ph.T.dot(gvv[:, None] * Ph)
        self.red[2][:] = np.concatenate((H.ravel(), q, [E, gt]))

    def backward(self):
        K = self.K
        H = self.red[2][:K * K].reshape(K, K)
        q = self.red[2][K * K:K * K + K]
        t = self.Bm.dot(q)
        al = self.alpha
        self.Abar = self.Bm - self.Bm.dot(H).dot(self.Bm) - 0.5 * (np.outer(t, al) + np.outer(al, t)) \
            + np.outer(al, al) / self.sig2n
        self.pbar = t - 2 * al / self.sig2n

    def pass3(self):
        J = self.J
        Wbar = np.zeros((self.D, J)); phibar = np.zeros(J); db = 0.0
        for i, Xc, yc in self._chunks():
            Ph = _features(Xc, self.W, self.phi, self.s)
            n = len(yc)
            Phbar = Ph.dot(2 * self.Abar) + (2 * self.gv[i:i + n])[:, None] * Ph.dot(self.Bm) \
                + np.outer(yc, self.pbar) + np.outer(self.gmu[i:i + n], self.alpha)
            db += (Phbar * Ph).sum()
            Zbar = -Phbar[:, :J] * Ph[:, J:] + Phbar[:, J:] * Ph[:, :J]
            Wbar += Xc.T.dot(Zbar); phibar += Zbar.sum(0)
        self.red[3][:] = np.concatenate((Wbar.ravel(), phibar, [db]))

    def finish(self, want_grad=True):
        K, J, D, S, M, Ng = self.K, self.J, self.D, self.S, self.M, self.Ng
        E, gt = self.red[2][-2], self.red[2][-1]
        nlml = 2 * np.log(np.diagonal(self.L)).sum() + E + (self.yy - self.beta.dot(self.beta)) / self.sig2n \
            + 2 * (Ng - M) * self.a
        pen, pen_dlF, pen_dF = penalty_and_grad(self.l_F, self.F, S, M)
        cost = (nlml + pen) / Ng
        if not want_grad:
            return float(cost), None, self.alpha[:, None], self.Li
        Wbar = self.red[3][:D * J].reshape(D, J)
        phibar = self.red[3][D * J:D * J + J]
        d_b = self.red[3][-1]
        d_a = 2 * self.sig2n * (np.trace(self.Abar) - (self.yy - self.p.dot(self.alpha)) / self.sig2n ** 2) \
            + 2 * (Ng - M)
        d_c = gt / (1 + np.exp(-self.c))
        Wl = Wbar[:, :S] - phibar[None, :S] / D + pen_dlF
        WF = Wbar[:, S:] - phibar[None, S:] / D + pen_dF
        grad = np.concatenate(([d_a, d_b, d_c], (Wl + WF.dot(self.r_F)).ravel(), WF.T.dot(self.l_F).ravel(),
                               phibar[:S], phibar[S:])) / Ng
        return float(cost), grad, self.alpha[:, None], self.Li
