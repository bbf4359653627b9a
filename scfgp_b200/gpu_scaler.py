"""Column-wise data preprocessing streamed on the GPU.

``GpuScaler`` has the interface, the fitted fields and the checkpoint format of the host
``Scaler`` (which restates SCFGP/Scaler.py:11-135) and can be swapped for it anywhere.  Every
pass over the data is a bandwidth-bound kernel behind the C ABI (``scfgp_scaler_*``):

* ``fit``: column min / max / constant-column detection in one pass; for the ``auto-*``
  algorithms each column is gathered once into a contiguous min-max scaled device buffer, the
  "fewer than 10 distinct values" test (Scaler.py:63) and every objective evaluation of the
  Box-Cox search (squared unbiased sample skewness, Scaler.py:64-66) run on that buffer (two
  reduction passes per evaluation), and the scalar search itself is the reference's: SciPy
  SLSQP on t in [-5, 5], lambda = softplus(t), ftol 1e-8 (Scaler.py:67-70).
* ``forward_transform``: one elementwise kernel; a CUDA tensor in gives a CUDA tensor out
  (ready for ``Engine.set_data_device`` without another host round trip), a NumPy array in
  gives a NumPy array out.
* ``backward_transform`` acts on predictions (T x 1) and stays on the host.

There is no CPU fallback: the class needs a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.optimize import minimize

from . import _lib
from .scaler import Scaler, _softplus

_ALGO_ID = {"min-max": 0, "normal": 1, "inv-normal": 2, "auto-normal": 3, "auto-inv-normal": 4}


class GpuScaler(Scaler):
    " Scaler (Data Preprocessing) on the device "

    def __init__(self, algo, device: int = 0):
        super().__init__(algo)
        self.device = int(device)
        self.lib = _lib.load()
        self.objective_evals = 0

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, what):
        if rc != _lib.OK:
            raise RuntimeError(f"{what} failed (rc={rc})")

    def _to_device(self, X):
        import torch
        if isinstance(X, torch.Tensor):
            if not X.is_cuda or X.dtype != torch.float64:
                X = X.to(device=f"cuda:{self.device}", dtype=torch.float64)
            return X.contiguous(), True
        Xn = _lib.as_f64(X)
        return torch.as_tensor(Xn).to(f"cuda:{self.device}"), False

    def _moments(self, col_t, n, lm):
        mean, var, g1 = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.scfgp_scaler_column_moments(self.device, C.c_void_p(col_t.data_ptr()), n, float(lm),
                                                         C.byref(mean), C.byref(var), C.byref(g1)), "column_moments")
        return mean.value, var.value, g1.value

    @staticmethod
    def _unbiased_skew(g1, n):
        # scipy.stats.skew(..., bias=False) (Scaler.py:64): G1 = sqrt(n (n - 1)) / (n - 2) * g1
        return np.sqrt(n * (n - 1.0)) / (n - 2.0) * g1 if n > 2 else g1

    # ------------------------------------------------------------------ fit
    def fit(self, X):
        import torch
        Xd, _ = self._to_device(X)
        n, D = int(Xd.shape[0]), int(Xd.shape[1])
        mn, mx = np.zeros(D), np.zeros(D)
        const = (C.c_int * D)()
        self._check(self.lib.scfgp_scaler_colstats(self.device, C.c_void_p(Xd.data_ptr()), n, D, D, _lib.dptr(mn),
                                                   _lib.dptr(mx), const), "colstats")
        cols = [d for d in range(D) if not const[d]]
        self.data["cols"] = cols
        col_t = torch.empty(n, dtype=torch.float64, device=Xd.device)
        ptr = C.c_void_p(Xd.data_ptr())

        def gather(d, shift, rng):
            self._check(self.lib.scfgp_scaler_gather_column(self.device, ptr, n, D, d, float(shift), float(rng),
                                                            C.c_void_p(col_t.data_ptr())), "gather_column")

        if self.algo == "min-max":
            self.data["min"], self.data["max"] = mn[cols], mx[cols]
            return
        if self.algo in ("normal", "inv-normal"):
            mu, sd = np.zeros(len(cols)), np.zeros(len(cols))
            for j, d in enumerate(cols):
                gather(d, 0.0, 1.0)
                m, v, _ = self._moments(col_t, n, -1.0)
                mu[j], sd[j] = m, np.sqrt(v)
            self.data["mu"], self.data["std"] = mu, sd
            return
        self.data["min"], self.data["max"] = mn[cols], mx[cols]
        lms, mu, sd = np.zeros(len(cols)), np.zeros(len(cols)), np.zeros(len(cols))
        for j, d in enumerate(cols):
            gather(d, mn[d], mx[d] - mn[d])
            nd = C.c_int()
            self._check(self.lib.scfgp_scaler_column_distinct(self.device, C.c_void_p(col_t.data_ptr()), n, 9,
                                                              C.byref(nd)), "column_distinct")
            if nd.value < 10:
                lms[j] = 1.0
            else:
                def obj(t):
                    self.objective_evals += 1
                    _, _, g1 = self._moments(col_t, n, _softplus(t[0]))
                    return self._unbiased_skew(g1, n) ** 2
                t = minimize(obj, [0.], method='SLSQP', bounds=[(-5, 5)],
                             options={'ftol': 1e-8, 'maxiter': 100, 'disp': False})['x']
                lms[j] = _softplus(t[0])
            m, v, _ = self._moments(col_t, n, lms[j])
            mu[j], sd[j] = m, np.sqrt(v)
        self.data["boxcox"], self.data["mu"], self.data["std"] = lms, mu, sd

    # ------------------------------------------------------------------ transforms
    def forward_transform(self, X):
        import torch
        Xd, was_tensor = self._to_device(X)
        d = self.data
        cols = np.asarray(d["cols"], dtype=np.int32)
        n, D, nc = int(Xd.shape[0]), int(Xd.shape[1]), len(cols)
        out = torch.empty((n, nc), dtype=torch.float64, device=Xd.device)
        nul = C.POINTER(C.c_double)()

        keep = [_lib.as_f64(np.broadcast_to(np.asarray(d[k], dtype=np.float64), (nc,))) if k in d else None
                for k in ("min", "max", "boxcox", "mu", "std")]
        ptrs = [(_lib.dptr(a) if a is not None else nul) for a in keep]
        self._check(self.lib.scfgp_scaler_forward(self.device, _ALGO_ID[self.algo], C.c_void_p(Xd.data_ptr()), n, D,
                                                  cols.ctypes.data_as(C.POINTER(C.c_int)), nc, *ptrs,
                                                  C.c_void_p(out.data_ptr()), nc), "forward")
        return out if was_tensor else out.cpu().numpy()
