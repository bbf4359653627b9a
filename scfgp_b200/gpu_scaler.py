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

``distributed=True`` (one process per GPU under ``torch.distributed``): every rank passes ITS OWN rows of a table
that is sharded by rows; minima / maxima, the power sums behind every moment and the distinct-value sets are
combined with small all-reduces / all-gathers, so all ranks fit the same parameters as a single process would
on the whole table (the search runs replicated, in lock-step).  Meant for tables that do not fit one GPU
(cfg5: N = 1e8, D = 256 is 205 GB).

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

    def __init__(self, algo, device: int = 0, distributed: bool = False, group=None):
        super().__init__(algo)
        self.device = int(device)
        self.lib = _lib.load()
        self.objective_evals = 0
        self.distributed = bool(distributed)
        self.group = group

    # ------------------------------------------------------------------ rows sharded over ranks
    def _td(self):
        if not self.distributed:
            return None
        import torch.distributed as td
        if td.is_available() and td.is_initialized() and td.get_world_size(self.group) > 1:
            return td
        return None

    def _allreduce(self, values, op="sum"):
        """All-reduce a few host doubles over the ranks (identity in a single process)."""
        v = np.atleast_1d(np.asarray(values, dtype=np.float64)).copy()
        td = self._td()
        if td is None:
            return v
        import torch
        dev = f"cuda:{self.device}" if td.get_backend(self.group) == "nccl" else "cpu"
        t = torch.as_tensor(v).to(dev)
        td.all_reduce(t, op={"sum": td.ReduceOp.SUM, "min": td.ReduceOp.MIN, "max": td.ReduceOp.MAX}[op],
                      group=self.group)
        return t.cpu().numpy()

    def _sums(self, col_t, n, lm, central, centre):
        a, b = C.c_double(), C.c_double()
        self._check(self.lib.scfgp_scaler_column_sums(self.device, C.c_void_p(col_t.data_ptr()), n, float(lm),
                                                      int(central), float(centre), C.byref(a), C.byref(b)),
                    "column_sums")
        return a.value, b.value

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
        if self._td() is not None:           # this rank's power sums, all-reduced; n_total set by fit()
            s1, _ = self._sums(col_t, n, lm, 0, 0.0)
            mean = float(self._allreduce([s1])[0]) / self.n_total
            a, b = self._sums(col_t, n, lm, 1, mean)
            a, b = self._allreduce([a, b])
            m2, m3 = a / self.n_total, b / self.n_total
            return mean, m2, (m3 / (m2 * np.sqrt(m2)) if m2 > 0 else 0.0)
        mean, var, g1 = C.c_double(), C.c_double(), C.c_double()
        self._check(self.lib.scfgp_scaler_column_moments(self.device, C.c_void_p(col_t.data_ptr()), n, float(lm),
                                                         C.byref(mean), C.byref(var), C.byref(g1)), "column_moments")
        return mean.value, var.value, g1.value

    @staticmethod
    def _unbiased_skew(g1, n):
        # scipy.stats.skew(..., bias=False) (Scaler.py:64): G1 = sqrt(n (n - 1)) / (n - 2) * g1
        return np.sqrt(n * (n - 1.0)) / (n - 2.0) * g1 if n > 2 else g1

    def _distinct(self, col_t, n):
        """min(number of distinct values of the column, 10) -- over all ranks when distributed."""
        nd = C.c_int()
        vals = np.zeros(9)
        self._check(self.lib.scfgp_scaler_column_distinct_values(self.device, C.c_void_p(col_t.data_ptr()), n, 9,
                                                                 C.byref(nd), _lib.dptr(vals)), "column_distinct")
        td = self._td()
        if td is None:
            return nd.value
        import torch
        dev = f"cuda:{self.device}" if td.get_backend(self.group) == "nccl" else "cpu"
        mine = torch.as_tensor(np.concatenate(([float(nd.value)], vals))).to(dev)
        parts = [torch.empty_like(mine) for _ in range(td.get_world_size(self.group))]
        td.all_gather(parts, mine, group=self.group)
        union = set()
        for p in parts:
            p = p.cpu().numpy()
            k = int(p[0])
            if k >= 10:
                return 10
            union.update(p[1:1 + k].tolist())
        return min(len(union), 10)

    # ------------------------------------------------------------------ fit
    def fit(self, X):
        import torch
        Xd, _ = self._to_device(X)
        n, D = int(Xd.shape[0]), int(Xd.shape[1])
        mn, mx = np.zeros(D), np.zeros(D)
        const = (C.c_int * D)()
        self._check(self.lib.scfgp_scaler_colstats(self.device, C.c_void_p(Xd.data_ptr()), n, D, D, _lib.dptr(mn),
                                                   _lib.dptr(mx), const), "colstats")
        self.n_total = n
        if self._td() is not None:
            self.n_total = int(round(float(self._allreduce([n])[0])))
            mn, mx = self._allreduce(mn, "min"), self._allreduce(mx, "max")
            cols = [d for d in range(D) if mn[d] != mx[d]]
        else:
            cols = [d for d in range(D) if not const[d]]
        nt = self.n_total
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
            if self._distinct(col_t, n) < 10:
                lms[j] = 1.0
            else:
                def obj(t):
                    self.objective_evals += 1
                    _, _, g1 = self._moments(col_t, n, _softplus(t[0]))
                    return self._unbiased_skew(g1, nt) ** 2
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
