"""Host wrapper around the C ABI: one ``Engine`` = one ``scfgp_handle`` on one GPU.

Multi-GPU: one process per GPU (torchrun); training rows are sharded over ranks, each
pass's partial sums are combined with ``torch.distributed.all_reduce`` (NCCL over
NVLink) on the device buffer the phase API exposes, and the small dense phases are
replicated on every rank (identical inputs -> identical factors, no broadcast).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class ScfgpError(RuntimeError):
    pass


class _DevBuf:
    """Expose a raw device pointer to torch via __cuda_array_interface__ (tensor hand-off)."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {
            "shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
            "strides": None,
        }


class Engine:
    def __init__(self, D: int, S: int, M: int, fftype: int = _lib.FF_PHASE, device: int = 0,
                 chunk_rows: int | None = None, precision: str | None = None, pinned_results: bool = False):
        self.lib = _lib.load()
        self.D, self.S, self.M = int(D), int(S), int(M)
        self.J = self.S + self.M
        self.K = 2 * self.J
        self.P = 3 + self.D * self.S + self.M * self.S + self.S + self.M
        self.device = int(device)
        self._h = C.c_void_p()
        rc = self.lib.scfgp_create(self.D, self.S, self.M, int(fftype), self.device, C.byref(self._h))
        if rc != _lib.OK:
            msg = self._err()
            if self._h:
                self.lib.scfgp_destroy(self._h)
                self._h = C.c_void_p()
            raise ScfgpError(f"scfgp_create failed ({rc}): {msg}")
        if chunk_rows:
            self._check(self.lib.scfgp_set_chunk_rows(self._h, int(chunk_rows)))
        if precision is not None:
            self.set_precision(precision)
        self.N = 0
        # pinned_results: eval() returns grad / alpha / Li as views of page-locked buffers owned by the
        # engine (the 8 K^2-byte copy of Li runs at PCIe speed instead of through the driver's bounce
        # buffers); the views are overwritten by the next eval() -- copy them if they must survive it
        self._pinned = {} if pinned_results else None
        self._keep = None   # keeps borrowed torch tensors alive
        self._red = {}

    # ------------------------------------------------------------------ plumbing
    def _err(self) -> str:
        if not self._h:
            return "no handle"
        return (self.lib.scfgp_last_error(self._h) or b"").decode()

    def _check(self, rc: int):
        if rc == _lib.OK:
            return
        msg = self._err()
        if rc == _lib.ERR_NOT_PD:
            # what Theano's Cholesky thunk raises at SCFGP/SCFGP.py:106
            raise np.linalg.LinAlgError(msg)
        if rc == _lib.ERR_ARG:
            raise ValueError(msg or "bad argument")
        if rc == _lib.ERR_NONFINITE:
            raise FloatingPointError(msg)
        raise ScfgpError(f"CUDA error: {msg}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.scfgp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.scfgp_launch_count(self._h))

    def pass_ms(self):
        f = [C.c_float() for _ in range(4)]
        self.lib.scfgp_last_pass_ms(self._h, *[C.byref(x) for x in f])
        return tuple(float(x.value) for x in f)

    def set_precision(self, precision: str):
        """'fp64' (FP64 tensor path) or 'emu' / 'emu<digits>' (float64 emulated on the int8
        tensor cores with 4..8 base-256 digits; 'emu' = 5)."""
        p = str(precision).lower()
        if p == "fp64":
            mode, digits, digits2 = _lib.PREC_FP64, 0, 0
        elif p.startswith("emu"):
            # 'emu5/4': 5 digits for the Gram that feeds the Cholesky, 4 for the other products
            a, _, b = p[3:].partition("/")
            try:
                mode, digits, digits2 = _lib.PREC_EMU_I8, int(a or 0), int(b or 0)
            except ValueError:
                raise ValueError(f"unknown precision {precision!r}") from None
        else:
            raise ValueError(f"unknown precision {precision!r}")
        self._check(self.lib.scfgp_set_precision2(self._h, mode, digits, digits2))

    @property
    def precision(self) -> str:
        m, d, d2 = C.c_int(), C.c_int(), C.c_int()
        self.lib.scfgp_get_precision2(self._h, C.byref(m), C.byref(d), C.byref(d2))
        if m.value == _lib.PREC_FP64:
            return "fp64"
        return f"emu{d.value}" if d2.value == d.value else f"emu{d.value}/{d2.value}"

    def set_phi_cache(self, on: bool = True, max_gb: float = 0.0):
        """Digit cache of Phi for all rows (emulated mode; see include/scfgp_b200.h): allow / forbid it and bound
        its size; ``phi_cache_bytes`` tells what the last evaluation actually holds (0 = no cache)."""
        self._check(self.lib.scfgp_set_phi_cache(self._h, int(bool(on)), float(max_gb)))

    @property
    def phi_cache_bytes(self) -> int:
        return int(self.lib.scfgp_phi_cache_bytes(self._h))

    def set_guard(self, on: bool = True, cost_tol: float = 0.0, alpha_tol: float = 0.0):
        """A-posteriori guard of the emulated arithmetic (include/scfgp_b200.h, scfgp_set_guard): thresholds on the
        estimated error of `cost` (absolute) and of alpha (relative to max(1, |alpha|_inf)); 0 keeps the default."""
        self._check(self.lib.scfgp_set_guard(self._h, int(bool(on)), float(cost_tol), float(alpha_tol)))

    def guard_info(self) -> dict:
        """Estimates of the last factor phase and the escalations this engine has made so far."""
        ec, ea, fro = C.c_double(), C.c_double(), C.c_double()
        rows, exc, esc = C.c_longlong(), C.c_int(), C.c_int()
        self.lib.scfgp_guard_last(self._h, C.byref(ec), C.byref(ea), C.byref(fro), C.byref(rows), C.byref(exc),
                                  C.byref(esc))
        return {"est_cost": ec.value, "est_alpha": ea.value, "dG_fro_sample": fro.value, "sample_rows": rows.value,
                "exceeded": bool(exc.value), "escalations": esc.value, "precision": self.precision}

    def set_streams(self, n: int):
        self._check(self.lib.scfgp_set_streams(self._h, int(n)))

    def set_tracing(self, on: bool):
        self._check(self.lib.scfgp_set_tracing(self._h, int(bool(on))))

    def trace_dump(self, path: str):
        """Write the launch trace recorded since set_tracing(True) as Chrome-trace JSON."""
        self._check(self.lib.scfgp_trace_dump(self._h, str(path).encode()))

    def set_profiling(self, family: int):
        self._check(self.lib.scfgp_set_profiling(self._h, int(family)))

    def get_profile(self):
        """(total ms, launches) of the profiled kernel family since set_profiling."""
        ms, n = C.c_double(), C.c_longlong()
        self.lib.scfgp_get_profile(self._h, C.byref(ms), C.byref(n))
        return float(ms.value), int(n.value)

    def set_stream(self, stream_ptr: int):
        self._check(self.lib.scfgp_set_stream(self._h, C.c_void_p(stream_ptr)))

    # ------------------------------------------------------------------ data
    def set_data(self, X, y, n_global: int | None = None, overlap: bool = False):
        """X (N,D), y (N,) or (N,1) host arrays -> device copies owned by the handle.

        ``overlap=True`` returns at once; the upload runs in row blocks during the first pass of the next
        ``eval``, which consumes the blocks as they land (use pinned host memory; the copies are queued behind
        that evaluation's parameter upload).  The arrays are kept alive by the engine and must not be modified
        until that ``eval`` returns."""
        X = _lib.as_f64(X)
        y = _lib.as_f64(np.asarray(y).reshape(-1))
        if X.ndim != 2 or X.shape[1] != self.D or y.shape[0] != X.shape[0]:
            raise ValueError(f"shape mismatch: X {X.shape}, y {y.shape}, D={self.D}")
        self.N = X.shape[0]
        fn = self.lib.scfgp_set_data_host_async if overlap else self.lib.scfgp_set_data_host
        self._check(fn(self._h, _lib.dptr(X), _lib.dptr(y), self.N, int(n_global or self.N)))
        self._keep = (X, y) if overlap else None

    def set_data_device(self, X_t, y_t, n_global: int | None = None):
        """Borrow float64 CUDA torch tensors (zero-copy)."""
        import torch
        assert X_t.is_cuda and y_t.is_cuda and X_t.dtype == torch.float64 and y_t.dtype == torch.float64
        X_t = X_t.contiguous()
        y_t = y_t.contiguous().reshape(-1)
        if X_t.ndim != 2 or X_t.shape[1] != self.D or y_t.shape[0] != X_t.shape[0]:
            raise ValueError("shape mismatch")
        self.N = int(X_t.shape[0])
        self._keep = (X_t, y_t)
        self._check(self.lib.scfgp_set_data_device(self._h, C.c_void_p(X_t.data_ptr()),
                                                   C.c_void_p(y_t.data_ptr()), self.N, int(n_global or self.N)))

    def select_rows(self, idx=None, n_global: int | None = None):
        """Minibatch mode (SCFGP.py:172-182): evaluate on the rows ``idx`` (LOCAL indices into this engine's resident
        training set; gathered on the device, nothing is uploaded again).  ``None`` returns to all rows."""
        if idx is None:
            self._check(self.lib.scfgp_select_rows(self._h, None, 0, 0))
            return
        idx = np.ascontiguousarray(idx, dtype=np.int64).reshape(-1)
        buf = idx if idx.size else np.zeros(1, dtype=np.int64)
        self._check(self.lib.scfgp_select_rows(self._h, buf.ctypes.data_as(C.POINTER(C.c_longlong)), idx.size,
                                               int(n_global or idx.size)))

    # ------------------------------------------------------------------ evaluation
    def _reduce_tensor(self, which: int):
        import torch
        if which not in self._red:
            p, n = C.c_void_p(), C.c_longlong()
            self._check(self.lib.scfgp_reduce_buffer(self._h, which, C.byref(p), C.byref(n)))
            self._red[which] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{self.device}")
        return self._red[which]

    def eval(self, params, want_grad: bool = True, want_factors: bool = True, group=None):
        """cost, grad (P,) or None, alpha (K,1) or None, Li (K,K) or None.

        If ``torch.distributed`` is initialised with world_size > 1 the three partial
        reductions are all-reduced between phases."""
        params = _lib.as_f64(params).reshape(-1)
        if params.shape[0] != self.P:
            raise ValueError(f"params has {params.shape[0]} entries, expected {self.P}")
        dist = None
        try:
            import torch.distributed as td
            if td.is_available() and td.is_initialized() and td.get_world_size(group) > 1:
                dist = td
        except ImportError:
            pass
        cost = np.zeros(1)
        if self._pinned is not None:
            if not self._pinned:
                import torch
                for name, shape in (("grad", (self.P,)), ("alpha", (self.K,)), ("Li", (self.K, self.K))):
                    self._pinned[name] = torch.zeros(shape, dtype=torch.float64, pin_memory=True).numpy()
            grad = self._pinned["grad"] if want_grad else None
            alpha = self._pinned["alpha"] if want_factors else None
            Li = self._pinned["Li"] if want_factors else None
        else:
            grad = np.zeros(self.P) if want_grad else None
            alpha = np.zeros(self.K) if want_factors else None
            Li = np.zeros((self.K, self.K)) if want_factors else None
        nul = C.POINTER(C.c_double)()
        gp = _lib.dptr(grad) if want_grad else nul
        ap = _lib.dptr(alpha) if want_factors else nul
        lp = _lib.dptr(Li) if want_factors else nul
        if dist is None:
            self._check(self.lib.scfgp_eval(self._h, _lib.dptr(params), _lib.dptr(cost), gp, ap, lp))
        else:
            import torch
            from .sharding import run_eval_phases
            lib, h, eng = self.lib, self._h, self

            class _Api:
                def pass1(s): eng._check(lib.scfgp_eval_pass1(h))
                def factor(s):
                    eng._check(lib.scfgp_eval_factor(h))
                    eng._check(lib.scfgp_eval_prefetch_factors(h, ap, lp))   # alpha / Li leave during passes 2-3
                def pass2(s): eng._check(lib.scfgp_eval_pass2(h))
                def backward(s): eng._check(lib.scfgp_eval_backward(h))
                def pass3(s): eng._check(lib.scfgp_eval_pass3(h))
                def reduce_buffer(s, which): return eng._reduce_tensor(which)

                def needs_more_precision(s):
                    # every rank measured its own row sample: the decision is taken on the maximum over ranks
                    g = eng.guard_info()
                    t = torch.tensor([1.0 if g["exceeded"] else 0.0], dtype=torch.float64, device=f"cuda:{eng.device}")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
                    if float(t[0]) == 0.0:
                        return False
                    eng._red.clear()                                  # (buffers may be re-allocated)
                    eng._check(lib.scfgp_guard_escalate(h))
                    eng._check(lib.scfgp_eval_begin(h, _lib.dptr(params), int(want_grad)))
                    return True

            def _allreduce(buf):
                # stream-ordered, no host synchronisation: torch's current stream (which NCCL orders its
                # own stream against) waits for the pass, the handle's stream waits for the reduction
                ts = C.c_void_p(torch.cuda.current_stream(eng.device).cuda_stream)
                eng._check(lib.scfgp_stream_handoff(h, ts, 0))
                dist.all_reduce(buf, group=group)
                eng._check(lib.scfgp_stream_handoff(h, ts, 1))

            self._check(lib.scfgp_eval_begin(h, _lib.dptr(params), int(want_grad)))
            run_eval_phases(_Api(), _allreduce, want_grad)
            self._check(lib.scfgp_eval_finish(h, _lib.dptr(cost), gp, ap, lp))
        return (float(cost[0]), grad, alpha.reshape(-1, 1) if want_factors else None, Li)

    def predict(self, params, Xs, alpha=None, Li=None):
        """mu_pred (T,1), std_pred (T,)  -- the shapes pred_func returns (SCFGP.py:145-148)."""
        params = _lib.as_f64(params).reshape(-1)
        Xs = _lib.as_f64(Xs)
        if Xs.ndim != 2 or Xs.shape[1] != self.D:
            raise ValueError(f"Xs has shape {Xs.shape}, expected (T,{self.D})")
        T = Xs.shape[0]
        mu, sd = np.zeros(T), np.zeros(T)
        nul = C.POINTER(C.c_double)()
        if alpha is not None:
            a = _lib.as_f64(alpha).reshape(-1)
            L = _lib.as_f64(Li)
            if a.shape[0] != self.K or L.shape != (self.K, self.K):
                raise ValueError("alpha/Li shape mismatch")
            ap, lp = _lib.dptr(a), _lib.dptr(L)
        else:
            ap, lp = nul, nul
        self._check(self.lib.scfgp_predict(self._h, _lib.dptr(params), _lib.dptr(Xs), T, ap, lp,
                                           _lib.dptr(mu), _lib.dptr(sd)))
        return mu.reshape(-1, 1), sd

    def set_factors(self, alpha, Li):
        """Make host ``alpha`` (K,) / (K,1) and ``Li`` (K,K) the device-resident factors that ``predict_device`` /
        ``predict_sharded`` / ``predict(..., None, None)`` use (the (alpha, Li) arguments of the reference's
        pred_func, SCFGP/SCFGP.py:145-148)."""
        alpha = _lib.as_f64(np.asarray(alpha).reshape(-1))
        Li = _lib.as_f64(Li)
        if alpha.shape[0] != self.K or Li.shape != (self.K, self.K):
            raise ValueError(f"alpha/Li must have K={self.K} rows")
        self._check(self.lib.scfgp_set_factors(self._h, _lib.dptr(alpha), _lib.dptr(Li)))

    def predict_device(self, params, Xs_t):
        """Device-resident batched predict: returns (mu, std) CUDA tensors of length T."""
        import torch
        params = _lib.as_f64(params).reshape(-1)
        Xs_t = Xs_t.contiguous()
        T = int(Xs_t.shape[0])
        mu = torch.empty(T, dtype=torch.float64, device=Xs_t.device)
        sd = torch.empty(T, dtype=torch.float64, device=Xs_t.device)
        self._check(self.lib.scfgp_predict_device(self._h, _lib.dptr(params), C.c_void_p(Xs_t.data_ptr()), T,
                                                  C.c_void_p(mu.data_ptr()), C.c_void_p(sd.data_ptr())))
        return mu, sd

    def predict_sharded(self, params, Xs_t, group=None):
        """Batched prediction over all ranks of ``torch.distributed``: the rows of ``Xs_t`` (a CUDA
        tensor every rank holds, or at least its own shard of) are split with ``shard_bounds``, each
        rank predicts its slice on its GPU and the results are all-gathered (no other collective).
        Needs the factors of the last evaluation on every rank (they are replicated bit for bit)."""
        import torch
        import torch.distributed as td
        from .sharding import predict_sharded
        world = td.get_world_size(group) if td.is_initialized() else 1
        rank = td.get_rank(group) if td.is_initialized() else 0

        def local(lo, hi):
            if hi <= lo:
                e = torch.empty(0, dtype=torch.float64, device=Xs_t.device)
                return e, e
            return self.predict_device(params, Xs_t[lo:hi])

        def gather(t, n):
            if world == 1:
                return [t]
            parts = [torch.empty_like(t) for _ in range(world)]
            td.all_gather(parts, t, group=group)
            return parts

        return predict_sharded(local, int(Xs_t.shape[0]), rank, world, gather)

    def features(self, params, X):
        """Debug hook: materialise Phi (N,K) through the feature kernel."""
        params = _lib.as_f64(params).reshape(-1)
        X = _lib.as_f64(X)
        Phi = np.zeros((X.shape[0], self.K))
        self._check(self.lib.scfgp_dbg_features(self._h, _lib.dptr(params), _lib.dptr(X), X.shape[0],
                                                _lib.dptr(Phi)))
        return Phi


def dbg_gemm(layout: int, A, B, device: int = 0):
    lib = _lib.load()
    A, B = _lib.as_f64(A), _lib.as_f64(B)
    if layout % 10 == 0:
        m, k = A.shape; n = B.shape[0]
    elif layout % 10 == 1:
        k, m = A.shape; n = B.shape[1]
    else:
        m, k = A.shape; n = B.shape[1]
    Cm = np.zeros((m, n))
    rc = lib.scfgp_dbg_gemm(device, layout, m, n, k, _lib.dptr(A), _lib.dptr(B), _lib.dptr(Cm))
    if rc:
        raise ScfgpError(f"scfgp_dbg_gemm rc={rc}")
    return Cm


def dbg_factor(A, device: int = 0):
    lib = _lib.load()
    A = _lib.as_f64(A)
    K = A.shape[0]
    L, Li = np.zeros((K, K)), np.zeros((K, K))
    info = C.c_longlong(0)
    rc = lib.scfgp_dbg_factor(device, K, _lib.dptr(A), _lib.dptr(L), _lib.dptr(Li), C.byref(info))
    if rc:
        raise ScfgpError(f"scfgp_dbg_factor rc={rc}")
    return L, Li, int(info.value)


def dbg_emu_gemm(A, B, digits: int = 6, device: int = 0):
    """C = A . B^T through digit extraction + the int8 tensor-core mainloop."""
    lib = _lib.load()
    A, B = _lib.as_f64(A), _lib.as_f64(B)
    m, k = A.shape
    n = B.shape[0]
    Cm = np.zeros((m, n))
    rc = lib.scfgp_dbg_emu_gemm(device, int(digits), m, n, k, _lib.dptr(A), _lib.dptr(B), _lib.dptr(Cm))
    if rc:
        raise ScfgpError(f"scfgp_dbg_emu_gemm rc={rc}")
    return Cm
