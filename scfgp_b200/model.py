"""``SCFGP`` -- the reference's model facade over the B200 hot path.

Mirrors ``SCFGP/SCFGP.py:21-317`` (constructor, ``set_data``, ``optimize``, ``predict``,
``get_compiled_funcs``, ``save``/``load``) and the README-generation spelling
``SCFGP(rank, feature_size, fftype, msg).fit/predict`` (README.md:44-53,
experiments/kin8nm/kin8nm.py:53-74).  The three "compiled functions" are closures over a
``scfgp_b200.engine.Engine`` (C ABI -> sm_100a kernels); the optimiser and the scalers
stay on the host.  No CPU fallback exists: constructing the functions without the CUDA
library raises.
"""
from __future__ import annotations

import string
import sys
import time

import numpy as np
import numpy.random as npr

from . import _lib
from .engine import Engine
from .optimizer import Optimizer as OPT
from .scaler import Scaler
from .sharding import shard_bounds, weighted_bounds

_PHASE_NAMES = {None: _lib.FF_PHASE, True: _lib.FF_PHASE, "phase": _lib.FF_PHASE, "ph": _lib.FF_PHASE,
                False: _lib.FF_NOPHASE, "non-phase": _lib.FF_NOPHASE, "nonphase": _lib.FF_NOPHASE,
                "no-phase": _lib.FF_NOPHASE, "f": _lib.FF_NOPHASE}


def _new_evals():
    return {
        "SCORE": ["Model Selection Score", []],
        "COST": ["Hyperparameter Selection Cost", []],
        "MAE": ["Mean Absolute Error", []],
        "NMAE": ["Normalized Mean Absolute Error", []],
        "MSE": ["Mean Square Error", []],
        "NMSE": ["Normalized Mean Square Error", []],
        "MNLP": ["Mean Negative Log Probability", []],
        "TIME(s)": ["Training Time", []],
    }


class CompiledFuncs(tuple):
    """(train_func, train_iter_func, pred_func) -- what ``get_compiled_funcs`` returns
    (SCFGP/SCFGP.py:150-151).  Carries the engine so another model of the same shape can
    reuse the handle (``optimize(funcs=...)``, SCFGP/SCFGP.py:209-210)."""
    engine = None


class SCFGP(object):
    """Sparsely Correlated Fourier Features Based Gaussian Process (B200-native hot path)."""

    def __init__(self, rank=None, feature_size=None, fftype=None, msg=None, *, sparsity=None, nfeats=None,
                 evals=None, X_scaling_method='auto-inv-normal', y_scaling_method='auto-normal',
                 verbose=None, device=None, chunk_rows=None, precision=None, scaler_device='cpu',
                 presharded=False, shard_weights=None):
        # G2 spelling: SCFGP(sparsity=20, nfeats=18, ...) (SCFGP.py:33-35);
        # G1 spelling: SCFGP(rank, feature_size, fftype, msg) (README.md:47-50).
        if device is None:             # one process per GPU under torchrun: LOCAL_RANK picks the device
            import os
            device = int(os.environ.get("LOCAL_RANK", 0))
        S = sparsity if sparsity is not None else (rank if rank is not None else 20)
        M = nfeats if nfeats is not None else (feature_size if feature_size is not None else 18)
        self.S_requested = int(S)      # rank <= 0 ("-1" in the G1 scripts): resolved to D in set_data
        self.S = int(S)
        self.M = int(M)
        key = fftype.lower() if isinstance(fftype, str) else fftype
        if key not in _PHASE_NAMES:
            raise ValueError(f"unknown fftype {fftype!r}; use 'phase' or 'non-phase'")
        self.fftype = _PHASE_NAMES[key]
        self.verbose = bool(verbose if verbose is not None else (msg if msg is not None else False))
        # scaler_device='gpu': fit / apply the column scalers with the streaming kernels of
        # gpu_scaler.py (same interface, fields and checkpoint format as the host Scaler); meant for
        # N >= 1e6 rows where the reference's per-column host search dominates set_data
        if scaler_device not in ('cpu', 'gpu'):
            raise ValueError("scaler_device must be 'cpu' or 'gpu'")
        # presharded=True (torch.distributed, one process per GPU): set_data receives THIS rank's rows of a
        # table that is sharded by rows (too large to replicate); the scalers then combine their statistics
        # over the ranks, which needs the streaming GPU scalers
        self.presharded = bool(presharded)
        # shard_weights (torch.distributed, replicated table): one positive number per rank, e.g. the GPUs' measured
        # rows/s -- the row shards are then proportional to them (sharding.weighted_bounds) instead of equal
        self.shard_weights = None if shard_weights is None else [float(w) for w in shard_weights]
        if self.presharded and scaler_device != 'gpu':
            raise ValueError("presharded=True needs scaler_device='gpu' (the scalers all-reduce their statistics)")
        if scaler_device == 'gpu':
            from .gpu_scaler import GpuScaler
            self.X_scaler = GpuScaler(X_scaling_method, device, distributed=self.presharded)
            self.y_scaler = GpuScaler(y_scaling_method, device, distributed=self.presharded)
        else:
            self.X_scaler = Scaler(X_scaling_method)
            self.y_scaler = Scaler(y_scaling_method)
        self.evals = _new_evals() if evals is None else evals
        self.device = device
        self.chunk_rows = chunk_rows
        self.precision = precision     # None/'fp64' (FP64 tensor path) or 'emu'/'emu<digits>' (int8 tensor cores)
        self.N = self.D = -1
        self.X = self.y = self.params = self.alpha = self.Li = None
        self.train_func = self.train_iter_func = self.pred_func = None
        self._engine = None
        self._engine_data_id = None
        self._engine_data_held = None
        self._rule, self._rule_cfg, self._rule_resumed = None, None, False
        self.min_obj_ind = 0
        self.generate_ID()

    # ------------------------------------------------------------------ misc (SCFGP.py:53-62)
    def message(self, *arg):
        if self.verbose:
            print(" ".join(map(str, arg)))
            sys.stdout.flush()

    def generate_ID(self):
        chars = string.ascii_uppercase + string.digits
        self.ID = ''.join(chars[i] for i in npr.randint(0, len(chars), 5))
        self.NAME = "SCFGP (Sparsity=%d, Fourier Features=%d)" % (self.S, self.M)

    # ------------------------------------------------------------------ parameters (SCFGP.py:64-72)
    def init_params(self, rng=None):
        r = npr if rng is None else rng
        randn = (lambda n: r.randn(n)) if rng is None else (lambda n: r.standard_normal(n))
        rand = (lambda n: r.rand(n)) if rng is None else (lambda n: r.random(n))
        a, b, c = randn(1), randn(1), randn(1)
        l_f = randn(self.D * self.S)
        r_f = rand(self.M * self.S)
        l_p = 2 * np.pi * rand(self.S)
        p = 2 * np.pi * rand(self.M)
        self.params = np.concatenate([a, b, c, l_f, r_f, l_p, p])

    # ------------------------------------------------------------------ compiled functions
    def _ensure_engine(self, engine=None):
        if engine is not None:
            if (engine.D, engine.S, engine.M) != (self.D, self.S, self.M):
                raise ValueError("injected funcs were built for a different (D, S, M)")
            self._engine = engine
        if self._engine is None:
            self._engine = Engine(self.D, self.S, self.M, self.fftype, self.device, self.chunk_rows,
                                  precision=self.precision)
        return self._engine

    @staticmethod
    def _is_cuda_tensor(a):
        return hasattr(a, "is_cuda") and bool(a.is_cuda)

    @staticmethod
    def _ranks():
        """(rank, world) of ``torch.distributed`` when it is initialised (one process per GPU), else (0, 1)."""
        try:
            import torch.distributed as td
            if td.is_available() and td.is_initialized():
                return td.get_rank(), td.get_world_size()
        except ImportError:
            pass
        return 0, 1

    def _sync_params(self):
        """Every rank must evaluate the same hyper-parameters: rank 0's draw wins (the reference draws
        from the unseeded global NumPy state, SCFGP.py:65-71)."""
        rank, world = self._ranks()
        if world == 1 or self.params is None:
            return
        import torch
        import torch.distributed as td
        dev = f"cuda:{self.device}" if td.get_backend() == "nccl" else "cpu"
        t = torch.as_tensor(np.ascontiguousarray(self.params, dtype=np.float64)).to(dev)
        td.broadcast(t, src=0)
        self.params = t.cpu().numpy().copy()

    @staticmethod
    def _fingerprint(a):
        """Cheap content fingerprint of a training array: address, layout and a strided sample of the
        values (host arrays) or the in-place version counter (CUDA tensors)."""
        if hasattr(a, "data_ptr"):
            return ("t", a.data_ptr(), tuple(a.shape), tuple(a.stride()), a._version)
        a = np.asarray(a)
        flat = a.reshape(-1)
        sample = flat[::max(1, flat.size // 509)]
        edge = (flat[:64].tobytes(), flat[-64:].tobytes())
        return ("n", a.__array_interface__["data"][0], a.shape, a.strides, sample.tobytes(), edge)

    def _upload(self, X, y):
        """Hand the (scaled) training set to the engine.  Under ``torch.distributed`` every rank holds the
        same X, y and uploads only its own contiguous row shard (``shard_bounds``); the evaluation then
        all-reduces the three partial reductions.

        The reference's Theano functions evaluate whatever arrays they are passed (SCFGP.py:132-137).  The
        device copy is reused only while the SAME array objects (held here, so their ids cannot be
        recycled) still carry the same fingerprint; anything else is uploaded again.
        ``invalidate_data()`` forces the next call to upload."""
        key = (self._fingerprint(X), self._fingerprint(y))
        held = self._engine_data_held
        if held is None or held[0] is not X or held[1] is not y or self._engine_data_id != key:
            held_new = (X, y)
            rank, world = self._ranks()
            n = int(X.shape[0])
            if world > 1 and self.presharded:
                n = int(getattr(self.X_scaler, "n_total", n))      # rows of all ranks (all-reduced by the scaler)
            elif world > 1:
                lo, hi = self._my_rows(n, rank, world)
                X, y = X[lo:hi], y[lo:hi]
            if self._is_cuda_tensor(X):
                self._engine.set_data_device(X, y, n_global=n)   # device-resident training set: zero-copy hand-off
            else:
                self._engine.set_data(np.ascontiguousarray(X), np.ascontiguousarray(y), n_global=n)
            self._engine_data_id = key
            self._engine_data_held = held_new

    def invalidate_data(self):
        """Forget the device copy of the training set: the next train_func / train_iter_func call uploads."""
        self._engine_data_id = None
        self._engine_data_held = None

    def build_models(self, algo='adam', algo_params=None, engine=None):
        """Counterpart of ``build_theano_models`` (SCFGP.py:92-148): binds the three callables."""
        eng = self._ensure_engine(engine)
        self.invalidate_data()
        cfg = (algo, dict(algo_params or {}))
        if getattr(self, "_rule_resumed", False) and self._rule_cfg == cfg:
            self._rule_resumed = False          # checkpointed slots / step counter carry on (load())
        else:
            self._rule, self._rule_cfg = OPT.build(algo, algo_params, nesterov=0.9), cfg

        def train_func(X, y):
            self._upload(X, y)
            cost, _, alpha, Li = eng.eval(self.params, want_grad=False)
            return [np.float64(cost), alpha, Li]

        def train_iter_func(X, y):
            self._upload(X, y)
            cost, grad, alpha, Li = eng.eval(self.params, want_grad=True)
            self.last_grad = grad
            self.params = self._rule.step(self.params, grad)
            return [np.float64(cost), alpha, Li]

        def pred_func(Xs, alpha, Li):
            mu, sd = eng.predict(self.params, Xs, alpha, Li)
            return [mu, sd]

        self.train_func, self.train_iter_func, self.pred_func = train_func, train_iter_func, pred_func

    def get_compiled_funcs(self):
        f = CompiledFuncs((self.train_func, self.train_iter_func, self.pred_func))
        f.engine = self._engine
        return f

    # ------------------------------------------------------------------ data (SCFGP.py:153-170)
    def set_data(self, X, y):
        """X (N, D), y (N,) or (N, 1): NumPy arrays, or -- with ``scaler_device='gpu'`` -- float64 CUDA
        tensors, in which case scaling, the scaled training set and every evaluation stay on the
        device (no host copy of X is ever made; meant for N >= 1e7 rows)."""
        self.message("-" * 60, "\nNormalizing SCFGP training data...")
        if self._is_cuda_tensor(X):
            if not hasattr(self.X_scaler, "device"):
                raise ValueError("CUDA tensors need scaler_device='gpu'")
            import torch
            y = y if self._is_cuda_tensor(y) else torch.as_tensor(np.asarray(y, dtype=np.float64), device=X.device)
            if y.ndim == 1:
                y = y[:, None]
            self.X_scaler.fit(X)
            self.y_scaler.fit(y)
            self.X = self.X_scaler.forward_transform(X).contiguous()
            self.y = self.y_scaler.forward_transform(y).contiguous()
        else:
            X = np.asarray(X, dtype=np.float64)
            y = np.asarray(y, dtype=np.float64)
            if y.ndim == 1:
                y = y[:, None]
            self.X_scaler.fit(X)
            self.y_scaler.fit(y)
            self.X = np.ascontiguousarray(self.X_scaler.forward_transform(X))
            self.y = np.ascontiguousarray(self.y_scaler.forward_transform(y))
        self.message("done.")
        self.N, self.D = (int(v) for v in self.X.shape)
        if self.presharded:
            self.N = int(getattr(self.X_scaler, "n_total", self.N))     # rows of all ranks
        if self.S_requested <= 0:
            self.S = self.D
            self.NAME = "SCFGP (Sparsity=%d, Fourier Features=%d)" % (self.S, self.M)
        if self.train_func is None:
            self.message("-" * 60, "\nInitializing SCFGP hyperparameters...")
            if self.params is None or len(self.params) != 3 + self.D * self.S + self.M * self.S + self.S + self.M:
                self.init_params()
            self._sync_params()
            self.message("done.")
        else:
            cost, self.alpha, self.Li = self.train_func(self.X, self.y)

    def minibatches(self, X, y, batchsize, shuffle=True):
        """The reference's generator (SCFGP.py:172-182): yields host copies (X[batch], y[batch])."""
        assert len(X) == len(y)
        for batch in self._minibatch_indices(len(X), batchsize, shuffle):
            yield np.ascontiguousarray(X[batch]), np.ascontiguousarray(y[batch])

    def _minibatch_indices(self, n, batchsize, shuffle=True):
        """Row indices of the batches of one sweep, drawn exactly as SCFGP.minibatches does (one
        ``np.random.shuffle`` of arange(n) per sweep, SCFGP.py:174-176; trailing rows that do not fill a batch are
        dropped).  Under ``torch.distributed`` rank 0's permutation is broadcast so that all ranks cut the same
        batches."""
        if shuffle:
            inds = np.arange(n)
            np.random.shuffle(inds)
            rank, world = self._ranks()
            if world > 1:
                import torch
                import torch.distributed as td
                dev = f"cuda:{self.device}" if td.get_backend() == "nccl" else "cpu"
                t = torch.as_tensor(inds).to(dev)
                td.broadcast(t, src=0)
                inds = t.cpu().numpy()
        for start_ind in range(0, n - batchsize + 1, batchsize):
            yield inds[start_ind:start_ind + batchsize] if shuffle else np.arange(start_ind, start_ind + batchsize)

    def _row_range(self):
        """[lo, hi) of the global row indices this rank's engine holds."""
        rank, world = self._ranks()
        n = int(self.X.shape[0])
        if world == 1:
            return 0, n
        if self.presharded:                     # every rank passed its own rows: offsets from the row counts
            import torch
            import torch.distributed as td
            dev = f"cuda:{self.device}" if td.get_backend() == "nccl" else "cpu"
            counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
            td.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev))
            lo = int(sum(int(c) for c in counts[:rank]))
            return lo, lo + n
        return self._my_rows(n, rank, world)

    def _my_rows(self, n, rank, world):
        """[lo, hi) of the rows of a replicated table that this rank holds: equal shares, or ``shard_weights``."""
        if self.shard_weights is not None:
            if len(self.shard_weights) != world:
                raise ValueError(f"shard_weights has {len(self.shard_weights)} entries for {world} ranks")
            return weighted_bounds(n, rank, self.shard_weights)
        return shard_bounds(n, rank, world)

    def _train_iter_rows(self, batch):
        """One ``train_iter_func`` step on the rows ``batch`` (global indices) of the resident training set: the
        rows are gathered on the device (``scfgp_select_rows``), nothing is uploaded again."""
        self._upload(self.X, self.y)
        lo, hi = self._batch_range
        local = batch[(batch >= lo) & (batch < hi)] - lo
        self._engine.select_rows(local, n_global=len(batch))
        cost, grad, alpha, Li = self._engine.eval(self.params, want_grad=True)
        self.last_grad = grad
        self.params = self._rule.step(self.params, grad)
        return [np.float64(cost), alpha, Li]

    # ------------------------------------------------------------------ training loop (SCFGP.py:184-276)
    def optimize(self, Xv=None, yv=None, funcs=None, visualizer=None, **args):
        obj = args.get('obj', 'COST').upper()
        obj = obj if obj in self.evals else 'COST'
        algo = args.get('algo', {'algo': None})
        nbatches = args.get('nbatches', 1)
        batchsize = args.get('batchsize', 150)
        cvrg_tol = args.get('cvrg_tol', 1e-4)
        max_cvrg = args.get('max_cvrg', 18)
        max_iter = args.get('max_iter', 500)
        restore_best = args.get('restore_best', False)  # the reference's restore is a no-op
        if not isinstance(algo, dict) or algo.get('algo') not in OPT.algos:
            algo = {'algo': 'adam',
                    'algo_params': {'learning_rate': 0.01, 'beta1': 0.9, 'beta2': 0.999, 'epsilon': 1e-8}}
        for metric in self.evals:
            self.evals[metric][1] = []
        if funcs is None:
            self.message("-" * 50, "\nBinding SCFGP B200 kernels...")
            self.build_models(algo['algo'], algo.get('algo_params'))
            self.message("done.")
        else:
            # Reuse the injected engine; unlike the reference the functions act on
            # *this* model's parameters, not on the first model's captured shared variable.
            self.build_models(algo['algo'], algo.get('algo_params'), engine=getattr(funcs, "engine", None))
        if visualizer is not None:
            raise NotImplementedError("Visualizer (matplotlib live plots) is out of scope of the hot path")
        self._sync_params()
        if Xv is None or yv is None:
            obj = 'COST'
            for m in ('MAE', 'NMAE', 'MSE', 'NMSE', 'MNLP', 'SCORE'):
                self.evals[m][1].append(0)
        self.min_obj_ind = 0
        train_start_time = time.time()
        min_obj_val, argmin_params, cvrg_iter = np.inf, self.params.copy(), 0
        report_every = max(1, max_iter // 10)
        for it in range(max_iter):
            if nbatches > 1:
                # SCFGP.py:226-235 with the batches gathered on the device from the resident training set (host
                # arrays or CUDA tensors, one process or one per GPU).  The reference's per-sweep median of the
                # parameter snapshots is rebound to a variable its compiled functions never read:
                # the effective iterate is the last one, which is what is kept here.
                self._batch_range = self._row_range()
                cost_sum, batch_count = 0, 0
                for batch in self._minibatch_indices(self.N, batchsize):
                    cost, self.alpha, self.Li = self._train_iter_rows(batch)
                    cost_sum += cost
                    batch_count += 1
                    if batch_count == nbatches:
                        break
                self._engine.select_rows(None)
                self.evals['COST'][1].append(np.double(cost_sum / batch_count))
            else:
                cost, self.alpha, self.Li = self.train_iter_func(self.X, self.y)
                self.evals['COST'][1].append(cost)
            self.evals['TIME(s)'][1].append(time.time() - train_start_time)
            if Xv is not None and yv is not None:
                self.predict(Xv, yv)
            if it % report_every == 1:
                self.message("-" * 17, "VALIDATION ITERATION", it, "-" * 17)
                self._print_current_evals()
            obj_val = self.evals[obj][1][-1]
            if obj_val < min_obj_val:
                cvrg_iter = cvrg_iter + 1 if min_obj_val - obj_val < cvrg_tol else 0
                min_obj_val = obj_val
                self.min_obj_ind = len(self.evals['COST'][1]) - 1
                argmin_params = self.params.copy()
            else:
                cvrg_iter += 1
            if it > 30 and cvrg_iter > max_cvrg:
                break
        if restore_best:
            self.params = argmin_params.copy()
        cost, self.alpha, self.Li = self.train_func(self.X, self.y)
        self.evals['COST'][1].append(np.double(cost))
        self.evals['TIME(s)'][1].append(time.time() - train_start_time)
        if Xv is not None and yv is not None:
            self.predict(Xv, yv)
        self.min_obj_ind = len(self.evals['COST'][1]) - 1
        disp, self.verbose = self.verbose, True if args.get('print_result', self.verbose) else False
        self.message("-" * 19, "OPTIMIZATION RESULT", "-" * 20)
        self._print_current_evals()
        self.message("-" * 60)
        self.verbose = disp

    # ------------------------------------------------------------------ prediction (SCFGP.py:278-294)
    def predict(self, Xs, ys=None):
        """``Xs`` (T, D): a NumPy array, or -- with ``scaler_device='gpu'`` -- a float64 CUDA tensor; the test rows
        then stay on the device and, under ``torch.distributed``, are predicted in row shards over the ranks
        (every rank passes the same ``Xs`` and gets the full result).  Returns NumPy (T, 1) mean and std."""
        if self.pred_func is None:
            self.build_models()
        if self._is_cuda_tensor(Xs):
            if not hasattr(self.X_scaler, "device"):
                raise ValueError("CUDA tensors need scaler_device='gpu'")
            self.Xs = self.X_scaler.forward_transform(Xs).contiguous()
            eng = self._engine
            eng.set_factors(self.alpha, self.Li)            # pred_func(Xs, alpha, Li): SCFGP.py:280
            if self._ranks()[1] > 1:
                mu_t, sd_t = eng.predict_sharded(self.params, self.Xs)
            else:
                mu_t, sd_t = eng.predict_device(self.params, self.Xs)
            mu_f, std_f = mu_t.cpu().numpy().reshape(-1, 1), sd_t.cpu().numpy().reshape(-1)
        else:
            self.Xs = np.ascontiguousarray(self.X_scaler.forward_transform(np.asarray(Xs, dtype=np.float64)))
            mu_f, std_f = self.pred_func(self.Xs, self.alpha, self.Li)
        mu_y = self.y_scaler.backward_transform(mu_f)
        up_bnd_y = self.y_scaler.backward_transform(mu_f + std_f[:, None])
        dn_bnd_y = self.y_scaler.backward_transform(mu_f - std_f[:, None])
        std_y = 0.5 * (up_bnd_y - dn_bnd_y)
        if ys is not None:
            ys = np.asarray(ys.cpu().numpy() if self._is_cuda_tensor(ys) else ys, dtype=np.float64)
            if ys.ndim == 1:
                ys = ys[:, None]
            ev = self.evals
            ev['MAE'][1].append(np.mean(np.abs(mu_y - ys)))
            ev['NMAE'][1].append(ev['MAE'][1][-1] / np.std(ys))
            ev['MSE'][1].append(np.mean((mu_y - ys) ** 2.))
            ev['NMSE'][1].append(ev['MSE'][1][-1] / np.var(ys))
            ev['MNLP'][1].append(0.5 * np.mean(((ys - mu_y) / std_y) ** 2 + np.log(2 * np.pi * std_y ** 2)))
            ev['SCORE'][1].append(ev['NMSE'][1][-1] / (1 + np.exp(-ev['MNLP'][1][-1])))
            # README-generation result attributes (experiments/kin8nm/kin8nm.py:67-74)
            self.TsMAE, self.TsMSE = ev['MAE'][1][-1], ev['MSE'][1][-1]
            self.TsRMSE = float(np.sqrt(self.TsMSE))
            self.TsNMSE, self.TsMNLP, self.SCORE = ev['NMSE'][1][-1], ev['MNLP'][1][-1], ev['SCORE'][1][-1]
        return mu_y, std_y

    # ------------------------------------------------------------------ README-generation API
    def fit(self, X_train, y_train, X_test=None, y_test=None, funcs=None, plot=False, plot_training=False,
            **opt):
        """``model.fit(X_train, y_train, {X_test}, {y_test})`` (README.md:51).  ``plot*`` are
        accepted and ignored (matplotlib visualisation is out of scope)."""
        self.set_data(X_train, y_train)
        t0 = time.time()
        self.optimize(X_test, y_test, funcs, None, **opt)
        self.TrTime = time.time() - t0
        self.COST = self.evals['COST'][1][-1]
        return self

    # ------------------------------------------------------------------ checkpoint (SCFGP.py:296-310)
    def save(self, path):
        """``.npz`` checkpoint: unlike the reference's pickle (SCFGP.py:296-302) it also stores S, D, the
        update rule (name, settings, slots, step counter) and the scaler kind, and it contains no compiled
        functions.  Fields that do not exist yet (alpha / Li before the first evaluation) are left out."""
        out = dict(ID=self.ID, M=self.M, S=self.S, D=self.D, fftype=self.fftype,
                   scaler_device='gpu' if hasattr(self.X_scaler, "device") else 'cpu')
        for k in ("params", "alpha", "Li"):
            v = getattr(self, k)
            if v is not None:
                out[k] = np.asarray(v.cpu().numpy() if self._is_cuda_tensor(v) else v, dtype=np.float64)
        for k, v in self.X_scaler.state_dict().items():
            out["xs_" + k] = v
        for k, v in self.y_scaler.state_dict().items():
            out["ys_" + k] = v
        for k, v in self.evals.items():
            out["evals_" + k] = np.asarray(v[1], dtype=np.float64)
        if self._rule is not None:
            out["optcfg_algo"] = self._rule_cfg[0]
            for k, v in (self._rule_cfg[1] or {}).items():
                out["optkw_" + k] = float(v)
            for k, v in self._rule.state_dict().items():
                out["opt_" + k] = v
        with open(path, "wb") as f:
            np.savez_compressed(f, **out)

    def load(self, path):
        with np.load(path, allow_pickle=False) as z:
            d = {k: z[k] for k in z.files}
        self.ID = str(d["ID"])
        self.M, self.S, self.D = int(d["M"]), int(d["S"]), int(d["D"])
        self.S_requested = self.S
        self.fftype = int(d.get("fftype", _lib.FF_PHASE))
        self.params, self.alpha, self.Li = d.get("params"), d.get("alpha"), d.get("Li")
        scaler_cls = Scaler
        if str(d.get("scaler_device", "cpu")) == "gpu":
            from .gpu_scaler import GpuScaler
            scaler_cls = GpuScaler
        self.X_scaler = scaler_cls.from_state_dict({k[3:]: v for k, v in d.items() if k.startswith("xs_")})
        self.y_scaler = scaler_cls.from_state_dict({k[3:]: v for k, v in d.items() if k.startswith("ys_")})
        for k in self.evals:
            if "evals_" + k in d:
                self.evals[k][1] = list(d["evals_" + k])
        self._engine = None
        self.invalidate_data()
        self.train_func = self.train_iter_func = self.pred_func = None
        # the update rule resumes where it stopped: build_models() picks this up instead of a fresh rule
        self._rule, self._rule_cfg, self._rule_resumed = None, None, False
        if "optcfg_algo" in d:
            algo = str(d["optcfg_algo"])
            kw = {k[6:]: float(v) for k, v in d.items() if k.startswith("optkw_")}
            self._rule = OPT.build(algo, kw, nesterov=0.9)
            self._rule_cfg = (algo, kw)
            self._rule.load_state_dict({k[4:]: v for k, v in d.items() if k.startswith("opt_")})
            self._rule_resumed = True
        self.NAME = "SCFGP (Sparsity=%d, Fourier Features=%d)" % (self.S, self.M)
        return self

    def load_reference_golden(self, npz_path):
        """Load the arrays extracted from the reference's shipped pickle
        (tests/golden/make_golden.py) so the shipped Boston model can predict."""
        with np.load(npz_path, allow_pickle=False) as z:
            d = {k: z[k] for k in z.files}
        self.M, self.S, self.D = int(d["M"]), int(d["S"]), int(d["D"])
        self.S_requested = self.S
        self.params, self.alpha, self.Li = d["params"], d["alpha"], d["Li"]
        self.X_scaler = Scaler.from_state_dict({k[3:]: v for k, v in d.items() if k.startswith("xs_")})
        self.y_scaler = Scaler.from_state_dict({k[3:]: v for k, v in d.items() if k.startswith("ys_")})
        self._engine = None
        self.invalidate_data()
        self.train_func = self.train_iter_func = self.pred_func = None
        return self

    def _print_current_evals(self):
        for metric in sorted(self.evals.keys()):
            if len(self.evals[metric][1]) < len(self.evals['COST'][1]):
                continue
            best_perform_eval = self.evals[metric][1][self.min_obj_ind]
            self.message(self.NAME, "%7s = %.4e" % (metric, best_perform_eval))
