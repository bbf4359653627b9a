"""N>1 protocol on CPU: two gloo ranks, rows sharded with shard_bounds, each rank runs the
phase API (NumPy PhaseOracle standing in for the CUDA phases) through the SAME
run_eval_phases driver Engine.eval uses; the all-reduced result must equal the
single-process evaluation."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, shape, out):
    sys.path.insert(0, ROOT)
    from oracle import scfgp_oracle as O
    from scfgp_b200 import synth, shard_bounds, run_eval_phases
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=11)
    lo, hi = shard_bounds(N, rank, world)
    po = O.PhaseOracle(X[lo:hi], y[lo:hi], S, M, n_global=N, chunk=64)
    po.begin(h)

    def allreduce(buf):
        t = torch.from_numpy(buf)       # shares memory: in-place sum like the NCCL path
        dist.all_reduce(t)

    run_eval_phases(po, allreduce, want_grad=True)
    cost, grad, alpha, Li = po.finish()
    if rank == 0:
        np.savez(out, cost=cost, grad=grad, alpha=alpha, Li=Li)
    # every rank must hold identical replicated results
    t = torch.tensor([cost], dtype=torch.float64); lst = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(lst, t)
    assert all(float(x) == float(lst[0]) for x in lst)
    dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(301, 6, 3, 20)])
def test_two_rank_gloo_matches_single(tmp_path, shape):
    from oracle import scfgp_oracle as O
    from scfgp_b200 import synth
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), shape, out), nprocs=2, join=True)
    N, D, S, M = shape
    X, y, h = synth.make_problem(N, D, S, M, seed=11)
    c0, g0, a0, L0 = O.streaming_eval(h, X, y, S, M)
    z = np.load(out)
    assert abs(float(z["cost"]) - c0) <= 1e-12 * abs(c0)
    np.testing.assert_allclose(z["grad"], g0, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(z["alpha"], a0, rtol=1e-8, atol=1e-10)


def _pred_worker(rank, world, port, T):
    sys.path.insert(0, ROOT)
    from scfgp_b200 import predict_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base = torch.arange(T, dtype=torch.float64)

    def local(lo, hi):                      # stand-in for Engine.predict_device on this rank's slice
        return base[lo:hi] * 2.0, base[lo:hi] + 0.5

    def gather(t, n):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return parts

    mu, sd = predict_sharded(local, T, rank, world, gather)
    assert torch.equal(mu, base * 2.0) and torch.equal(sd, base + 0.5)
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [7, 10, 1])
def test_sharded_predict_gathers_in_row_order(T):
    """Batched predict over ranks: uneven shards (and an empty one) come back in row order."""
    mp.spawn(_pred_worker, args=(2, _free_port(), T), nprocs=2, join=True)


class _PhaseEngine:
    """Stand-in for the CUDA Engine on a CPU rank: the oracle's NumPy phases on this rank's row shard,
    driven by the product's own run_eval_phases with a gloo all-reduce (what Engine.eval does over NCCL)."""

    def __init__(self, D, S, M):
        self.D, self.S, self.M = D, S, M
        self.rows = None

    def set_data(self, X, y, n_global=None):
        self.X, self.y, self.n_global = X, np.asarray(y).reshape(-1), n_global
        self.fullX, self.fully, self.full_global = self.X, self.y, n_global
        self.rows = X.shape[0]
        self.batch_rows = []

    def select_rows(self, idx=None, n_global=None):      # Engine.select_rows: LOCAL indices of the resident shard
        if idx is None:
            self.X, self.y, self.n_global = self.fullX, self.fully, self.full_global
        else:
            self.X, self.y, self.n_global = self.fullX[idx], self.fully[idx], n_global
            self.batch_rows.append(len(idx))

    def eval(self, h, want_grad=True, **_):
        from oracle import scfgp_oracle as O
        from scfgp_b200 import run_eval_phases
        po = O.PhaseOracle(self.X, self.y, self.S, self.M, n_global=self.n_global, chunk=64)
        po.begin(h)
        run_eval_phases(po, lambda buf: dist.all_reduce(torch.from_numpy(buf)), want_grad=want_grad)
        return po.finish(want_grad=want_grad)

    def predict(self, h, Xs, a, L):
        from oracle import scfgp_oracle as O
        return O.literal_predict(h, Xs, a, L, self.S, self.M)


def _facade_data():
    rng = np.random.default_rng(3)
    X = rng.lognormal(size=(161, 5))
    y = np.tanh(X @ rng.standard_normal(5) / 4) * 5 + 20 + 0.3 * rng.standard_normal(161)
    return X[:121], y[:121, None], X[121:], y[121:, None]


def _facade_worker(rank, world, port, out, weights=None):
    sys.path.insert(0, ROOT)
    from scfgp_b200 import SCFGP, shard_bounds, weighted_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Xtr, ytr, Xte, yte = _facade_data()
    np.random.seed(100 + rank)                  # different draws per rank: rank 0's must win
    m = SCFGP(sparsity=3, nfeats=12, shard_weights=weights)
    m.set_data(Xtr, ytr)
    m._engine = _PhaseEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, None, None, max_iter=6)
    lo, hi = shard_bounds(len(Xtr), rank, world) if weights is None else weighted_bounds(len(Xtr), rank, weights)
    assert m._engine.rows == hi - lo and m._engine.n_global == len(Xtr)
    mu, sd = m.predict(Xte)
    if rank == 0:
        np.savez(out, params=m.params, cost=np.array(m.evals["COST"][1]), mu=mu, sd=sd)
    t = torch.from_numpy(m.params.copy()); lst = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(lst, t)
    assert all(torch.equal(x, lst[0]) for x in lst)      # replicated optimizer state stays in lock-step
    dist.destroy_process_group()


@pytest.mark.parametrize("weights", [None, [1.0, 2.5]])
def test_facade_fit_shards_rows_over_two_ranks(tmp_path, weights):
    """SCFGP.set_data/optimize/predict under torch.distributed: every rank passes the same training set,
    uploads only its row shard (equal shares, or proportional to ``shard_weights``) and follows the
    single-process fit started from rank 0's parameters."""
    from oracle import scfgp_oracle as O
    from scfgp_b200 import SCFGP

    class OracleEngine:                                  # single-process stand-in (as in test_facade_cpu.py)
        def __init__(self, D, S, M):
            self.D, self.S, self.M = D, S, M

        def set_data(self, X, y, n_global=None):
            self.X, self.y = X, np.asarray(y).reshape(-1)

        def eval(self, h, want_grad=True, **_):
            return O.streaming_eval(h, self.X, self.y, self.S, self.M, want_grad=want_grad)

        def predict(self, h, Xs, a, L):
            return O.literal_predict(h, Xs, a, L, self.S, self.M)

    out = str(tmp_path / "facade.npz")
    mp.spawn(_facade_worker, args=(2, _free_port(), out, weights), nprocs=2, join=True)
    Xtr, ytr, Xte, yte = _facade_data()
    np.random.seed(100)
    m = SCFGP(sparsity=3, nfeats=12)
    m.set_data(Xtr, ytr)
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, None, None, max_iter=6)
    mu, sd = m.predict(Xte)
    z = np.load(out)
    np.testing.assert_allclose(z["cost"], np.array(m.evals["COST"][1]), rtol=1e-10)
    np.testing.assert_allclose(z["params"], m.params, atol=1e-9)
    np.testing.assert_allclose(z["mu"], mu, atol=1e-8)
    np.testing.assert_allclose(z["sd"], sd, atol=1e-8)


def _minibatch_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from scfgp_b200 import SCFGP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Xtr, ytr, Xte, yte = _facade_data()
    np.random.seed(100 + rank)                  # different shuffles per rank: rank 0's permutation must win
    m = SCFGP(sparsity=3, nfeats=12)
    m.set_data(Xtr, ytr)
    m._engine = _PhaseEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, None, None, max_iter=4, nbatches=2, batchsize=40)
    rows = torch.tensor(m._engine.batch_rows)
    lst = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(lst, rows)
    assert torch.equal(sum(lst), torch.full_like(rows, 40))   # every batch is split over the ranks, none lost
    if rank == 0:
        np.savez(out, params=m.params, cost=np.array(m.evals["COST"][1]), alpha=m.alpha)
    dist.destroy_process_group()


def test_facade_minibatch_mode_over_two_ranks(tmp_path):
    """f5 under torch.distributed (SCFGP.py:172-182,226-235): rank 0's permutation is broadcast, every rank
    gathers the rows of the batch that lie in its shard (local indices), and the all-reduced evaluation of a
    batch equals the single-process evaluation of the same batch."""
    from oracle import scfgp_oracle as O
    from scfgp_b200 import SCFGP

    class OracleEngine:
        def __init__(self, D, S, M):
            self.D, self.S, self.M = D, S, M

        def set_data(self, X, y, n_global=None):
            self.X, self.y = X, np.asarray(y).reshape(-1)
            self.fullX, self.fully = self.X, self.y

        def select_rows(self, idx=None, n_global=None):
            self.X, self.y = (self.fullX, self.fully) if idx is None else (self.fullX[idx], self.fully[idx])

        def eval(self, h, want_grad=True, **_):
            return O.streaming_eval(h, self.X, self.y, self.S, self.M, want_grad=want_grad)

        def predict(self, h, Xs, a, L):
            return O.literal_predict(h, Xs, a, L, self.S, self.M)

    out = str(tmp_path / "minibatch.npz")
    mp.spawn(_minibatch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    Xtr, ytr, Xte, yte = _facade_data()
    np.random.seed(100)
    m = SCFGP(sparsity=3, nfeats=12)
    m.set_data(Xtr, ytr)
    m._engine = OracleEngine(m.D, m.S, m.M)
    m.optimize(Xte, yte, None, None, max_iter=4, nbatches=2, batchsize=40)
    z = np.load(out)
    np.testing.assert_allclose(z["cost"], np.array(m.evals["COST"][1]), rtol=1e-10)
    np.testing.assert_allclose(z["params"], m.params, atol=1e-9)
    np.testing.assert_allclose(z["alpha"], m.alpha, atol=1e-8)
