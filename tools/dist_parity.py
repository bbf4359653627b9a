"""Multi-GPU parity (torchrun, one rank per GPU): the row-sharded, all-reduced evaluation must
equal the single-GPU evaluation of the same problem.  Rank 0 prints the differences.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_parity.py
"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scfgp_b200 import Engine, SCFGP, synth, shard_bounds

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
out = []
for (N, D, S, M) in [(5001, 8, 4, 100), (20000, 64, 16, 1024)]:
    X, y, h = synth.make_problem(N, D, S, M, seed=21)
    lo, hi = shard_bounds(N, rank, world)
    eng = Engine(D, S, M, device=lr)
    eng.set_data(X[lo:hi], y[lo:hi], n_global=N)
    c, g, a, L = eng.eval(h)
    # batched predict over ranks (rows of Xs sharded, results all-gathered) against this rank's own full predict
    Xs = torch.rand((1001, D), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    mu_s, sd_s = eng.predict_sharded(h, Xs)
    mu_f, sd_f = eng.predict_device(h, Xs)
    assert float((mu_s - mu_f).abs().max()) == 0.0 and float((sd_s - sd_f).abs().max()) == 0.0, "sharded predict differs"
    cs = torch.tensor([c], dtype=torch.float64, device="cuda"); lst = [torch.zeros_like(cs) for _ in range(world)]
    dist.all_gather(lst, cs)
    same = all(float(x) == float(lst[0]) for x in lst)
    dist.barrier()
    eng.close()
    out.append((N, D, S, M, c, g, a, L, same))


def facade_fit():
    """SCFGP.fit through the facade: the same call on every rank (rows sharded inside) or on one GPU."""
    rng = np.random.default_rng(4)
    X = rng.lognormal(size=(3000, 8))
    y = (np.tanh(X @ rng.standard_normal(8) / 4) * 5 + 20 + 0.3 * rng.standard_normal(3000))[:, None]
    np.random.seed(17 + rank)                         # rank 0's initial parameters are broadcast
    m = SCFGP(rank=4, feature_size=100, fftype="phase", msg=False, device=lr)
    m.fit(X[:2500], y[:2500], X[2500:], y[2500:], max_iter=15)
    mu, sd = m.predict(X[2500:])
    return np.array(m.evals["COST"][1]), m.params.copy(), mu, sd


def facade_fit_presharded(shard):
    """Table sharded by rows BEFORE the facade sees it (presharded=True, GPU scalers all-reduce their statistics);
    shard=False: the whole table in one process."""
    rng = np.random.default_rng(6)
    X = rng.lognormal(size=(4001, 6)); X[:, 2] = np.round(X[:, 2])[:] % 4          # one column with < 10 distinct values
    y = (np.tanh(X @ rng.standard_normal(6) / 4) * 5 + 20 + 0.3 * rng.standard_normal(4001))[:, None]
    Xte, yte = X[3500:], y[3500:]
    lo, hi = shard_bounds(3500, rank, world) if shard else (0, 3500)
    np.random.seed(23 + rank)
    m = SCFGP(rank=3, feature_size=60, fftype="phase", msg=False, device=lr, scaler_device="gpu", presharded=shard)
    m.fit(X[lo:hi], y[lo:hi], Xte, yte, max_iter=10)
    mu, sd = m.predict(Xte)
    return (np.array(m.evals["COST"][1]), np.asarray(m.X_scaler.data["boxcox"]), np.asarray(m.X_scaler.data["mu"]),
            np.asarray(mu), np.asarray(sd), m.N)


def facade_fit_minibatch():
    """nbatches > 1 under torch.distributed (row f5): rank 0's permutation is broadcast, every rank gathers the
    rows of the batch that live in its shard on the device (scfgp_select_rows), the partial sums are all-reduced."""
    rng = np.random.default_rng(8)
    X = rng.lognormal(size=(2400, 6))
    y = (np.sin(X.sum(1) / 3) + 0.1 * rng.standard_normal(2400))[:, None]
    np.random.seed(31)                                # same seed everywhere: the single-GPU rerun draws the same batches
    m = SCFGP(rank=3, feature_size=40, fftype="phase", msg=False, device=lr)
    m.fit(X, y, None, None, max_iter=12, nbatches=3, batchsize=200)
    return np.array(m.evals["COST"][1]), m.params.copy()


fit_sharded = facade_fit()
fit_pre = facade_fit_presharded(True)
fit_mb = facade_fit_minibatch()
dist.barrier()
if rank == 0:
    # now evaluate the same problems on one GPU with no process group in play
    dist.destroy_process_group()
    for (N, D, S, M, c, g, a, L, same) in out:
        X, y, h = synth.make_problem(N, D, S, M, seed=21)
        e1 = Engine(D, S, M, device=lr); e1.set_data(X, y)
        c1, g1, a1, L1 = e1.eval(h)
        print(json.dumps({"shape": [N, D, S, M], "world": world, "ranks_identical": same,
                          "cost_rel": abs(c - c1) / abs(c1), "grad_rel": float(np.linalg.norm(g - g1) / np.linalg.norm(g1)),
                          "alpha_abs": float(np.abs(a - a1).max()), "Li_abs": float(np.abs(L - L1).max())}), flush=True)
        e1.close()
    c1, p1, mu1, sd1 = facade_fit()
    c0, p0, mu0, sd0 = fit_sharded
    print(json.dumps({"facade_fit": "SCFGP.fit, 2500 rows, 15 iterations", "world": world,
                      "cost_traj_rel": float(np.abs(c0 - c1).max() / np.abs(c1).max()),
                      "params_abs": float(np.abs(p0 - p1).max()), "mu_abs": float(np.abs(mu0 - mu1).max()),
                      "sd_abs": float(np.abs(sd0 - sd1).max())}), flush=True)
    c1, l1, m1, mu1, sd1, n1 = facade_fit_presharded(False)
    c0, l0, m0, mu0, sd0, n0 = fit_pre
    print(json.dumps({"facade_fit_presharded": "SCFGP(scaler_device='gpu', presharded=True).fit, 3500 rows over the ranks",
                      "world": world, "rows_seen": [n0, n1], "boxcox_lambda_abs": float(np.abs(l0 - l1).max()),
                      "scaler_mu_abs": float(np.abs(m0 - m1).max()),
                      "cost_traj_rel": float(np.abs(c0 - c1).max() / np.abs(c1).max()),
                      "mu_abs": float(np.abs(mu0 - mu1).max()), "sd_abs": float(np.abs(sd0 - sd1).max())}), flush=True)
    c1, p1 = facade_fit_minibatch()
    c0, p0 = fit_mb
    print(json.dumps({"facade_fit_minibatch": "SCFGP.fit(nbatches=3, batchsize=200), 2400 rows over the ranks", "world": world,
                      "cost_traj_rel": float(np.abs(c0 - c1).max() / np.abs(c1).max()),
                      "params_abs": float(np.abs(p0 - p1).max())}), flush=True)
else:
    dist.destroy_process_group()
