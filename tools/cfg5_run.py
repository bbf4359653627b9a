"""BASELINE.json configs[4] at full size: synthetic N=100M, D=256, rank=32, feature_size=2048,
fit (a few optimizer steps) + batched predict, rows sharded over the ranks of torchrun.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/cfg5_run.py [--rows 100000000] [--iters 3] [--test-rows 2000000] [--fp64]

Every rank generates its own row shard on its GPU (25.6 GB of X per rank at 8 ranks), so nothing of
size N ever lives on the host.  Rank 0 prints one JSON line per stage.  Size-independent checks:
the ranks hold bit-identical costs/factors, the cost decreases along the optimizer steps, the sharded
predict equals this rank's own predict of the same rows, and (``--fp64``) the emulated evaluation
matches one evaluation of the FP64 DMMA path on the same data within north_star's tolerances.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_device_problem  # noqa: E402
from scfgp_b200 import Engine, Optimizer, shard_bounds  # noqa: E402


def facade(a, X, y, N, D, S, M, rank, world, lr, say, sync):
    """The same job through the reference-shaped API: every rank hands ITS rows to SCFGP.fit (the GPU scalers
    combine their statistics over the ranks), validation and test rows are CUDA tensors every rank holds."""
    from scfgp_b200 import SCFGP
    Xv, yv = X[:8192].clone(), y[:8192].clone()          # rank 0's first rows serve as the validation set
    if world > 1:
        dist.broadcast(Xv, src=0); dist.broadcast(yv, src=0)
    m = SCFGP(rank=S, feature_size=M, fftype="phase", msg=False, device=lr, precision=a.precision,
              scaler_device="gpu", presharded=world > 1)
    sync(); t0 = time.time()
    m.set_data(X, y)
    sync(); t_scale = time.time() - t0
    say(stage="facade.set_data", seconds=round(t_scale, 2), rows_total=m.N,
        boxcox_objective_evals=m.X_scaler.objective_evals + m.y_scaler.objective_evals)
    del X, y
    sync(); t0 = time.time()
    m.optimize(Xv, yv, None, None, max_iter=a.iters)
    sync(); t_fit = time.time() - t0
    costs = [float(c) for c in m.evals["COST"][1]]
    say(stage="facade.optimize", iters=a.iters, seconds=round(t_fit, 2), costs=costs, nmse=m.evals["NMSE"][1],
        mnlp=m.evals["MNLP"][1], decreasing=bool(costs[-1] < costs[0]))
    Xs = torch.rand((a.test_rows, D), dtype=torch.float64, device="cuda",
                    generator=torch.Generator(device="cuda").manual_seed(5))
    sync(); t0 = time.time()
    mu, sd = m.predict(Xs)
    sync(); dt = time.time() - t0
    chk = torch.tensor([float(mu.sum()), float(sd.sum())], dtype=torch.float64, device="cuda")
    lst = [torch.zeros_like(chk) for _ in range(world)]
    if world > 1:
        dist.all_gather(lst, chk)
    say(stage="facade.predict", test_rows=a.test_rows, seconds=round(dt, 3), rows_per_sec=round(a.test_rows / dt, 1),
        shapes=[list(mu.shape), list(sd.shape)], finite=bool(np.isfinite(mu).all() and np.isfinite(sd).all()),
        ranks_identical=all(bool((x == chk).all()) for x in lst) if world > 1 else True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--dims", type=int, nargs=3, default=[256, 32, 2048], metavar=("D", "S", "M"))
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--test-rows", type=int, default=2_000_000)
    ap.add_argument("--precision", default="emu5")
    ap.add_argument("--fp64", action="store_true", help="also evaluate once on the FP64 path and report parity")
    ap.add_argument("--facade", action="store_true",
                    help="run through SCFGP(scaler_device='gpu', presharded=True).fit / .predict instead of Engine")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    N = a.rows; D, S, M = a.dims
    lo, hi = shard_bounds(N, rank, world)

    def say(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    t0 = time.time()
    X, y, h = make_device_problem(hi - lo, D, S, M, 1000 + rank, f"cuda:{lr}", dist if world > 1 else None)
    sync()
    say(stage="data", rows=N, D=D, S=S, M=M, K=2 * (S + M), world=world, rows_per_rank=hi - lo,
        x_gb_per_rank=round((hi - lo) * D * 8 / 1e9, 2), seconds=round(time.time() - t0, 2))

    if a.facade:
        facade(a, X, y, N, D, S, M, rank, world, lr, say, sync)
        if world > 1:
            dist.destroy_process_group()
        return

    eng = Engine(D, S, M, device=lr, precision=a.precision)
    eng.set_data_device(X, y, n_global=N)
    opt = Optimizer.build("adam", {"learning_rate": 0.01})
    costs = []
    for it in range(a.iters + 1):                       # the first evaluation is the warm-up of every buffer
        sync(); t0 = time.time()
        c, g, al, Li = eng.eval(h)
        sync(); dt = time.time() - t0
        cs = torch.tensor([c, float(np.abs(Li).sum()), float(np.abs(g).sum())], dtype=torch.float64, device="cuda")
        lst = [torch.zeros_like(cs) for _ in range(world)]
        if world > 1:
            dist.all_gather(lst, cs)
        same = all(bool((x == cs).all()) for x in lst) if world > 1 else True
        costs.append(c)
        say(stage="eval", iter=it, cost=c, grad_norm=float(np.linalg.norm(g)), seconds=round(dt, 3),
            rows_per_sec=round(N / dt, 1), pass_ms=[round(v, 1) for v in eng.pass_ms()], ranks_identical=same,
            finite=bool(np.isfinite(g).all() and np.isfinite(Li).all()))
        if it == 0:
            c0, g0, al0, Li0 = c, g.copy(), al.copy(), Li.copy()
            h0 = h.copy()
        if it < a.iters:
            h = opt.step(h, g)
    say(stage="fit", iters=a.iters, costs=costs, decreasing=bool(all(b < a_ for a_, b in zip(costs[1:], costs[2:]))
                                                                 and costs[-1] < costs[0]))

    # batched predict: test rows sharded over the ranks, results all-gathered
    T = a.test_rows
    Xs = torch.rand((T, D), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
    eng.eval(h, want_grad=False)                        # factors at the final parameters
    sync(); t0 = time.time()
    mu, sd = eng.predict_sharded(h, Xs)
    sync(); dt = time.time() - t0
    tl, th = shard_bounds(T, rank, world)
    k = min(th - tl, 4096)
    mu_l, sd_l = eng.predict_device(h, Xs[tl:tl + k])
    ok = bool((mu[tl:tl + k] == mu_l.reshape(-1)).all() and (sd[tl:tl + k] == sd_l.reshape(-1)).all())
    say(stage="predict", test_rows=T, seconds=round(dt, 3), rows_per_sec=round(T / dt, 1), matches_local=ok,
        finite=bool(torch.isfinite(mu).all() and torch.isfinite(sd).all()), sd_min=float(sd.min()), sd_max=float(sd.max()))
    del Xs, mu, sd
    eng.close()

    if a.fp64:
        e2 = Engine(D, S, M, device=lr, precision="fp64")
        e2.set_data_device(X, y, n_global=N)
        sync(); t0 = time.time()
        c1, g1, al1, Li1 = e2.eval(h0)
        sync(); dt = time.time() - t0
        say(stage="fp64_parity", seconds=round(dt, 3), cost_rel=abs(c0 - c1) / abs(c1),
            grad_rel=float(np.linalg.norm(g0 - g1) / np.linalg.norm(g1)),
            grad_max_rel_to_max=float(np.abs(g0 - g1).max() / np.abs(g1).max()),
            alpha_abs=float(np.abs(al0 - al1).max()), Li_abs=float(np.abs(Li0 - Li1).max()),
            tolerance="cost 1e-6 relative; gradient, alpha, Li 1e-5 (north_star)")
        e2.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
