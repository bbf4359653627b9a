#!/usr/bin/env python
"""Benchmark of the SCFGP hot path: one "step" = one NLML + gradient evaluation
(cost, full gradient, alpha, Li) over the whole synthetic training set.

Workload (BASELINE.json configs[3], the configuration the metric is quoted on):
N = 10M rows total, D = 64, rank S = 16, feature_size M = 1024 (K = 2080), float64.
Rows are sharded over the N GPUs (strong scaling: the total is fixed), each pass's
partial sums are all-reduced over NCCL.

    python bench.py --gpus 1 --steps 3 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29500 bench.py --gpus 8 --steps 3 --warmup 3
    python bench.py --impl reference --steps 2 --warmup 1     # CPU restatement of the reference

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel: read from the committed
# summary of the ncu --set full capture (tools/ncu_traffic.py writes it from the raw CSV export)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")

# CPU arms: a FIXED row sample of the cfg4 geometry, identical in `cpu_baseline` and `--impl reference`, at every
# --gpus N, so that the driver's ratios are reproducible (50 000 rows: the dense K^3 work is ~5 % of an evaluation)
CPU_SAMPLE_ROWS = 50_000

WORKLOAD = dict(name="cfg4: synthetic N=10M, D=64, rank=16, feature_size=1024", N=10_000_000, D=64, S=16, M=1024)


def flops_per_row(D, S, M):
    """Algorithmic flops of one evaluation per training row:
    F_eval = 7 N K^2 + 8 N D J."""
    J = S + M
    K = 2 * J
    return 7.0 * K * K + 8.0 * D * J


# ---------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# synthetic data on the device (torch used for data generation only)
# ---------------------------------------------------------------------------------------
def make_device_problem(n_local, D, S, M, seed, device, dist=None, lf_scale=None):
    import torch
    from scfgp_b200 import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.rand((n_local, D), dtype=torch.float64, device=device, generator=g)
    rng = np.random.default_rng(0)          # teacher and evaluation point are shared by all ranks
    # The spread of Z = x.W grows like sqrt(D): the frequency scale shrinks with it (0.3 at the D = 64 of cfg4),
    # otherwise wide tables (cfg5: D = 256) give random-phase features -- A ~ c I and nothing to learn.
    if lf_scale is None:
        lf_scale = 0.3 * (64.0 / D) ** 0.5 if D > 64 else 0.3
    teacher = synth.draw_params(D, S, M, rng, lf_scale=lf_scale)
    w = torch.as_tensor(rng.standard_normal(2 * (S + M)), device=device)
    params0 = synth.draw_params(D, S, M, rng, lf_scale=lf_scale)
    t = 3
    l_F = teacher[t:t + D * S].reshape(D, S); t += D * S
    r_F = teacher[t:t + M * S].reshape(M, S); t += M * S
    F = l_F @ r_F.T
    l_p = teacher[t:t + S]; t += S
    p = teacher[t:t + M]
    W = torch.as_tensor(np.concatenate([l_F, F], 1), device=device)
    phi = torch.as_tensor(np.concatenate([l_p - l_F.mean(0), p - F.mean(0)]), device=device)
    sc = float(np.exp(teacher[1]) * np.sqrt(2.0 / M))
    y = torch.empty(n_local, dtype=torch.float64, device=device)
    step = 262144
    for i in range(0, n_local, step):
        Z = X[i:i + step] @ W + phi
        y[i:i + step] = sc * (torch.cos(Z) @ w[:S + M] + torch.sin(Z) @ w[S + M:])
        del Z
    y += 0.05 * y.std() * torch.randn(n_local, dtype=torch.float64, device=device, generator=g)
    # standardise globally so that sum(y^2) = N like the auto-normal scaler (Scaler.py:73-79)
    stats = torch.stack([y.sum(), (y * y).sum(), torch.tensor(float(n_local), dtype=torch.float64, device=device)])
    if dist is not None:
        dist.all_reduce(stats)
    mean = stats[0] / stats[2]
    std = torch.sqrt(stats[1] / stats[2] - mean * mean)
    y = (y - mean) / std
    return X, y, params0


# ---------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle's reference-shaped evaluation on host cores
# ---------------------------------------------------------------------------------------
def cpu_reference_run(D, S, M, steps, warmup, rows=CPU_SAMPLE_ROWS, slope=True):
    """Times the literal float64 restatement of the reference's Theano graph (materialised
    Phi, Gauss-Hermite term, LAPACK Cholesky + inverse, reverse-mode gradient) with all host
    threads on a FIXED row sample of the workload.  Returns (rows/s, s/eval, info dict).

    ``slope``: one more evaluation at half the rows gives the per-row cost with the N-independent
    dense K^3 work excluded (two-point fit) -- the rate the CPU path would approach at N = 1e7."""
    import torch
    from oracle import scfgp_oracle as O
    from scfgp_b200 import synth
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(cores)
    n = int(rows)
    try:                                     # Phi and its autograd copies need ~0.2 MB per row at K = 2080
        avail = int(next(l for l in open("/proc/meminfo") if l.startswith("MemAvailable")).split()[1]) * 1024
        while n > 5000 and n * 2.5e5 > 0.6 * avail:
            n //= 2
    except Exception:
        pass
    X, y, h = synth.make_problem(n, D, S, M, seed=1)
    for _ in range(warmup):
        O.cpu_eval_reference_shaped(h, X, y, S, M, threads=cores)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        O.cpu_eval_reference_shaped(h, X, y, S, M, threads=cores)
        ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    info = {"cores": cores, "kind": "port", "sample_rows": n,
            "sample": f"{n} rows (fixed) of the cfg4 geometry (D={D}, S={S}, M={M}; K={2 * (S + M)}); literal float64 "
                      f"restatement of SCFGP.py:92-129 (materialised Phi + autograd), torch-CPU {cores} threads; "
                      f"{sec:.2f} s/eval including the N-independent dense K^3 work"}
    if slope and n >= 10000:
        half = n // 2
        t0 = time.perf_counter()
        O.cpu_eval_reference_shaped(h, X[:half], y[:half], S, M, threads=cores)
        sec_half = time.perf_counter() - t0
        per_row = max((min(ts) - sec_half) / (n - half), 1e-12)
        info["rows_per_sec_asymptotic"] = 1.0 / per_row
        info["fixed_seconds_per_eval"] = max(min(ts) - per_row * n, 0.0)
        info["sample"] += (f"; two-point fit with {half} rows: {per_row * 1e6:.1f} us/row + "
                           f"{info['fixed_seconds_per_eval']:.2f} s fixed => {1.0 / per_row:.0f} rows/s as N grows")
    return n / sec, sec, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=WORKLOAD["N"], help="total training rows (default: the cfg4 workload)")
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--precision", default="emu5",
                    help="arithmetic of the streaming products: 'emu<digits>' = float64 emulated on the int8 "
                         "tensor cores (tcgen05, default 5 base-256 digits), 'fp64' = FP64 DMMA path")
    ap.add_argument("--no-fp64-compare", action="store_true",
                    help="skip the one-evaluation comparison run of the FP64 DMMA path (N = 1 GPU only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--balance", action="store_true",
                    help="N > 1: row shards proportional to each GPU's measured rows/s instead of equal shards "
                         "(sharding.weighted_bounds).  Off by default: on the pool's 8-GPU boxes the per-evaluation "
                         "speed of a power-capped GPU jitters by about as much (+-3 %%) as the GPUs differ, and the "
                         "gain was 0.3 %% (profiles/r2_bench_N10M_8gpu_emu5_balanced.json)")
    args = ap.parse_args()

    D, S, M = WORKLOAD["D"], WORKLOAD["S"], WORKLOAD["M"]
    N = args.rows
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    fpr = flops_per_row(D, S, M)
    config = {"workload": WORKLOAD["name"] if N == WORKLOAD["N"] else f"cfg4 geometry at N={N}",
              "N": N, "D": D, "rank": S, "feature_size": M, "K": 2 * (S + M),
              "parallelism": f"rows sharded over {args.gpus} GPU(s), 3 all-reduces/eval",
              "l2": "inputs (X = N*D*8 bytes per pass) far exceed the 126 MB L2; no explicit flush"}

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        rows_s, sec, info = cpu_reference_run(D, S, M, max(1, args.steps), args.warmup)
        config = dict(config, reference_sample_rows=info["sample_rows"])
        line = {"impl": "reference", "metric": "nlml_grad_rows_per_sec", "value": rows_s, "unit": "rows/s",
                "evals_per_sec_at_N": rows_s / N, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": dict(info, value=rows_s, unit="rows/s"),
                "e2e": {"value": rows_s, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm
    import torch
    from scfgp_b200 import Engine, shard_bounds, weighted_bounds, _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    lo, hi = shard_bounds(N, rank, world)
    n_local = hi - lo

    # (N > 1: 15 % more rows than the equal share are generated per rank, so that the shards can be re-cut in
    #  proportion to the ranks' measured speeds without regenerating anything; the evaluation only sees n_local rows)
    n_gen = n_local if world == 1 or (not args.balance) else int(n_local * 1.15) // 128 * 128 + 128
    Xfull, yfull, params0 = make_device_problem(n_gen, D, S, M, seed=1000 + rank, device=device, dist=dist)
    Xd, yd = Xfull[:n_local], yfull[:n_local]
    eng = Engine(D, S, M, device=local_rank, chunk_rows=args.chunk_rows or None, precision=args.precision,
                 pinned_results=True)   # results land in page-locked host buffers owned by the engine
    eng.set_data_device(Xd, yd, n_global=N)
    emulated = eng.precision.startswith("emu")
    # 'emu5' or 'emu5/4' (Gram digits / digits of the other products)
    digits_gram, _, digits_rest = eng.precision[3:].partition("/") if emulated else ("0", "", "")
    digits_gram = int(digits_gram or 0)
    digits = int(digits_rest or digits_gram)          # pass 3 (the roofline kernel) runs with these
    config["precision"] = (f"float64 emulated on the int8 tensor cores: {digits_gram} base-256 digits per operand for the "
                           f"Gram that feeds the Cholesky, {digits} for the other products "
                           f"({8 * digits - 2} fractional bits of each row's power-of-two scale, "
                           f"{digits * (digits + 1) // 2} digit products), exact int32 accumulation (tcgen05.mma kind::i8); "
                           "feature map, Cholesky, inverse and small dense backward in true FP64"
                           if emulated else "FP64 tensor path (DMMA.8x8x4)")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.eval(params0)
    # ---- N > 1: row shards in proportion to each GPU's own speed.  The B200s of a node run this workload at their
    #      1 kW power cap and settle at clocks that differ by a few per cent; every all-reduce of an evaluation waits
    #      for the slowest rank.  One more untimed evaluation measures each rank's device time in the three streaming
    #      passes (CUDA events inside the passes: waiting for other ranks is not in them), the shards are re-cut with
    #      sharding.weighted_bounds and the synthetic rows regenerated.  Total rows and total work are unchanged.
    shards = None
    if dist is not None:
        eq_rows = [shard_bounds(N, r, world)[1] - shard_bounds(N, r, world)[0] for r in range(world)]
        shards = {"policy": "equal", "rows": eq_rows}

        def rank_speeds():
            eng.eval(params0)                      # (settle at the new shard sizes)
            eng.eval(params0)
            p1, p2, p3, _ = eng.pass_ms()
            mine = torch.tensor([n_local / max(p1 + p2 + p3, 1e-6)], dtype=torch.float64, device=device)
            allsp = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allsp, mine)
            return [float(t[0]) for t in allsp]

        speeds = rank_speeds()
        shards["rank_rows_per_ms_equal_shards"] = speeds
        if args.balance:
            # a fast rank that idles at the all-reduces boosts higher than it sustains under full load, so the
            # measurement is repeated at the re-cut shard sizes (up to 3 rounds)
            shards["policy"] = "proportional to each rank's measured rows/ms (sharding.weighted_bounds, granule 128, <= 3 rounds)"
            for _round in range(3):
                lo2, hi2 = weighted_bounds(N, rank, speeds, granule=128)
                n_new = hi2 - lo2
                # [largest relative move, 1 if some rank would need more rows than it generated]
                moved = torch.tensor([abs(n_new - n_local) / max(n_local, 1), float(n_new > n_gen)],
                                     dtype=torch.float64, device=device)
                dist.all_reduce(moved, op=dist.ReduceOp.MAX)
                if float(moved[0]) < 2e-3 or float(moved[1]) > 0:
                    break
                n_local = n_new
                Xd, yd = Xfull[:n_local], yfull[:n_local]
                eng.set_data_device(Xd, yd, n_global=N)
                speeds = rank_speeds()
            cnt = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
            allc = [torch.zeros_like(cnt) for _ in range(world)]
            dist.all_gather(allc, cnt)
            shards["rows"] = [int(t[0]) for t in allc]
            shards["rank_rows_per_ms_final_shards"] = speeds
            N_eff = sum(shards["rows"])
            assert N_eff == N or world == 1, (N_eff, N)
    # ---- timed region: K evaluations, inputs resident in HBM
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.set_profiling(_lib.FAM_BACKWARD)
    launches0 = eng.launches
    barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    pass_ms = []
    for _ in range(args.steps):
        cost, grad, alpha, Li = eng.eval(params0)
        pass_ms.append(eng.pass_ms())
    e1.record()
    barrier()
    grad, alpha, Li = grad.copy(), alpha.copy(), Li.copy()   # (views of the engine's pinned buffers: later evals overwrite them)
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches - launches0
    bw_ms, bw_launches = eng.get_profile()
    eng.set_profiling(_lib.FAM_NONE)
    t = torch.tensor([ms_total, float(launches)], dtype=torch.float64, device=device)
    if dist is not None:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        if shards is not None:
            shards["rank_ms_per_step"] = [float(x[0]) / args.steps for x in allt]   # (each includes its waits for the others)
            config["parallelism"] += "; shards " + shards["policy"]
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_total = float(tmax[0]); launches = int(tsum[1])
    ms_step = ms_total / args.steps
    rows_per_s = N / (ms_step * 1e-3)

    # ---- the same evaluation without the digit cache of Phi (emulated mode keeps Phi's int8 digits of all rows
    #      in HBM when they fit; without it passes 2 and 3 recompute the feature map): 2 timed steps
    phi_cache = None
    if emulated:
        cache_bytes = eng.phi_cache_bytes
        eng.set_phi_cache(False)
        eng.eval(params0)
        barrier()
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            eng.eval(params0)
        c1.record()
        barrier()
        tc = torch.tensor([c0.elapsed_time(c1) / 2], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        eng.set_phi_cache(True)
        eng.eval(params0)                                    # (re-build it for what follows)
        phi_cache = {"bytes_per_gpu": cache_bytes, "in_use_for_value": cache_bytes > 0,
                     "ms_per_step_without_cache": float(tc[0]), "ms_per_step_with_cache": ms_step if cache_bytes > 0 else None,
                     "note": "digits x ceil128(rows) x ceil16(K) bytes of HBM per GPU; bounded by SCFGP_PHI_CACHE_MAX_GB "
                             "(default 120) and free memory; declined caches are reported on stderr"}

    # ---- end to end through the public call with HOST buffers (train_func(X, y) semantics):
    #      pinned host X/y -> device inside the timed region, results read back to host
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((n_local, D), dtype=torch.float64, pin_memory=True); Xh.copy_(Xd)
        yh = torch.empty((n_local,), dtype=torch.float64, pin_memory=True); yh.copy_(yd)
        Xn, yn = Xh.numpy(), yh.numpy()
        eng.set_data(Xn, yn, n_global=N)
        nrep = max(1, args.steps)                            # the same K steps as the device-timed value
        barrier()
        t0 = time.perf_counter()
        for _ in range(nrep):
            eng.set_data(Xn, yn, n_global=N, overlap=True)   # H2D of this step's inputs (row blocks, consumed by pass 1 as they land)
            c2, g2, a2, L2 = eng.eval(params0)               # D2H of cost, grad, alpha, Li
        barrier()
        sec = (time.perf_counter() - t0) / nrep
        tt = torch.tensor([sec], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt[0])
        K = eng.K
        e2e = {"value": N / sec, "unit": "rows/s", "ms_per_step": sec * 1e3,
               "pass_ms_last": dict(zip(["pass1", "pass2", "pass3", "dense"], [float(v) for v in eng.pass_ms()])),
               "h2d_bytes_per_step": int(N * D * 8 + N * 8 + eng.P * 8 * world),
               "d2h_bytes_per_step": int(8 * (1 + eng.P + K + K * K)) * world,
               "api": "Engine.set_data(X_host, y_host, overlap=True) + Engine.eval(params) == "
                      "scfgp_set_data_host_async + scfgp_eval (pinned host X/y uploaded in row blocks inside the "
                      "timed region, pass 1 starts on each block as it lands; cost, grad, alpha, Li read back)"}
        assert abs(c2 - cost) <= 1e-9 * abs(cost)
        eng.set_data_device(Xd, yd, n_global=N)

    if rank != 0:
        if dist is not None:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (pass 3: 4 N K^2 of the 7 N K^2 flops)
    K = eng.K
    peak_burst = peak_sus = None
    try:
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=device); b = torch.randn(n, n, dtype=torch.float64, device=device)
        for _ in range(2):
            a @ b
        best = 1e9
        for _ in range(4):
            s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
            s0.record(); a @ b; s1.record(); torch.cuda.synchronize()
            best = min(best, s0.elapsed_time(s1))
        peak_burst = 2.0 * n ** 3 / best * 1e-9
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(60):
            a @ b
        s1.record(); torch.cuda.synchronize()
        peak_sus = 2.0 * n ** 3 * 60 / s0.elapsed_time(s1) * 1e-9
        del a, b
    except Exception as ex:  # pragma: no cover
        peak_burst = None
    dgemm_peak = peak_sus or 35.4
    # (1) over the timed region: pass 3 = [features, digits, backward GEMM, dW contraction, column sums] on
    #     overlapped streams, bracketed by CUDA events on the handle's main stream.  Its algorithmic work is
    #     4 N K^2 float64 flops (Phi.(2Abar) + Psi.(2B)); the helpers are charged to the kernel (conservative).
    p3_ms = sum(p[2] for p in pass_ms) / len(pass_ms)
    flops_p3 = 4.0 * K * K * n_local
    achieved_f64 = flops_p3 / (p3_ms * 1e-3) * 1e-12
    # (2) one launch in isolation (single stream, CUDA events around every backward-GEMM launch)
    iso_ms, iso_n = 0.0, 0
    if world == 1:                      # (an evaluation is collective: rank 0 cannot run one alone)
        eng.set_streams(1)
        eng.set_profiling(_lib.FAM_BACKWARD)
        eng.eval(params0)
        iso_ms, iso_n = eng.get_profile()
        eng.set_profiling(_lib.FAM_NONE)
        eng.set_streams(4)
    rows_per_launch = n_local / max(iso_n, 1) if iso_n else (8192 if emulated else 4096)
    flops_per_launch = 4.0 * K * K * rows_per_launch
    iso_avg = iso_ms / max(iso_n, 1)
    whole_tf = fpr * N / (ms_step * 1e-3) * 1e-12 / max(world, 1)
    common = {"avg_launch_ms_isolated": iso_avg, "launches_isolated": iso_n,
              "overlapped_launch_ms_sum": bw_ms, "overlapped_launches": bw_launches,
              "peak_dgemm_burst": peak_burst, "peak_dgemm_sustained": peak_sus,
              "whole_eval_fp64_equivalent_tflops": whole_tf,
              "whole_eval_vs_dgemm_peak": whole_tf / dgemm_peak,
              "pass_ms_last": dict(zip(["pass1", "pass2", "pass3", "dense"], pass_ms[-1]))}
    if emulated:
        # digit-product operations actually asked of the int8 tensor cores: 2 * rows * K * 2K per digit product
        nprod = digits * (digits + 1) // 2
        int8_ops_p3 = flops_p3 * nprod
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_sus = peaks.get("bf16_tflops_sustained") or 1400.0
        bf16_burst = peaks.get("bf16_tflops") or 1590.0
        peak_i8 = 2.0 * bf16_sus
        achieved = int8_ops_p3 / (p3_ms * 1e-3) * 1e-12
        # measured in this run: what the tensor pipe sustains on this very instruction mix when nothing else
        # happens (no operand traffic, no epilogue): every SM for ~0.5 s, random int8 operands (power-capped)
        import ctypes as C
        stream_only = None
        tile_n = 96 if digits == 5 else 64
        tops, clk, mhz = C.c_double(), C.c_double(), C.c_double()
        if digits in (4, 5, 6) and eng.lib.scfgp_dbg_mma_stream_peak(
                local_rank, digits, tile_n, 1_200_000, 1, C.byref(tops), C.byref(clk), C.byref(mhz)) == 0:
            stream_only = {"sustained_tops": tops.value, "clk_per_k32_step": clk.value, "sm_mhz": mhz.value,
                           "frac_of_it": achieved / tops.value,
                           "how": f"scfgp_dbg_mma_stream_peak: the {digits}-digit / 128 x {tile_n} tile tcgen05.mma stream "
                                  "of the product kernels without TMA traffic or epilogue, 1.2 M K=32 steps on every SM, "
                                  "random int8 operands, measured in this run"}
        traffic, traffic_note = None, "no ncu capture summarised in profiles/ncu_traffic.json"
        try:
            tj = json.load(open(TRAFFIC_FILE))
            key = f"k_backward_i8<{digits},{tile_n}>"
            if key in tj:
                traffic = tj[key]["dram_bytes_per_launch"]
                traffic_note = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch of {key} "
                                f"({tj[key]['rows']} rows) from {tj[key]['source']}; algorithmic bytes per launch "
                                f"{tj[key].get('algorithmic_bytes')}: {tj[key].get('algorithmic_note', '')}")
        except Exception:
            pass
        roofline = dict(common, **{
            "bound": "tensor", "kernel": "k_backward_i8 (tcgen05.mma kind::i8, int32 accumulators in TMEM, TMA-fed)",
            "achieved": achieved, "peak": peak_i8, "unit": "TOP/s", "frac": achieved / peak_i8,
            "traffic": traffic, "traffic_note": traffic_note,
            "how": f"int8 digit-product operations of pass 3 (4 N K^2 x {nprod} digit products) / pass-3 device time "
                   "(CUDA events on the launching stream) averaged over the timed steps; the pass includes the feature "
                   "map, digit extraction and dW kernels, so the GEMM kernel alone is faster than this",
            "algorithmic_int8_ops_per_launch": flops_per_launch * nprod,
            "algorithmic_flops_per_launch": flops_per_launch,
            "achieved_isolated_launch": flops_per_launch * nprod / (iso_avg * 1e-3) * 1e-12 if iso_avg > 0 else None,
            "peak_source": ("2 x the sustained dense bf16 rate of MEASURED_PEAKS.json (" + str(bf16_sus) + " TFLOP/s; "
                            "int8 dense = 2 x bf16 dense on sm_100a; the file has no int8 entry; burst would be 2 x "
                            + str(bf16_burst) + "); the evaluation runs for seconds at the 1 kW power cap, hence sustained"
                            if peaks else "fallback 2 x 1.4 PFLOP/s sustained bf16 (MEASURED_PEAKS.json absent)"),
            "fp64_equivalent_tflops_pass3": achieved_f64,
            "fp64_equivalent_vs_dgemm_peak": achieved_f64 / dgemm_peak,
            "int8_mma_stream_only": stream_only,
            "frac_vs_measured_mma_stream": (achieved / stream_only["sustained_tops"]) if stream_only else None,
        })
    else:
        roofline = dict(common, **{
            "bound": "tensor", "kernel": "k_backward_tma (FP64 DMMA.8x8x4 fed by TMA)", "achieved": achieved_f64,
            "peak": dgemm_peak, "unit": "TFLOP/s", "frac": achieved_f64 / dgemm_peak, "traffic": 562.2e6,
            "traffic_note": "dram__bytes_read+write of one 4096-row launch from ncu --set full "
                            "(profiles/r1_prof_backward_ncu_full_N8192.txt); algorithmic operand bytes per "
                            "launch: Phi+Psi slabs 136 MB + 2Abar,2B 69 MB",
            "how": "4 N K^2 algorithmic flops of pass 3 / pass-3 device time (CUDA events on the launching "
                   "stream) averaged over the timed steps",
            "algorithmic_flops_per_launch": flops_per_launch,
            "achieved_isolated_launch": flops_per_launch / (iso_avg * 1e-3) * 1e-12 if iso_avg > 0 else None,
            "peak_source": "cublasDgemm 8192^3 sustained, measured in this run (MEASURED_PEAKS.json holds no "
                           "fp64 entry; register-resident DMMA.8x8x4 issue peak measured 37.0 TFLOP/s, "
                           "profiles/r1_fp64_pipe_microbench.txt)",
        })

    # ---- the FP64 DMMA path on the same data: one evaluation, timing + parity of the emulated results
    fp64_compare = None
    if emulated and world == 1 and not args.no_fp64_compare:
        eng.set_precision("fp64")
        eng.eval(params0)
        n64 = 3
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(n64):
            c64, g64, a64, L64 = eng.eval(params0)
        torch.cuda.synchronize(); sec64 = (time.perf_counter() - t0) / n64
        fp64_compare = {"ms_per_step": sec64 * 1e3, "steps": n64, "rows_per_sec": N / sec64,
                        "fp64_tflops": fpr * N / sec64 * 1e-12,
                        "speedup_of_emulated": sec64 * 1e3 / ms_step,
                        "parity_of_emulated": {
                            "cost_rel": abs(cost - c64) / abs(c64),
                            "grad_rel": float(np.linalg.norm(grad - g64) / np.linalg.norm(g64)),
                            "alpha_abs": float(np.abs(alpha - a64).max()), "Li_abs": float(np.abs(Li - L64).max()),
                            "tolerance": "cost 1e-6 relative; gradient, alpha, Li 1e-5 (north_star)"}}
        eng.set_precision(args.precision)

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        rs, sec, info = cpu_reference_run(D, S, M, steps=2, warmup=1)
        cpu_baseline = dict(info, value=rs, unit="rows/s")

    line = {"metric": "nlml_grad_rows_per_sec", "value": rows_per_s, "unit": "rows/s",
            "evals_per_sec": 1e3 / ms_step, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 emulated on int8 tensor cores (s8 x s8 -> s32 digit products)" if emulated else "f64",
            "data": "synthetic", "config": config, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "phi_cache": phi_cache, "shards": shards, "guard": eng.guard_info(),
            "fp64_dmma_path": fp64_compare,
            "cpu_baseline": cpu_baseline,
            "cost": cost, "grad_norm": float(np.linalg.norm(grad))}
    print(json.dumps(line))
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
