"""Row sharding and the phase sequence of one multi-GPU likelihood evaluation.

Training rows shard over ranks (every O(N) quantity of the path is a sum over rows,
SURVEY.md section 8e); X never moves.  ``run_eval_phases`` is the one place that knows
the order of phases and reductions; ``Engine.eval`` drives the CUDA phase API with it
(NCCL), and the CPU tests drive a NumPy phase implementation with it (gloo).
"""
from __future__ import annotations


def shard_bounds(n_rows: int, rank: int, world: int):
    """Contiguous, balanced row range [lo, hi) of `rank`; the first n_rows % world ranks
    get one extra row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def weighted_bounds(n_rows: int, rank: int, weights, granule: int = 1):
    """Contiguous row range [lo, hi) of `rank` when the ranks' shares are proportional to `weights` (one positive
    number per rank, e.g. each GPU's measured rows/s: power-capped B200s of one node differ by a few per cent in
    sustained clock, and every all-reduce of an evaluation waits for the slowest rank).  Boundaries are rounded to
    multiples of `granule` rows (the last rank takes the remainder); equal weights reproduce ``shard_bounds`` up to
    that rounding.  Deterministic: every rank computes the same table from the same weights."""
    w = [float(x) for x in weights]
    world = len(w)
    if world <= 0 or not (0 <= rank < world) or any(not (x > 0.0) for x in w) or granule <= 0:
        raise ValueError("bad rank/weights")
    total = sum(w)
    cuts = [0]
    acc = 0.0
    for r in range(world - 1):
        acc += w[r]
        c = int(round(n_rows * acc / total / granule)) * granule
        cuts.append(min(max(c, cuts[-1]), int(n_rows)))
    cuts.append(int(n_rows))
    return cuts[rank], cuts[rank + 1]


def run_eval_phases(api, allreduce, want_grad: bool):
    """Sequence: pass1 -> allreduce(RED1) -> factor -> pass2 -> allreduce(RED2)
    [-> backward -> pass3 -> allreduce(RED3)].  `api` exposes pass1/factor/pass2/backward/
    pass3 and reduce_buffer(which); `allreduce(buf)` sums `buf` in place over ranks.

    Optional `api.needs_more_precision()`: called after the factor phase; when it returns True (the
    a-posteriori guard of the emulated arithmetic fired on some rank and every rank has switched to the
    next precision and re-begun the evaluation) the sequence starts over from pass 1."""
    while True:
        api.pass1()
        allreduce(api.reduce_buffer(1))
        api.factor()
        again = getattr(api, "needs_more_precision", None)
        if again is None or not again():
            break
    api.pass2()
    allreduce(api.reduce_buffer(2))
    if want_grad:
        api.backward()
        api.pass3()
        allreduce(api.reduce_buffer(3))


def predict_sharded(predict_local, n_rows: int, rank: int, world: int, all_gather):
    """Batched prediction over ranks (test rows shard independently, the only
    exchange is the gather of the results).  ``predict_local(lo, hi)`` returns (mu, std) for the
    rows [lo, hi) of this rank (1-D tensors); ``all_gather(t, padded_len)`` returns the list of every
    rank's tensor padded to ``padded_len``.  Every rank gets the full (mu, std) of length n_rows."""
    import torch
    lo, hi = shard_bounds(n_rows, rank, world)
    mu, sd = predict_local(lo, hi)
    longest = max(b - a for a, b in (shard_bounds(n_rows, r, world) for r in range(world)))
    outs = []
    for t in (mu, sd):
        pad = torch.zeros(longest, dtype=t.dtype, device=t.device)
        pad[: hi - lo] = t.reshape(-1)
        parts = all_gather(pad, longest)
        outs.append(torch.cat([parts[r][: b - a] for r, (a, b) in
                               enumerate(shard_bounds(n_rows, q, world) for q in range(world))]))
    return outs[0], outs[1]
