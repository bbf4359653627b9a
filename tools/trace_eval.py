"""Record the launch trace of one evaluation (Chrome-trace JSON) and print a per-stream
utilisation summary.  python tools/trace_eval.py [N] [precision] [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    from scfgp_b200 import Engine, synth
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    prec = sys.argv[2] if len(sys.argv) > 2 else "emu5"
    out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/trace.json"
    D, S, M = 64, 16, 1024
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
    eng = Engine(D, S, M, precision=prec)
    eng.set_data_device(X, y)
    eng.eval(h)
    eng.set_tracing(True)
    eng.eval(h)
    eng.trace_dump(out)
    eng.set_tracing(False)
    ev = json.load(open(out))["traceEvents"]
    t_end = max(e["ts"] + e["dur"] for e in ev)
    print(f"{len(ev)} launches, span {t_end / 1e3:.2f} ms, pass ms {eng.pass_ms()}")
    by = {}
    for e in ev:
        a = by.setdefault(e["name"], [0, 0.0])
        a[0] += 1
        a[1] += e["dur"]
    for k, (n, d) in sorted(by.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k:12s} n={n:5d} total {d / 1e3:8.2f} ms  avg {d / n:8.1f} us  ({d / t_end:.2f} of the span)")
    # union of intervals where a tensor-core GEMM kernel is running
    gem = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev if e["name"] in ("gram", "tgemm", "backward"))
    cov, cur0, cur1 = 0.0, None, None
    for a, b in gem:
        if cur1 is None or a > cur1:
            if cur1 is not None:
                cov += cur1 - cur0
            cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    if cur1 is not None:
        cov += cur1 - cur0
    print(f"  some GEMM kernel running during {cov / t_end:.2f} of the span; sum of GEMM durations {sum(b - a for a, b in gem) / t_end:.2f}")


if __name__ == "__main__":
    main()
