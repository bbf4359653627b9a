"""Developer probe: time one evaluation of the cfg4 geometry at reduced N for several
settings given as env-var combos.  python tools/perf_probe.py [N] [combo ...]
combo example: SCFGP_STREAMS=3,SCFGP_NO_TMA=1,R=2048"""
import os, sys, time, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child(N, R):
    import torch, numpy as np
    from scfgp_b200 import Engine, synth
    D, S, M = 64, 16, 1024
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    X = torch.rand((N, D), dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    h = synth.draw_params(D, S, M, np.random.default_rng(0), lf_scale=0.3)
    eng = Engine(D, S, M, chunk_rows=R)
    eng.set_data_device(X, y)
    eng.eval(h)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c, g_, a, L = eng.eval(h)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    fl = (7 * 2080 ** 2 + 8 * 64 * 1040) * N
    print(json.dumps({"env": os.environ.get("PROBE_TAG"), "R": R, "N": N, "sec": round(t1 - t0, 4),
                      "tflops": round(fl / (t1 - t0) * 1e-12, 2), "pass_ms": [round(x, 1) for x in eng.pass_ms()],
                      "cost": c, "gnorm": float(np.linalg.norm(g_))}), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]), int(sys.argv[3]))
    else:
        N = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
        combos = sys.argv[2:] or ["SCFGP_STREAMS=3,R=2048"]
        for combo in combos:
            env = dict(os.environ, PROBE_TAG=combo)
            R = 2048
            for kv in combo.split(","):
                k, v = kv.split("=")
                if k == "R": R = int(v)
                else: env[k] = v
            subprocess.run([sys.executable, __file__, "child", str(N), str(R)], env=env)
