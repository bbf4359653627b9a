"""cuBLAS DGEMM probe (roofline denominator for the FP64 tensor path): torch.matmul float64."""
import json, sys, torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained: back-to-back for ~3 s
import time
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
while time.time() - t0 < 3.0:
    for _ in range(4): c = a @ b
    k += 4; torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / k
print(json.dumps({"n": n, "dgemm_tflops_burst": 2 * n**3 / best * 1e-9, "dgemm_tflops_sustained": 2 * n**3 / sus * 1e-9}))
