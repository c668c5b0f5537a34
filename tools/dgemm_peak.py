"""cuBLAS DGEMM 8192^3 burst and sustained throughput (the FP64 roofline denominator;
MEASURED_PEAKS.json has no FP64 entry)."""
import json
import time
import torch

n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n ** 3 / best * 1e-9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(4):
        c = a @ b
    k += 4
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n ** 3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"dgemm_tflops_burst": round(burst, 2), "dgemm_tflops_sustained": round(sust, 2), "n": n}))
