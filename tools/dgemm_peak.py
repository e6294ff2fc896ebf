"""Measure the cuBLAS DGEMM (FP64) peak on this B200: the roofline denominator for the dense fold.
Same recipe as MEASURED_PEAKS.json (8192^3, best of 10 burst; back-to-back 4 s sustained)."""
import json, sys, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
c = torch.empty_like(a)
for _ in range(3): torch.matmul(a, b, out=c)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best / 1e9
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); k = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(5): torch.matmul(a, b, out=c)
    k += 5; torch.cuda.synchronize()
e1.record(); e1.synchronize()
sus = 2 * n**3 * k / e0.elapsed_time(e1) / 1e9
out = {"fp64_tflops": round(burst, 2), "fp64_tflops_sustained": round(sus, 2), "how": "torch.matmul f64 8192^3 best of 10 / 4 s loop"}
# kernel name
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    torch.matmul(a, b, out=c); torch.cuda.synchronize()
out["kernels"] = [e.key for e in prof.key_averages()][:4]
print(json.dumps(out))
