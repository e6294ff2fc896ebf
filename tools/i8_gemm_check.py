"""GPU check + timing of the sliced-integer fold engine on its own (elisa_b200_sliced_gemm_dev):
accuracy against torch FP64 matmul and int8-tensor-core throughput.  usage: python tools/i8_gemm_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import sliced_gemm

torch.manual_seed(0)
dev = 'cuda'

def check(m, n, k, ks, signed=True, dyn=3.0):
    a = torch.randn(m, k, dtype=torch.float64, device=dev) * torch.exp(dyn * torch.randn(m, 1, dtype=torch.float64, device=dev))
    b = torch.randn(n, k, dtype=torch.float64, device=dev) * torch.exp(dyn * torch.randn(n, 1, dtype=torch.float64, device=dev))
    if not signed:
        a, b = a.abs(), b.abs()
    d, ms = sliced_gemm(a, b, ks)
    ref = a @ b.T
    bound = a.abs().amax(1, keepdim=True) * b.abs().sum(1)[None, :] + a.abs().sum(1, keepdim=True) * b.abs().amax(1)[None, :]
    err = ((d - ref).abs() / bound).max().item()
    rel = ((d - ref).abs() / ref.abs().clamp_min(1e-300)).median().item()
    print(f'm {m} n {n} k {k} ks {ks} signed {signed}: normwise err {err:.3e} (2^{torch.log2(torch.tensor(err)).item():.1f}), median rel {rel:.2e}, {ms:.3f} ms', flush=True)
    return err

ok = True
for (m, n, k) in [(128, 64, 128), (128, 64, 256), (256, 128, 512), (300, 200, 1000), (1000, 1024, 2048)]:
    for ks in (4, 5, 6, 7):
        e = check(m, n, k, ks)
        ok &= e < 2.0 ** (-8 * ks + 6)
e = check(512, 512, 2048, 6, signed=False)
ok &= e < 2.0 ** (-42)
print('ACCURACY', 'OK' if ok else 'FAIL', flush=True)

# throughput at the bench shape: 18944 x 1024 x 2048
for ks in (4, 6, 7):
    a = torch.rand(18944, 2048, dtype=torch.float64, device=dev)
    b = torch.rand(1024, 2048, dtype=torch.float64, device=dev)
    best = 1e9
    for _ in range(5):
        d, ms = sliced_gemm(a, b, ks)
        best = min(best, ms)
    flop = 2.0 * 18944 * 1024 * 2048
    nprod = ks * (ks + 1) // 2
    print(f'ks {ks}: {best:.3f} ms  -> {flop / best / 1e9:.1f} TFLOP/s FP64-equivalent, {nprod * flop / best / 1e9:.0f} TOP/s int8', flush=True)
a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
c = a @ b
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
print(f'cuBLAS DGEMM 8192^3: {2 * 8192**3 / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s')
ai = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev); bi = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
try:
    ci = torch._int_mm(ai, bi.T); torch.cuda.synchronize()
    e0.record(); ci = torch._int_mm(ai, bi.T); e1.record(); torch.cuda.synchronize()
    print(f'cuBLAS int8 GEMM 8192^3: {2 * 8192**3 / e0.elapsed_time(e1) / 1e9:.0f} TOP/s')
except Exception as ex:
    print('int8 cuBLAS reference unavailable:', ex)
