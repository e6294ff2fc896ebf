"""Bare sliced-integer GEMM at the bench shape, for ncu: python tools/i8_gemm_prof.py [ks] [m n k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import sliced_gemm
ks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
m, n, k = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (18944, 1024, 2048)
a = torch.rand(m, k, dtype=torch.float64, device='cuda'); b = torch.rand(n, k, dtype=torch.float64, device='cuda')
best = 1e9
for _ in range(4):
    d, ms = sliced_gemm(a, b, ks); best = min(best, ms)
print(f'ks {ks} {m}x{n}x{k}: {best:.3f} ms, {2.0*m*n*k/best/1e9:.1f} TFLOP/s FP64-equivalent')
