"""Run-to-run bit reproducibility of the sliced-integer GEMM (detects accumulation races)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import sliced_gemm
torch.manual_seed(1)
for (m, n, k) in [(18944, 1024, 2048), (4096, 2048, 1024), (1000, 320, 777)]:
    a = torch.randn(m, k, dtype=torch.float64, device='cuda'); b = torch.randn(n, k, dtype=torch.float64, device='cuda')
    ref, _ = sliced_gemm(a, b, 6)
    bad = 0
    for it in range(30):
        d, _ = sliced_gemm(a, b, 6)
        nd = (d != ref).sum().item()
        bad += nd > 0
        if nd:
            dd = (d - ref).abs()
            print('  run', it, 'differs in', nd, 'elements; max abs diff', dd.max().item(), 'rel', (dd / ref.abs().clamp_min(1e-300)).max().item())
    exact = a @ b.T
    print((m, n, k), 'runs differing:', bad, 'of 30; max rel err vs fp64', ((ref - exact).abs() / (a.abs() @ b.abs().T)).max().item())
# row-block invariance: rows are independent, so any row partition must give identical bits
m, n, k = 40000, 2048, 1024
a = torch.randn(m, k, dtype=torch.float64, device='cuda'); b = torch.randn(n, k, dtype=torch.float64, device='cuda')
full, _ = sliced_gemm(a, b, 6)
for blk in (4096, 18944, 128, 1000):
    parts = [sliced_gemm(a[i:i + blk], b, 6)[0] for i in range(0, m, blk)]
    d = torch.cat(parts)
    nd = (d != full).sum().item()
    print('row blocks of', blk, ': differing elements', nd, 'max rel', ((d - full).abs() / full.abs().clamp_min(1e-300)).max().item())
