"""Bitwise cross-check of the two-issuer fold kernel against the single-issuer one (and repeated
runs against each other) over random shapes: any lost or doubled accumulator update shows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import sliced_gemm
torch.manual_seed(7)
g = torch.Generator().manual_seed(3)
bad = 0
cases = [(18944, 1024, 2048), (18944, 2048, 1024), (4096, 4096, 4096), (128, 64, 128)]
for _ in range(28):
    m = int(torch.randint(1, 6000, (1,), generator=g)); n = int(torch.randint(1, 3000, (1,), generator=g)); k = int(torch.randint(1, 5000, (1,), generator=g))
    cases.append((m, n, k))
for (m, n, k) in cases:
    a = torch.randn(m, k, dtype=torch.float64, device='cuda') * torch.exp(3 * torch.randn(m, 1, dtype=torch.float64, device='cuda'))
    b = torch.randn(n, k, dtype=torch.float64, device='cuda')
    for ks in (4, 6, 7):
        os.environ['ELISA_B200_I8_ISSUERS'] = '1'
        ref, _ = sliced_gemm(a, b, ks)
        os.environ['ELISA_B200_I8_ISSUERS'] = '2'
        for rep in range(4):
            d, _ = sliced_gemm(a, b, ks)
            nd = int((d != ref).sum())
            if nd:
                bad += 1
                print('MISMATCH', (m, n, k), 'ks', ks, 'rep', rep, nd, 'elements, max rel', float(((d - ref).abs() / ref.abs().clamp_min(1e-300)).max()))
print('cases', len(cases), 'mismatching runs', bad)
