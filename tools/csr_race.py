"""cfg 4 (4096 x 4096 band-sparse CCD-like response, <= 49 nnz per channel row): the row-wise CSR fold
(`elisa_b200_csr_fold_dev`, FP64 FMA pipe, one thread per channel row -- the north star's band-sparse
kernel) raced against the block-banded DMMA fold the evaluation path uses, on the same unfolded
models.  Prints a markdown table.  usage: python tools/csr_race.py [replicas]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from elisa_b200 import Likelihood, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
cfg = synthetic.config4(n=8)  # data only: the replicas' counts are not needed for the fold
d = cfg['data'][0]
theta = torch.as_tensor(synthetic._theta_batch(cfg['truth'], n, 32, rel=0.01), dtype=torch.float64, device='cuda')
nnz = len(d.sparse_matrix[2])
lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='dmma')
U = lk.unfold(0, theta)  # (n, 4096)
# dense reference of the fold for correctness (a 4096 x 4096 double matrix: 134 MB)
ip, ix, v = d.sparse_matrix
R = np.zeros((d.n_energies, d.n_channels))
rows = np.repeat(np.arange(d.n_channels), np.diff(ip))
np.add.at(R, (ix, rows), v)
Rt = torch.as_tensor(R, device='cuda')
ref = U[:4096] @ Rt
X = lk.csr_fold(0, U[:4096])
err = float(((X - ref).abs() / ref.abs().clamp_min(1e-300)).max())


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms_csr = timed(lambda: lk.csr_fold(0, U))
# the DMMA fold alone: the plain-store fold of the per-channel sites path, timed by the plan's own events
lk.sites(0, theta[:1024])
torch.cuda.synchronize()
lk.set_profiling(True)
for _ in range(3):
    lk.sites(0, theta)
prof = lk.get_profile()
lk.set_profiling(False)
ms_dmma = prof['fold'][0] / 3
ms_unfold = prof['unfold'][0] / 3
flop = 2.0 * nnz * n
print(f'| fold of {n} unfolded models through the cfg 4 response (nnz {nnz}) | ms | useful TFLOP/s (2 nnz per model) | GB/s of U read + X written |')
print('|---|---|---|---|')
io = 8.0 * (d.n_energies + d.n_channels) * n
print(f'| row-wise CSR kernel (FP64 FMA, thread per channel row, 8 replicas per block) | {ms_csr:.3f} | {flop / ms_csr / 1e9:.2f} | {io / ms_csr / 1e6:.0f} |')
print(f'| block-banded DMMA fold (the path in use) | {ms_dmma:.3f} | {flop / ms_dmma / 1e9:.2f} | {io / ms_dmma / 1e6:.0f} |')
print(f'\ncomponent evaluation of the same batch (for scale): {ms_unfold:.3f} ms; CSR vs dense FP64 reference on 4096 models: max rel dev {err:.1e}')
lk.close()
