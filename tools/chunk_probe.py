"""Chunk size of the pipeline on the bench configuration: ms per 303 104 draws for chunks of 1, 2 and 3 row tiles per SM."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import Likelihood, synthetic
n = 18944 * 24
cfg = synthetic.config2(n=n)
th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
for rep in range(2):
    for tiles in (1, 2, 3):
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], chunk=18944 * tiles, guard='off')
        for _ in range(2):
            lk.loglike_and_grad(th)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ll, g = lk.loglike_and_grad(th)
        e1.record()
        torch.cuda.synchronize()
        print(f'chunk = {tiles} x 18944 rows: {e0.elapsed_time(e1) / 3:.3f} ms per {n} draws ({n * 3 / e0.elapsed_time(e1) * 1e3:.4g} evals/s), workspace {lk.workspace_bytes / 1e9:.1f} GB', flush=True)
        lk.close()
