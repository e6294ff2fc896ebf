"""Per-call latency of small batches (cfg3: Band, 4 GBM-like detectors, pgstat, 64 / 8 chains)
with and without CUDA-graph replay."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import Likelihood, synthetic
for n in (64, 8):
    cfg = synthetic.config3(n=n)
    for graphs in ('0', '1'):
        os.environ['ELISA_B200_GRAPHS'] = graphs
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
        th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
        for _ in range(5):
            lk.loglike_and_grad(th)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(200):
            ll, g = lk.loglike_and_grad(th)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 200
        print(f'cfg3 n={n} graphs={graphs}: {dt * 1e6:.1f} us per value+gradient call of 4 spectra', flush=True)
        lk.close()
