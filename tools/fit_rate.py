"""Batched re-fit throughput (fits/s) on the bench shape: CutoffPL + Blackbody, wstat, 1024 x 2048
dense response, one simulated data set per replica (the reference's boot()/ppc() workload,
infer/helper.py:601-806)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from elisa_b200 import Likelihood, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
cfg = synthetic.config2(n=n)
d = cfg['data'][0]
lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
truth = cfg['truth']
probe = Likelihood(cfg['data'], cfg['model'], ['cstat'])
mu = probe.sites(0, truth[None])['cfg2_Non_model'][0]
probe.close()
g = torch.Generator(device='cuda').manual_seed(1)
on = torch.poisson((mu + 0.5 * 30.0).expand(n, -1).contiguous(), generator=g)
off = torch.poisson(torch.full((n, d.n_channels), 30.0, dtype=torch.float64, device='cuda'), generator=g)
start = np.tile(truth, (n, 1)) * (1.0 + 0.05 * np.random.default_rng(0).standard_normal((n, truth.size)))
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = lk.fit(start, counts_on=on, counts_off=off)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{n} replicas: {dt:.3f} s -> {n / dt:.0f} fits/s; iterations max {int(out["n_iter"].max())}, median {int(out["n_iter"].median())}; '
          f'valid {float(out["valid"].double().mean()):.4f}; median deviance {float(out["deviance"].median()):.1f}', flush=True)
print('theta median', out['theta'].median(0).values.tolist(), 'truth', truth.tolist())
