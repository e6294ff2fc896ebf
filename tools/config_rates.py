"""Throughput (likelihood+gradient evals/s, device-resident inputs) of the five BASELINE
configurations with both fold engines on one GPU.  usage: python tools/config_rates.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import Likelihood, synthetic

CASES = [('cfg1 PowerLaw cstat 1024x2048 dense', 'config1', dict(n=200000)),
         ('cfg2 CutoffPL+BB wstat 1024x2048 dense', 'config2', dict(n=200000)),
         ('cfg3 Band pgstat 4 x (128 x 140), 64 chains', 'config3', dict(n=64)),
         ('cfg3 same, 65536 vectors', 'config3', dict(n=65536)),
         ('cfg4 Gauss+PL chi2 4096x4096 band-sparse, replica counts', 'config4', dict(n=20000)),
         ('cfg5 PowerLaw 16 spectra cstat/chi2 1024x2048', 'config5', dict(n=60000))]
for label, name, kw in CASES:
    cfg = getattr(synthetic, name)(**kw)
    n = len(cfg['theta']); S = len(cfg['data'])
    row = [label]
    for fold in ('i8', 'dmma'):
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold=fold)
        th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
        on = None if cfg.get('counts_on') is None else torch.as_tensor(cfg['counts_on'], dtype=torch.float64, device='cuda')
        for _ in range(2):
            lk.loglike_and_grad(th, counts_on=on)
        torch.cuda.synchronize()
        reps = 3 if n > 1000 else 50
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            lk.loglike_and_grad(th, counts_on=on)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row.append(f'{fold}: {n * S / ms * 1e3:.3e} evals/s ({ms:.3f} ms per call)')
        lk.close(); del lk
    print(' | '.join(row), flush=True)
