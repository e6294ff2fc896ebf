"""Per-kernel-class device time of one BASELINE configuration (both engines): python tools/config_profile.py config4 20000"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import Likelihood, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else 'config4'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
cfg = getattr(synthetic, name)(n=n)
for fold in ('i8', 'dmma'):
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold=fold)
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    on = None if cfg.get('counts_on') is None else torch.as_tensor(cfg['counts_on'], dtype=torch.float64, device='cuda')
    for _ in range(2):
        lk.loglike_and_grad(th, counts_on=on)
    torch.cuda.synchronize()
    lk.set_profiling(True)
    for _ in range(3):
        lk.loglike_and_grad(th, counts_on=on)
    prof = lk.get_profile()
    lk.set_profiling(False)
    tot = sum(v[0] for v in prof.values())
    print(name, n, fold, {k: round(v[0] / 3, 3) for k, v in prof.items()}, 'total ms', round(tot / 3, 3), flush=True)
    lk.close()
