"""Accuracy-guard estimates of the integer fold on the BASELINE configurations (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elisa_b200 import Likelihood, synthetic
for name, kw in [('config1', dict(n=2000)), ('config2', dict(n=4000)), ('config3', dict(n=64)), ('config4', dict(n=1000)), ('config5', dict(n=500))]:
    cfg = getattr(synthetic, name)(**kw)
    for ks in (6, 7):
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8', slices=ks)
        lk.loglike_and_grad(cfg['theta'], counts_on=cfg.get('counts_on'), counts_off=cfg.get('counts_off'))
        print(name, 'slices', ks, 'guard (flagged, worst estimate):', lk.fold_guard(), flush=True)
        lk.close()
