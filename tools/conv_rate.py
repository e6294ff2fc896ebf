"""Cost of the convolution components on the bench shape (cfg 2: 1024 ch x 2048 E dense RSP, wstat):
the plain CutoffPL + Blackbody against a redshifted one and a flux-normalised one (1000-point flux
grid, the reference's default).  usage: python tools/conv_rate.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from elisa_b200 import Likelihood, synthetic
from elisa_b200 import models as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cfg = synthetic.config2(n=n)
base = np.asarray(cfg['theta'])  # alpha, Ec, K, kT, K
fine = np.geomspace(1.0, 50.0, 2001)
mid = np.sqrt(fine[:-1] * fine[1:])
cpl = synthetic._cutoffpl(fine, 1.2, 20.0, 5.0)
bb = synthetic._blackbody(fine, 1.5, 2.0)
band = (mid >= 2.0) & (mid <= 10.0)
F_en = float(np.sum(1.602176634e-9 * mid[band] * cpl[band]))  # erg cm^-2 s^-1 of the CutoffPL in 2-10 keV
F_ph = float(np.sum(cpl + bb))                                # ph cm^-2 s^-1 of the sum in 1-50 keV
cases = [
    ('CutoffPL + Blackbody', M.CutoffPL() + M.Blackbody(), base),
    ('ZAShift(z=0.5)(CutoffPL + Blackbody)', M.ZAShift(z=0.5)(M.CutoffPL() + M.Blackbody()), base),
    ('EnFlux(2, 10)(CutoffPL(K=1)) + Blackbody', M.EnFlux(2.0, 10.0)(M.CutoffPL(K=1.0)) + M.Blackbody(),
     np.column_stack([base[:, 0], base[:, 1], F_en * base[:, 2] / 5.0, base[:, 3], base[:, 4]])),
    ('PhFlux(1, 50)(CutoffPL(K=1) + Blackbody)', M.PhFlux(1.0, 50.0)(M.CutoffPL(K=1.0) + M.Blackbody()),
     np.column_stack([base[:, 0], base[:, 1], base[:, 3], base[:, 4] / base[:, 2], F_ph * base[:, 2] / 5.0])),
]
for label, model, theta in cases:
    lk = Likelihood(cfg['data'], model, cfg['stat'])
    assert theta.shape[1] == lk.nparam, (label, lk.nparam)
    th = torch.as_tensor(theta, dtype=torch.float64, device='cuda')
    for _ in range(2):
        ll, g = lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(ll).all()) and bool(torch.isfinite(g).all()), label
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lk.loglike_and_grad(th)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    lk.set_profiling(True)
    for _ in range(3):
        lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    prof = {k: v[0] for k, v in lk.get_profile().items()}
    print(f'{label:48s} {n / ms * 1e3:10.3e} evals/s  {ms:8.2f} ms/call  guard fallbacks {lk.guard_fallbacks}', 
          {k: round(v / 3, 2) for k, v in prof.items()})
    lk.close()
