"""Edge-wise contraction in tree form against the dual-number contraction on an absorbed continuum
(PhAbs * (PowerLaw + Blackbody), wstat, the bench's 1024 x 2048 response): ELISA_B200_CONTRACT = edges | generic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from elisa_b200 import Likelihood, synthetic, models as M
n = 18944 * 8
cfg = synthetic.config2(n=n)
te = np.geomspace(0.05, 500.0, 300)
model = M.PhAbs(te, 0.02 * te ** -2.5) * (M.PowerLaw() + M.Blackbody())
cm = model.compile()
rng = np.random.default_rng(1)
theta = torch.as_tensor(cm.params_default * (1.0 + 0.05 * rng.standard_normal((n, cm.nparam))), dtype=torch.float64, device='cuda')
lk = Likelihood(cfg['data'], model, cfg['stat'])
for _ in range(2):
    lk.loglike_and_grad(theta)
torch.cuda.synchronize()
lk.set_profiling(True)
for _ in range(3):
    ll, g = lk.loglike_and_grad(theta)
torch.cuda.synchronize()
prof = lk.get_profile()
print(os.environ.get('ELISA_B200_CONTRACT', 'edges'), {k: round(v[0] / 3, 3) for k, v in prof.items() if v[1]}, 'ms per', n, 'draws; checksum', float(g.sum()))
lk.close()
