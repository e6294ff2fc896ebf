"""Full-shape check of the tree-form contraction: PhAbs * (PowerLaw + Blackbody), wstat, 1024 x 2048 response;
edges vs generic elementwise, both vs the oracle on 48 rows."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from elisa_b200 import synthetic, models as M
from parity import gpu_eval, oracle_eval, rel_err_grad, rel_err_grad_rownorm, rel_err_grad_rownorm_floor, rel_err_value
n = 4096
cfg = synthetic.config2(n=n)
te = np.geomspace(0.05, 500.0, 300)
model = M.PhAbs(te, 0.02 * te ** -2.5) * (M.PowerLaw() + M.Blackbody())
cm = model.compile()
rng = np.random.default_rng(1)
cfg = dict(data=cfg['data'], model=[model], stat=cfg['stat'], theta=cm.params_default * (1.0 + 0.05 * rng.standard_normal((n, cm.nparam))))
out = {}
for mode in ('edges', 'generic'):
    os.environ['ELISA_B200_CONTRACT'] = mode
    out[mode] = gpu_eval(cfg, specialise=True, small_batch=False)
a, b = out['edges'][1], out['generic'][1]
print('edges vs generic: column-normwise %.2e, row-normwise (floored) %.2e, raw %.2e; sum|g| %.3e, sum g %.6e / %.6e' % (
    rel_err_grad(a, b), rel_err_grad_rownorm_floor(a, b), rel_err_grad_rownorm(a, b), np.abs(b).sum(), a.sum(), b.sum()))
rows = np.sort(rng.choice(n, 48, replace=False))
sub = dict(cfg, theta=cfg['theta'][rows])
ll_ref, g_ref = oracle_eval(sub)
for mode in ('edges', 'generic'):
    print(mode, 'vs oracle: value %.2e, grad column-normwise %.2e, row-normwise %.2e' % (
        rel_err_value(out[mode][0][rows], ll_ref), rel_err_grad(out[mode][1][rows], g_ref), rel_err_grad_rownorm_floor(out[mode][1][rows], g_ref)))
