"""Where does the bench's parity sample deviate?  First 16384 draws of cfg2 against the oracle with
the integer fold (6/6, 7/7), the DMMA fold; worst rows / columns, and the engines against each other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from elisa_b200 import Likelihood, synthetic
from parity import oracle_eval

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = synthetic.config2(n=n)
ll_ref, g_ref = oracle_eval(cfg)
col = np.max(np.abs(g_ref), axis=0, keepdims=True)
res = {}
for label, kw in (('i8 6/6', dict(fold='i8')), ('i8 7/7', dict(fold='i8', slices=7, slices_bwd=7)), ('dmma', dict(fold='dmma'))):
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], **kw)
    ll, g = lk.loglike_and_grad(cfg['theta'])
    res[label] = (ll.cpu().numpy(), g.cpu().numpy())
    lk.close()
    d = np.abs(res[label][1] - g_ref) / col
    r, s, p = np.unravel_index(np.argmax(d), d.shape)
    print(f'{label}: value {np.max(np.abs(res[label][0] - ll_ref) / np.maximum(np.abs(ll_ref), 1)):.2e}  grad(col) {d.max():.2e} at row {r} param {p}'
          f'  per-column max {np.max(d, axis=(0, 1))}')
    print('   theta of that row', cfg['theta'][r], ' g_ref', g_ref[r, 0], ' g', res[label][1][r, 0], ' col scale', col[0, 0])
    worst = np.argsort(-d.max(axis=(1, 2)))[:5]
    print('   five worst rows', worst, d.max(axis=(1, 2))[worst])
a, b = res['i8 6/6'][1], res['dmma'][1]
print('i8 6/6 vs dmma grad(col):', np.max(np.abs(a - b) / col), ' i8 7/7 vs dmma:', np.max(np.abs(res['i8 7/7'][1] - b) / col))
