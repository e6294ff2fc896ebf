import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import test_gpu_parity as T
from parity import oracle_eval, gpu_eval
from elisa_b200 import models as M
np.set_printoptions(linewidth=200, precision=6)

def report(name, cfg):
    ll_ref, g_ref = oracle_eval(cfg)
    ll, g = gpu_eval(cfg)
    scale = np.maximum(np.abs(g_ref), np.nanmax(np.abs(g_ref), axis=0, keepdims=True))
    err = np.abs(g - g_ref) / scale
    print('==', name, 'value err', np.nanmax(np.abs(ll - ll_ref) / np.maximum(np.abs(ll_ref), 1)), 'grad err', np.nanmax(err),
          'nan ref', np.isnan(g_ref).sum(), 'nan gpu', np.isnan(g).sum())
    i = np.unravel_index(np.nanargmax(err), err.shape)
    print('   worst at', i, 'theta', cfg['theta'][i[0]], 'g', g[i[0], i[1]], 'ref', g_ref[i[0], i[1]])
    bad = np.argwhere(np.isnan(g_ref))
    if len(bad):
        n = bad[0][0]
        print('   nan ref row', n, cfg['theta'][n], 'gpu', g[n], 'll', ll[n], ll_ref[n])

K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
m = M.Constant() * (M.ExpAbs(Ec=0.7) * M.PowerLaw(K=K) + M.HighECut() * (M.Blackbody(K=K) + M.Gauss(El=6.4)))
d9 = T._one_spectrum(E=90, C=50, seed=9)
cm = m.compile(); print(cm.params_name)
report('composite', dict(data=[d9], model=[cm], stat=['cstat'], theta=T._theta_for(m, 50, 7)))
