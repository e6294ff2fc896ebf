"""Accuracy of the UNGUARDED integer fold with and without balancing of the contraction index, on
the inputs that are hard for a normwise-accurate engine: the seeded random sweep of
tests/test_gpu_random.py, a line through a diagonal response (starved channels), a heavily grouped
OGIP spectrum far from its optimum, and the BASELINE shapes.  Prints value / gradient deviation from
the oracle and the guard's estimates.  usage: python tools/balance_probe.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from elisa_b200 import FixedData, Likelihood, models as M
from parity import oracle_eval, rel_err_grad, rel_err_value
from test_gpu_parity import _theta_for, small_config
from test_gpu_random import _random_data, _random_model
import ogip_cases
from elisa_b200.ogip import load_data


def cases():
    for seed in range(24):
        rng = np.random.default_rng(1000 + seed)
        stat = ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'][seed % 5]
        E = int(rng.choice([33, 64, 127, 128, 129, 200, 257, 500])); C = int(rng.choice([17, 63, 64, 65, 128, 190, 300]))
        n = int(rng.choice([1, 5, 127, 128, 129, 300, 700]))
        model = _random_model(rng); cm = model.compile()
        if cm.nparam > 16:
            continue
        d = _random_data(rng, E, C, stat)
        yield f'random{seed:02d} {stat} {E}x{C} n={n}', dict(data=[d], model=[cm], stat=[stat], theta=_theta_for(model, n, seed))
    E = C = 256
    rng = np.random.default_rng(3)
    d = FixedData('starved', np.linspace(1.0, 10.0, E + 1), rng.poisson(20.0, C).astype(float), np.ones(C), 100.0,
                  response_matrix=np.eye(E) * 50.0)
    U = M.UniformParameter
    model = M.Gauss(El=U('El', 5.0, 1.0, 10.0), sigma=U('sigma', 0.05, 0.01, 1.0), K=U('K', 10.0, 1e-3, 1e3))
    yield 'line through a diagonal response', dict(data=[d], model=[model], stat=['cstat'], theta=_theta_for(model, 40, 11))
    tmp = tempfile.mkdtemp(); cs = ogip_cases.build(tmp); os.chdir(tmp)
    d = load_data(**{k: v for k, v in cs['c_type2_row2_gaussian_back'].items() if not k.startswith('_')})
    model = M.CutoffPL(K=M.UniformParameter('K', 0.05, 1e-6, 1e3)) + M.Gauss(El=6.4, sigma=0.3)
    yield 'grouped OGIP 30x6 pgstat', dict(data=[d], model=[model], stat=['pgstat'], theta=_theta_for(model, 24, 9))
    for k in (1, 2, 3, 5):
        yield f'small cfg{k}', small_config(k)


print(f'{"case":44s} {"balance off: value grad flagged est":>44s}   {"balance on: value grad flagged est":>44s}')
for label, cfg in cases():
    ll_ref, g_ref = oracle_eval(cfg)
    row = []
    for bal in ('0', '1'):
        os.environ['ELISA_B200_BALANCE'] = bal
        models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
        lk = Likelihood(cfg['data'], models, cfg['stat'], fold='i8')
        ll, g = lk.loglike_and_grad(cfg['theta'], counts_on=cfg.get('counts_on'), counts_off=cfg.get('counts_off'))
        flagged, worst = lk.fold_guard()
        lk.close()
        row.append(f'{rel_err_value(ll.cpu().numpy(), ll_ref):9.1e} {rel_err_grad(g.cpu().numpy(), g_ref):9.1e} {flagged:5d} {worst:9.1e}')
    print(f'{label:44s} {row[0]:>44s}   {row[1]:>44s}')
