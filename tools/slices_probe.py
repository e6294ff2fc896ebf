"""Accuracy and time of the integer fold as a function of the digits per operand, separately for
the forward fold (value side) and the transposed fold (gradient side): the five BASELINE shapes
(scaled), the seeded random sweep, the grouped OGIP case.  Unguarded integer fold (fold='i8')
against the CPU oracle.  usage: python tools/slices_probe.py  -> markdown on stdout"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from elisa_b200 import Likelihood, synthetic
from parity import oracle_eval, rel_err_grad, rel_err_grad_rownorm_floor, rel_err_value
from test_gpu_parity import _theta_for, small_config
from test_gpu_random import _random_data, _random_model

VARIANTS = [(6, 6), (6, 5), (6, 4), (7, 5), (5, 5)]


def cases():
    for k in (1, 2, 3, 5):
        yield f'small cfg{k}', small_config(k)
    yield 'cfg1 full shape, 192 rows', synthetic.config1(n=192)
    yield 'cfg2 full shape, 192 rows', synthetic.config2(n=192)
    yield 'cfg5 full shape (4 of 16 spectra), 64 rows', synthetic.config5(n=64, S=4)
    for seed in range(24):
        rng = np.random.default_rng(1000 + seed)
        stat = ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'][seed % 5]
        E = int(rng.choice([33, 64, 127, 128, 129, 200, 257, 500])); C = int(rng.choice([17, 63, 64, 65, 128, 190, 300]))
        n = int(rng.choice([1, 5, 127, 128, 129, 300, 700]))
        model = _random_model(rng); cm = model.compile()
        if cm.nparam > 16 or min(E, C) < 64:
            continue
        d = _random_data(rng, E, C, stat)
        yield f'random{seed:02d} {stat} {E}x{C} n={n}', dict(data=[d], model=[cm], stat=[stat], theta=_theta_for(model, n, seed))


TIME_ONLY = '--time-only' in sys.argv
print('| case | ' + ' | '.join(f'{f}/{b}: value, grad (col), grad (row), flagged' for f, b in VARIANTS) + ' |')
print('|---|' + '---|' * len(VARIANTS))
worst = {v: [0.0, 0.0, 0.0] for v in VARIANTS}
for label, cfg in ([] if TIME_ONLY else cases()):
    ll_ref, g_ref = oracle_eval(cfg)
    row = []
    for f, b in VARIANTS:
        models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
        lk = Likelihood(cfg['data'], models, cfg['stat'], fold='i8', slices=f, slices_bwd=b)
        ll, g = lk.loglike_and_grad(cfg['theta'], counts_on=cfg.get('counts_on'), counts_off=cfg.get('counts_off'))
        flagged, est = lk.fold_guard()
        lk.close()
        e = (rel_err_value(ll.cpu().numpy(), ll_ref), rel_err_grad(g.cpu().numpy(), g_ref), rel_err_grad_rownorm_floor(g.cpu().numpy(), g_ref))
        if not flagged:
            for i in range(3):
                worst[(f, b)][i] = max(worst[(f, b)][i], e[i])
        row.append(f'{e[0]:.1e}, {e[1]:.1e}, {e[2]:.1e}, {flagged}')
    print(f'| {label} | ' + ' | '.join(row) + ' |', flush=True)
print('| **worst unflagged** | ' + ' | '.join(f'{w[0]:.1e}, {w[1]:.1e}, {w[2]:.1e}' for w in worst.values()) + ' |')

# time of the bench chunk per variant
print()
print('| forward/transposed digits | digit products | fold ms/chunk | tfold ms/chunk | step ms (8 chunks) |')
print('|---|---|---|---|---|')
cfg = synthetic.config2(n=18944 * 8)
th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
for f, b in VARIANTS:
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8', slices=f, slices_bwd=b)
    for _ in range(2):
        lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    lk.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        lk.loglike_and_grad(th)
    e1.record(); torch.cuda.synchronize()
    prof = lk.get_profile()
    print(f'| {f}/{b} | {f * (f + 1) // 2} + {b * (b + 1) // 2} | {prof["fold"][0] / 24:.3f} | {prof["tfold"][0] / 24:.3f} | {e0.elapsed_time(e1) / 3:.2f} |', flush=True)
    lk.close()
