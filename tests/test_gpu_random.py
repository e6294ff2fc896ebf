"""Seeded random parity sweep: random model trees over every built-in component, random grid /
channel / batch sizes (ragged against every tile size of the kernels), random statistic,
engine, chunking and per-replica data -- value and gradient against the oracle."""

import numpy as np
import pytest

from elisa_b200 import models as M
from elisa_b200.data import FixedData

pytestmark = pytest.mark.gpu

ADD = [c for c in M.ADDITIVE if c not in (M.PLEnFlux, M.PLPhFlux)]


def _random_model(rng):
    def additive():
        cls = ADD[rng.integers(len(ADD))]
        kw = {}
        if cls._numerical and rng.random() < 0.3:
            kw['method'] = 'simpson'
        return cls(**kw)

    def multiplicative():
        if rng.random() < 0.2:
            te = np.geomspace(0.05, 400.0, 64)
            return M.PhAbs(te, 0.3 * te**-2.4, nH=M.UniformParameter('nH', 0.5, 0.0, 10.0))
        cls = M.MULTIPLICATIVE[rng.integers(len(M.MULTIPLICATIVE))]
        return cls()

    terms = []
    for _ in range(int(rng.integers(1, 4))):
        t = additive()
        if rng.random() < 0.5:
            t = multiplicative() * t
        terms.append(t)
    model = terms[0]
    for t in terms[1:]:
        model = model + t
    if rng.random() < 0.3:
        model = multiplicative() * model
    return model


def _random_data(rng, E, C, stat):
    eg = np.geomspace(0.4, 150.0, E + 1) if rng.random() < 0.7 else np.linspace(0.4, 150.0, E + 1)
    if rng.random() < 0.5:
        R = np.abs(rng.standard_normal((E, C))) + 0.05
    else:  # banded response: exercises the per-tile k ranges
        R = np.zeros((E, C))
        centre = np.linspace(0, C - 1, E)
        for e in range(E):
            lo, hi = max(0, int(centre[e]) - 6), min(C, int(centre[e]) + 7)
            R[e, lo:hi] = np.abs(rng.standard_normal(hi - lo)) + 0.1
    counts = rng.poisson(60.0, C).astype(float)
    back = rng.poisson(25.0, C).astype(float)
    counts[rng.random(C) < 0.05] = 0.0
    back[rng.random(C) < 0.05] = 0.0
    return FixedData('d', eg, counts, 0.5 + rng.random(C), float(rng.uniform(0.5, 5.0)), response_matrix=R,
                     net_counts=counts - 0.4 * back, net_errors=np.sqrt(counts + 1.0), back_counts=back,
                     back_errors=np.sqrt(back + 1.0), back_ratio=np.full(C, 0.4))


@pytest.mark.parametrize('seed', range(24))
def test_random_model_shape_statistic_engine(lib, seed):
    from parity import RTOL, compare
    from test_gpu_parity import _theta_for

    rng = np.random.default_rng(1000 + seed)
    stat = ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'][seed % 5]
    E = int(rng.choice([33, 64, 127, 128, 129, 200, 257, 500]))
    C = int(rng.choice([17, 63, 64, 65, 128, 190, 300]))
    n = int(rng.choice([1, 5, 127, 128, 129, 300, 700]))
    model = _random_model(rng)
    cm = model.compile()
    if cm.nparam > 16:
        pytest.skip('more than 16 free parameters')
    d = _random_data(rng, E, C, stat)
    theta = _theta_for(model, n, seed)
    cfg = dict(data=[d], model=[cm], stat=[stat], theta=theta)
    if rng.random() < 0.4:
        cfg['counts_on'] = rng.poisson(60.0, (n, C)).astype(float)
        if stat in ('pgstat', 'wstat') and rng.random() < 0.5:
            cfg['counts_off'] = rng.poisson(25.0, (n, C)).astype(float)
    # None = the default: integer fold under its accuracy guard (a flagged call is re-evaluated by
    # the FP64 DMMA twin, see Likelihood._guard_tripped); 'dmma' = the FP64 fold
    fold = [None, 'dmma'][int(rng.integers(2))]
    chunk = int(rng.choice([0, 128, 256]))
    spec = bool(rng.integers(2))
    err = compare(cfg, fold=fold, chunk=chunk, specialise=spec)
    assert np.isfinite(err['value']) and np.isfinite(err['grad']), (seed, repr(model), err)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, (seed, repr(model), stat, E, C, n, fold, err)
    # the unguarded integer fold: accurate, or flagged by its guard (bad-fit rows whose starved
    # channels dominate W keep fewer digits of the gradient -- seeds 6 and 17)
    from parity import gpu_eval, oracle_eval, rel_err_grad, rel_err_value

    from elisa_b200 import Likelihood

    lk = Likelihood([d], cm, [stat], fold='i8', chunk=chunk, specialise=spec)
    ll, g = lk.loglike_and_grad(theta, counts_on=cfg.get('counts_on'), counts_off=cfg.get('counts_off'))
    flagged, worst = lk.fold_guard()
    lk.close()
    ll_ref, g_ref = oracle_eval(cfg)
    ok = rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL and rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
    assert ok or flagged > 0, (seed, repr(model), stat, worst)
