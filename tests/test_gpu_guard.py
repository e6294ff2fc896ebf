"""Accuracy guard of the sliced-integer fold: per-row fallback, the guard of the normal-equations
path, and a randomized sweep holding the guard's estimate against the true error (the integer
fold against the FP64 DMMA fold of the same plan) at small and medium shapes."""

import numpy as np
import pytest
import torch

from elisa_b200 import FixedData, Likelihood, models as M

pytestmark = pytest.mark.gpu


def _sweep_case(seed):
    from test_gpu_parity import _theta_for
    from test_gpu_random import _random_data, _random_model

    rng = np.random.default_rng(5000 + seed)
    stat = ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'][seed % 5]
    E = int(rng.choice([64, 65, 96, 128, 160, 200, 257, 384]))
    C = int(rng.choice([64, 65, 100, 128, 190, 256, 300]))
    n = int(rng.choice([5, 127, 129, 300]))
    model = _random_model(rng)
    cm = model.compile()
    d = _random_data(rng, E, C, stat)
    theta = _theta_for(model, n, seed)
    # far from the optimum: a normalisation off by up to two decades starves / floods the channels
    for j, name in enumerate(cm.params_name):
        if name.split('.')[-1] in ('K', 'norm') and rng.random() < 0.5:
            theta[:, j] *= 10.0 ** rng.uniform(-2.0, 2.0)
    return dict(data=[d], model=[cm], stat=[stat], theta=theta), f'{stat} {E}x{C} n={n}'


@pytest.mark.parametrize('block', range(4))
def test_guard_estimate_covers_the_true_error(lib, block):
    """No case may be off by more than 1e-10 (integer fold vs DMMA fold, same plan constants) and
    unflagged.  40 seeded cases per block."""
    from parity import rel_err_grad, rel_err_value

    worst_unflagged = 0.0
    for seed in range(40 * block, 40 * block + 40):
        cfg, label = _sweep_case(seed)
        if cfg['model'][0].nparam > 16 or cfg['model'][0].nparam == 0:
            continue
        ref_lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='dmma')
        ll_ref, g_ref = ref_lk.loglike_and_grad(cfg['theta'])
        ref_lk.close()
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8')
        ll, g = lk.loglike_and_grad(cfg['theta'])
        flagged, est = lk.fold_guard()
        lk.close()
        ok = torch.isfinite(ll_ref).all(dim=1) & torch.isfinite(g_ref).flatten(1).all(dim=1)
        if not bool(ok.any()):
            continue
        e_v = rel_err_value(ll[ok].cpu().numpy(), ll_ref[ok].cpu().numpy())
        e_g = rel_err_grad(g[ok].cpu().numpy(), g_ref[ok].cpu().numpy())
        if flagged == 0:
            worst_unflagged = max(worst_unflagged, e_v, e_g)
            assert e_v <= 1e-10 and e_g <= 1e-10, (seed, label, e_v, e_g, est)
    assert worst_unflagged <= 1e-10


def test_only_flagged_rows_take_the_dmma_fold(lib):
    """A batch of steep power laws (index 4 over three decades) with a few flat ones (index 0.3)
    mixed in, none of them among the pilot rows: the envelopes fit the steep rows, the guard flags
    the flat ones, and the guarded entry point re-evaluates exactly those rows on the DMMA fold."""
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    rng = np.random.default_rng(5)
    E, C, n = 256, 128, 2000
    eg = np.geomspace(0.1, 100.0, E + 1)
    R = np.abs(rng.standard_normal((E, C))) + 0.1
    d = FixedData('d', eg, rng.poisson(50.0, C).astype(float), np.ones(C), 10.0, response_matrix=R)
    model = M.PowerLaw(alpha=M.UniformParameter('alpha', 2.0, -3.0, 10.0), K=M.UniformParameter('K', 1.0, 1e-10, 1e10))
    theta = np.column_stack([4.0 + 0.01 * rng.random(n), np.full(n, 5.0)])
    odd = 7 + 15 * rng.choice(n // 15 - 1, size=17, replace=False)  # the pilot takes rows 0, 15, 30, ...
    theta[odd] = np.column_stack([0.3 + 0.01 * rng.random(17), np.full(17, 0.05)])
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    lk = Likelihood([d], model, ['cstat'])
    assert lk.fold_slices == [6]
    ll, g = lk.loglike_and_grad(theta)
    assert rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL and rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
    assert lk.fallback_rows == len(odd) and lk.guard_fallbacks == 1 and lk.rebalances == 0, \
        (lk.fallback_rows, lk.guard_fallbacks, lk.rebalances)
    # the host entry point does the same
    ll_h, g_h = lk.loglike_and_grad_host(theta)
    assert np.array_equal(ll_h, ll.cpu().numpy()) and np.array_equal(g_h, g.cpu().numpy())
    assert lk.fallback_rows == 2 * len(odd)
    lk.close()


def test_normal_equations_consult_the_guard(lib):
    """the re-fit path folds U and its tangents through the integer engine; the kernel
    accumulates the guard estimate and a flagged call is repeated on the DMMA fold; fit() reports
    deviance and gradient from a guarded evaluation of the final parameters."""
    E = C = 256
    rng = np.random.default_rng(3)
    d = FixedData('starved', np.linspace(1.0, 10.0, E + 1), rng.poisson(20.0, C).astype(float), np.ones(C), 100.0,
                  response_matrix=np.eye(E) * 50.0)
    U = M.UniformParameter
    model = M.Gauss(El=U('El', 5.0, 1.0, 10.0), sigma=U('sigma', 0.05, 0.01, 1.0), K=U('K', 10.0, 1e-3, 1e3))
    theta = np.column_stack([np.linspace(2.0, 9.0, 40), np.full(40, 0.05), np.full(40, 10.0)])
    lk = Likelihood([d], model, ['cstat'])
    ll, g, h = lk.normal_equations(theta)
    assert getattr(lk, 'normal_eq_fallbacks', 0) == 1
    ref = Likelihood([d], model, ['cstat'], fold='dmma')
    ll_r, g_r, h_r = ref.normal_equations(theta)
    assert torch.equal(ll, ll_r) and torch.equal(g, g_r) and torch.equal(h, h_r)  # the same DMMA kernels answered
    ll2, g2 = ref.loglike_and_grad(theta)
    assert torch.allclose(ll, ll2.sum(-1), rtol=1e-12)
    ref.close()
    lk.close()


def test_plans_are_released(lib):
    """a Likelihood that is dropped (not closed) gives its device memory back, and
    CompiledModel.eval closes the plans it evicts from its cache."""
    import gc

    from elisa_b200 import synthetic

    cfg = synthetic.config2(n=8, E=512, C=256)
    torch.cuda.synchronize()
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    lk.loglike_and_grad(cfg['theta'])
    del lk
    gc.collect()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
        lk.loglike_and_grad(cfg['theta'])
        del lk
        gc.collect()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)
    cm = M.PowerLaw().compile()
    for k in range(6):
        cm.eval(np.geomspace(0.1, 10.0 + k, 300), cfg['theta'][:4, :2])
    gc.collect()
    torch.cuda.synchronize()
    assert torch.cuda.mem_get_info()[0] >= free0 - (64 << 20)
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk2:
        lk2.loglike(cfg['theta'])
    assert lk2._h is None
