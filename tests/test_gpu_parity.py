"""GPU parity: the CUDA path through the C ABI against the CPU oracle on identical inputs
(scaled-down versions of the five BASELINE.json configurations + every built-in component)."""

import numpy as np
import pytest

from elisa_b200 import models as M
from elisa_b200 import synthetic
from elisa_b200.data import FixedData

pytestmark = pytest.mark.gpu


def small_config(k):
    if k == 1:
        return synthetic.config1(n=300, E=200, C=96)
    if k == 2:
        return synthetic.config2(n=500, E=300, C=150)
    if k == 3:
        return synthetic.config3(n=64)
    if k == 4:
        return synthetic.config4(n=200, size=512)
    return synthetic.config5(n=257, S=3, E=128, C=64)


KERNELS = pytest.mark.parametrize('specialise', [True, False], ids=['nvrtc', 'interp'])


@KERNELS
@pytest.mark.parametrize('k', [1, 2, 3, 4, 5])
def test_config_parity(lib, k, specialise):
    from parity import RTOL, compare

    err = compare(small_config(k), specialise=specialise)
    assert err['value'] <= RTOL, err
    assert err['grad'] <= RTOL, err


@pytest.mark.parametrize('k', [1, 2, 4])
def test_value_only_and_host_paths(lib, k):
    from parity import RTOL, compare, gpu_eval

    cfg = small_config(k)
    assert compare(cfg, grad=False)['value'] <= RTOL
    ll_d, g_d = gpu_eval(cfg)
    ll_h, g_h = gpu_eval(cfg, host=True)
    assert np.array_equal(ll_d, ll_h) and np.array_equal(g_d, g_h)
    # chunking must not change a single bit
    ll_c, g_c = gpu_eval(cfg, chunk=128)
    assert np.array_equal(ll_d, ll_c) and np.array_equal(g_d, g_c)


def _one_spectrum(E=70, C=40, seed=0, stat='cstat'):
    rng = np.random.default_rng(seed)
    eg = np.geomspace(0.5, 200.0, E + 1)
    R = np.abs(rng.standard_normal((E, C))) + 0.1
    counts = rng.poisson(50.0, C).astype(float)
    back = rng.poisson(20.0, C).astype(float)
    return FixedData('d', eg, counts, 0.5 + rng.random(C), 3.0, response_matrix=R, net_counts=counts - 0.4 * back,
                     net_errors=np.sqrt(counts + 1.0), back_counts=back, back_errors=np.sqrt(back + 1.0),
                     back_ratio=0.4 + 0.0 * back)


def _theta_for(model, n, seed):
    """Parameter vectors scattered 10 % around the defaults.  Power-law indices are kept
    >= 0.05 away from 1: E^(1-a)/(1-a) (models/add.py:40-50, mul.py:266-271) is a
    removable singularity there and the reference formula itself loses ~|1-a|^-2 digits
    of the gradient, so no two implementations agree to 1e-10."""
    cm = model.compile() if not hasattr(model, 'params_default') else model
    rng = np.random.default_rng(seed)
    theta = cm.params_default * (1.0 + 0.1 * rng.standard_normal((n, cm.nparam)))
    for j, name in enumerate(cm.params_name):
        if 'alpha' in name:
            close = np.abs(theta[:, j] - 1.0) < 0.05
            theta[close, j] = 1.0 + np.where(theta[close, j] >= 1.0, 0.05, -0.05) + (theta[close, j] - 1.0)
    return theta


ADD = [c for c in M.ADDITIVE if c not in (M.PLEnFlux, M.PLPhFlux)]


@KERNELS
@pytest.mark.parametrize('method', ['trapz', 'simpson'])
@pytest.mark.parametrize('comp', ADD, ids=lambda c: c.__name__)
def test_every_additive_component(lib, comp, method, specialise):
    from parity import RTOL, compare

    model = comp(method=method) if comp._numerical else comp()
    if comp._numerical is False and method == 'simpson':
        pytest.skip('analytic integral: method is ignored')
    d = _one_spectrum()
    theta = _theta_for(model, 33, 1)
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    err = compare(cfg, specialise=specialise)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err


@KERNELS
def test_powerlaw_alpha_exactly_one(lib, specialise):
    """the `where(alpha != 1, ...)` branch of _powerlaw_integral (models/add.py:40-50)."""
    from parity import RTOL, compare

    d = _one_spectrum()
    for model in (M.PowerLaw(), M.BrokenPL(), M.PLPhFlux(1.0, 10.0)):
        theta = _theta_for(model, 8, 1)
        theta[::2, 0] = 1.0
        if isinstance(model, M.BrokenPL):
            theta[:, 2] = 1.0
        err = compare(dict(data=[d], model=[model], stat=['cstat'], theta=theta), specialise=specialise)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (model, err)


@KERNELS
@pytest.mark.parametrize('energy', [False, True])
def test_plflux_components(lib, energy, specialise):
    from parity import RTOL, compare

    model = (M.PLEnFlux if energy else M.PLPhFlux)(1.0, 10.0)
    d = _one_spectrum()
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=_theta_for(model, 17, 2))
    if energy:
        cfg['theta'][:, 1] *= 1e4  # bring the counts to a sensible range
    err = compare(cfg, specialise=specialise)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err


@KERNELS
@pytest.mark.parametrize('method', ['trapz', 'simpson'])
@pytest.mark.parametrize('comp', M.MULTIPLICATIVE, ids=lambda c: c.__name__)
def test_every_multiplicative_component(lib, comp, method, specialise):
    from parity import RTOL, compare

    mul = comp(method=method) if comp._numerical else comp()
    model = mul * M.PowerLaw()
    d = _one_spectrum()
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=_theta_for(model, 33, 3))
    err = compare(cfg, specialise=specialise)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err


@pytest.mark.parametrize('stat', ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'])
def test_every_statistic_with_edge_channels(lib, stat):
    """zero counts, zero background and clipped source counts exercise every branch of
    the profile backgrounds (likelihood.py:41-133)."""
    from parity import RTOL, compare, gpu_eval, oracle_eval, rel_err_value

    d = _one_spectrum(seed=5)
    d.spec_counts[::7] = 0.0
    d.back_counts[::5] = 0.0
    d.spec_counts[3] = 0.0
    d.back_counts[3] = 0.0
    model = M.PowerLaw()
    theta = _theta_for(model, 40, 4)
    theta[:5, 1] = 1e-9  # tiny norm: source counts well below the data
    cfg = dict(data=[d], model=[model], stat=[stat], theta=theta)
    if stat != 'wstat':
        err = compare(cfg)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, err
        return
    # wstat with n_off == 0 and b == 0: the reference's autodiff goes through
    # xlogy(0, 0) and returns NaN gradients (0 * log 0; torch autograd of the oracle does
    # the same).  The CUDA path returns the finite one-sided derivative instead (SURVEY.md
    # section 7: "match values, do not reproduce NaNs"): check it against central
    # differences of the oracle's VALUE.
    ll_ref, g_ref = oracle_eval(cfg)
    ll, g = gpu_eval(cfg)
    assert rel_err_value(ll, ll_ref) <= RTOL
    assert np.isnan(g_ref).all() and np.isfinite(g).all()
    for j in range(2):
        h = 1e-6 * np.abs(theta[:, j])
        lo = 0
        if j == 1:
            h[:5] = 0.25 * np.abs(theta[:5, j])  # tiny-norm rows: loglike is linear in K there
        else:
            lo = 5  # ... and its alpha-derivative is below the resolution of differences
        tp, tm = theta.copy(), theta.copy()
        tp[:, j] += h
        tm[:, j] -= h
        fd = (oracle_eval(cfg, tp, grad=False)[0] - oracle_eval(cfg, tm, grad=False)[0])[:, 0] / (2 * h)
        assert np.allclose(g[lo:, 0, j], fd[lo:], rtol=2e-5, atol=1e-6 * np.abs(fd).max()), j


@KERNELS
def test_composite_tree_linked_and_fixed(lib, specialise):
    from parity import RTOL, compare

    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    model = M.Constant() * (M.ExpAbs(Ec=0.7) * M.PowerLaw(K=K) + M.HighECut() * (M.Blackbody(K=K) + M.Gauss(El=6.4)))
    d = _one_spectrum(E=90, C=50, seed=9)
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=_theta_for(model, 50, 7))
    err = compare(cfg, specialise=specialise)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err


def test_sites_and_unfold(lib):
    import torch

    import oracle
    from elisa_b200 import Likelihood

    for stat in ('chi2', 'cstat', 'pstat', 'pgstat', 'wstat'):
        d = _one_spectrum(seed=11)
        d.channel_width = np.linspace(0.5, 2.0, d.n_channels)
        model = (M.CutoffPL() + M.Blackbody()).compile()
        theta = _theta_for(model.model, 9, 8)
        lk = Likelihood([d], model, [stat])
        got = lk.sites(0, theta)
        ref = oracle.sites(stat, d.as_mapping(), model.to_spec(), torch.as_tensor(theta))
        names = {'d': 'rate', 'd_Non_model': 'Non_model', 'd_Non_loglike': 'Non_loglike', 'd_loglike': 'loglike',
                 'd_Noff_model': 'Noff_model', 'd_Noff_loglike': 'Noff_loglike'}
        assert set(got) == {k for k, v in names.items() if v in ref}
        for k, t in got.items():
            r = ref[names[k]].numpy()
            assert np.allclose(t.cpu().numpy(), r, rtol=1e-11, atol=1e-11 * np.abs(r).max()), (stat, k)
        u = lk.unfold(0, theta).cpu().numpy()
        u_ref = oracle.eval_model(model.to_spec(), torch.as_tensor(d.photon_egrid), torch.as_tensor(theta)).numpy()
        assert np.allclose(u, u_ref, rtol=1e-12, atol=0)
        lk.close()


@KERNELS
@pytest.mark.parametrize('method', ['trapz', 'simpson'])
@pytest.mark.parametrize('fold', ['i8', 'dmma'])
def test_photon_absorption_in_a_fit_model(lib, method, fold, specialise):
    """PhAbs * PowerLaw (the reference's quick-start model) and TBAbs * (CutoffPL + Gauss)
    with a second, different table: value and gradient incl. d/d nH against the oracle."""
    import os

    from parity import RTOL, compare

    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'photonabs.npz'))
    te, tx = g['table_energy'], g['table_xsect']
    d = _one_spectrum(E=120, C=60, seed=13)
    m1 = M.PhAbs(te, tx, method=method) * M.PowerLaw()
    m2 = M.TBAbs(te, 0.7 * tx * te**0.1, nH=M.UniformParameter('nH', 0.3, 0.0, 10.0), method=method) * (
        M.CutoffPL() + M.Gauss(El=20.0, sigma=3.0)) + M.WAbs(te, tx, nH=0.2) * M.Blackbody()
    for model in (m1, m2):
        cfg = dict(data=[d], model=[model], stat=['cstat'], theta=_theta_for(model, 33, 5))
        err = compare(cfg, specialise=specialise, fold=fold)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (model, err)


def test_row_wise_csr_fold_matches_the_dense_product(lib):
    """elisa_b200_csr_fold_dev (the band-sparse row kernel) against a dense FP64 product, for CSR input
    and for a dense response that is mostly zeros; refused for a genuinely dense one."""
    import torch

    from elisa_b200 import Likelihood, _lib

    cfg = small_config(4)
    d = cfg['data'][0]
    ip, ix, v = d.sparse_matrix
    R = np.zeros((d.n_energies, d.n_channels))
    np.add.at(R, (ix, np.repeat(np.arange(d.n_channels), np.diff(ip))), v)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    u = lk.unfold(0, cfg['theta'][:37])
    x = lk.csr_fold(0, u)
    ref = u @ torch.as_tensor(R, device='cuda')
    assert float(((x - ref).abs() / ref.abs().clamp_min(1e-300)).max()) < 1e-13
    lk.close()
    dense_like = FixedData('d', d.photon_egrid, d.spec_counts, d.area_scale, d.spec_exposure, response_matrix=R,
                           net_counts=d.net_counts, net_errors=d.net_errors)
    lk = Likelihood([dense_like], cfg['model'], cfg['stat'])
    assert torch.equal(lk.csr_fold(0, u), x)
    lk.close()
    full = small_config(1)
    lk = Likelihood(full['data'], full['model'], full['stat'])
    with pytest.raises(_lib.ElisaB200Error, match='dense'):
        lk.csr_fold(0, lk.unfold(0, full['theta'][:4]))
    lk.close()


@pytest.mark.parametrize('fold', [None, 'dmma'])
def test_joint_fit_of_16_spectra_with_per_spectrum_constants(lib, fold, monkeypatch):
    """18 free parameters (a shared power law + one cross-calibration constant per spectrum): more than the
    16 columns round 1 stopped at.  Value and gradient against the oracle with both gradient paths (for the
    tangent-plane path the integer fold's gradient epilogue runs twice, 16 + 2 columns)."""
    from parity import RTOL, compare

    base = synthetic.config5(n=97, S=16, E=128, C=64)
    a = M.UniformParameter('alpha', 1.7, -3.0, 10.0)
    K = M.UniformParameter('K', 10.0, 1e-10, 1e10)
    models = [M.Constant(f=M.UniformParameter(f'f{i}', 1.0, 0.1, 10.0)) * M.PowerLaw(alpha=a, K=K) for i in range(16)]
    rng = np.random.default_rng(3)
    theta = np.column_stack([1.0 + 0.05 * rng.standard_normal((97, 1)), 1.7 + 0.05 * rng.standard_normal((97, 1)),
                             10.0 * (1.0 + 0.05 * rng.standard_normal((97, 1))), 1.0 + 0.05 * rng.standard_normal((97, 15))])
    cfg = dict(base, model=models, theta=theta)
    from elisa_b200 import Likelihood

    lk = Likelihood(cfg['data'], models, cfg['stat'])
    assert lk.nparam == 18 and [p.name for p in lk.free][:4] == ['f0', 'alpha', 'K', 'f1']
    lk.close()
    for mode in ('contract', 'planes'):
        monkeypatch.setenv('ELISA_B200_GRAD', mode)
        err = compare(cfg, fold=fold) if fold else compare(cfg)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (mode, err)
