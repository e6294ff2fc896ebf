"""CPU tests: the oracle against the golden vectors produced by executing the reference's
own source (tools/make_golden.py), and against the closed forms the reference's test-suite
pins."""

import os

import numpy as np
import pytest
import torch

import oracle
from oracle.components import ADDITIVE, MULTIPLICATIVE

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
F64 = torch.float64


def _params(name, theta):
    names = (ADDITIVE.get(name) or MULTIPLICATIVE.get(name))[0]
    return {k: torch.tensor([[v]], dtype=F64) for k, v in zip(names, theta)}


@pytest.fixture(scope='module')
def comp_gold():
    return np.load(os.path.join(GOLD, 'components.npz'))


@pytest.fixture(scope='module')
def like_gold():
    return np.load(os.path.join(GOLD, 'likelihood.npz'))


def test_golden_covers_every_builtin(comp_gold):
    assert set(comp_gold['names']) == set(ADDITIVE) | set(MULTIPLICATIVE)


@pytest.mark.parametrize('name', sorted(set(ADDITIVE) | set(MULTIPLICATIVE)))
def test_component_bins_match_reference_source(comp_gold, name):
    thetas = comp_gold[f'{name}/theta']
    for gname in ('wide', 'kev'):
        grid = torch.tensor(comp_gold[f'grid_{gname}'])
        for method in ('trapz', 'simpson'):
            key = f'{name}/{gname}/{method}'
            if key not in comp_gold:
                continue
            gold = comp_gold[key]
            kw = dict(emin=1.0, emax=10.0) if name in ('PLPhFlux', 'PLEnFlux') else {}
            for t, g in zip(thetas, gold):
                with np.errstate(all='ignore'):
                    got = oracle.eval_component(name, grid.unsqueeze(0), _params(name, t), method=method, **kw)[0].numpy()
                # differences of E^(1-a)/(1-a) amplify the last-ulp libm difference by
                # |F| / |dF| (~1e4 on these grids at the default alpha = 1.01)
                scale = np.nanmax(np.abs(g[np.isfinite(g)])) if np.isfinite(g).any() else 1.0
                np.testing.assert_allclose(got, g, rtol=2e-11, atol=1e-14 * scale, equal_nan=True, err_msg=key)


def _data(g):
    keys = ('photon_egrid', 'response_matrix', 'spec_counts', 'net_counts', 'net_errors', 'back_counts', 'back_errors',
            'back_ratio', 'channel_width', 'area_scale')
    d = {k: g[k] for k in keys}
    d['spec_exposure'] = float(g['spec_exposure'])
    return d


SPEC = ('+', ('comp', 'CutoffPL', {'alpha': ('theta', 0), 'Ec': ('theta', 1), 'K': ('theta', 2)}, {'method': 'trapz'}),
        ('comp', 'Blackbody', {'kT': ('theta', 3), 'K': ('theta', 4)}, {'method': 'trapz'}))
SITES = {'g': 'rate', 'g_Non_model': 'Non_model', 'g_Non_loglike': 'Non_loglike', 'g_loglike': 'loglike',
         'g_Noff_model': 'Noff_model', 'g_Noff_loglike': 'Noff_loglike'}


@pytest.mark.parametrize('stat', oracle.STATISTICS)
def test_statistic_sites_match_reference_closures(like_gold, stat):
    theta = torch.tensor(like_gold['theta'])
    got = oracle.sites(stat, _data(like_gold), SPEC, theta)
    seen = 0
    for site, mine in SITES.items():
        key = f'{stat}/{site}'
        if key not in like_gold:
            assert mine not in got or mine == 'loglike'
            continue
        ref = like_gold[key]
        np.testing.assert_allclose(got[mine].numpy(), ref, rtol=1e-12, atol=1e-13 * np.abs(ref).max(), err_msg=key)
        seen += 1
    assert seen >= 4


@pytest.mark.parametrize('stat', oracle.STATISTICS)
def test_autograd_matches_differences_of_reference(like_gold, stat):
    theta = like_gold['theta']
    _, g = oracle.loglike_and_grad([stat], [_data(like_gold)], SPEC, torch.tensor(theta))
    fd = like_gold[f'{stat}/fd_grad']
    g = g.numpy()[1:, 0]  # row 0 sits on the lower clip: its finite differences are noise
    fd = fd[1:]
    if stat == 'wstat':
        # channels with n_off = 0 and b = 0 make autodiff return NaN (xlogy(0, 0)), in
        # JAX as in torch; the CUDA path returns the finite derivative (tests/test_gpu_*)
        assert np.isnan(g).all()
        return
    np.testing.assert_allclose(g, fd, rtol=5e-6, atol=1e-7 * np.abs(fd).max())


def test_profile_backgrounds_match_reference(like_gold):
    s, n_on, n_off = (torch.tensor(like_gold[f'bg/{k}']) for k in ('s', 'n_on', 'n_off'))
    w = oracle.wstat_background(s, n_on, n_off, torch.tensor(0.4, dtype=F64)).numpy()
    np.testing.assert_allclose(w, like_gold['bg/wstat'], rtol=1e-14, atol=0)
    p = oracle.pgstat_background(s, n_on, torch.tensor(like_gold['bg/b_est']), torch.tensor(2.0), torch.tensor(0.4, dtype=F64)).numpy()
    np.testing.assert_allclose(p, like_gold['bg/pgstat'], rtol=1e-14, atol=0)


# ------------------------------------------------------------------ closed forms the reference tests pin
def test_powerlaw_integral_closed_form():
    """tests/conftest.py:23-31 (powerlaw_fn): K (E1^(1-a) - E0^(1-a)) / (1-a), log form at a = 1."""
    eg = np.linspace(1.0, 100.0, 201)
    for alpha in (0.0, 0.5, 1.0, 1.7, 2.0, 3.5):
        got = oracle.eval_component('PowerLaw', torch.tensor(eg).unsqueeze(0),
                                    {'alpha': torch.tensor([[alpha]], dtype=F64), 'K': torch.tensor([[10.0]], dtype=F64)})[0].numpy()
        if alpha == 1.0:
            ref = 10.0 * np.log(eg[1:] / eg[:-1])
        else:
            ref = 10.0 * (eg[1:] ** (1 - alpha) - eg[:-1] ** (1 - alpha)) / (1 - alpha)
        np.testing.assert_allclose(got, ref, rtol=1e-12)


def _trivial_fit():
    """tests/conftest.py:45-71: PowerLaw(alpha = 0 fixed, K = 10), 200 linear bins 1-100 keV,
    identity response, exposure 50, Poisson counts, seed 42."""
    eg = np.linspace(1.0, 100.0, 201)
    de = np.diff(eg)
    expo = 50.0
    counts = np.random.default_rng(42).poisson(10.0 * de * expo).astype(float)
    data = dict(photon_egrid=eg, response_matrix=np.eye(200), area_scale=np.ones(200), spec_exposure=expo,
                channel_width=de, spec_counts=counts)
    spec = ('comp', 'PowerLaw', {'alpha': ('const', 0.0), 'K': ('theta', 0)}, {})
    return data, spec, counts, de, expo


def test_trivial_cstat_fit_closed_form():
    """tests/test_fit.py:24-42: K_hat = mean(counts / dE / exposure); sigma = sqrt(K_hat / nbins / exposure / dE)."""
    data, spec, counts, de, expo = _trivial_fit()
    k_hat = np.mean(counts / de / expo)
    ll, g = oracle.loglike_and_grad(['cstat'], [data], spec, torch.tensor([[k_hat]]))
    assert abs(g[0, 0, 0].item()) < 1e-9 * counts.sum()  # stationary at the closed-form MLE
    h = 1e-3 * k_hat
    gp = oracle.loglike_and_grad(['cstat'], [data], spec, torch.tensor([[k_hat + h]]))[1][0, 0, 0].item()
    gm = oracle.loglike_and_grad(['cstat'], [data], spec, torch.tensor([[k_hat - h]]))[1][0, 0, 0].item()
    sigma = 1.0 / np.sqrt(-(gp - gm) / (2 * h))
    assert np.isclose(sigma, np.sqrt(k_hat / 200 / expo / de[0]), rtol=1e-5)


def test_deviance_identities():
    """-2 loglike of cstat is the C-stat deviance 2 sum[mu - n + n ln(n / mu)]
    (plot/residuals.py:406-427); of chi2 it is sum ((n - mu) / sigma)^2."""
    data, spec, counts, de, expo = _trivial_fit()
    theta = torch.tensor([[8.0], [10.0], [13.0]])
    mu = theta.numpy() * de * expo
    ll = oracle.loglike(['cstat'], [data], spec, theta).numpy()[:, 0]
    with np.errstate(all='ignore'):
        term = np.where(counts > 0, counts * np.log(counts / mu), 0.0)
    np.testing.assert_allclose(-2 * ll, 2 * np.sum(mu - counts + term, axis=1), rtol=1e-12)
    data2 = dict(data, net_counts=counts, net_errors=np.sqrt(counts + 1.0))
    ll2 = oracle.loglike(['chi2'], [data2], spec, theta).numpy()[:, 0]
    np.testing.assert_allclose(-2 * ll2, np.sum(((counts - mu) / np.sqrt(counts + 1.0)) ** 2, axis=1), rtol=1e-12)


def test_profile_background_is_stationary_in_the_interior():
    """wstat / pgstat: the profiled b zeroes dL/db where no boundary branch is active."""
    s = torch.tensor([3.0, 8.0, 40.0], dtype=F64)
    n_on = torch.tensor([5.0, 12.0, 30.0], dtype=F64)
    n_off = torch.tensor([7.0, 9.0, 25.0], dtype=F64)
    a = torch.tensor(0.4, dtype=F64)
    b = oracle.wstat_background(s, n_on, n_off, a).clone().requires_grad_(True)
    L = (torch.special.xlogy(n_on, s + a * b) - (s + a * b) + torch.special.xlogy(n_off, b) - b).sum()
    (db,) = torch.autograd.grad(L, b)
    assert torch.all(db.abs() < 1e-12)
    sig = torch.tensor(2.0, dtype=F64)
    b = oracle.pgstat_background(s, n_on, n_off, sig, a).clone().requires_grad_(True)
    L = (torch.special.xlogy(n_on, s + a * b) - (s + a * b) - 0.5 * ((n_off - b) / sig) ** 2).sum()
    (db,) = torch.autograd.grad(L, b)
    assert torch.all(db.abs() < 1e-12)


def test_every_builtin_is_finite_on_the_reference_grid():
    """tests/test_model.py:178-195: all add / mul built-ins evaluate finite on geomspace(1e-4, 1e9, 1301)."""
    grid = torch.tensor(np.geomspace(1e-4, 1e9, 1301)).unsqueeze(0)
    for table in (ADDITIVE, MULTIPLICATIVE):
        for name, (pnames, defaults, *_rest) in table.items():
            kw = dict(emin=1.0, emax=10.0) if name in ('PLPhFlux', 'PLEnFlux') else {}
            p = {k: torch.tensor([[v]], dtype=F64) for k, v in zip(pnames, defaults)}
            v = oracle.eval_component(name, grid, p, **kw)
            assert torch.isfinite(v).all(), name


def test_photon_absorption_matches_reference_source():
    """PhAbs / TBAbs / WAbs: `_make_interp` + `PhotonAbsorption.continuum` of the reference
    (models/mul.py:274-283, 360-388) executed on a synthetic cross-section table, inside and
    beyond the table's energy range (log-log extrapolation), trapz and Simpson."""
    g = np.load(os.path.join(GOLD, 'photonabs.npz'))
    nH = torch.as_tensor(g['nH']).unsqueeze(-1)
    for gname in ('inside', 'beyond'):
        eg = torch.as_tensor(g[f'grid_{gname}'])
        for method in ('trapz', 'simpson'):
            out = oracle.components.eval_component('PhotonAbsorption', eg, {'nH': nH}, method=method,
                                                   table_energy=g['table_energy'], table_xsect=g['table_xsect'])
            ref = g[f'{gname}_{method}']
            assert np.allclose(out.numpy(), ref, rtol=1e-13, atol=1e-300), (gname, method)
            # float32 table with repeated energies (the storage of the reference's xsect.hdf5)
            out = oracle.components.eval_component('PhotonAbsorption', eg, {'nH': nH}, method=method,
                                                   table_energy=g['table32_energy'], table_xsect=g['table32_xsect'])
            ref = g[f'f32_{gname}_{method}']
            assert np.allclose(out.numpy(), ref, rtol=1e-13, atol=1e-300), ('f32', gname, method)
            assert not np.allclose(ref, g[f'{gname}_{method}'], rtol=1e-7, atol=0)  # the float32 log matters


def _conv_cases():
    import conv_cases

    return conv_cases


@pytest.mark.parametrize('name', ['zashift_sum', 'zmshift_mul', 'vashift_gauss_simpson', 'vmshift_edge',
                                  'same_component_in_and_out', 'phflux_pl', 'phflux_linear_grid_sum',
                                  'enflux_zashift_plus_gauss', 'zashift_of_phflux', 'two_norms_fixed_flux'])
def test_convolution_components_match_reference_source(name):
    """ZAShift / ZMShift / VAShift / VMShift / PhFlux / EnFlux: the oracle against bins produced
    by the reference's own `convolve` methods (models/conv.py, tools/make_golden.py --conv)."""
    from oracle.likelihood import eval_model

    cc = _conv_cases()
    g = np.load(os.path.join(GOLD, 'conv.npz'))
    model, theta = cc.cases()[name]
    assert np.array_equal(theta, g[f'{name}/theta']), 'tests/conv_cases.py changed: re-run tools/make_golden.py --conv'
    out = eval_model(model.compile().to_spec(), torch.as_tensor(g['grid']), torch.as_tensor(theta)).numpy()
    ref = g[f'{name}/bins']
    assert out.shape == ref.shape and np.all(np.isfinite(ref))
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-300), float(np.max(np.abs(out / ref - 1.0)))


def test_convolution_host_rules():
    """Type rules and parameter order of the reference (models/model.py:1884-1923)."""
    from elisa_b200 import models as M

    with pytest.raises(TypeError):
        M.ZAShift()(M.ExpAbs())  # additive-only
    with pytest.raises(TypeError):
        M.ZMShift()(M.PowerLaw())  # multiplicative-only
    with pytest.raises(TypeError):
        M.PhFlux(1.0, 10.0)(M.ExpAbs())
    with pytest.raises(TypeError):
        M.PowerLaw() + M.ZAShift()  # not applied to a model yet
    with pytest.raises(ValueError):
        M.PhFlux(10.0, 1.0)
    with pytest.raises(NotImplementedError):
        M.PhFlux(1.0, 10.0)(M.EnFlux(1.0, 10.0)(M.PowerLaw(K=1.0))).compile()  # nested normalisation
    # a FREE shift parameter is a column of theta after the wrapped model's own (model.py:1917-1919) and a
    # dual-number factor of the leaves below it; ZAShift's division reaches the additive leaves only
    from elisa_b200 import _lib

    z = M.UniformParameter('z', 0.5, 0.0, 3.0)
    v = M.UniformParameter('v', 100.0, -1e4, 1e4)
    cm = M.ZAShift(z=z)(M.VMShift(v=v)(M.ExpAbs()) * M.PowerLaw() + M.ZAShift(z=0.25)(M.CutoffPL())).compile()
    assert cm.params_name == ['ExpAbs.Ec', 'VMShift.v', 'PowerLaw.alpha', 'PowerLaw.K', 'CutoffPL.alpha', 'CutoffPL.Ec', 'CutoffPL.K',
                              'ZAShift_2.z']  # the inner, fixed ZAShift is 'ZAShift'
    kinds = [[(k, p.name) for k, p in e[6]] for e in cm.entries]
    assert kinds == [[(_lib.SHIFT_Z, 'z'), (_lib.SHIFT_V, 'v')], [(_lib.SHIFT_ZA, 'z')], [(_lib.SHIFT_ZA, 'z')]]
    assert cm.entries[2][1:3] == (1.25, 1.0 / 1.25) and cm.entries[0][1:3] == (1.0, 1.0)  # the fixed shift stays a scale
    md, _keep = cm.to_desc()
    assert [md.components[0].shift_kind[k] for k in range(2)] == [_lib.SHIFT_Z, _lib.SHIFT_V]
    assert [md.components[0].shift_src[k] for k in range(2)] == [7, 1]
    assert md.components[2].shift_kind[0] == _lib.SHIFT_ZA and md.components[2].shift_kind[1] == 0
    with pytest.raises(NotImplementedError):  # more than two nested free shifts
        z2, z3 = M.UniformParameter('z', 0.1, 0.0, 3.0), M.UniformParameter('z', 0.2, 0.0, 3.0)
        M.ZAShift(z=z)(M.ZAShift(z=z2)(M.ZAShift(z=z3)(M.PowerLaw()))).compile()
    with pytest.raises(NotImplementedError):  # together with a flux normalisation
        M.PhFlux(1.0, 10.0)(M.ZAShift(z=z)(M.PowerLaw(K=1.0))).compile()
    cm = M.EnFlux(1.0, 10.0)(M.ZAShift(z=0.2)(M.CutoffPL(K=1.0))).compile()
    # the wrapped model's parameters first, then the convolution's own; fixed z is not a column
    assert cm.params_name == ['CutoffPL.alpha', 'CutoffPL.Ec', 'EnFlux.F']
    assert cm.ops == [0, -16] and cm.entries[0][1:5] == (1.2, 1.0 / 1.2, 1.2, 1.0 / 1.2)
    pl = M.PowerLaw()
    cm = (pl + M.ZAShift(z=1.0)(pl)).compile()
    assert cm.nparam == 2 and len(cm.entries) == 2 and cm.ops == [0, 1, -1]


REF_TABLES = '/root/reference/src/elisa/models/tables/xsect.hdf5'


def test_hdf5_reader_rejects_what_it_does_not_read(tmp_path):
    from elisa_b200.tables import read_hdf5

    p = tmp_path / 'x.hdf5'
    p.write_bytes(b'not an hdf5 file at all')
    with pytest.raises(ValueError):
        read_hdf5(p)
    p.write_bytes(b'\x89HDF\r\n\x1a\n' + bytes([2]) + bytes(64))  # version-2 superblock
    with pytest.raises(NotImplementedError):
        read_hdf5(p)


@pytest.mark.skipif(not os.path.exists(REF_TABLES), reason='the reference tree is only present in the build container')
def test_hdf5_reader_on_the_reference_tables():
    """models/tables/xsect.hdf5 read without h5py: 38 float32 datasets; the energy dataset is
    checked against an independent reconstruction from the recipe in
    tables/generate_xsect_table.py (geomspace(0.1, 20, 10000) refined around three edges)."""
    from elisa_b200 import models as M
    from elisa_b200.tables import load_xsect, read_hdf5

    flat = read_hdf5(REF_TABLES)
    assert len(flat) == 38 and all(a.dtype == np.float32 and a.shape == flat['energy'].shape for a in flat.values())
    e = flat['energy'].astype(np.float64)
    eg = np.geomspace(0.1, 20.0, 10000)
    extra = [np.geomspace(eg[i], eg[j], (j - i) * 5 + 1) for i, j in [[3123, 3155], [3182, 3219], [3202, 3215]]]
    eg = np.unique(np.hstack([eg, *extra]))
    idx = np.clip(np.searchsorted(eg, e), 1, eg.size - 1)
    near = np.where(np.abs(eg[idx] - e) < np.abs(eg[idx - 1] - e), eg[idx], eg[idx - 1])
    assert np.max(np.abs(near / e - 1.0)) < 1e-7  # every stored energy is a float32 rounding of the recipe's
    idx = np.clip(np.searchsorted(e, eg), 1, e.size - 1)
    near = np.where(np.abs(e[idx] - eg) < np.abs(e[idx - 1] - eg), e[idx], e[idx - 1])
    assert np.max(np.abs(near / eg - 1.0)) < 1e-7  # and every recipe energy is stored
    assert np.all(np.diff(e) >= 0) and np.sum(np.diff(e) == 0) == 87
    t = load_xsect(REF_TABLES)
    assert sorted(k for k in t if k != 'energy') == ['phabs', 'tbabs', 'wabs']
    assert sorted(t['phabs']) == ['bcmc', 'obcm', 'vern'] and len(t['phabs']['vern']) == 9
    assert all(np.all(a > 0) for m in ('phabs', 'tbabs', 'wabs') for xs in t[m].values() for a in xs.values())
    # photo-electric absorption: at 0.1 keV of order 1e2 - 1e3 x 1e-22 cm^2 per H, four decades less at 20 keV
    s = t['phabs']['vern']['angr']
    assert 1e2 < s[0] < 1e4 and 1e-4 < s[-1] < 1e-1
    for cls in (M.PhAbs, M.TBAbs, M.WAbs):  # the reference's default abund / xsect of each class
        c = cls.from_tables(REF_TABLES)
        assert c.table_energy.dtype == np.float32 and c.table_log_energy.dtype == np.float64
        assert np.array_equal(c.table_log_energy, np.log(flat['energy']).astype(np.float64))
    with pytest.raises(ValueError):
        M.PhAbs.from_tables(REF_TABLES, abund='nope')


def test_phflux_of_a_powerlaw_is_plphflux():
    """Closed form: the bins of a power law are exact integrals, so their sum over the flux grid
    is the exact band flux and PhFlux(emin, emax)(PowerLaw) must equal PLPhFlux(emin, emax)
    (the identity behind the reference's tests/test_model.py::test_simulation)."""
    from elisa_b200 import models as M
    from oracle.likelihood import eval_model

    eg = torch.as_tensor(np.geomspace(0.3, 80.0, 101))
    theta = torch.tensor([[1.7, 3.0], [0.4, 12.0], [2.6, 0.5]], dtype=torch.float64)
    a = eval_model(M.PhFlux(1.0, 10.0)(M.PowerLaw(K=1.0)).compile().to_spec(), eg, theta).numpy()
    b = eval_model(M.PLPhFlux(1.0, 10.0).compile().to_spec(), eg, theta).numpy()
    assert np.allclose(a, b, rtol=1e-12, atol=0.0)


def test_oracle_free_shift_rows_equal_fixed_shift():
    """A FREE z / v in the oracle (one factor per row, models/conv.py:294-300 under the reference's vmap over rows)
    gives, row by row, what the FIXED shift gives -- the form tests/golden/conv.npz pins to the reference."""
    import sys

    sys.path.insert(0, os.path.dirname(__file__))
    from parity import oracle_eval
    from elisa_b200 import FixedData
    from elisa_b200 import models as M

    rng = np.random.default_rng(4)
    g = np.geomspace(0.5, 120.0, 97)
    C = 24
    R = np.abs(rng.standard_normal((g.size - 1, C))) + 0.05
    counts = rng.poisson(60.0, C).astype(float)
    d = FixedData('d', g, counts, 0.5 + rng.random(C), 2.0, response_matrix=R)
    base = np.array([1.2, 40.0, 2.0, 6.4, 0.3, 1.5])
    for zz, vv in ((0.0, 0.0), (0.7, -3000.0), (2.5, 9000.0)):
        z = M.UniformParameter('z', 0.5, -0.9, 5.0)
        v = M.UniformParameter('v', 100.0, -1e4, 1e4)
        free = M.ZAShift(z=z)(M.CutoffPL() + M.VAShift(v=v)(M.Gauss()))
        fixed = M.ZAShift(z=zz)(M.CutoffPL() + M.VAShift(v=vv)(M.Gauss()))
        names = free.compile().params_name
        assert names[-2:] == ['VAShift.v', 'ZAShift.z'], names
        ll_free, g_free = oracle_eval(dict(data=[d], model=[free], stat=['cstat'], theta=np.concatenate([base, [vv, zz]])[None]))
        ll_fix, g_fix = oracle_eval(dict(data=[d], model=[fixed], stat=['cstat'], theta=base[None]))
        assert np.allclose(ll_free, ll_fix, rtol=1e-13, atol=0.0)
        assert np.allclose(g_free[..., :6], g_fix, rtol=1e-11, atol=1e-9)
        # d/dz against a central difference of the fixed-shift value
        h = 1e-5
        lp, _ = oracle_eval(dict(data=[d], model=[M.ZAShift(z=zz + h)(M.CutoffPL() + M.VAShift(v=vv)(M.Gauss()))], stat=['cstat'],
                                 theta=base[None]), grad=False)
        lm, _ = oracle_eval(dict(data=[d], model=[M.ZAShift(z=zz - h)(M.CutoffPL() + M.VAShift(v=vv)(M.Gauss()))], stat=['cstat'],
                                 theta=base[None]), grad=False)
        fd = (lp - lm) / (2 * h)
        assert np.allclose(g_free[..., 7], fd, rtol=2e-6), (g_free[..., 7], fd)
