"""GPU tests against the committed golden vectors (outputs of the reference's own source,
tools/make_golden.py) -- the CUDA path checked without the oracle in between."""

import os

import numpy as np
import pytest

from elisa_b200 import models as M
from elisa_b200.data import FixedData

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
CLASSES = {c.__name__: c for c in M.ADDITIVE + M.MULTIPLICATIVE}


@pytest.mark.parametrize('name', sorted(CLASSES))
def test_component_eval_matches_reference_source(lib, name):
    gold = np.load(os.path.join(GOLD, 'components.npz'))
    cls = CLASSES[name]
    thetas = gold[f'{name}/theta']
    for gname in ('wide', 'kev'):
        grid = gold[f'grid_{gname}']
        for method in ('trapz', 'simpson'):
            key = f'{name}/{gname}/{method}'
            if key not in gold:
                continue
            # rho of SmoothlyBrokenPL is fixed by default upstream: free every parameter here
            free = {c.name: M.UniformParameter(c.name, c.default, -1e300, 1e300) for c in cls._config}
            args = (1.0, 10.0) if name in ('PLPhFlux', 'PLEnFlux') else ()
            model = cls(*args, method=method, **free).compile()
            with np.errstate(all='ignore'):
                got = model.eval(grid, thetas)
            ref = gold[key]
            scale = np.nanmax(np.abs(ref[np.isfinite(ref)])) if np.isfinite(ref).any() else 1.0
            # E^x is evaluated as exp(x ln E) with tabulated ln E: relative deviation <= |x ln E| 2^-53,
            # amplified by |F| / |dF| in the differences of E^(1-a)/(1-a) (see test_cpu_oracle_golden)
            np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-13 * scale, equal_nan=True, err_msg=key)


def _golden_data(g):
    return FixedData('g', g['photon_egrid'], g['spec_counts'], g['area_scale'], float(g['spec_exposure']),
                     channel_width=g['channel_width'], response_matrix=g['response_matrix'], net_counts=g['net_counts'],
                     net_errors=g['net_errors'], back_counts=g['back_counts'], back_errors=g['back_errors'],
                     back_ratio=g['back_ratio'])


@pytest.mark.parametrize('stat', ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'])
def test_statistic_sites_and_gradient_match_reference_closures(lib, stat):
    from elisa_b200 import Likelihood

    g = np.load(os.path.join(GOLD, 'likelihood.npz'))
    lk = Likelihood([_golden_data(g)], M.CutoffPL() + M.Blackbody(), [stat])
    theta = g['theta']
    sites = lk.sites(0, theta)
    n = 0
    for k, t in sites.items():
        ref = g[f'{stat}/{k}']
        np.testing.assert_allclose(t.cpu().numpy(), ref, rtol=1e-11, atol=1e-12 * np.abs(ref).max(), err_msg=k)
        n += 1
    assert n == (6 if stat in ('pgstat', 'wstat') else 4)
    ll, grad = lk.loglike_and_grad(theta)
    ref_ll = g[f'{stat}/g_loglike'].sum(-1)
    np.testing.assert_allclose(ll.cpu().numpy()[:, 0], ref_ll, rtol=1e-11)
    fd = g[f'{stat}/fd_grad'][1:]
    np.testing.assert_allclose(grad.cpu().numpy()[1:, 0], fd, rtol=5e-6, atol=1e-7 * np.abs(fd).max())
    lk.close()


@pytest.mark.parametrize('cls', ['PhAbs', 'TBAbs', 'WAbs'])
def test_photon_absorption_eval_matches_reference_source(lib, cls):
    """the table component on the GPU against the reference's own `_make_interp` +
    `PhotonAbsorption.continuum` output (tests/golden/photonabs.npz)"""
    g = np.load(os.path.join(GOLD, 'photonabs.npz'))
    for gname in ('inside', 'beyond'):
        for method in ('trapz', 'simpson'):
            comp = getattr(M, cls)(g['table_energy'], g['table_xsect'], method=method)
            out = comp.compile().eval(g[f'grid_{gname}'], g['nH'][:, None])
            ref = g[f'{gname}_{method}']
            assert np.allclose(out, ref, rtol=2e-13, atol=1e-300), (gname, method, np.abs(out - ref).max())
            # the float32 table with repeated energies, as the reference's xsect.hdf5 stores it:
            # float32 logarithms (1e-5 away from the float64 ones) and jnp.interp's dx0 branch
            comp = getattr(M, cls)(g['table32_energy'], g['table32_xsect'], method=method)
            out = comp.compile().eval(g[f'grid_{gname}'], g['nH'][:, None])
            ref = g[f'f32_{gname}_{method}']
            assert np.allclose(out, ref, rtol=2e-13, atol=1e-300), ('f32', gname, method, np.abs(out - ref).max())
