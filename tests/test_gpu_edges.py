"""The edge-wise gradient contraction (unfold_contract_edges, csrc/unfold_skel.cuh): for sums of trapezoid
continuum leaves and power laws the specialised kernel contracts G = d loglike / d u with CLOSED-FORM
edge tangents (Comp::basis / Comp::finish) instead of dual numbers.  Checked here against the oracle's
autograd AND against the dual-number contraction of the same library (ELISA_B200_CONTRACT=generic), on
the cases the per-component tests do not reach: both limits of the blackbody branches, shifted and
repeated leaves, fixed and linked parameters, a power law at alpha = 1, clipped bins."""

import ctypes as C

import numpy as np
import pytest

from elisa_b200 import models as M
from test_gpu_parity import _one_spectrum, _theta_for

pytestmark = pytest.mark.gpu


def _free(name, value):
    return M.UniformParameter(name, value, value / 100.0, value * 100.0)


def _cases():
    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    bb = M.Blackbody()
    return {
        'cpl+bb': M.CutoffPL() + M.Blackbody(),
        'gauss+pl': M.Gauss(El=_free('El', 20.0), sigma=_free('sigma', 4.0)) + M.PowerLaw(),
        'pl': M.PowerLaw(),
        'all_closed': M.Compt() + M.OTTB() + M.BlackbodyRad() + M.Gauss(El=_free('El', 30.0), sigma=_free('sigma', 6.0)),
        # Band, OTTS: no closed form -> Dual<NP> on the edge
        'dual_fallback': M.Band(Ec=_free('Ec', 40.0)) + M.OTTS() + M.CutoffPL(),
        # one K in three leaves (one theta column), numbers = fixed parameters
        'linked_and_fixed': M.CutoffPL(K=K, Ec=25.0) + M.Blackbody(K=K) + M.PowerLaw(K=K, alpha=1.7),
        'same_leaf_twice': bb + M.CutoffPL() + bb,
        'shifted': M.ZAShift(z=0.7)(M.CutoffPL() + M.PowerLaw())
                   + M.VAShift(v=20000.0)(M.Gauss(El=_free('El', 10.0), sigma=_free('sigma', 2.0))),
    }


def test_edge_model_is_generated_only_where_it_applies(lib):
    """ModelE appears in the translation unit of plain sums of trapezoid continuum leaves / power laws,
    and nowhere else (products, Simpson's rule, flux normalisations, other analytic integrals)."""
    buf = C.create_string_buffer(1 << 18)

    def src(m):
        cm = m.compile()
        md, _keep = cm.to_desc()
        assert lib.elisa_b200_specialise_check(C.byref(md), max(cm.nparam, 1), buf, len(buf)) > 0
        return buf.value.decode()

    for name, m in _cases().items():
        assert 'unfold_contract_edges<ModelE>' in src(m), name
    # trees of sums and products of such leaves and multiplicative trapezoid leaves: the tree form (ModelT)
    for name, m in _tree_cases().items():
        s = src(m)
        assert 'ModelE' not in s and 'unfold_contract_tree<ModelT>' in s, name
    for m in (M.CutoffPL(method='simpson') + M.Blackbody(), M.BrokenPL() + M.CutoffPL(),
              M.PhFlux(1.0, 10.0, ngrid=100)(M.CutoffPL(K=1.0)), M.PLAbs() * M.CutoffPL(), M.ExpAbs(method='simpson') * M.PowerLaw(),
              M.ZAShift(z=M.UniformParameter('z', 0.5, 0.0, 3.0))(M.ExpAbs() * M.PowerLaw())):
        s = src(m)
        assert 'ModelE' not in s and 'ModelT' not in s and 'unfold_contract_generic<ModelG' in s, repr(m)


def _tree_cases():
    ea = M.ExpAbs()
    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    return {
        'expabs*pl': M.ExpAbs() * M.PowerLaw(),
        'const*(cpl+bb)': M.Constant() * (M.CutoffPL() + M.Blackbody()),
        # one multiplicative leaf in two terms; one K in two additive leaves
        'leaf_in_two_terms': ea * M.PowerLaw(K=K) + ea * M.Blackbody(K=K) + M.Gauss(El=_free('El', 20.0), sigma=_free('sigma', 4.0)),
        'two_factors': M.HighECut(Ec=_free('Ec', 8.0), Ef=_free('Ef', 30.0)) * M.Edge(Ec=_free('Ec', 3.0)) * M.CutoffPL(),
        'dual_leaf_under_product': M.GAbs(El=_free('El', 15.0), sigma=_free('sigma', 3.0)) * M.Band(Ec=_free('Ec', 40.0)),
        'shifted_product': M.ZAShift(z=0.4)(M.ExpFac() * (M.CutoffPL() + M.PowerLaw())) + M.ZMShift(z=1.0)(M.ExpAbs()) * M.OTTB(),
    }


@pytest.mark.parametrize('name', list(_tree_cases()))
@pytest.mark.parametrize('stat', ['cstat', 'wstat'])
def test_tree_form_matches_oracle_and_dual_contraction(lib, name, stat, monkeypatch):
    from parity import RTOL, gpu_eval, oracle_eval, rel_err_grad, rel_err_grad_rownorm, rel_err_value

    model = _tree_cases()[name]
    d = _one_spectrum(E=97, C=40, seed=4)
    theta = _theta_for(model, 65, 9)
    cfg = dict(data=[d], model=[model], stat=[stat], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    out = {}
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        out[mode] = gpu_eval(cfg, specialise=True)
        assert rel_err_value(out[mode][0], ll_ref) <= RTOL, (mode, name)
        assert rel_err_grad(out[mode][1], g_ref) <= RTOL, (mode, name)
    assert np.array_equal(out['edges'][0], out['generic'][0])
    assert rel_err_grad_rownorm(out['edges'][1], out['generic'][1]) <= 1e-11, name


def test_tree_form_with_table_absorption_and_clipped_bins(lib, monkeypatch):
    """PhAbs * (PowerLaw + Blackbody) -- the table leaf reads sigma(E) of its own table on the edge -- and a product whose
    bins underflow below 1e-300 (the clip passes no derivative; a zero weight skips the accumulation)."""
    from parity import RTOL, gpu_eval, oracle_eval, rel_err_grad, rel_err_value

    d = _one_spectrum(E=97, C=40, seed=6)
    rng = np.random.default_rng(2)
    te = np.geomspace(0.05, 500.0, 300)
    tx = 1e-2 * te ** -2.5 * (1.0 + 0.3 * rng.random(300))
    model = M.PhAbs(te, tx) * (M.PowerLaw() + M.Blackbody())
    theta = _theta_for(model, 33, 3)
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        ll, g = gpu_eval(cfg, specialise=True)
        assert rel_err_value(ll, ll_ref) <= RTOL and rel_err_grad(g, g_ref) <= RTOL, mode
    model = M.ExpAbs() * M.CutoffPL() + M.Gauss(El=1.0, sigma=0.05)
    cm = model.compile()
    theta = np.tile(cm.params_default, (6, 1))
    theta[:, list(cm.params_name).index('CutoffPL.Ec')] = [0.05, 0.08, 0.12, 0.2, 0.3, 15.0]
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        ll, g = gpu_eval(cfg, specialise=True)
        assert rel_err_value(ll, ll_ref) <= RTOL and rel_err_grad(g, g_ref) <= RTOL, mode
        assert np.isfinite(g).all()


@pytest.mark.parametrize('name', list(_cases()))
@pytest.mark.parametrize('stat', ['cstat', 'chi2'])
def test_edges_match_oracle_and_dual_contraction(lib, name, stat, monkeypatch):
    from parity import RTOL, gpu_eval, oracle_eval, rel_err_grad, rel_err_grad_rownorm, rel_err_value

    model = _cases()[name]
    d = _one_spectrum(E=97, C=40, seed=3)
    theta = _theta_for(model, 65, 7)
    cfg = dict(data=[d], model=[model], stat=[stat], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    out = {}
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        out[mode] = gpu_eval(cfg, specialise=True)
        assert rel_err_value(out[mode][0], ll_ref) <= RTOL, (mode, name)
        assert rel_err_grad(out[mode][1], g_ref) <= RTOL, (mode, name)
    assert np.array_equal(out['edges'][0], out['generic'][0])  # the values come from the same forward kernels
    assert rel_err_grad_rownorm(out['edges'][1], out['generic'][1]) <= 1e-11, name


def test_blackbody_limits_powerlaw_at_one_and_clipped_bins(lib, monkeypatch):
    """x = E / kT <= 1e-4 (Rayleigh-Jeans branch: dg/dkT = -3 g / kT), x >= 50 (zero), alpha = 1 (no
    alpha-derivative, as jnp.where gives it), and rows whose unfolded model leaves [1e-300, 1e300]
    in some bins (clip(...) of likelihood.py:204 passes no derivative there)."""
    from parity import RTOL, gpu_eval, oracle_eval, rel_err_grad, rel_err_value

    d = _one_spectrum(E=97, C=40, seed=5)  # 0.5 - 200 keV
    model = M.Blackbody() + M.BlackbodyRad() + M.PowerLaw()
    cm = model.compile()
    names = cm.params_name
    theta = np.tile(cm.params_default, (12, 1))
    col = {n: j for j, n in enumerate(names)}
    theta[0:3, col['Blackbody.kT']] = [3e6, 5e6, 1e7]         # x <= 1e-4 on the whole grid
    theta[0:3, col['Blackbody.K']] = 1e18
    theta[3:6, col['BlackbodyRad.kT']] = [3e6, 5e6, 1e7]
    theta[6:9, col['Blackbody.kT']] = [0.004, 0.006, 0.009]    # x >= 50 on the whole grid
    theta[6:9, col['BlackbodyRad.kT']] = [0.004, 0.02, 0.05]   # ... and on the upper part of it
    theta[9:12, col['PowerLaw.alpha']] = 1.0
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        ll, g = gpu_eval(cfg, specialise=True)
        assert rel_err_value(ll, ll_ref) <= RTOL and rel_err_grad(g, g_ref) <= RTOL, mode
    # clipped bins: a cutoff power law with a tiny cutoff energy underflows below 1e-300 in the upper bins
    model = M.CutoffPL() + M.Gauss(El=1.0, sigma=0.05)
    cm = model.compile()
    theta = np.tile(cm.params_default, (6, 1))
    theta[:, list(cm.params_name).index('CutoffPL.Ec')] = [0.05, 0.08, 0.12, 0.2, 0.3, 15.0]
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    got = {}
    for mode in ('edges', 'generic'):
        monkeypatch.setenv('ELISA_B200_CONTRACT', mode)
        got[mode] = gpu_eval(cfg, specialise=True)
        assert rel_err_value(got[mode][0], ll_ref) <= RTOL and rel_err_grad(got[mode][1], g_ref) <= RTOL, mode
    assert np.isfinite(got['edges'][1]).all()
