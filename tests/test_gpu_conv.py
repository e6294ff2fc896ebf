"""Convolution components on the GPU: ZAShift / ZMShift / VAShift /
VMShift with fixed shift and PhFlux / EnFlux flux normalisations (models/conv.py), against bins
produced by the reference's own `convolve` methods (tests/golden/conv.npz) and, through the
fold and the statistic, against the oracle's value and autograd gradient."""

import os

import numpy as np
import pytest

import conv_cases
from elisa_b200 import FixedData, Likelihood

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
NAMES = sorted(conv_cases.cases())
KERNELS = pytest.mark.parametrize('specialise', [True, False], ids=['nvrtc', 'interp'])


def _data(seed=5, C=48):
    g = conv_cases.grid()
    rng = np.random.default_rng(seed)
    E = g.size - 1
    R = np.abs(rng.standard_normal((E, C))) + 0.05
    counts = rng.poisson(60.0, C).astype(float)
    back = rng.poisson(15.0, C).astype(float)
    return FixedData('d', g, counts, 0.5 + rng.random(C), 2.0, response_matrix=R, net_counts=counts - 0.4 * back,
                     net_errors=np.sqrt(counts + 1.0), back_counts=back, back_errors=np.sqrt(back + 1.0),
                     back_ratio=0.4 + 0.0 * back)


@KERNELS
@pytest.mark.parametrize('name', NAMES)
def test_bins_match_reference_convolve(lib, name, specialise):
    gold = np.load(os.path.join(GOLD, 'conv.npz'))
    model, theta = conv_cases.cases()[name]
    assert np.array_equal(theta, gold[f'{name}/theta'])
    lk = Likelihood([_data()], model, ['cstat'], specialise=specialise)
    assert lk.specialised == [specialise]
    u = lk.unfold(0, theta).cpu().numpy()
    lk.close()
    ref = gold[f'{name}/bins']
    assert np.allclose(u, ref, rtol=1e-11, atol=0.0), float(np.max(np.abs(u / ref - 1.0)))


def test_compiled_model_eval(lib):
    gold = np.load(os.path.join(GOLD, 'conv.npz'))
    for name in ('zashift_of_phflux', 'two_norms_fixed_flux'):
        model, theta = conv_cases.cases()[name]
        out = model.compile().eval(gold['grid'], theta)
        assert np.allclose(out, gold[f'{name}/bins'], rtol=1e-11, atol=0.0), name


@KERNELS
@pytest.mark.parametrize('stat', ['cstat', 'wstat', 'chi2'])
@pytest.mark.parametrize('name', NAMES)
def test_value_and_gradient_against_oracle(lib, name, stat, specialise):
    """Through the fold and the statistic: d/d theta includes d/dF and the derivative of the
    normalising flux with respect to the wrapped model's parameters."""
    from parity import RTOL, compare

    model, theta = conv_cases.cases()[name]
    rng = np.random.default_rng(3)
    theta = np.concatenate([theta, theta[:1] * (1.0 + 0.05 * rng.standard_normal((14, theta.shape[1])))])
    cfg = dict(data=[_data()], model=[model], stat=[stat], theta=theta)
    err = compare(cfg, specialise=specialise)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, (name, stat, err)


def test_joint_fit_with_shared_flux(lib):
    """Two spectra, one model each, sharing the EnFlux parameter and the photon index."""
    from elisa_b200 import models as M
    from parity import RTOL, compare

    F = M.UniformParameter('F', 3e-8, 1e-12, 1e-3, log=True)
    a = M.UniformParameter('alpha', 1.6, -3.0, 10.0)
    m1 = M.EnFlux(2.0, 40.0, F=F)(M.PowerLaw(alpha=a, K=1.0))
    m2 = M.Constant() * M.EnFlux(2.0, 40.0, F=F, ngrid=333)(M.ZAShift(z=0.4)(M.PowerLaw(alpha=a, K=1.0)))
    rng = np.random.default_rng(8)
    theta = np.array([1.6, 3e-8, 1.0]) * (1.0 + 0.1 * rng.standard_normal((20, 3)))
    cfg = dict(data=[_data(1), _data(2, C=33)], model=[m1, m2], stat=['cstat', 'pgstat'], theta=theta)
    for fold in ('i8', 'dmma'):
        err = compare(cfg, fold=fold)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (fold, err)


def test_fit_recovers_flux(lib):
    """simulate_and_fit on a flux-normalised model: the batched LM re-fit works through the
    normalisation (the reference's boot() on PhFlux(...)(PowerLaw))."""
    from elisa_b200 import models as M

    g = np.geomspace(1.0, 50.0, 129)
    E, C = g.size - 1, 64
    R = np.zeros((E, C))
    for e in range(E):
        R[e, max(0, e // 2 - 1): e // 2 + 2] = 40.0
    d = FixedData('d', g, np.ones(C), np.ones(C), 100.0, response_matrix=R)
    model = M.PhFlux(1.0, 50.0)(M.PowerLaw(K=1.0))
    lk = Likelihood([d], model, ['cstat'])
    truth = np.array([1.7, 2.0])
    out = lk.simulate_and_fit(truth, n=400, seed=4)
    th = out['theta'].cpu().numpy()[out['valid'].cpu().numpy()]
    lk.close()
    assert th.shape[0] >= 380
    assert np.all(np.abs(th.mean(axis=0) / truth - 1.0) < 0.01), th.mean(axis=0)


def test_phflux_of_a_powerlaw_is_plphflux(lib):
    """Closed form (see the CPU twin): PhFlux(emin, emax)(PowerLaw) == PLPhFlux(emin, emax)."""
    from elisa_b200 import models as M

    eg = np.geomspace(0.3, 80.0, 101)
    theta = np.array([[1.7, 3.0], [0.4, 12.0], [2.6, 0.5]])
    a = M.PhFlux(1.0, 10.0)(M.PowerLaw(K=1.0)).compile().eval(eg, theta)
    b = M.PLPhFlux(1.0, 10.0).compile().eval(eg, theta)
    assert np.allclose(a, b, rtol=1e-12, atol=0.0)
