"""FREE energy shifts on the GPU (missing item 5): ZAShift / ZMShift / VAShift / VMShift
whose z / v is an ordinary fit parameter (models/conv.py:294-300, 335-341, 378-386, 424-432 -- `z` and `v`
are ParamConfig entries like any other, `fixed=True` being only their default).  The kernels evaluate the
wrapped components on dual-number ENERGIES, so d/dz comes with the other derivatives; checked against the
oracle, which follows the reference: model_fn(egrid * factor) under torch autograd, one factor per row."""

import numpy as np
import pytest

from elisa_b200 import Likelihood
from elisa_b200 import models as M

pytestmark = pytest.mark.gpu


def _theta(model, n, seed, shift_spread=0.3):
    from test_gpu_parity import _theta_for

    cm = model.compile()
    theta = _theta_for(model, n, seed)
    rng = np.random.default_rng(seed + 100)
    for j, name in enumerate(cm.params_name):
        if name.endswith('.z'):
            theta[:, j] = 0.5 + shift_spread * rng.standard_normal(n).clip(-1.5, 3.0)
        if name.endswith('.v'):
            theta[:, j] = 4000.0 * rng.standard_normal(n)
    return theta


ADD = [c for c in M.ADDITIVE if c not in (M.PLEnFlux, M.PLPhFlux)]
MUL = [c for c in M.MULTIPLICATIVE if not issubclass(c, M.PhotonAbsorption)]


@pytest.mark.parametrize('method', ['trapz', 'simpson'])
@pytest.mark.parametrize('comp', ADD, ids=lambda c: c.__name__)
def test_zashift_of_every_additive_component(lib, comp, method):
    from parity import RTOL, compare
    from test_gpu_parity import _one_spectrum

    if comp._numerical is False and method == 'simpson':
        pytest.skip('analytic integral: method is ignored')
    z = M.UniformParameter('z', 0.5, -0.9, 5.0)
    model = M.ZAShift(z=z)(comp(method=method) if comp._numerical else comp())
    assert model.compile().params_name[-1] == 'ZAShift.z'
    cfg = dict(data=[_one_spectrum()], model=[model], stat=['cstat'], theta=_theta(model, 33, 1))
    for kw in (dict(), dict(small_batch=False)):  # the one-kernel path and the chunk pipeline
        err = compare(cfg, **kw)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (kw, err)


@pytest.mark.parametrize('shift', ['z', 'v'])
@pytest.mark.parametrize('comp', MUL, ids=lambda c: c.__name__)
def test_shift_of_every_multiplicative_component(lib, comp, shift):
    from parity import RTOL, compare
    from test_gpu_parity import _one_spectrum

    if shift == 'z':
        op = M.ZMShift(z=M.UniformParameter('z', 0.5, -0.9, 5.0))
    else:
        op = M.VMShift(v=M.UniformParameter('v', 100.0, -1e4, 1e4))
    model = op(comp()) * M.PowerLaw()
    cfg = dict(data=[_one_spectrum(seed=3)], model=[model], stat=['wstat'], theta=_theta(model, 33, 2))
    for kw in (dict(), dict(small_batch=False)):
        err = compare(cfg, **kw)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (kw, err)


def test_nested_and_shared_shifts(lib):
    """One z shared by two spectra (a joint fit of one source seen by two detectors), a velocity shift of a line
    inside it, a fixed shift next to the free one, and one component instance inside and outside the shift."""
    from parity import RTOL, compare
    from test_gpu_parity import _one_spectrum

    z = M.UniformParameter('z', 0.5, -0.9, 5.0)
    v = M.UniformParameter('v', 100.0, -1e4, 1e4)
    pl = M.PowerLaw()
    m1 = M.ZAShift(z=z)(M.Band() + M.VAShift(v=v)(M.Gauss())) + pl
    m2 = M.Constant() * M.ZAShift(z=z)(M.ZMShift(z=0.25)(M.ExpAbs()) * pl + M.ZAShift(z=0.1)(M.CutoffPL(method='simpson')))
    lk = Likelihood([_one_spectrum(seed=1), _one_spectrum(E=55, C=31, seed=2)], [m1, m2], ['cstat', 'pgstat'])
    names = lk.params_name
    lk.close()
    assert sum(n.endswith('.z') for n in names) == 1 and sum(n.endswith('.v') for n in names) == 1, names
    rng = np.random.default_rng(9)
    default = []
    for m in (m1.compile(), m2.compile()):
        for p, d in zip(m.free, m.params_default):
            if all(p is not q for q, _ in default):
                default.append((p, d))
    theta = np.array([d for _, d in default]) * (1.0 + 0.08 * rng.standard_normal((40, len(default))))
    for j, n in enumerate(names):
        if n.endswith('.v'):
            theta[:, j] = 3000.0 * rng.standard_normal(40)
    cfg = dict(data=[_one_spectrum(seed=1), _one_spectrum(E=55, C=31, seed=2)], model=[m1, m2], stat=['cstat', 'pgstat'], theta=theta)
    for kw in (dict(), dict(small_batch=False), dict(fold='dmma'), dict(fold='i8')):
        err = compare(cfg, **kw)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (kw, err)


def test_free_shift_equals_fixed_shift_row_by_row(lib):
    """The bins of a FREE shift at z_n are, bit for bit in the formula's terms, the bins of the FIXED shift z = z_n
    -- which tests/golden/conv.npz pins to the reference's own `convolve`."""
    eg = np.geomspace(0.5, 120.0, 193)
    zs = [0.0, 0.7, 2.5]
    z = M.UniformParameter('z', 0.5, -0.9, 5.0)
    free = M.ZAShift(z=z)(M.CutoffPL() + M.Blackbody()).compile()
    base = np.array([1.2, 40.0, 2.0, 3.0, 0.5])
    theta = np.array([np.concatenate([base, [zz]]) for zz in zs])
    out = free.eval(eg, theta)
    for i, zz in enumerate(zs):
        fixed = M.ZAShift(z=zz)(M.CutoffPL() + M.Blackbody()).compile().eval(eg, base[None])[0]
        assert np.allclose(out[i], fixed, rtol=1e-13, atol=0.0), (zz, float(np.max(np.abs(out[i] / fixed - 1.0))))


def test_refusals(lib):
    from test_gpu_parity import _one_spectrum

    z = M.UniformParameter('z', 0.5, -0.9, 5.0)
    with pytest.raises(Exception, match='runtime-specialised'):
        Likelihood([_one_spectrum()], M.ZAShift(z=z)(M.PowerLaw()), ['cstat'], specialise=False)
