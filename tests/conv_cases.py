"""Convolution-component cases shared by tools/make_golden.py (which evaluates them with the
reference's own ``convolve`` methods, models/conv.py) and the parity tests."""

import numpy as np


def grid():
    return np.geomspace(0.5, 120.0, 193)


def cases():
    """{name: (model, theta (N, P))}; parameter order = CompiledModel.params_name."""
    from elisa_b200 import models as M

    rng = np.random.default_rng(77)

    def jitter(base, n=3):
        base = np.asarray(base, dtype=np.float64)
        return base * (1.0 + 0.15 * (rng.random((n, base.size)) - 0.5))

    out = {}
    out['zashift_sum'] = (M.ZAShift(z=0.7)(M.CutoffPL() + M.Blackbody()), jitter([1.2, 40.0, 2.0, 3.0, 0.5]))
    out['zmshift_mul'] = (M.ZMShift(z=0.5)(M.ExpAbs() * M.Edge()) * M.PowerLaw(), jitter([2.0, 7.0, 0.8, 1.5, 1.0]))
    out['vashift_gauss_simpson'] = (M.VAShift(v=5000.0)(M.Gauss(method='simpson')), jitter([6.5, 0.4, 2.0]))
    out['vmshift_edge'] = (M.VMShift(v=-3000.0)(M.Edge()) * M.PowerLaw(), jitter([7.0, 1.0, 1.6, 3.0]))
    out['same_component_in_and_out'] = None  # filled below: one component instance inside and outside a shift
    pl = M.PowerLaw()
    out['same_component_in_and_out'] = (pl + M.ZAShift(z=1.5)(pl), jitter([1.7, 2.0]))
    out['phflux_pl'] = (M.PhFlux(1.0, 10.0)(M.PowerLaw(K=1.0)), jitter([1.8, 3.0]))
    out['phflux_linear_grid_sum'] = (M.PhFlux(2.0, 30.0, ngrid=257, elog=False)(M.CutoffPL(K=1.0) + M.Blackbody()),
                                     jitter([1.1, 30.0, 2.5, 0.3, 4.0]))
    out['enflux_zashift_plus_gauss'] = (M.EnFlux(2.0, 50.0, ngrid=300)(M.ZAShift(z=0.3)(M.CutoffPL(K=1.0))) + M.Gauss(),
                                        jitter([1.3, 60.0, 2e-8, 6.4, 0.3, 1.0]))
    out['zashift_of_phflux'] = (M.ZAShift(z=1.0)(M.PhFlux(1.0, 10.0, ngrid=400)(M.CutoffPL(K=1.0) * M.ZMShift(z=1.0)(M.ExpAbs()))),
                                jitter([1.4, 25.0, 1.5, 5.0]))
    out['two_norms_fixed_flux'] = (M.PhFlux(1.0, 10.0, F=2.5)(M.PowerLaw(K=1.0)) +
                                   M.EnFlux(10.0, 100.0, ngrid=500)(M.Blackbody(K=1.0, method='simpson')),
                                   jitter([1.5, 8.0, 3e-9]))
    return out
