"""Oracle restatement of the reference's built-in spectral components (torch, CPU, float64).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Every function follows the cited
reference lines op for op (``jnp.X`` -> ``torch.X``), quirks included, so that
torch autograd differentiates the same graph ``jax.grad`` would.

Conventions: ``egrid`` is a float64 tensor whose LAST axis runs over energy edges
(E+1 of them); each parameter is a scalar tensor or a tensor broadcastable against
``egrid`` with a trailing singleton axis (shape (N, 1) for a batch of N parameter
vectors -- the stand-in for ``jax.vmap`` over params, models/model.py:406-411).
"""

from __future__ import annotations

import math

import numpy as np
import torch

F64 = torch.float64


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=F64)


# --------------------------------------------------------------------------- integrators
def _powerlaw_integral(egrid, alpha):
    """models/add.py:40-50."""
    alpha = _t(alpha)
    cond = alpha != 1.0
    one_minus_alpha = torch.where(cond, 1.0 - alpha, torch.ones_like(alpha))
    f1 = torch.pow(egrid, one_minus_alpha) / one_minus_alpha
    f1 = f1[..., 1:] - f1[..., :-1]
    f2 = torch.log(egrid)
    f2 = f2[..., 1:] - f2[..., :-1]
    return torch.where(cond, f1, f2)


def _integrate(continuum, egrid, params, mtype, method):
    """NumericalIntegral._make_integral, models/model.py:1731-1761."""
    if method == 'trapz':
        factor = 0.5 * (egrid[..., 1:] - egrid[..., :-1]) if mtype == 'add' else 0.5
        f_grid = continuum(egrid, params)
        return factor * (f_grid[..., :-1] + f_grid[..., 1:])
    if method == 'simpson':
        factor = (egrid[..., 1:] - egrid[..., :-1]) / 6.0 if mtype == 'add' else 1.0 / 6.0
        e_mid = 0.5 * (egrid[..., :-1] + egrid[..., 1:])
        f_grid = continuum(egrid, params)
        f_mid = continuum(e_mid, params)
        return factor * (f_grid[..., :-1] + 4.0 * f_mid + f_grid[..., 1:])
    raise NotImplementedError(f"integration method '{method}'")


# --------------------------------------------------------------------------- additive continua
def _band(egrid, p):
    """Band.continuum, models/add.py:100-121."""
    alpha, beta, Ec, K = p['alpha'], p['beta'], p['Ec'], p['K']
    e0 = 100.0
    amb_ = alpha - beta
    inv_Ec = 1.0 / Ec
    amb = torch.where(amb_ < inv_Ec, inv_Ec, amb_)
    Ebreak = Ec * amb
    log = torch.where(
        egrid < Ebreak,
        alpha * torch.log(egrid / e0) - egrid / Ec,
        amb * torch.log(amb * Ec / e0) - amb + beta * torch.log(egrid / e0),
    )
    return K * torch.exp(log)


def _bandep(egrid, p):
    """BandEp.continuum, models/add.py:177-200."""
    alpha, beta, Ep, K = p['alpha'], p['beta'], p['Ep'], p['K']
    e0 = 100.0
    Ec = Ep / (2.0 + alpha)
    Ebreak = (alpha - beta) * Ec
    amb_ = alpha - beta
    inv_Ec = 1.0 / Ec
    amb = torch.where(amb_ < inv_Ec, inv_Ec, amb_)
    log = torch.where(
        egrid < Ebreak,
        alpha * torch.log(egrid / e0) - egrid / Ec,
        amb * torch.log(amb * Ec / e0) - amb + beta * torch.log(egrid / e0),
    )
    return K * torch.exp(log)


def _bentpl(egrid, p):
    """BentPL.continuum, models/add.py:231-236."""
    return p['K'] * 2.0 / (1.0 + torch.pow(egrid / p['Eb'], p['alpha']))


def _blackbody(egrid, p):
    """Blackbody.continuum, models/add.py:266-286."""
    kT, K = p['kT'], p['K']
    x = egrid / kT
    tmp = 8.0525 * K * egrid / (kT * kT * kT)
    x_ = torch.where(x >= 50.0, torch.ones_like(x), x)
    return torch.where(
        x <= 1e-4,
        tmp,
        torch.where(x >= 50.0, torch.zeros_like(x), tmp * x / torch.expm1(x_)),
    )


def _blackbodyrad(egrid, p):
    """BlackbodyRad.continuum, models/add.py:317-338."""
    kT, K = p['kT'], p['K']
    x = egrid / kT
    tmp = 1.0344e-3 * K * egrid
    x_ = torch.where(x >= 50.0, torch.ones_like(x), x)
    return torch.where(
        x <= 1e-4,
        tmp * kT,
        torch.where(x >= 50.0, torch.zeros_like(x), tmp * egrid / torch.expm1(x_)),
    )


def _cutoffpl(egrid, p):
    """CutoffPL.continuum, models/add.py:434-439."""
    return p['K'] * torch.pow(egrid, -p['alpha']) * torch.exp(-egrid / p['Ec'])


def _compt(egrid, p):
    """Compt.continuum, models/add.py:473-479."""
    neg_inv_Ec = (p['alpha'] - 2.0) / p['Ep']
    return p['K'] * torch.pow(egrid, -p['alpha']) * torch.exp(egrid * neg_inv_Ec)


def _gauss(egrid, p):
    """Gauss.continuum, models/add.py:511-516; jax.scipy.stats.norm.pdf = exp(logpdf),
    logpdf = -(log(2 pi sigma^2) + (x - loc)^2 / sigma^2) / 2."""
    El, sigma, K = p['El'], p['sigma'], p['K']
    scale_sqrd = sigma * sigma
    log_normalizer = torch.log(2.0 * math.pi * scale_sqrd)
    quadratic = (egrid - El) ** 2 / scale_sqrd
    return K * torch.exp(-(log_normalizer + quadratic) / 2.0)


def _ottb(egrid, p):
    """OTTB.continuum, models/add.py:590-594."""
    return p['K'] * torch.exp((1.0 - egrid) / p['kT']) / egrid


def _otts(egrid, p):
    """OTTS.continuum, models/add.py:625-629."""
    return p['K'] * torch.exp(-torch.pow(egrid / p['Ec'], 1.0 / 3.0))


def _sbpl(egrid, p):
    """SmoothlyBrokenPL.continuum, models/add.py:846-861."""
    alpha1, Eb, alpha2, rho, K = p['alpha1'], p['Eb'], p['alpha2'], p['rho'], p['K']
    e = egrid / Eb
    x = (alpha2 - alpha1) / rho * torch.log(e)
    threshold = 30
    alpha = torch.where(x > threshold, alpha2 + 0.0 * x, alpha1 + 0.0 * x)
    r = torch.where(torch.abs(x) > threshold, torch.full_like(x, 0.5), 0.5 * (1.0 + torch.exp(x)))
    return K * torch.pow(e, -alpha) * torch.pow(r, -rho)


# --------------------------------------------------------------------------- additive integrals
def _powerlaw(egrid, p, **_):
    """PowerLaw.integral, models/add.py:655-657."""
    return p['K'] * _powerlaw_integral(egrid, p['alpha'])


def _brokenpl(egrid, p, **_):
    """BrokenPL.integral, models/add.py:378-393.  The break-bin fix-up on line 392
    (``f.at[idx].set(...)``) discards its result, so it is NOT applied here either;
    the integral is taken on ``egrid / Eb`` (lines 385-387)."""
    alpha1, Eb, alpha2, K = p['alpha1'], p['Eb'], p['alpha2'], p['K']
    mask = egrid[..., :-1] <= Eb
    egrid_ = egrid / Eb
    p1 = _powerlaw_integral(egrid_, alpha1)
    p2 = _powerlaw_integral(egrid_, alpha2)
    return K * torch.where(mask, p1, p2)


def _lorentz(egrid, p, **_):
    """Lorentz.integral, models/add.py:553-561."""
    El, sigma, K = p['El'], p['sigma'], p['K']
    x = 2.0 * (egrid - El) / sigma
    integral = torch.arctan(x)
    norm = 2.0 * K / (math.pi + 2 * torch.arctan(2 * El / sigma))
    return norm * (integral[..., 1:] - integral[..., :-1])


def _plfluxnorm(egrid, p, energy, emin, emax, **_):
    """PLFluxNorm.eval, models/add.py:682-708 (PLPhFlux: energy=False; PLEnFlux: True)."""
    band = torch.tensor([float(emin), float(emax)], dtype=F64)
    if energy:
        keV_to_erg = 1.602176634e-9
        f = _powerlaw_integral(band, p['alpha'] - 1.0)
        factor = p['F'] / (f * keV_to_erg)
    else:
        f = _powerlaw_integral(band, p['alpha'])
        factor = p['F'] / f
    return factor * _powerlaw_integral(egrid, p['alpha'])


# --------------------------------------------------------------------------- multiplicative
def _constant(egrid, p, **_):
    """Constant.integral, models/mul.py:54-56."""
    f = _t(p['f'])
    return f + torch.zeros_like(egrid[..., 1:])


def _edge(egrid, p):
    """Edge.continuum, models/mul.py:88-94."""
    Ec, D = p['Ec'], p['D']
    return torch.where(egrid >= Ec, torch.exp(-D * torch.pow(egrid / Ec, 3.0)), torch.ones_like(egrid * Ec))


def _expabs(egrid, p):
    """ExpAbs.continuum, models/mul.py:116-118."""
    return torch.exp(-p['Ec'] / egrid)


def _expfac(egrid, p):
    """ExpFac.continuum, models/mul.py:154-159."""
    A, f, Ec = p['A'], p['f'], p['Ec']
    return torch.where(egrid >= Ec, 1.0 + A * torch.exp(-f * egrid), torch.ones_like(egrid * Ec))


def _gabs(egrid, p):
    """GAbs.continuum, models/mul.py:195-204."""
    El, sigma, tau = p['El'], p['sigma'], p['tau']
    return torch.exp(
        -tau / (math.sqrt(2 * math.pi) * sigma) * torch.exp(-0.5 * torch.pow((egrid - El) / sigma, 2))
    )


def _highecut(egrid, p):
    """HighECut.continuum, models/mul.py:236-240."""
    Ec, Ef = p['Ec'], p['Ef']
    return torch.where(egrid >= Ec, torch.exp((Ec - egrid) / Ef), torch.ones_like(egrid * Ec))


def _plabs(egrid, p, **_):
    """PLAbs.integral, models/mul.py:266-271 (alpha = 1 is ignored there too)."""
    one_minus_alpha = 1.0 - p['alpha']
    f = p['K'] / one_minus_alpha * torch.pow(egrid, one_minus_alpha)
    return (f[..., 1:] - f[..., :-1]) / (egrid[..., 1:] - egrid[..., :-1])


def _make_interp(table_energy, table_xsect):
    """_make_interp, models/mul.py:274-283: exp(jnp.interp(log e, log energy, log xsect,
    left='extrapolate', right='extrapolate')).  jnp.interp: i = clip(searchsorted(xp, x,
    'right'), 1, n - 1); f = fp[i-1] + (x - xp[i-1]) / (xp[i] - xp[i-1]) * (fp[i] - fp[i-1]);
    'extrapolate' keeps that line outside [xp[0], xp[-1]]."""
    # np.log of the arrays AS STORED (mul.py:275-276): float32 tables give float32 logarithms,
    # which jnp.interp then promotes to float64
    xp = torch.as_tensor(np.log(np.asarray(table_energy)).astype(np.float64))
    fp = torch.as_tensor(np.log(np.asarray(table_xsect)).astype(np.float64))
    eps = float(np.spacing(np.finfo(np.float64).eps))

    def interp(e):
        x = torch.log(e)
        i = torch.clamp(torch.searchsorted(xp, x.contiguous(), right=True), 1, xp.numel() - 1)
        df = fp[i] - fp[i - 1]
        dx = xp[i] - xp[i - 1]
        dx0 = dx.abs() <= eps  # repeated table energies (jnp.interp's dx0 branch)
        return torch.exp(torch.where(dx0, fp[i - 1], fp[i - 1] + ((x - xp[i - 1]) / torch.where(dx0, 1.0, dx)) * df))

    return interp


def _photonabs(egrid, p, sigma_of=None):
    """PhotonAbsorption.continuum, models/mul.py:360-388."""
    return torch.exp(-p['nH'] * sigma_of(egrid))


# name -> (parameter names in _config order, defaults, 'continuum'|'integral', function, extra kwargs)
ADDITIVE = {
    'Band': (('alpha', 'beta', 'Ec', 'K'), (-1.0, -2.0, 300.0, 1.0), 'continuum', _band, {}),
    'BandEp': (('alpha', 'beta', 'Ep', 'K'), (-1.0, -2.0, 300.0, 1.0), 'continuum', _bandep, {}),
    'BentPL': (('alpha', 'Eb', 'K'), (1.5, 5.0, 1.0), 'continuum', _bentpl, {}),
    'Blackbody': (('kT', 'K'), (3.0, 1.0), 'continuum', _blackbody, {}),
    'BlackbodyRad': (('kT', 'K'), (3.0, 1.0), 'continuum', _blackbodyrad, {}),
    'BrokenPL': (('alpha1', 'Eb', 'alpha2', 'K'), (1.0, 5.0, 2.0, 1.0), 'integral', _brokenpl, {}),
    'Compt': (('alpha', 'Ep', 'K'), (1.0, 15.0, 1.0), 'continuum', _compt, {}),
    'CutoffPL': (('alpha', 'Ec', 'K'), (1.0, 15.0, 1.0), 'continuum', _cutoffpl, {}),
    'Gauss': (('El', 'sigma', 'K'), (6.5, 0.1, 1.0), 'continuum', _gauss, {}),
    'Lorentz': (('El', 'sigma', 'K'), (6.5, 0.1, 1.0), 'integral', _lorentz, {}),
    'OTTB': (('kT', 'K'), (30.0, 1.0), 'continuum', _ottb, {}),
    'OTTS': (('Ec', 'K'), (100.0, 1.0), 'continuum', _otts, {}),
    'PLEnFlux': (('alpha', 'F'), (1.01, 1e-12), 'integral', _plfluxnorm, {'energy': True}),
    'PLPhFlux': (('alpha', 'F'), (1.01, 1.0), 'integral', _plfluxnorm, {'energy': False}),
    'PowerLaw': (('alpha', 'K'), (1.01, 1.0), 'integral', _powerlaw, {}),
    'SmoothlyBrokenPL': (('alpha1', 'Eb', 'alpha2', 'rho', 'K'), (1.0, 5.0, 2.0, 1.0, 1.0), 'continuum', _sbpl, {}),
}

MULTIPLICATIVE = {
    'Constant': (('f',), (1.0,), 'integral', _constant, {}),
    'Edge': (('Ec', 'D'), (7.0, 1.0), 'continuum', _edge, {}),
    'ExpAbs': (('Ec',), (2.0,), 'continuum', _expabs, {}),
    'ExpFac': (('A', 'f', 'Ec'), (1.0, 1.0, 0.5), 'continuum', _expfac, {}),
    'GAbs': (('El', 'sigma', 'tau'), (1.0, 0.01, 1.0), 'continuum', _gabs, {}),
    'HighECut': (('Ec', 'Ef'), (10.0, 15.0), 'continuum', _highecut, {}),
    'PLAbs': (('alpha', 'K'), (2.0, 1.0), 'integral', _plabs, {}),
}


def eval_component(name, egrid, params, method='trapz', **extra):
    """Bin values (..., E) of one component: Component.eval (models/model.py:1693-1698,
    1722-1761).  ``params`` maps parameter name -> tensor."""
    if name == 'PhotonAbsorption':  # PhAbs / TBAbs / WAbs: the table comes with the call
        sigma_of = _make_interp(extra['table_energy'], extra['table_xsect'])
        return _integrate(lambda e, p: _photonabs(e, p, sigma_of), egrid, params, 'mul', method)
    if name in ADDITIVE:
        mtype, entry = 'add', ADDITIVE[name]
    else:
        mtype, entry = 'mul', MULTIPLICATIVE[name]
    _, _, how, fn, kwargs = entry
    if how == 'integral':
        return fn(egrid, params, **{**kwargs, **extra})
    return _integrate(fn, egrid, params, mtype, method)


# --------------------------------------------------------------------------- convolution components
def _shift_factor(x):
    """z or v: one value (fixed shift) or one per row (a free shift parameter; the row axis broadcasts
    against the energy axis like every other parameter)."""
    x = x if isinstance(x, torch.Tensor) else _t(x)
    v = x.reshape(-1)
    if not x.requires_grad and bool(torch.all(v == v[0])):
        return float(v[0])  # a fixed shift: the grid stays one-dimensional, as in the reference
    return x


def eval_convolution(name, egrid, params, model_fn, **extra):
    """``convolve`` of the convolution components (models/conv.py); ``model_fn(egrid)`` is the
    wrapped model's bin values (..., len(egrid) - 1)."""
    egrid = _t(egrid)
    if name == 'ZAShift':  # conv.py:294-300
        factor = 1.0 + _shift_factor(params['z'])
        return model_fn(egrid * factor) / factor
    if name == 'ZMShift':  # conv.py:335-341
        factor = 1.0 + _shift_factor(params['z'])
        return model_fn(egrid * factor)
    if name in ('VAShift', 'VMShift'):  # conv.py:378-386, 424-432
        c = 299792.458
        f = 1.0 - _shift_factor(params['v']) / c
        return model_fn(egrid * f)
    if name in ('PhFlux', 'EnFlux'):
        # NormConvolution.eval, conv.py:76-98: the flux grid
        if extra.get('elog', True):
            flux_egrid = np.geomspace(extra['emin'], extra['emax'], extra.get('ngrid', 1000))
        else:
            flux_egrid = np.linspace(extra['emin'], extra['emax'], extra.get('ngrid', 1000))
        flux_egrid = _t(flux_egrid)
        F = params['F']
        if name == 'PhFlux':  # conv.py:185-195
            mflux = torch.sum(model_fn(flux_egrid), dim=-1, keepdim=True)
        else:  # conv.py:244-256
            keV_to_erg = 1.602176634e-9
            mid = torch.sqrt(flux_egrid[:-1] * flux_egrid[1:])
            mflux = torch.sum(keV_to_erg * mid * model_fn(flux_egrid), dim=-1, keepdim=True)
        return F / mflux * model_fn(egrid)
    raise ValueError(f'unknown convolution component {name}')
