"""Synthetic spectra, responses and parameter batches of the BASELINE.json configurations
(SURVEY.md section 8d).  NumPy only; used by tests/ and bench.py to build inputs -- the
shapes, seeds and statistics are the ones the survey names.  The "truth" spectra used to
draw counts are simple midpoint-rule evaluations: they only have to be plausible data.
"""

from __future__ import annotations

import math

import numpy as np

from . import models as M
from .data import FixedData

_erf = np.vectorize(math.erf, otypes=[np.float64])


def _erf_arr(x):
    try:
        from scipy.special import erf

        return erf(x)
    except Exception:  # pragma: no cover
        return _erf(x)


# --------------------------------------------------------------------------- responses
def dense_response(photon_egrid, channel_egrid, sigma_scale=1.0, peak=3.0, fwhm=(0.05, 0.03), seed=1):
    """(E, C) dense redistribution x effective area: A(E) * int_chan N(E'; E, sigma(E)) dE'
    plus a flat 1e-4 * A tail so that no element is zero (cfg 1)."""
    rng = np.random.default_rng(seed)
    e = np.sqrt(photon_egrid[:-1] * photon_egrid[1:])
    sigma = sigma_scale * (fwhm[0] + fwhm[1] * np.sqrt(e)) / 2.355
    area = 500.0 * np.exp(-0.5 * np.log(e / peak) ** 2)
    z = (channel_egrid[None, :] - e[:, None]) / (math.sqrt(2.0) * sigma[:, None])
    cdf = 0.5 * (1.0 + _erf_arr(z))
    rmf = cdf[:, 1:] - cdf[:, :-1]
    C = channel_egrid.size - 1
    R = area[:, None] * (rmf + 1e-4 / C)
    R *= 1.0 + 0.01 * rng.standard_normal(R.shape)
    return np.ascontiguousarray(np.abs(R))


def broad_response(photon_egrid, channel_egrid, rel_sigma, seed):
    """GBM-like broad dense response: sigma / E = rel_sigma (cfg 3)."""
    rng = np.random.default_rng(seed)
    e = np.sqrt(photon_egrid[:-1] * photon_egrid[1:])
    sigma = rel_sigma * e
    area = 100.0 * np.exp(-0.5 * (np.log(e / np.sqrt(channel_egrid[0] * channel_egrid[-1])) / 2.0) ** 2)
    z = (channel_egrid[None, :] - e[:, None]) / (math.sqrt(2.0) * sigma[:, None])
    cdf = 0.5 * (1.0 + _erf_arr(z))
    R = area[:, None] * (cdf[:, 1:] - cdf[:, :-1] + 1e-5)
    R *= 1.0 + 0.01 * rng.standard_normal(R.shape)
    return np.ascontiguousarray(np.abs(R))


def band_sparse_response(n, sigma_ch=6.0, nsig=4.0, seed=30):
    """(indptr, indices, values): CSR by channel row of R.T for an n x n CCD-like response,
    Gaussian kernel of sigma_ch channels truncated at +-nsig sigma (<= 49 nnz/row) (cfg 4)."""
    rng = np.random.default_rng(seed)
    half = int(nsig * sigma_ch)
    offs = np.arange(-half, half + 1)
    kern = np.exp(-0.5 * (offs / sigma_ch) ** 2)
    kern /= kern.sum()
    indptr = np.zeros(n + 1, dtype=np.int32)
    idx, val = [], []
    area = 300.0 * (1.0 + 0.3 * np.sin(np.linspace(0.0, 6.0, n)))
    for c in range(n):
        e = c + offs
        ok = (e >= 0) & (e < n)
        idx.append(e[ok])
        val.append(kern[ok] * area[e[ok]] * (1.0 + 0.01 * rng.standard_normal(ok.sum())))
        indptr[c + 1] = indptr[c] + ok.sum()
    return indptr, np.concatenate(idx).astype(np.int32), np.abs(np.concatenate(val))


# --------------------------------------------------------------------------- truth spectra
def _mid(egrid):
    return 0.5 * (egrid[:-1] + egrid[1:]), np.diff(egrid)


def _powerlaw(egrid, alpha, K):
    e, de = _mid(egrid)
    return K * e ** (-alpha) * de


def _cutoffpl(egrid, alpha, Ec, K):
    e, de = _mid(egrid)
    return K * e ** (-alpha) * np.exp(-e / Ec) * de


def _blackbody(egrid, kT, K):
    e, de = _mid(egrid)
    x = np.minimum(e / kT, 50.0)
    return 8.0525 * K * e * e / kT**4 / np.expm1(x) * de


def _band(egrid, alpha, beta, Ec, K):
    e, de = _mid(egrid)
    eb = (alpha - beta) * Ec
    lo = (e / 100.0) ** alpha * np.exp(-e / Ec)
    hi = (eb / 100.0) ** (alpha - beta) * np.exp(beta - alpha) * (e / 100.0) ** beta
    return K * np.where(e < eb, lo, hi) * de


def _gauss(egrid, El, sigma, K):
    z = (egrid - El) / (math.sqrt(2.0) * sigma)
    cdf = 0.5 * (1.0 + _erf_arr(z))
    return K * np.diff(cdf)


def _theta_batch(truth, n, seed, rel=0.05):
    rng = np.random.default_rng(seed)
    truth = np.asarray(truth, dtype=np.float64)
    return truth * (1.0 + rel * rng.standard_normal((n, truth.size)))


# --------------------------------------------------------------------------- configurations
def config1(n=1, E=2048, C=1024):
    """PowerLaw, one spectrum, dense RSP, cstat (BASELINE configs[0])."""
    eg = np.geomspace(0.3, 80.0, E + 1)
    cg = np.geomspace(0.3, 12.0, C + 1)
    R = dense_response(eg, cg, seed=1)
    truth = (1.7, 10.0)
    mu = 1e3 * (_powerlaw(eg, *truth) @ R)
    counts = np.random.default_rng(2).poisson(mu).astype(np.float64)
    d = FixedData('cfg1', eg, counts, np.ones(C), 1e3, channel_width=np.diff(cg), response_matrix=R)
    return dict(data=[d], model=[M.PowerLaw()], stat=['cstat'], truth=np.array(truth), theta=_theta_batch(truth, n, 5))


def config2(n=1_000_000, E=2048, C=1024):
    """CutoffPL + Blackbody with background, wstat (BASELINE configs[1], the bench workload)."""
    eg = np.geomspace(0.3, 80.0, E + 1)
    cg = np.geomspace(0.3, 12.0, C + 1)
    R = dense_response(eg, cg, seed=1)
    truth = (1.2, 20.0, 5.0, 1.5, 2.0)
    src = 50.0 * ((_cutoffpl(eg, *truth[:3]) + _blackbody(eg, *truth[3:])) @ R)
    off = np.random.default_rng(4).poisson(30.0, C).astype(np.float64)
    on = np.random.default_rng(3).poisson(src + 0.5 * 30.0).astype(np.float64)
    d = FixedData('cfg2', eg, on, np.ones(C), 50.0, channel_width=np.diff(cg), response_matrix=R,
                  back_counts=off, back_errors=np.sqrt(np.maximum(off, 1.0)), back_ratio=0.5)
    return dict(data=[d], model=[M.CutoffPL() + M.Blackbody()], stat=['wstat'], truth=np.array(truth),
                theta=_theta_batch(truth, n, 5))


def config3(n=64):
    """Band, joint fit of 4 GBM-like detectors (140 energies x 128 channels), pgstat (configs[2])."""
    truth = (-1.0, -2.3, 300.0, 0.02)
    eg = np.geomspace(5.0, 5e4, 141)
    model = M.Band()
    datas = []
    for i in range(4):
        bgo = i == 3
        cg = np.geomspace(250.0, 4e4, 129) if bgo else np.geomspace(8.0, 900.0, 129)
        R = broad_response(eg, cg, 0.25 if bgo else 0.15, seed=10 + i)
        expo = 10.0
        mu = expo * (_band(eg, *truth) @ R)
        b_true = 40.0 * np.ones(128)
        b_err = np.full(128, 3.0)
        b_est = np.random.default_rng(18 + i).normal(b_true, b_err)
        on = np.random.default_rng(14 + i).poisson(mu + b_true).astype(np.float64)
        datas.append(FixedData(f'det{i}', eg, on, np.ones(128), expo, channel_width=np.diff(cg), response_matrix=R,
                               back_counts=b_est, back_errors=b_err, back_ratio=1.0))
    return dict(data=datas, model=[model] * 4, stat=['pgstat'] * 4, truth=np.array(truth),
                theta=_theta_batch(truth, n, 22, rel=0.03))


def config4(n=100_000, size=4096):
    """Gauss + PowerLaw on a band-sparse CCD-like RSP, chi2, per-replica counts (configs[3])."""
    truth = (6.4, 0.05, 1e-3, 1.8, 0.05)
    eg = np.linspace(0.2, 12.0, size + 1)
    sp = band_sparse_response(size, seed=30)
    indptr, indices, values = sp
    u = _gauss(eg, *truth[:3]) + _powerlaw(eg, *truth[3:])
    expo = 2e4
    mu = np.zeros(size)
    rows = np.repeat(np.arange(size), np.diff(indptr))
    np.add.at(mu, rows, values * u[indices])
    mu *= expo
    err = np.sqrt(np.maximum(mu, 1.0))
    counts = np.random.default_rng(31).normal(mu, err, (n, size))  # infer/helper.py:235-237
    d = FixedData('cfg4', eg, np.round(mu), np.ones(size), expo, channel_width=np.diff(eg), sparse_matrix=sp,
                  response_sparse=True, net_counts=mu.copy(), net_errors=err)
    return dict(data=[d], model=[M.Gauss() + M.PowerLaw()], stat=['chi2'], truth=np.array(truth),
                theta=_theta_batch(truth, n, 32, rel=0.01), counts_on=counts)


def config5(n=1_000_000, S=16, E=2048, C=1024, heavy=False):
    """Joint fit of S spectra with distinct dense responses, mixed cstat / chi2 (configs[4])."""
    eg = np.geomspace(0.3, 80.0, E + 1)
    cg = np.geomspace(0.3, 12.0, C + 1)
    if heavy:
        model = M.CutoffPL() + M.Blackbody()
        truth = (1.2, 20.0, 5.0, 1.5, 2.0)
        u = _cutoffpl(eg, *truth[:3]) + _blackbody(eg, *truth[3:])
    else:
        model = M.PowerLaw()
        truth = (1.7, 10.0)
        u = _powerlaw(eg, *truth)
    datas, stats = [], []
    for i in range(S):
        t = i / max(S - 1, 1)
        R = dense_response(eg, cg, sigma_scale=0.8 + 0.8 * t, peak=1.0 + 5.0 * t, seed=40 + i)
        expo = 1e3
        mu = expo * (u @ R)
        rng = np.random.default_rng(60 + i)
        if i % 2 == 0:
            counts = rng.poisson(mu).astype(np.float64)
            datas.append(FixedData(f's{i}', eg, counts, np.ones(C), expo, channel_width=np.diff(cg), response_matrix=R))
            stats.append('cstat')
        else:
            err = np.sqrt(np.maximum(mu, 1.0))
            net = rng.normal(mu, err)
            datas.append(FixedData(f's{i}', eg, np.round(np.abs(net)), np.ones(C), expo, channel_width=np.diff(cg),
                                   response_matrix=R, net_counts=net, net_errors=err))
            stats.append('chi2')
    return dict(data=datas, model=[model] * S, stat=stats, truth=np.array(truth), theta=_theta_batch(truth, n, 5))


CONFIGS = {1: config1, 2: config2, 3: config3, 4: config4, 5: config5}
