"""Oracle restatement of the fold + fit-statistic + reduction path (torch, CPU, float64).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows infer/likelihood.py and
infer/helper.py:494-512,571-594 of the reference op for op.

Neutral input formats (produced by ``elisa_b200.models`` / ``elisa_b200.data`` but
expressed in plain Python so the oracle imports nothing from the product):

* model spec -- nested tuples
      ('comp', name, {pname: ref, ...}, {'method': 'trapz'|'simpson', 'emin':..., 'emax':...})
      ('+', lhs, rhs) | ('*', lhs, rhs)
      ('conv', name, {pname: ref}, {'emin':..., 'emax':..., 'ngrid':..., 'elog':...}, inner)
  with ``ref = ('theta', j)`` (column j of the parameter matrix) or ``('const', v)``.
* data -- a mapping with the ``FixedData`` field names (data/base.py:1893-1975):
  photon_egrid (E+1,), response_matrix (E, C), area_scale, spec_exposure, channel_width,
  spec_counts, net_counts, net_errors, back_counts, back_errors, back_ratio.
"""

from __future__ import annotations

import numpy as np
import torch

from .components import F64, eval_component, eval_convolution

STATISTICS = ('chi2', 'cstat', 'pstat', 'pgstat', 'wstat')
_WITH_BACK = ('pgstat', 'wstat')


def _as(x):
    if isinstance(x, torch.Tensor):
        return x.to(F64)
    return torch.tensor(np.asarray(x, dtype=np.float64))  # copies: inputs may be read-only views


# --------------------------------------------------------------------------- model tree
def eval_model(spec, egrid, theta):
    """CompiledModel.eval for a batch: models/model.py:202-219 (routing of parameter
    values to components), :1188-1197 (leaf), :1285-1295 (composite, op on BIN values).
    ``theta``: (N, P) tensor.  Returns (N, E)."""
    kind = spec[0]
    if kind == 'comp':
        _, name, refs, extra = spec
        params = {}
        for pname, ref in refs.items():
            if ref[0] == 'theta':
                params[pname] = theta[:, ref[1]].unsqueeze(-1)
            else:
                params[pname] = torch.full((theta.shape[0], 1), float(ref[1]), dtype=F64)
        extra = dict(extra)
        method = extra.pop('method', 'trapz')
        return eval_component(name, egrid, params, method=method, **extra)
    if kind == 'conv':  # ConvolvedModel.eval, models/model.py:1945-1955
        _, name, refs, extra, inner = spec
        params = {}
        for pname, ref in refs.items():
            if ref[0] == 'theta':
                params[pname] = theta[:, ref[1]].unsqueeze(-1)
            else:
                params[pname] = torch.full((theta.shape[0], 1), float(ref[1]), dtype=F64)
        return eval_convolution(name, egrid, params, lambda e: eval_model(inner, e, theta), **extra)
    lhs = eval_model(spec[1], egrid, theta)
    rhs = eval_model(spec[2], egrid, theta)
    if kind == '+':
        return lhs + rhs
    if kind == '*':
        return lhs * rhs
    raise ValueError(kind)


# --------------------------------------------------------------------------- profile backgrounds
def pgstat_background(s, n, b_est, b_err, a):
    """infer/likelihood.py:41-86."""
    variance = b_err * b_err
    e = b_est - a * variance
    f = a * variance * n + e * s
    c = a * e - s
    d = torch.sqrt(c * c + 4.0 * a * f)
    return torch.where(
        (e >= 0.0) | (f >= 0.0),
        torch.where(n > 0.0, (c + d) / (2 * a), e + 0.0 * d),
        torch.zeros_like(d),
    )


def wstat_background(s, n_on, n_off, a):
    """infer/likelihood.py:89-133."""
    c = a * (n_on + n_off) - (a + 1) * s
    d = torch.sqrt(c * c + 4 * a * (a + 1) * n_off * s)
    return torch.where(
        n_on == 0,
        n_off / (1 + a) + 0.0 * d,
        torch.where(
            n_off == 0,
            torch.where(s <= a / (a + 1) * n_on, n_on / (1 + a) - s / a, torch.zeros_like(d)),
            (c + d) / (2 * a * (a + 1)),
        ),
    )


def _normal_logprob(value, loc, scale):
    """BetterNormal.log_prob, infer/likelihood.py:136-140."""
    value_scaled = (value - loc) / scale
    return -0.5 * value_scaled * value_scaled


def _poisson_logprob(value, rate):
    """BetterPoisson.log_prob, dense branch, infer/likelihood.py:171-174."""
    xlogy = torch.special.xlogy
    logp = xlogy(value, rate) - rate
    gof = xlogy(value, value) - value
    return torch.clamp(logp - gof, max=0.0)


# --------------------------------------------------------------------------- per-statistic wiring
def sites(stat, data, spec, theta, counts_on=None, counts_off=None):
    """All per-channel sites of one dataset for a batch of parameter vectors.

    Restates the likelihood closures chi2/cstat/pstat/pgstat/wstat
    (infer/likelihood.py:184-450).  ``counts_on`` / ``counts_off`` (N, C) replace the
    observed data per replica, which is what substituting ``{name}_Non_data`` /
    ``{name}_Noff_data`` does upstream (infer/helper.py:608-613).

    Returns a dict: 'rate' ({name} site = source_rate / channel_width), 'Non_model',
    'Non_loglike', optionally 'Noff_model', 'Noff_loglike', and 'loglike' (N, C).
    """
    theta = _as(theta)
    if theta.dim() == 1:
        theta = theta.unsqueeze(0)
    egrid = _as(data['photon_egrid'])
    resp = _as(data['response_matrix']).T  # _get_resp_matrix, likelihood.py:177-181 -> (C, E)
    area_scale = _as(data['area_scale'])
    exposure = _as(data['spec_exposure'])
    channel_width = _as(data['channel_width'])

    unfold = eval_model(spec, egrid, theta)
    unfold = torch.clamp(unfold, min=1e-300, max=1e300)  # :204
    source_rate = (unfold @ resp.T) * area_scale  # :205  resp_matrix @ unfold * area_scale
    out = {'rate': source_rate / channel_width}

    if stat == 'chi2':
        spec_data = _as(data['net_counts']) if counts_on is None else _as(counts_on)
        error = _as(data['net_errors'])
        source_counts = torch.clamp(source_rate * exposure, min=1e-30, max=1e15)
        out['Non_model'] = source_counts
        out['Non_loglike'] = _normal_logprob(spec_data, source_counts, error)
        out['loglike'] = out['Non_loglike']
    elif stat == 'cstat':
        spec_data = _as(data['spec_counts']) if counts_on is None else _as(counts_on)
        source_counts = torch.clamp(source_rate * exposure, min=1e-30, max=1e15)
        out['Non_model'] = source_counts
        out['Non_loglike'] = _poisson_logprob(spec_data, source_counts)
        out['loglike'] = out['Non_loglike']
    elif stat == 'pstat':
        spec_data = _as(data['spec_counts']) if counts_on is None else _as(counts_on)
        back = _as(data['back_counts'])
        back_ratio = _as(data['back_ratio'])
        model_counts = source_rate * exposure + back_ratio * back  # :301 (before the clip)
        model_counts = torch.clamp(model_counts, min=1e-30, max=1e15)
        out['Non_model'] = model_counts
        out['Non_loglike'] = _poisson_logprob(spec_data, model_counts)
        out['loglike'] = out['Non_loglike']
    elif stat in _WITH_BACK:
        spec_data = _as(data['spec_counts']) if counts_on is None else _as(counts_on)
        back_data = _as(data['back_counts']) if counts_off is None else _as(counts_off)
        back_ratio = _as(data['back_ratio'])
        source_counts = torch.clamp(source_rate * exposure, min=1e-30, max=1e15)
        if stat == 'pgstat':
            back_error = _as(data['back_errors'])
            b = pgstat_background(source_counts, spec_data, back_data, back_error, back_ratio)
            out['Noff_loglike'] = _normal_logprob(back_data, b, back_error)
        else:
            b = wstat_background(source_counts, spec_data, back_data, back_ratio)
            out['Noff_loglike'] = _poisson_logprob(back_data, b)
        model_counts = source_counts + back_ratio * b
        out['Non_model'] = model_counts
        out['Noff_model'] = b
        out['Non_loglike'] = _poisson_logprob(spec_data, model_counts)
        out['loglike'] = out['Non_loglike'] + out['Noff_loglike']
    else:
        raise ValueError(f'unknown statistic {stat!r}')
    return out


def loglike(stats, datas, specs, theta, counts_on=None, counts_off=None):
    """(N, S) log-likelihood, one column per dataset: get_loglike 'group'
    (infer/helper.py:494-512).  ``deviance = -2 * loglike`` (:571-586); the joint total
    is the row sum.  ``specs`` may be one spec shared by all datasets or a list."""
    S = len(datas)
    if not isinstance(specs, list):
        specs = [specs] * S
    cols = []
    for s in range(S):
        on = None if counts_on is None else counts_on[s]
        off = None if counts_off is None else counts_off[s]
        cols.append(sites(stats[s], datas[s], specs[s], theta, on, off)['loglike'].sum(-1))
    return torch.stack(cols, dim=-1)


def loglike_and_grad(stats, datas, specs, theta, counts_on=None, counts_off=None):
    """Values (N, S) and d loglike[n, s] / d theta[n, :] (N, S, P) by reverse-mode
    autograd -- the stand-in for jax.value_and_grad (infer/fit.py:487; numpyro NUTS)."""
    theta = _as(theta).clone().requires_grad_(True)
    ll = loglike(stats, datas, specs, theta, counts_on, counts_off)
    grads = []
    for s in range(ll.shape[1]):
        (g,) = torch.autograd.grad(ll[:, s].sum(), theta, retain_graph=True, allow_unused=True)
        grads.append(torch.zeros_like(theta) if g is None else g)
    return ll.detach(), torch.stack(grads, dim=1)
