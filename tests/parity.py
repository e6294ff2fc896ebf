"""Shared helpers of the parity tests: run the CUDA path and the CPU oracle on the same
inputs and measure the deviation.  (Test infrastructure; imports oracle/.)"""

from __future__ import annotations

import numpy as np
import torch

import oracle
from elisa_b200 import Likelihood

RTOL = 1e-10  # BASELINE.json north_star: 1e-10 relative in FP64 on the statistic and its gradient


def oracle_eval(cfg, theta=None, grad=True):
    theta = cfg['theta'] if theta is None else theta
    models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
    # global free list as Likelihood builds it
    free = []
    for m in models:
        for p in m.free:
            if all(p is not q for q in free):
                free.append(p)
    specs = [m.to_spec(free) for m in models]
    datas = [d.as_mapping() for d in cfg['data']]
    on = off = None
    if cfg.get('counts_on') is not None:
        on = np.split(np.asarray(cfg['counts_on']), np.cumsum([d.n_channels for d in cfg['data']])[:-1], axis=1)
    if cfg.get('counts_off') is not None:
        off = np.split(np.asarray(cfg['counts_off']), np.cumsum([d.n_channels for d in cfg['data']])[:-1], axis=1)
    t = torch.as_tensor(theta, dtype=torch.float64)
    if grad:
        ll, g = oracle.loglike_and_grad(cfg['stat'], datas, specs, t, on, off)
        return ll.numpy(), g.numpy()
    return oracle.loglike(cfg['stat'], datas, specs, t, on, off).numpy(), None


def gpu_eval(cfg, theta=None, grad=True, chunk=0, host=False, specialise=None, fold=None, slices=0, slices_bwd=0,
             small_batch=True):
    theta = cfg['theta'] if theta is None else theta
    models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
    lk = Likelihood(cfg['data'], models, cfg['stat'], chunk=chunk, specialise=specialise, fold=fold, slices=slices,
                    slices_bwd=slices_bwd, small_batch=small_batch)
    if fold is not None:
        assert all((s > 0) == (fold == 'i8') for s in lk.fold_slices), lk.fold_slices
    if specialise is not None:
        assert all(s == specialise for s in lk.specialised)
    kw = dict(counts_on=cfg.get('counts_on'), counts_off=cfg.get('counts_off'))
    try:
        if host:
            out = lk.loglike_and_grad_host(np.asarray(theta), grad=grad, **kw)
            return out if grad else (out, None)
        if grad:
            ll, g = lk.loglike_and_grad(theta, **kw)
            return ll.cpu().numpy(), g.cpu().numpy()
        return lk.loglike(theta, **kw).cpu().numpy(), None
    finally:
        lk.close()


def rel_err_value(a, ref):
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), 1.0)))


def rel_err_grad(g, ref):
    """Deviation of the gradient relative to the size of each (spectrum, parameter)
    column over the batch: |g - ref| / max(|ref|, max_n |ref[n, s, p]|)."""
    scale = np.maximum(np.abs(ref), np.max(np.abs(ref), axis=0, keepdims=True))
    scale = np.where(scale == 0.0, 1.0, scale)
    return float(np.max(np.abs(g - ref) / scale))


def rel_err_grad_rownorm(g, ref):
    """Row-normwise deviation of the gradient: each row's error is measured against THAT row's
    own gradient (its largest component in units of the column scales), not against the largest
    gradient of the batch -- a row whose gradient is 1e-6 of the batch maximum has to be right
    in its own leading digits.  |g - ref|[n, s, p] / (col[s, p] * max_q |ref[n, s, q]| / col[s, q])."""
    col = np.max(np.abs(ref), axis=0, keepdims=True)
    col = np.where(col == 0.0, 1.0, col)
    row = np.max(np.abs(ref) / col, axis=-1, keepdims=True)
    row = np.where(row == 0.0, 1.0, row)
    return float(np.max(np.abs(g - ref) / (col * row)))


# Tolerance of the row-normwise criterion.  The gradient is a sum over E energies of terms of
# both signs; for a row whose gradient is s_n times the batch scale the conditioning of that sum
# alone costs 1 / s_n digits in ANY arithmetic (the oracle's torch-autograd path and XLA differ by
# that much from each other), so the bar is 1e-10 relative to the row down to s_n = 1e-2 and
# 1e-12 of the column scale below that.
def rel_err_grad_rownorm_floor(g, ref, floor=1e-2):
    col = np.max(np.abs(ref), axis=0, keepdims=True)
    col = np.where(col == 0.0, 1.0, col)
    row = np.maximum(np.max(np.abs(ref) / col, axis=-1, keepdims=True), floor)
    return float(np.max(np.abs(g - ref) / (col * row)))


def compare(cfg, grad=True, **kw):
    ll_ref, g_ref = oracle_eval(cfg, grad=grad)
    ll, g = gpu_eval(cfg, grad=grad, **kw)
    out = {'value': rel_err_value(ll, ll_ref)}
    if grad:
        out['grad'] = rel_err_grad(g, g_ref)
        out['grad_row'] = rel_err_grad_rownorm_floor(g, g_ref)
        out['grad_row_raw'] = rel_err_grad_rownorm(g, g_ref)
    return out
