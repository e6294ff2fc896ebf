"""GPU tests at BASELINE.json's full shapes: a random sub-sample of rows against the
oracle, plus size-independent properties over the whole batch (bitwise chunk invariance
and determinism, directional-derivative consistency, sign of the deviance)."""

import numpy as np
import pytest
import torch

from elisa_b200 import Likelihood, synthetic

pytestmark = pytest.mark.gpu


def _check(cfg, n_oracle, chunk_alt):
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    theta = cfg['theta']
    n = len(theta)
    kw = dict(counts_on=cfg.get('counts_on'))
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    ll, g = lk.loglike_and_grad(theta, **kw)
    ll2, g2 = lk.loglike_and_grad(theta, **kw)
    assert torch.equal(ll, ll2) and torch.equal(g, g2)  # deterministic (fixed-order reductions)
    # value-only entry point: same numbers up to FMA contraction in the dual-number pass
    assert torch.allclose(lk.loglike(theta, **kw), ll, rtol=1e-10, atol=0)
    assert bool((ll <= 0).all())  # every statistic is referenced to the saturated model or is -chi2/2
    # directional derivative: (ll(t + h v) - ll(t - h v)) / 2h == g . v
    rng = np.random.default_rng(0)
    v = rng.standard_normal(theta.shape) * np.abs(theta)
    h = 1e-6
    lp = lk.loglike(theta + h * v, **kw)
    lm = lk.loglike(theta - h * v, **kw)
    fd = ((lp - lm) / (2 * h)).cpu().numpy()
    gv = (g.cpu().numpy() * v[:, None, :]).sum(-1)
    scale = np.abs(g.cpu().numpy() * v[:, None, :]).sum(-1) + 1e-6 * np.abs(ll.cpu().numpy())
    assert np.max(np.abs(fd - gv) / scale) < 1e-4
    lk.close()
    # chunk invariance, bit for bit
    lk2 = Likelihood(cfg['data'], cfg['model'], cfg['stat'], chunk=chunk_alt)
    ll3, g3 = lk2.loglike_and_grad(theta, **kw)
    assert torch.equal(ll, ll3) and torch.equal(g, g3)
    lk2.close()
    # oracle on a sub-sample of rows
    rows = np.sort(rng.choice(n, size=min(n_oracle, n), replace=False))
    sub = dict(cfg, theta=theta[rows])
    if cfg.get('counts_on') is not None:
        sub['counts_on'] = cfg['counts_on'][rows]
    ll_ref, g_ref = oracle_eval(sub)
    assert rel_err_value(ll.cpu().numpy()[rows], ll_ref) <= RTOL
    assert rel_err_grad(g.cpu().numpy()[rows], g_ref) <= RTOL


def test_config1_full_shape(lib):
    _check(synthetic.config1(n=20000), 192, 4096)


def test_config2_full_shape(lib):
    _check(synthetic.config2(n=40000), 192, 4096)


def test_config3_full_shape(lib):
    _check(synthetic.config3(n=64), 64, 128)


def test_config4_full_shape_band_sparse(lib):
    _check(synthetic.config4(n=3000, size=4096), 48, 1024)


def test_config5_sixteen_spectra(lib):
    _check(synthetic.config5(n=6000, S=16), 64, 2048)
