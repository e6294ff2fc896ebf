"""Batched Levenberg-Marquardt re-fit (reference: fit_once / batch_fit,
infer/helper.py:601-806): the normal equations against the oracle, and the fitted optimum
against an independent CPU optimiser driving the oracle on the same simulated data."""

import numpy as np
import pytest
import torch

from elisa_b200 import models as M

pytestmark = pytest.mark.gpu

from test_gpu_parity import _one_spectrum, _theta_for  # noqa: E402


def _oracle_residuals(cfg, theta, on, off=None):
    import oracle

    model = cfg['model'][0].compile() if not hasattr(cfg['model'][0], 'to_spec') else cfg['model'][0]
    d = cfg['data'][0].as_mapping()
    s = oracle.sites(cfg['stat'][0], d, model.to_spec(), theta, on, off)
    return torch.sqrt(-2.0 * s['loglike'])


@pytest.mark.parametrize('fold', ['i8', 'dmma'])
@pytest.mark.parametrize('stat', ['chi2', 'cstat', 'pgstat', 'wstat'])
def test_normal_equations_match_autograd_of_the_oracle_residuals(lib, stat, fold):
    from elisa_b200 import Likelihood

    d = _one_spectrum(E=90, C=48, seed=31)
    model = M.CutoffPL() + M.Gauss(El=20.0, sigma=3.0)
    cm = model.compile()
    theta = _theta_for(model, 12, 9)
    rng = np.random.default_rng(4)
    on = rng.poisson(40.0, (12, 48)).astype(float)
    off = rng.poisson(15.0, (12, 48)).astype(float)
    cfg = dict(data=[d], model=[cm], stat=[stat])
    lk = Likelihood([d], cm, [stat], fold=fold)
    kw = dict(counts_on=on) if stat in ('chi2', 'cstat') else dict(counts_on=on, counts_off=off)
    ll, g, h = lk.normal_equations(theta, **kw)
    lk.close()
    t = torch.as_tensor(theta, dtype=torch.float64)
    for i in range(len(theta)):
        ti = t[i : i + 1].clone().requires_grad_(True)
        fn = lambda x: _oracle_residuals(cfg, x, on[i : i + 1], off[i : i + 1] if 'counts_off' in kw else None)[0]  # noqa: E731
        r = fn(ti)
        J = torch.autograd.functional.jacobian(fn, ti)[:, 0, :]  # (C, P)
        ok = (r > 0).numpy()  # a channel at r = 0 has no defined dr/dtheta (and adds nothing here)
        J = torch.nan_to_num(J)[ok]
        r = r.detach()[ok]
        assert np.isclose(ll[i].item(), -0.5 * float((r**2).sum()), rtol=1e-10)
        gref = -(J.T @ r).numpy()
        href = (J.T @ J).numpy()
        assert np.allclose(g[i].cpu().numpy(), gref, rtol=1e-8, atol=1e-8 * np.abs(gref).max()), (stat, i)
        assert np.allclose(h[i].cpu().numpy(), href, rtol=1e-8, atol=1e-8 * np.abs(href).max()), (stat, i)


def _resolved_spectrum(E, C, exposure):
    """a spectrum with an energy-resolving response (synthetic.dense_response), so that spectral
    shape parameters are actually constrained by the data"""
    from elisa_b200 import synthetic
    from elisa_b200.data import FixedData

    eg = np.geomspace(0.3, 80.0, E + 1)
    cg = np.geomspace(0.3, 12.0, C + 1)
    R = synthetic.dense_response(eg, cg, seed=1)
    z = np.zeros(C)
    return FixedData('d', eg, z.copy(), np.ones(C), exposure, channel_width=np.diff(cg), response_matrix=R,
                     net_counts=z.copy(), net_errors=np.ones(C), back_counts=np.full(C, 30.0),
                     back_errors=np.full(C, 5.5), back_ratio=0.5)


@pytest.mark.parametrize('case', ['powerlaw_cstat', 'cutoffpl_gauss_chi2', 'absorbed_pl_wstat'])
def test_batched_fit_reaches_the_optimum_of_an_independent_optimiser(lib, case):
    import scipy.optimize

    import oracle
    from elisa_b200 import Likelihood

    rng = np.random.default_rng(17)
    E, C, n = 160, 96, 24
    d = _resolved_spectrum(E, C, 20.0)
    U = M.UniformParameter
    if case == 'powerlaw_cstat':
        model, stat, truth = M.PowerLaw(alpha=U('alpha', 1.6, 0.0, 4.0), K=U('K', 3.0, 1e-3, 1e3)), 'cstat', [1.6, 3.0]
    elif case == 'cutoffpl_gauss_chi2':
        model = M.CutoffPL(alpha=U('alpha', 1.3, -1.0, 4.0), Ec=U('Ec', 6.0, 0.5, 500.0), K=U('K', 5.0, 1e-3, 1e3)) + \
            M.Gauss(El=6.4, sigma=0.3, K=U('Kg', 1.0, 1e-3, 1e3))
        stat, truth = 'chi2', [1.3, 6.0, 5.0, 1.0]
    else:
        te = np.geomspace(0.1, 500.0, 200)
        model = M.PhAbs(te, 0.25 * te**-2.5, nH=U('nH', 1.0, 0.0, 10.0)) * M.PowerLaw(alpha=U('alpha', 1.8, 0.0, 4.0), K=U('K', 6.0, 1e-3, 1e3))
        stat, truth = 'wstat', [1.0, 1.8, 6.0]
    cm = model.compile()
    truth = np.asarray(truth)
    probe = Likelihood([d], cm, ['cstat'])  # cstat's Non_model site is the folded source counts
    mu = probe.sites(0, truth[None])['d_Non_model'][0].cpu().numpy()
    probe.close()
    off = None
    if stat == 'chi2':
        d.net_errors[:] = np.sqrt(np.maximum(mu, 1.0))
        on = rng.normal(mu, d.net_errors, (n, C))
    elif stat == 'cstat':
        on = rng.poisson(mu, (n, C)).astype(float)
    else:
        on = rng.poisson(mu + 0.5 * 30.0, (n, C)).astype(float)
        off = rng.poisson(30.0, (n, C)).astype(float)
    lk = Likelihood([d], cm, [stat])
    start = truth * (1.0 + 0.2 * rng.standard_normal((n, truth.size)))
    start = np.clip(start, [p.min + 1e-3 * (p.max - p.min) for p in cm.free], [p.max - 1e-3 * (p.max - p.min) for p in cm.free])
    out = lk.fit(start, counts_on=on, counts_off=off)
    lk.close()
    assert bool(out['valid'].all()), out['grad_norm']
    theta = out['theta'].cpu().numpy()
    dev = out['deviance'].cpu().numpy()
    # independent optimiser on the oracle: BFGS from the truth, in the same unconstrained space
    lo = np.array([p.min for p in cm.free]); hi = np.array([p.max for p in cm.free])
    spec, dm = cm.to_spec(), d.as_mapping()
    for i in range(0, n, 6):
        oi = None if off is None else [off[i : i + 1]]

        def f(u):
            sg = 1.0 / (1.0 + np.exp(-u))
            th = lo + (hi - lo) * sg
            ll, g = oracle.loglike_and_grad([stat], [dm], [spec], torch.as_tensor(th[None]), [on[i : i + 1]], oi)
            return -2.0 * float(ll[0, 0]), -2.0 * g[0, 0].numpy() * (hi - lo) * sg * (1 - sg)

        z = (theta[i] - lo) / (hi - lo)
        res = scipy.optimize.minimize(f, np.log(z) - np.log1p(-z), jac=True, method='BFGS', options=dict(gtol=1e-9, maxiter=500))
        th_ref = lo + (hi - lo) / (1.0 + np.exp(-res.x))
        assert res.fun <= dev[i] * (1 + 1e-12) + 1e-9  # BFGS started at our optimum: it cannot be better than noise
        assert np.isclose(dev[i], res.fun, rtol=1e-9, atol=1e-7), (case, i, dev[i], res.fun)
        # the optimum itself: to 1e-3 (along a flat valley -- a cutoff energy far outside the data's range --
        # parameter values 1e-4 apart share the deviance to 1e-9)
        assert np.allclose(theta[i], th_ref, rtol=1e-3), (case, i, theta[i], th_ref)
    # ... and from a fresh start the independent optimiser lands on the same deviance
    for i in (1, 7):
        oi = None if off is None else [off[i : i + 1]]

        def f(u):
            sg = 1.0 / (1.0 + np.exp(-u))
            th = lo + (hi - lo) * sg
            ll, g = oracle.loglike_and_grad([stat], [dm], [spec], torch.as_tensor(th[None]), [on[i : i + 1]], oi)
            return -2.0 * float(ll[0, 0]), -2.0 * g[0, 0].numpy() * (hi - lo) * sg * (1 - sg)

        z = (truth - lo) / (hi - lo)
        res = scipy.optimize.minimize(f, np.log(z) - np.log1p(-z), jac=True, method='BFGS', options=dict(gtol=1e-8, maxiter=1000))
        assert np.isclose(dev[i], res.fun, rtol=1e-8, atol=1e-6), (case, i, dev[i], res.fun)


def test_device_simulator_and_bootstrap_loop(lib):
    """simulate (helper.py:221-278) + simulate_and_fit (:808-876) on the device: the sampled
    counts have the model's mean and the sampling distribution's variance, the bootstrap
    re-fits scatter around the generating parameters with the spread the curvature predicts."""
    from elisa_b200 import Likelihood

    U = M.UniformParameter
    n = 4000
    truth = np.array([1.6, 3.0])
    # Poisson on and off (wstat)
    d = _resolved_spectrum(160, 96, 20.0)
    model = M.PowerLaw(alpha=U('alpha', 1.6, 0.0, 4.0), K=U('K', 3.0, 1e-3, 1e3)).compile()
    lk = Likelihood([d], model, ['wstat'])
    on, off = lk.simulate(truth, n=n, seed=3)
    sites = lk.sites(0, truth[None])
    m_on, m_off = sites['d_Non_model'][0], sites['d_Noff_model'][0]
    assert on.shape == (n, 96) and off.shape == (n, 96)
    assert torch.all(on >= 0) and torch.all(on == on.round())
    z = (on.mean(0) - m_on) / torch.sqrt(m_on / n)
    assert float(z.abs().max()) < 5.0 and abs(float(z.mean())) < 0.5
    assert torch.allclose(on.var(0), m_on, rtol=0.2) and torch.allclose(off.var(0), m_off, rtol=0.2)
    on2, _ = lk.simulate(truth, n=n, seed=3)
    assert torch.equal(on, on2)  # reproducible per seed
    out = lk.simulate_and_fit(truth, n=n, seed=5)
    assert float(out['valid'].double().mean()) > 0.99
    th = out['theta'][out['valid']]
    _, _, h = lk.normal_equations(np.tile(truth, (256, 1)), counts_on=on[:256], counts_off=off[:256])
    cov = torch.linalg.inv(h.mean(0))  # Gauss-Newton (Fisher) covariance at the truth
    sd = torch.sqrt(torch.diagonal(cov))
    bias = (th.mean(0) - torch.as_tensor(truth, device=th.device)) / (sd / np.sqrt(len(th)))
    assert float(bias.abs().max()) < 6.0, bias
    assert torch.allclose(th.std(0), sd, rtol=0.15), (th.std(0), sd)
    lk.close()
    # Normal data (chi2) and per-row parameter vectors
    d.net_errors[:] = np.sqrt(np.maximum(m_on.cpu().numpy(), 1.0))
    lk = Likelihood([d], model, ['chi2'])
    thetas = truth * (1.0 + 0.01 * np.random.default_rng(0).standard_normal((500, 2)))
    on, off = lk.simulate(thetas, seed=1)
    assert on.shape == (500, 96) and off is None
    mu = lk.sites(0, thetas)['d_Non_model']
    pull = (on - mu) / torch.as_tensor(d.net_errors, device=on.device)
    assert abs(float(pull.mean())) < 0.02 and abs(float(pull.std()) - 1.0) < 0.02
    with pytest.raises(ValueError):
        lk.simulate(thetas, n=3)
    lk.close()


def test_normal_equations_on_asimov_data_use_the_curvature_limit(lib):
    """data equal to the model (r = 0 in every channel): J^T J must be the Fisher matrix
    sum_c s_p s_q / mu_c (Poisson), not 0/0"""
    from elisa_b200 import Likelihood

    d = _resolved_spectrum(160, 96, 20.0)
    U = M.UniformParameter
    model = M.PowerLaw(alpha=U('alpha', 1.6, 0.0, 4.0), K=U('K', 3.0, 1e-3, 1e3)).compile()
    truth = np.array([[1.6, 3.0]])
    lk = Likelihood([d], model, ['cstat'])
    mu = lk.sites(0, truth)['d_Non_model']
    ll, g, h = lk.normal_equations(truth, counts_on=mu)
    eps = 1e-6
    jac = []
    for j in range(2):
        tp, tm = truth.copy(), truth.copy()
        tp[0, j] *= 1 + eps
        tm[0, j] *= 1 - eps
        jac.append((lk.sites(0, tp)['d_Non_model'] - lk.sites(0, tm)['d_Non_model'])[0] / (2 * eps * truth[0, j]))
    J = torch.stack(jac, dim=1)  # d mu_c / d theta_p
    fisher = (J / mu[0][:, None]).T @ J
    assert abs(float(ll[0])) < 1e-9 and float(g.abs().max()) < 1e-6
    assert torch.allclose(h[0], fisher, rtol=2e-3), (h[0], fisher)
    lk.close()


def test_normal_equations_with_more_than_eight_parameters(lib):
    """P = 11 free parameters (three components): the 16-parameter instantiation against autograd"""
    from elisa_b200 import Likelihood

    d = _one_spectrum(E=90, C=48, seed=33)
    U = M.UniformParameter
    model = (M.Band() + M.CutoffPL() + M.SmoothlyBrokenPL(rho=U('rho', 1.0, 0.1, 10.0))).compile()
    assert model.nparam == 12
    theta = _theta_for(model.model, 4, 3)
    on = np.random.default_rng(8).poisson(60.0, (4, 48)).astype(float)
    lk = Likelihood([d], model, ['cstat'])
    ll, g, h = lk.normal_equations(theta, counts_on=on)
    ll2, g2 = lk.loglike_and_grad(theta, counts_on=on)
    lk.close()
    assert torch.allclose(ll, ll2[:, 0], rtol=1e-10) and torch.allclose(g, g2[:, 0], rtol=1e-7, atol=1e-7 * float(g2.abs().max()))
    cfg = dict(data=[d], model=[model], stat=['cstat'])
    t = torch.as_tensor(theta, dtype=torch.float64)
    for i in range(4):
        fn = lambda x: _oracle_residuals(cfg, x, on[i : i + 1])[0]  # noqa: E731
        J = torch.nan_to_num(torch.autograd.functional.jacobian(fn, t[i : i + 1].clone())[:, 0, :])
        href = (J.T @ J).numpy()
        assert np.allclose(h[i].cpu().numpy(), href, rtol=1e-7, atol=1e-8 * np.abs(href).max()), i


def test_sampling_kernel_distributions(lib):
    """elisa_b200_sample_dev (the repo's own Philox / Poisson / Normal kernel behind Likelihood.simulate):
    moments over 4e5 draws per mean on both sides of the multiplication / transformed-rejection switch at
    12, the probability mass function against scipy at small and moderate means, per-seed
    reproducibility, independence of the launch shape, Normal draws for chi2 data."""
    import ctypes as C

    from scipy import stats

    from elisa_b200 import FixedData, Likelihood, _lib

    means = np.array([0.0, 0.05, 0.7, 3.0, 11.9, 12.0, 12.1, 50.0, 1e3, 1e5, 3e7])
    Cn = means.size
    eg = np.geomspace(1.0, 10.0, 9)
    d = FixedData('d', eg, np.ones(Cn), np.ones(Cn), 1.0, response_matrix=np.ones((8, Cn)), back_counts=np.ones(Cn),
                  back_errors=np.full(Cn, 2.5), back_ratio=1.0, net_counts=np.ones(Cn), net_errors=np.linspace(0.5, 3.0, Cn))
    n = 400_000
    mean_t = torch.as_tensor(means[None, :], dtype=torch.float64, device='cuda')

    def draw(lk, which, seed, rows=n, mean=mean_t):
        out = torch.empty((rows, Cn), dtype=torch.float64, device='cuda')
        _lib.check(lk._lib.elisa_b200_sample_dev(lk._h, 0, which, C.c_void_p(mean.data_ptr()), mean.shape[0], rows, seed,
                                                 C.c_void_p(out.data_ptr()), Cn, lk._stream()))
        torch.cuda.synchronize()
        return out

    lk = Likelihood([d], M.PowerLaw(), ['wstat'])
    x = draw(lk, 0, 11).cpu().numpy()
    assert np.all(x >= 0) and np.all(x == np.round(x))
    m, v = x.mean(0), x.var(0)
    se_m = np.sqrt(np.maximum(means, 1e-300) / n)
    assert np.all(np.abs(m - means) <= 5 * se_m + 1e-12), (m, means)
    se_v = np.sqrt((means + 2 * means**2) / n) + 1e-12   # var of the sample variance of a Poisson, roughly
    assert np.all(np.abs(v - means) <= 6 * se_v), (v, means)
    for j, lam in enumerate(means):
        if 0.0 < lam <= 60.0:  # the whole pmf
            ks = np.arange(0, int(lam + 10 * np.sqrt(lam) + 10))
            cnt = np.array([(x[:, j] == k).sum() for k in ks])
            p = stats.poisson.pmf(ks, lam)
            keep = n * p > 20
            chi2 = float((((cnt - n * p) ** 2) / (n * p))[keep].sum())
            assert chi2 < keep.sum() + 6 * np.sqrt(2 * keep.sum()), (lam, chi2, keep.sum())
    # reproducible per seed, different across seeds and between the 'on' and 'off' arrays
    assert torch.equal(draw(lk, 0, 11, rows=1000), torch.as_tensor(x[:1000], device='cuda'))
    assert not torch.equal(draw(lk, 0, 12, rows=1000), torch.as_tensor(x[:1000], device='cuda'))
    assert not torch.equal(draw(lk, 1, 11, rows=1000), torch.as_tensor(x[:1000], device='cuda'))
    # one model row per replica
    per = torch.as_tensor(np.tile(means, (1000, 1)), device='cuda')
    assert torch.equal(draw(lk, 0, 11, rows=1000, mean=per), torch.as_tensor(x[:1000], device='cuda'))
    lk.close()
    # Normal draws: chi2 'on' data with net_errors, pgstat background with back_errors
    lk = Likelihood([d], M.PowerLaw(), ['chi2'])
    g = draw(lk, 0, 5).cpu().numpy()
    sig = np.linspace(0.5, 3.0, Cn)
    assert np.all(np.abs(g.mean(0) - means) <= 5 * sig / np.sqrt(n) + 1e-9 * means)
    assert np.all(np.abs(g.std(0) / sig - 1.0) <= 5 / np.sqrt(2 * n))
    z = (g[:, 3] - means[3]) / sig[3]
    assert abs(stats.kstest(z[:50000], 'norm').statistic) < 0.01
    with pytest.raises(_lib.ElisaB200Error):
        draw(lk, 1, 5, rows=10)  # chi2 has no sampled background
    lk.close()
    lk = Likelihood([d], M.PowerLaw(), ['pgstat'])
    b = draw(lk, 1, 5).cpu().numpy()
    assert np.all(np.abs(b.std(0) / 2.5 - 1.0) <= 5 / np.sqrt(2 * n))
    lk.close()
