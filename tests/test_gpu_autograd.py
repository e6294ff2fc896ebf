"""The autodiff binding of the C ABI (elisa_b200/autograd.py): the custom-JVP contract of
INTEGRATION.md section 4 executed through torch.autograd, and a 64-chain HMC on BASELINE cfg 3
(Band, 4 GBM-like detectors, pgstat -- the configuration the reference runs under numpyro NUTS,
infer/fit.py:833-843) whose trajectories agree with the same sampler on the CPU oracle."""

import numpy as np
import pytest
import torch

from elisa_b200 import Likelihood, autograd as AG, synthetic

pytestmark = pytest.mark.gpu


def _oracle_potential(cfg, bij):
    import oracle

    models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
    free = []
    for m in models:
        for p in m.free:
            if all(p is not q for q in free):
                free.append(p)
    specs = [m.to_spec(free) for m in models]
    datas = [d.as_mapping() for d in cfg['data']]

    def potential(u):
        theta, log_jac = bij(u)
        return -(oracle.loglike(cfg['stat'], datas, specs, theta).sum(-1) + log_jac)

    return potential


def test_backward_is_the_transposed_jacobian_product(lib):
    from test_gpu_parity import small_config

    cfg = small_config(2)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    theta = torch.as_tensor(cfg['theta'][:64], dtype=torch.float64, device='cuda').requires_grad_(True)
    w = torch.randn((64, 1), dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(1))
    ll = AG.loglike(lk, theta)
    (w * ll).sum().backward()
    ll2, jac = lk.loglike_and_grad(theta.detach())
    assert torch.equal(ll.detach(), ll2)
    assert torch.allclose(theta.grad, torch.einsum('ns,nsp->np', w, jac), rtol=0, atol=0)
    # forward-mode rule of the same contract: J . t against a central difference of the value
    t = torch.randn_like(theta) * theta.detach().abs()
    _, jt = AG.jvp(lk, theta.detach(), t)
    h = 1e-6
    fd = (lk.loglike(theta.detach() + h * t) - lk.loglike(theta.detach() - h * t)) / (2 * h)
    assert torch.allclose(jt, fd, rtol=1e-5, atol=1e-6 * float(ll2.abs().max()))
    lk.close()


def test_hmc_64_chains_cfg3_matches_the_oracle(lib):
    """Same sampler, same seeds: GPU likelihood through the autograd binding vs the CPU oracle
    differentiated by torch autograd.  Leapfrog integration amplifies differences, so the
    trajectories are short (4 iterations of 6 steps); positions agree to 1e-8, acceptances exactly."""
    cfg = synthetic.config3(n=64)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], guard='lazy', guard_every=4)
    lower = torch.tensor([p.min for p in lk.free], dtype=torch.float64)
    upper = torch.tensor([p.max for p in lk.free], dtype=torch.float64)
    theta0 = torch.as_tensor(cfg['theta'], dtype=torch.float64)
    z = ((theta0 - lower) / (upper - lower)).clamp(1e-9, 1 - 1e-9)
    u0 = torch.log(z) - torch.log1p(-z)
    bij = AG.interval_bijection(lower, upper)

    def gpu_potential(u):
        theta, log_jac = bij(u)
        return -(AG.loglike(lk, theta).sum(-1) + log_jac)

    # mass matrix from the gradient scale of the potential at the start (|grad_d| ~ 1 / sigma_d): a
    # step then moves every coordinate by about a tenth of its posterior width
    q = u0.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(_oracle_potential(cfg, bij)(q).sum(), q)
    mass = g.abs().amax(dim=0) ** 2
    kw = dict(n_iter=4, n_leapfrog=6, step=0.1, seed=7, mass=mass)
    ref = AG.hmc(_oracle_potential(cfg, bij), u0, **kw)
    out = AG.hmc(gpu_potential, u0.cuda(), **kw)
    lk.guard_poll(wait=True)
    assert lk.fold_guard()[0] == 0
    assert out['n_grad'] == ref['n_grad'] == 1 + 4 * 6
    assert torch.equal(out['accept'].cpu(), ref['accept'])
    assert bool(ref['accept'].any()) and float(((out['q'].cpu() - u0).abs() * mass.sqrt()).max()) > 0.05  # the chains did move
    dev = (out['q'].cpu() - ref['q']).abs() / ref['q'].abs().clamp_min(1.0)
    assert float(dev.max()) < 1e-8, float(dev.max())
    lk.close()


def test_lazy_guard_never_synchronises_and_reports(lib):
    """guard='lazy': calls only enqueue; a starved batch is reported on a LATER call."""
    from elisa_b200 import FixedData, models as M
    from elisa_b200.likelihood import GuardTripped

    E = C = 256
    rng = np.random.default_rng(3)
    d = FixedData('starved', np.linspace(1.0, 10.0, E + 1), rng.poisson(20.0, C).astype(float), np.ones(C), 100.0,
                  response_matrix=np.eye(E) * 50.0)
    U = M.UniformParameter
    model = M.Gauss(El=U('El', 5.0, 1.0, 10.0), sigma=U('sigma', 0.05, 0.01, 1.0), K=U('K', 10.0, 1e-3, 1e3))
    theta = np.column_stack([np.linspace(2.0, 9.0, 40), np.full(40, 0.05), np.full(40, 10.0)])
    lk = Likelihood([d], model, ['cstat'], guard='lazy', guard_every=2)
    with pytest.raises(GuardTripped):
        for _ in range(50):
            lk.loglike_and_grad(theta)
            torch.cuda.synchronize()  # let the snapshot land so that the next call sees it
    # the synchronous mode answers the same batch within tolerance (per-row DMMA fallback)
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    ll_ref, g_ref = oracle_eval(dict(data=[d], model=[model], stat=['cstat'], theta=theta))
    ll, g = lk.loglike_and_grad(theta, guard='sync')
    assert rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL and rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
    assert lk.fallback_rows > 0
    lk.close()
