"""Second-order / forward-mode consumers of the reference: the per-channel forward-mode Jacobian
the LM re-fit differentiates (infer/helper.py:588-594, 616-622; infer/fit.py:185-194) and the
Hessian of the joint deviance (infer/helper.py:533-540)."""

import numpy as np
import pytest
import torch

from elisa_b200 import Likelihood

pytestmark = pytest.mark.gpu


def _oracle_pieces(cfg):
    import oracle

    models = [m.compile() if not hasattr(m, 'to_spec') else m for m in cfg['model']]
    free = []
    for m in models:
        for p in m.free:
            if all(p is not q for q in free):
                free.append(p)
    return [m.to_spec(free) for m in models], [d.as_mapping() for d in cfg['data']], oracle


@pytest.mark.parametrize('k', [1, 2, 3])
def test_per_channel_jacobian_against_autograd_of_the_oracle(lib, k):
    from test_gpu_parity import small_config

    cfg = small_config(k)
    n = 6
    theta = torch.as_tensor(cfg['theta'][:n], dtype=torch.float64)
    specs, datas, oracle = _oracle_pieces(cfg)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    for s, name in enumerate(lk.names):
        x, ds, dl = lk.jacobian(s, theta)

        st = oracle.sites(cfg['stat'][s], datas[s], specs[s], theta)
        key, ll_key = 'Non_model', 'loglike'
        jac = torch.autograd.functional.jacobian(lambda t: oracle.sites(cfg['stat'][s], datas[s], specs[s], t)[key].sum(0), theta)
        # jac: (C, n, P) -- row n of the site only depends on row n of theta
        jac = jac.permute(1, 2, 0)
        if cfg['stat'][s] in ('chi2', 'cstat', 'pstat'):  # the site IS the clipped counts the statistic sees
            assert torch.allclose(x.cpu(), st[key], rtol=1e-10)
            scale = jac.abs().amax(dim=(0, 2), keepdim=True).clamp_min(1e-300)
            assert float(((ds.cpu() - jac).abs() / scale).max()) < 1e-9
        # d loglike_c / d theta through dl_dx * ds against autograd of the per-channel log-likelihood
        jll = torch.autograd.functional.jacobian(lambda t: oracle.sites(cfg['stat'][s], datas[s], specs[s], t)[ll_key].sum(0), theta)
        jll = jll.permute(1, 2, 0)
        mine = dl.cpu()[:, None, :] * ds.cpu()
        scale = jll.abs().amax(dim=(0, 2), keepdim=True).clamp_min(1e-300)
        assert float(((mine - jll).abs() / scale).max()) < 1e-9
    lk.close()


def test_hessian_against_autograd_of_the_oracle(lib):
    from test_gpu_parity import small_config

    cfg = small_config(2)
    theta = torch.as_tensor(cfg['truth'], dtype=torch.float64)[None, :] * torch.tensor([[1.0], [1.02]], dtype=torch.float64)
    specs, datas, oracle = _oracle_pieces(cfg)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    H = lk.hessian(theta).cpu()
    for i in range(2):
        ref = torch.autograd.functional.hessian(lambda t: oracle.loglike(cfg['stat'], datas, specs, t[None, :]).sum(), theta[i])
        scale = torch.sqrt(torch.outer(ref.diagonal().abs(), ref.diagonal().abs()))
        assert float(((H[i] - ref).abs() / scale).max()) < 1e-5  # the reference's own test asks rtol 1e-3
    lk.close()
