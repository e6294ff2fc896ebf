"""The one-kernel small-batch path (tiny_eval, csrc/unfold_skel.cuh): a call of <= 1024 parameter vectors on a
small spectrum (E * C * (1 + free parameters) <= 2^18) is ONE launch per spectrum -- model in dual numbers,
FP64 fold in forward mode, statistic, fixed-order reduction -- instead of the five-kernel chunk pipeline.
This is the NUTS shape of BASELINE configs[2] (64 chains x 4 GBM-like detectors).  Checked against the oracle
and against the pipeline (ELISA_B200_TINY=0) for every statistic, with per-replica counts, through the
graph-replayed entry point, and for which calls take it."""

import numpy as np
import pytest
import torch

from elisa_b200 import models as M
from elisa_b200 import synthetic
from test_gpu_parity import _one_spectrum, _theta_for

pytestmark = pytest.mark.gpu
SMALL_BATCH_KERNEL = True  # conftest.py: this module runs with the one-kernel path enabled (the library default)


def _tiny_launches(lk, fn):
    lk.set_profiling(True)
    fn()
    prof = lk.get_profile()
    lk.set_profiling(False)
    return prof['tiny'][1], sum(v[1] for k, v in prof.items() if k != 'tiny')


@pytest.mark.parametrize('n', [1, 8, 64, 1024])
def test_cfg3_small_batches_are_one_launch(lib, n, monkeypatch):
    from elisa_b200 import Likelihood
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_grad_rownorm, rel_err_value

    cfg = synthetic.config3(n=n)
    ll_ref, g_ref = oracle_eval(cfg)
    theta = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:
        tiny, other = _tiny_launches(lk, lambda: lk.loglike_and_grad(theta))
        assert tiny == 1 and other == 0  # ONE launch for the four detectors (same model and statistic), nothing else
        for _ in range(3):
            ll, g = lk.loglike_and_grad(theta)
        ll, g = ll.cpu().numpy(), g.cpu().numpy()
        ll_v = lk.loglike(theta).cpu().numpy()
    assert rel_err_value(ll, ll_ref) <= 1e-11 and rel_err_grad(g, g_ref) <= 1e-11  # plain FP64: far inside 1e-10
    assert np.array_equal(ll_v, ll)
    # the chunk pipeline on the same rows
    monkeypatch.setenv('ELISA_B200_TINY', '0')
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:
        tiny, other = _tiny_launches(lk, lambda: lk.loglike_and_grad(theta))
        assert tiny == 0 and other > 0
        ll_p, g_p = (t.cpu().numpy() for t in lk.loglike_and_grad(theta))
    assert rel_err_value(ll, ll_p) <= RTOL and rel_err_grad(g, g_p) <= RTOL
    if n > 1:
        assert rel_err_grad_rownorm(g_p, g_ref) <= RTOL


@pytest.mark.parametrize('stat', ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'])
def test_every_statistic_with_and_without_replica_counts(lib, stat):
    from parity import compare

    d = _one_spectrum(E=70, C=40, seed=2, stat=stat)
    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    model = M.ExpAbs(Ec=0.7) * M.PowerLaw(K=K) + M.HighECut() * (M.Blackbody(K=K, method='simpson') + M.Gauss(El=6.4))
    theta = _theta_for(model, 37, 4)
    cfg = dict(data=[d], model=[model], stat=[stat], theta=theta)
    err = compare(cfg, specialise=True)
    assert err["value"] <= 1e-11 and err["grad"] <= 1e-11, err
    rng = np.random.default_rng(1)
    cfg['counts_on'] = rng.poisson(40.0, (37, 40)).astype(float)
    if stat in ('pgstat', 'wstat'):
        cfg['counts_off'] = rng.poisson(15.0, (37, 40)).astype(float)
    err = compare(cfg, specialise=True)
    assert err["value"] <= 1e-11 and err["grad"] <= 1e-11, err


def test_which_calls_take_it(lib):
    """Per CALL, not per chunk: a batch above 1024 rows runs the pipeline even when it is evaluated in chunks of
    128 (re-chunking must not change which kernels compute a row); large spectra never take it; the
    interpreter kernels (no NVRTC) never take it; host and device entry points agree bit for bit."""
    from elisa_b200 import Likelihood

    cfg = synthetic.config3(n=1500)
    theta = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    with Likelihood(cfg['data'], cfg['model'], cfg['stat'], chunk=128) as lk:
        assert _tiny_launches(lk, lambda: lk.loglike_and_grad(theta))[0] == 0
        assert _tiny_launches(lk, lambda: lk.loglike_and_grad(theta[:1000]))[0] == 8  # 8 chunks of 128 rows
        ll_d, g_d = lk.loglike_and_grad(theta[:1000])
        ll_h, g_h = lk.loglike_and_grad_host(cfg['theta'][:1000])
        assert np.array_equal(ll_d.cpu().numpy(), ll_h) and np.array_equal(g_d.cpu().numpy(), g_h)
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:  # one chunk: the same numbers
        ll_1, g_1 = lk.loglike_and_grad(theta[:1000])
        assert torch.equal(ll_1, ll_d) and torch.equal(g_1, g_d)
    with Likelihood(cfg['data'], cfg['model'], cfg['stat'], specialise=False) as lk:
        assert _tiny_launches(lk, lambda: lk.loglike_and_grad(theta[:64]))[0] == 0
    big = synthetic.config2(n=64, E=512, C=256)
    with Likelihood(big['data'], big['model'], big['stat']) as lk:
        assert _tiny_launches(lk, lambda: lk.loglike_and_grad(big['theta']))[0] == 0


def test_nan_row_and_clipped_bins(lib):
    from elisa_b200 import Likelihood

    cfg = synthetic.config3(n=16)
    theta = cfg['theta'].copy()
    theta[5, 1] = np.nan
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:
        ll, g = (t.cpu().numpy() for t in lk.loglike_and_grad(theta))
    assert np.isnan(ll[5]).all() and np.isnan(g[5]).all()
    keep = np.arange(16) != 5
    assert np.isfinite(ll[keep]).all() and np.isfinite(g[keep]).all()


@pytest.mark.parametrize('k', [1, 2, 3, 4, 5])
def test_small_versions_of_the_baseline_configs(lib, k):
    """The scaled-down BASELINE configurations (tests/test_gpu_parity.py) through the library default: three of
    them are small enough for the one-kernel path; an explicitly requested fold engine switches it off."""
    from elisa_b200 import Likelihood
    from parity import RTOL, compare
    from test_gpu_parity import small_config

    cfg = small_config(k)
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:
        fits = [d.n_energies * d.n_channels * (1 + m.compile().nparam) <= 2 ** 18 for d, m in zip(cfg['data'], cfg['model'])]
        assert lk.small_batch_kernel == fits and any(fits) == (k in (1, 3, 5))  # cfg 2 / 4 are too large even scaled down
    with Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='dmma') as lk:
        assert not any(lk.small_batch_kernel)
    err = compare(cfg, specialise=True)
    # (the values sit at a few 1e-11 on cfg 1 / 2 / 4 / 5 under EITHER path: x log(mu) - mu - gof cancels four
    # to five digits per channel, so two correctly rounded FP64 evaluations differ by that much)
    assert err['value'] <= RTOL and err['grad'] <= 1e-11, err


def test_small_batch_call_inside_a_cuda_graph(lib):
    """A sampler that captures its leapfrog into a CUDA graph (torch.cuda.graph) can keep the likelihood call inside
    the capture: the one-kernel path is one launch on the capturing stream -- no allocation of its own, no
    synchronisation, no guard round trip -- and a replay with new parameter values in place gives bit for bit what a
    direct call gives."""
    from elisa_b200 import Likelihood

    cfg = synthetic.config3(n=8)
    rng = np.random.default_rng(5)
    theta = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    with Likelihood(cfg['data'], cfg['model'], cfg['stat']) as lk:
        assert all(lk.small_batch_kernel)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream, as torch's capture recipe asks
            for _ in range(3):
                lk.loglike_and_grad(theta)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ll_g, g_g = lk.loglike_and_grad(theta)
            step = theta + 1e-3 * g_g.sum(dim=1)  # something a sampler would do with the gradient, in the same graph
        for _ in range(3):
            new = torch.as_tensor(cfg['theta'] * (1.0 + 0.02 * rng.standard_normal(cfg['theta'].shape)), dtype=torch.float64, device='cuda')
            theta.copy_(new)
            graph.replay()
            torch.cuda.synchronize()
            ll_d, g_d = lk.loglike_and_grad(new)
            assert torch.equal(ll_g, ll_d) and torch.equal(g_g, g_d)
            assert torch.equal(step, new + 1e-3 * g_d.sum(dim=1))
