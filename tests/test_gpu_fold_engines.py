"""The two response-fold engines behind the same C ABI: the sliced-integer fold on the int8
tcgen05 tensor cores (default; csrc/fold_i8.cuh) and the native FP64 DMMA fold
(csrc/gemm_f64.cuh).  Both must hold the 1e-10 tolerance of BASELINE.json's north_star
against the CPU oracle; the engine on its own is checked as an FP64-class GEMM."""

import numpy as np
import pytest

from elisa_b200 import models as M
from elisa_b200 import synthetic

pytestmark = pytest.mark.gpu

from test_gpu_parity import _one_spectrum, _theta_for, small_config  # noqa: E402


@pytest.mark.parametrize('fold', ['i8', 'dmma'])
@pytest.mark.parametrize('k', [1, 2, 3, 4, 5])
def test_config_parity_both_engines(lib, k, fold):
    from parity import RTOL, compare

    cfg = small_config(k)
    err = compare(cfg, fold=fold)
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err
    assert compare(cfg, grad=False, fold=fold)['value'] <= RTOL


@pytest.mark.parametrize('stat', ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'])
def test_engines_agree_per_statistic(lib, stat):
    """same inputs through both engines: they agree far inside the tolerance"""
    from parity import gpu_eval, rel_err_grad, rel_err_value

    d = _one_spectrum(E=300, C=130, seed=21)
    d.spec_counts[::9] = 0.0
    d.back_counts[::4] = 0.0
    model = M.CutoffPL() + M.Gauss(El=20.0, sigma=2.0)
    cfg = dict(data=[d], model=[model], stat=[stat], theta=_theta_for(model, 300, 5))
    ll_a, g_a = gpu_eval(cfg, fold='i8')
    ll_b, g_b = gpu_eval(cfg, fold='dmma')
    assert rel_err_value(ll_a, ll_b) <= 2e-12
    ok = np.isfinite(g_b)
    assert np.array_equal(ok, np.isfinite(g_a))
    assert rel_err_grad(np.where(ok, g_a, 0.0), np.where(ok, g_b, 0.0)) <= 2e-11


@pytest.mark.parametrize('slices,tol', [(4, 1e-5), (5, 1e-7), (6, 1e-10), (7, 1e-11)])
def test_slices_set_the_accuracy(lib, slices, tol):
    """4 digits = the FP32-class variant of the north star (stated tolerance 1e-5),
    6 digits = the FP64 tolerance 1e-10, 7 digits = full FP64"""
    from parity import compare

    err = compare(small_config(2), fold='i8', slices=slices)
    assert err['value'] <= tol and err['grad'] <= tol, (slices, err)


def test_i8_chunking_host_path_and_replica_counts(lib):
    from parity import RTOL, compare, gpu_eval

    cfg = small_config(4)  # band-sparse response, per-replica counts
    assert cfg.get('counts_on') is not None
    err = compare(cfg, fold='i8')
    assert err['value'] <= RTOL and err['grad'] <= RTOL, err
    ll_d, g_d = gpu_eval(cfg, fold='i8')
    ll_h, g_h = gpu_eval(cfg, fold='i8', host=True)
    assert np.array_equal(ll_d, ll_h) and np.array_equal(g_d, g_h)
    ll_c, g_c = gpu_eval(cfg, fold='i8', chunk=128)  # chunking must not change a single bit
    assert np.array_equal(ll_d, ll_c) and np.array_equal(g_d, g_c)


def test_i8_nan_rows_stay_nan(lib):
    """a parameter vector that makes the model NaN yields NaN for that row only"""
    from parity import gpu_eval

    cfg = small_config(1)
    cfg['theta'] = cfg['theta'].copy()
    cfg['theta'][3, 1] = np.nan
    ll, g = gpu_eval(cfg, fold='i8')
    assert np.isnan(ll[3]).all() and np.isnan(g[3]).all()
    keep = np.ones(len(ll), bool)
    keep[3] = False
    assert np.isfinite(ll[keep]).all() and np.isfinite(g[keep]).all()


@pytest.mark.parametrize('shape', [(128, 64, 128), (300, 200, 1000), (77, 65, 129), (1, 1, 1), (513, 1024, 2048)])
@pytest.mark.parametrize('slices', [4, 5, 6, 7])
def test_sliced_gemm_is_an_fp64_class_gemm(lib, shape, slices):
    """normwise error <= 2^(-8 slices + 6) * (max|a| sum|b| + sum|a| max|b|), rows of wildly
    different magnitude, mixed signs"""
    import torch

    from elisa_b200 import sliced_gemm

    m, n, k = shape
    g = torch.Generator(device='cuda').manual_seed(m * 7 + n)
    a = torch.randn(m, k, dtype=torch.float64, device='cuda', generator=g)
    b = torch.randn(n, k, dtype=torch.float64, device='cuda', generator=g)
    a *= torch.exp(20.0 * torch.randn(m, 1, dtype=torch.float64, device='cuda', generator=g))
    b *= torch.exp(20.0 * torch.randn(n, 1, dtype=torch.float64, device='cuda', generator=g))
    d, ms = sliced_gemm(a, b, slices)
    ref = a @ b.T
    bound = a.abs().amax(1, keepdim=True) * b.abs().sum(1)[None, :] + a.abs().sum(1, keepdim=True) * b.abs().amax(1)[None, :]
    err = ((d - ref).abs() / bound).max().item()
    assert err <= max(2.0 ** (-8 * slices + 6), 4e-16), err
    assert ms > 0.0


def test_sliced_gemm_edge_rows(lib):
    import torch

    from elisa_b200 import sliced_gemm

    a = torch.rand(130, 300, dtype=torch.float64, device='cuda')
    b = torch.rand(70, 300, dtype=torch.float64, device='cuda')
    a[5] = 0.0            # an all-zero row
    a[6] *= 1e-300        # subnormal-adjacent row
    a[7] *= 1e300         # huge row
    a[8, 17] = float('nan')
    b[9] = 0.0
    d, _ = sliced_gemm(a, b, 6)
    ref = a @ b.T
    assert torch.all(d[5] == 0.0) and torch.all(d[:, 9][torch.arange(130, device='cuda') != 8] == 0.0)
    assert torch.isnan(d[8]).all()
    keep = torch.ones(130, dtype=torch.bool, device='cuda')
    keep[8] = False
    assert torch.allclose(d[keep], ref[keep], rtol=1e-12, atol=0.0)


def test_accuracy_guard_quiet_on_baseline_shapes_and_trips_on_starved_channels(lib):
    """The statistic kernel estimates the log-likelihood error the sliced fold may have caused.
    BASELINE-like inputs stay ~1e4 below the trip level; a model whose prediction in populated
    channels lies ~1e-30 below its brightest bin (line only, diagonal response) trips it, and
    the guarded Python host side re-evaluates with the FP64 DMMA twin."""
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    from elisa_b200 import Likelihood
    from elisa_b200.data import FixedData

    cfg = small_config(2)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8')
    lk.loglike_and_grad(cfg['theta'])
    flagged, worst = lk.fold_guard()
    assert flagged == 0 and 0.0 < worst < 3e-12, (flagged, worst)
    lk.close()
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    assert lk._auto_guard
    lk.loglike_and_grad(cfg['theta'])
    assert lk.guard_fallbacks == 0
    lk.close()

    # narrow line through a diagonal response, counts in every channel
    E = C = 256
    eg = np.linspace(1.0, 10.0, E + 1)
    rng = np.random.default_rng(3)
    d = FixedData('starved', eg, rng.poisson(20.0, C).astype(float), np.ones(C), 100.0,
                  response_matrix=np.eye(E) * 50.0)
    U = M.UniformParameter
    model = M.Gauss(El=U('El', 5.0, 1.0, 10.0), sigma=U('sigma', 0.05, 0.01, 1.0), K=U('K', 10.0, 1e-3, 1e3))
    theta = _theta_for(model, 40, 11)
    cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
    ll_ref, g_ref = oracle_eval(cfg)
    lk = Likelihood([d], model, ['cstat'], fold='i8')        # unguarded integer fold: flags rows
    ll_i8, _ = lk.loglike_and_grad(theta)
    flagged, worst = lk.fold_guard()
    assert flagged == len(theta) and worst > 1e-6, (flagged, worst)
    lk.close()
    lk = Likelihood([d], model, ['cstat'])                    # guarded: falls back to the plan's DMMA fold
    ll, g = lk.loglike_and_grad(theta)
    assert lk.guard_fallbacks == 1
    assert rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL
    assert rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
    ll_h, g_h = lk.loglike_and_grad_host(theta)               # the C host entry point does it itself
    assert lk.guard_fallbacks == 2 and rel_err_value(ll_h, ll_ref) <= RTOL and rel_err_grad(g_h, g_ref) <= RTOL
    assert all(s > 0 for s in lk.fold_slices)                 # ... and the plan is back on the integer fold
    ll2, _ = lk.loglike_and_grad(theta[:3])
    assert lk.guard_fallbacks == 3 and rel_err_value(ll2.cpu().numpy(), ll_ref[:3]) <= RTOL
    lk.close()


@pytest.mark.parametrize('fold', ['i8', 'dmma'])
def test_small_batches_replay_a_cuda_graph_with_identical_results(lib, fold, monkeypatch):
    """n <= 1024 without per-replica counts: the launch sequence is captured on the second call
    of a shape and replayed afterwards (new theta values, new output buffers): bit-identical to
    the direct launches (ELISA_B200_GRAPHS=0)."""
    import torch

    from elisa_b200 import Likelihood

    cfg = synthetic.config3(n=64)
    rng = np.random.default_rng(0)
    thetas = [cfg['theta'] * (1.0 + 0.01 * rng.standard_normal(cfg['theta'].shape)) for _ in range(5)]
    monkeypatch.setenv('ELISA_B200_GRAPHS', '0')
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold=fold)
    ref = [tuple(x.clone() for x in lk.loglike_and_grad(t)) for t in thetas]
    ref_v = [lk.loglike(t).clone() for t in thetas]
    lk.close()
    monkeypatch.delenv('ELISA_B200_GRAPHS')
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold=fold)
    n0 = lk.launch_count
    for rep in range(2):
        for t, (ll_ref, g_ref), v_ref in zip(thetas, ref, ref_v):
            ll, g = lk.loglike_and_grad(t)
            assert torch.equal(ll, ll_ref) and torch.equal(g, g_ref)
            assert torch.equal(lk.loglike(t), v_ref)
    assert lk.launch_count > n0
    ll, g = lk.loglike_and_grad(thetas[0][:7])  # another shape: its own graph
    assert torch.equal(ll, ref[0][0][:7]) and torch.equal(g, ref[0][1][:7])
    lk.close()


def test_narrow_band_sparse_response_keeps_the_dmma_fold(lib):
    """Default engine choice: under 5 % stored (cfg 4's CCD-like RSP: 2.7 %) the plan stays on the
    DMMA fold; fold='i8' overrides; both agree with the oracle."""
    from elisa_b200 import Likelihood
    from parity import RTOL, compare

    cfg = synthetic.config4(n=40, size=4096)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    assert lk.fold_slices == [0]
    lk.close()
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8')
    assert lk.fold_slices == [6]
    lk.close()
    dense = small_config(2)
    lk = Likelihood(dense['data'], dense['model'], dense['stat'])
    assert all(s == 6 for s in lk.fold_slices)
    lk.close()
    tiny = _one_spectrum(E=70, C=40)  # fewer than 64 channels: nothing for the integer engine to gain
    lk = Likelihood([tiny], M.PowerLaw(), ['cstat'])
    assert lk.fold_slices == [0]
    lk.close()
    lk = Likelihood([tiny], M.PowerLaw(), ['cstat'], fold='i8')
    assert lk.fold_slices == [6]
    lk.close()
    for fold in (None, 'i8'):
        err = compare(cfg, fold=fold) if fold else compare(cfg)
        assert err['value'] <= RTOL and err['grad'] <= RTOL, (fold, err)


def _random_case(seed):
    from test_gpu_random import _random_data, _random_model

    rng = np.random.default_rng(1000 + seed)
    stat = ['chi2', 'cstat', 'pstat', 'pgstat', 'wstat'][seed % 5]
    E = int(rng.choice([33, 64, 127, 128, 129, 200, 257, 500]))
    C = int(rng.choice([17, 63, 64, 65, 128, 190, 300]))
    n = int(rng.choice([1, 5, 127, 128, 129, 300, 700]))
    model = _random_model(rng)
    d = _random_data(rng, E, C, stat)
    return dict(data=[d], model=[model.compile()], stat=[stat], theta=_theta_for(model, n, seed))


@pytest.mark.parametrize('seed', [6, 17])
def test_balancing_removes_the_starved_channel_weakness(lib, seed, monkeypatch):
    """Bad-fit rows whose starved channels (n >> mu) tower over W: the unbalanced integer fold
    loses gradient digits (and says so), the balanced one is accurate without a flag."""
    from elisa_b200 import Likelihood
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    cfg = _random_case(seed)
    ll_ref, g_ref = oracle_eval(cfg)
    out = {}
    for bal in ('0', '1', 'flag'):
        monkeypatch.setenv('ELISA_B200_BALANCE', '1' if bal == 'flag' else bal)
        lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8', balance=bal != 'flag')
        ll, g = lk.loglike_and_grad(cfg['theta'])
        flagged, _ = lk.fold_guard()
        lk.close()
        out[bal] = (rel_err_value(ll.cpu().numpy(), ll_ref), rel_err_grad(g.cpu().numpy(), g_ref), flagged)
    assert out['0'][1] > RTOL and out['0'][2] > 0, out     # the weakness, detected
    assert out['flag'] == out['0']                          # ELISA_FLAG_NO_BALANCE == the env switch
    assert out['1'][0] <= 1e-11 and out['1'][1] <= 1e-11 and out['1'][2] == 0, out


def test_stale_envelopes_are_renewed(lib):
    """Envelopes taken on a steep spectrum (index 4 over three decades: columns 1e12 apart, floored
    28 bits below the peak) do not fit a flat one: the guard flags the call, the host side renews
    the envelopes from the call's own rows and repeats it on the integer fold -- no DMMA fallback."""
    from elisa_b200 import FixedData, Likelihood
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    rng = np.random.default_rng(5)
    E, C = 256, 128
    eg = np.geomspace(0.1, 100.0, E + 1)
    R = np.abs(rng.standard_normal((E, C))) + 0.1
    d = FixedData('d', eg, rng.poisson(50.0, C).astype(float), np.ones(C), 10.0, response_matrix=R)
    a = M.UniformParameter('alpha', 2.0, -3.0, 10.0)
    K = M.UniformParameter('K', 1.0, 1e-10, 1e10)
    model = M.PowerLaw(alpha=a, K=K)
    lk = Likelihood([d], model, ['cstat'])
    assert lk.fold_slices == [6]
    steep = np.column_stack([4.0 + 0.01 * rng.random(200), np.full(200, 5.0)])
    flat = np.column_stack([0.3 + 0.01 * rng.random(200), np.full(200, 0.05)])
    # (back on the steep spectrum the flat envelopes merely give the unbalanced accuracy, which
    # this dense response tolerates: no renewal needed)
    for theta, renew in ((steep, 0), (flat, 1), (flat, 1), (steep, 1)):
        cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
        ll_ref, g_ref = oracle_eval(cfg)
        ll, g = lk.loglike_and_grad(theta)
        assert rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL and rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
        assert lk.rebalances == renew and lk.guard_fallbacks == 0, (lk.rebalances, lk.guard_fallbacks)
    lk.close()
    # the synchronous C entry point answers the same way
    lk = Likelihood([d], model, ['cstat'])
    for theta in (steep, flat):
        cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta)
        ll_ref, g_ref = oracle_eval(cfg)
        ll, g = lk.loglike_and_grad_host(theta)
        assert rel_err_value(ll, ll_ref) <= RTOL and rel_err_grad(g, g_ref) <= RTOL
    assert lk.guard_fallbacks == 0
    lk.close()


def test_pilot_rows_are_spread_over_the_batch(lib):
    """A scan of a line energy over a parameter grid: the pilot is 128 rows gathered from the
    whole batch (theta and per-replica counts), but the rows differ in SHAPE -- every row has its
    line somewhere else on a continuum 1e5 below it -- which no column scaling can balance.  The
    guard sees it (renewal does not help, the call is repeated on the DMMA fold) and the result
    holds the tolerance."""
    from elisa_b200 import FixedData, Likelihood
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    rng = np.random.default_rng(12)
    E, C, n = 256, 128, 1500
    eg = np.linspace(1.0, 10.0, E + 1)
    R = np.zeros((E, C))
    for e in range(E):
        c = e // 2
        R[e, max(0, c - 3): c + 4] = 30.0 * (0.5 + rng.random(min(C, c + 4) - max(0, c - 3)))
    d = FixedData('d', eg, rng.poisson(40.0, C).astype(float), np.ones(C), 20.0, response_matrix=R)
    U = M.UniformParameter
    model = M.Gauss(El=U('El', 5.0, 1.0, 10.0), sigma=U('sigma', 0.15, 0.01, 1.0), K=U('K', 5.0, 1e-3, 1e3))
    model = model + M.PowerLaw(alpha=U('alpha', 1.5, -3.0, 10.0), K=U('Kp', 0.02, 1e-6, 1e3))
    theta = np.column_stack([np.linspace(2.0, 9.0, n), np.full(n, 0.15), np.full(n, 5.0), np.full(n, 1.5), np.full(n, 0.02)])
    counts = rng.poisson(40.0, (n, C)).astype(float)
    for kw in ({}, {'counts_on': counts}):
        cfg = dict(data=[d], model=[model], stat=['cstat'], theta=theta, **kw)
        ll_ref, g_ref = oracle_eval(cfg)
        lk = Likelihood([d], model, ['cstat'])
        assert lk.fold_slices == [6]
        ll, g = lk.loglike_and_grad(theta, **kw)
        assert rel_err_value(ll.cpu().numpy(), ll_ref) <= RTOL and rel_err_grad(g.cpu().numpy(), g_ref) <= RTOL
        assert lk.rebalances <= 1 and lk.guard_fallbacks <= 1, (lk.rebalances, lk.guard_fallbacks)
        lk.close()


@pytest.mark.parametrize('engine', ['i8', 'dmma'])
@pytest.mark.parametrize('k', [1, 2, 3, 4, 5])
def test_gradient_by_contraction_matches_tangent_planes(lib, k, engine, monkeypatch):
    """Both fold engines: the default gradient (transposed fold stores G, the component kernel contracts it
    with tangents it recomputes in registers) against the tangent-plane path (ELISA_B200_GRAD=planes:
    dU/dtheta written by the unfold kernel, contracted in the transposed fold's epilogue) -- with the
    specialised kernels and with the interpreter."""
    from elisa_b200 import Likelihood
    from parity import RTOL, oracle_eval, rel_err_grad, rel_err_value

    cfg = small_config(k)
    ll_ref, g_ref = oracle_eval(cfg)
    out = {}
    for mode in ('planes', 'contract'):
        monkeypatch.setenv('ELISA_B200_GRAD', mode)
        for spec in (True, False):
            lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold=engine, specialise=spec)
            ll, g = lk.loglike_and_grad(cfg['theta'], counts_on=cfg.get('counts_on'))
            lk.close()
            out[mode, spec] = (ll.cpu().numpy(), g.cpu().numpy())
            assert rel_err_value(out[mode, spec][0], ll_ref) <= RTOL and rel_err_grad(out[mode, spec][1], g_ref) <= RTOL
    for spec in (True, False):
        # (values: the planes path takes them from the dual-number kernel, the contraction path from the
        # value kernel -- the same numbers up to FMA contraction, which the E^(1-a) / (1-a) cancellation
        # of the power-law integral amplifies to a few 1e-11 in cfg 1 / 5; each is within 1e-10 of the oracle)
        assert rel_err_value(out['contract', spec][0], out['planes', spec][0]) <= 1e-10
        assert rel_err_grad(out['contract', spec][1], out['planes', spec][1]) <= 1e-10
