"""CPU tests of the host side: model algebra mirror, descriptors, data container, the C-ABI
library (loads, exports every declared symbol, refuses to run without a GPU), sharding."""

import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from elisa_b200 import _lib
from elisa_b200 import models as M
from elisa_b200.data import FixedData
from elisa_b200.parallel import shard_bounds

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ model algebra
def test_type_rules_follow_reference():
    """models/model.py:1228-1258."""
    pl, bb, ea, hc = M.PowerLaw(), M.Blackbody(), M.ExpAbs(), M.HighECut()
    assert (pl + bb).type == 'add' and (ea * pl).type == 'add' and (pl * ea).type == 'add' and (ea * hc).type == 'mul'
    with pytest.raises(TypeError):
        pl * bb
    with pytest.raises(TypeError):
        ea + pl
    with pytest.raises(TypeError):
        pl + ea
    with pytest.raises(TypeError):
        pl + 3.0


def test_compile_orders_free_parameters_and_links():
    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    m = (M.ExpAbs(Ec=0.7) * M.PowerLaw(K=K) + M.Blackbody(K=K) + M.PowerLaw()).compile()
    assert m.params_name == ['PowerLaw.alpha', 'PowerLaw.K', 'Blackbody.kT', 'PowerLaw_2.alpha', 'PowerLaw_2.K']
    assert m.nparam == 5 and np.allclose(m.params_default, [1.01, 2.0, 3.0, 1.01, 1.0])
    md, keep = m.to_desc()
    assert md.n_components == 4 and md.n_ops == 7
    comps = keep[0]
    assert comps[0].kind == 18 and comps[0].param_src[0] == -1 and comps[0].param_const[0] == 0.7
    assert list(comps[1].param_src[:2]) == [0, 1] and list(comps[2].param_src[:2]) == [2, 1]  # K linked
    assert list(keep[1]) == [0, 1, _lib.OP_MUL, 2, _lib.OP_ADD, 3, _lib.OP_ADD]
    spec = m.to_spec()
    assert spec[0] == '+' and spec[1][0] == '+' and spec[1][1][0] == '*'
    assert spec[1][1][2][2]['K'] == ('theta', 1) and spec[1][2][2]['K'] == ('theta', 1)


def test_component_constructor_semantics():
    """models/model.py:1426-1465: None -> free default, number -> fixed, Parameter -> as given."""
    c = M.CutoffPL(alpha=1.5, Ec=None)
    assert c.params['alpha'].fixed and c.params['alpha'].default == 1.5
    assert not c.params['Ec'].fixed and c.params['Ec'].default == 15.0
    assert M.SmoothlyBrokenPL().params['rho'].fixed  # fixed=True in the reference _config
    with pytest.raises(TypeError):
        M.PowerLaw(beta=1.0)
    with pytest.raises(NotImplementedError):
        M.CutoffPL(method='romberg')
    with pytest.raises(ValueError):
        M.PLPhFlux(10.0, 1.0)
    assert M.PLEnFlux(1.0, 10.0).aux == (1.0, 10.0)
    assert {c.__name__ for c in M.ADDITIVE} == set(M.__all__[:16]) and len(M.MULTIPLICATIVE) == 7


def test_eval_params_validation():
    m = (M.PowerLaw() + M.Gauss()).compile()
    with pytest.raises(ValueError):
        m._theta_from(np.zeros((3, 4)))
    with pytest.raises(ValueError):
        m._theta_from({'PowerLaw.alpha': 1.0})
    t = m._theta_from({n: np.full((2, 3), i) for i, n in enumerate(m.params_name)})
    assert t.shape == (2, 3, 5) and t[0, 0].tolist() == [0, 1, 2, 3, 4]
    assert m._theta_from(None).tolist() == m.params_default.tolist()


# ------------------------------------------------------------------ data container
def test_fixed_data_validation_and_sparse_view():
    eg = np.linspace(1, 2, 6)
    with pytest.raises(ValueError):
        FixedData('x', eg, np.zeros(4), 1.0, 1.0, response_matrix=np.zeros((4, 4)))
    with pytest.raises(ValueError):
        FixedData('x', eg, np.zeros(4), 1.0, 1.0)
    indptr = np.array([0, 2, 2, 3, 5]); idx = np.array([0, 1, 2, 3, 4]); val = np.arange(1.0, 6.0)
    d = FixedData('x', eg, np.zeros(4), 1.0, 1.0, sparse_matrix=(indptr, idx, val), response_sparse=True)
    R = d.dense_response()
    assert R.shape == (5, 4) and R[0, 0] == 1 and R[1, 0] == 2 and R[2, 2] == 3 and R[4, 3] == 5 and R.sum() == 15
    assert not d.has_back and d.as_mapping()['response_matrix'].shape == (5, 4)


# ------------------------------------------------------------------ the C-ABI library
def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'elisa_b200.h')).read()
    declared = set(re.findall(r'\b(elisa_b200_\w+)\s*\(', header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.elisa_b200_abi_version() == _lib.ABI_VERSION


def test_struct_layout_matches_header(lib):
    # 4 bytes of padding before the doubles; two table pointers, table_len + padding, four scales
    assert C.sizeof(_lib.ComponentDesc) == 4 + 4 + 5 * 4 + 4 + 5 * 8 + 16 + 8 + 8 + 8 + 32 + 2 * 4 + 2 * 4
    for which, st in enumerate((_lib.ComponentDesc, _lib.NormDesc, _lib.ModelDesc, _lib.SpectrumDesc, _lib.PlanDesc)):
        assert lib.elisa_b200_sizeof(which) == C.sizeof(st), st.__name__
    assert _lib.PlanDesc.chunk.offset == 24 and _lib.SpectrumDesc.exposure.offset % 8 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(lib):
    """Without a CUDA device plan creation fails loudly; nothing is computed on the CPU."""
    desc = _lib.PlanDesc(_lib.ABI_VERSION, 0, 1, 1, None, 0, 0)
    h = C.c_void_p()
    rc = lib.elisa_b200_plan_create(C.byref(desc), C.byref(h))
    assert rc < 0 and not h.value
    assert lib.elisa_b200_last_error()
    from elisa_b200 import ElisaB200Error, Likelihood

    d = FixedData('x', np.linspace(1, 2, 5), np.ones(3), 1.0, 1.0, response_matrix=np.ones((4, 3)))
    with pytest.raises(ElisaB200Error):
        Likelihood([d], M.PowerLaw(), ['cstat'])


def test_argument_errors_do_not_need_a_gpu(lib):
    h = C.c_void_p()
    assert lib.elisa_b200_plan_create(None, C.byref(h)) == -1
    assert b'desc is null' in lib.elisa_b200_last_error()
    bad = _lib.PlanDesc(99, 0, 1, 1, None, 0, 0)
    assert lib.elisa_b200_plan_create(C.byref(bad), C.byref(h)) == -1
    assert lib.elisa_b200_loglike_dev(None, None, None, None, 4, None, None) == -1
    assert lib.elisa_b200_plan_launch_count(None) == 0


def test_specialised_kernel_compiles_for_every_component(lib):
    """NVRTC needs no GPU: the generated translation unit of every built-in (and of a
    composite with fixed + linked parameters) must compile for sm_100a."""
    K = M.UniformParameter('K', 2.0, 1e-3, 1e3)
    models = [c(1.0, 10.0) if c in (M.PLEnFlux, M.PLPhFlux) else c() for c in M.ADDITIVE]
    models += [m() * M.PowerLaw() for m in M.MULTIPLICATIVE]
    models += [M.CutoffPL(method='simpson') + M.Blackbody(method='simpson'),
               M.Constant() * (M.ExpAbs(Ec=0.7) * M.PowerLaw(K=K) + M.HighECut() * (M.Blackbody(K=K) + M.Gauss(El=6.4)))]
    buf = C.create_string_buffer(1 << 16)
    for m in models:
        cm = m.compile()
        md, _keep = cm.to_desc()
        n = lib.elisa_b200_specialise_check(C.byref(md), max(cm.nparam, 1), buf, len(buf))
        assert n > 0, (repr(m), lib.elisa_b200_last_error().decode()[:500])
        src = buf.value.decode()
        assert 'unfold_generic<ModelG' in src and src.count('leaf_bin_pc<') == 2 * len(cm.leaves)  # ModelV + ModelG


def test_specialised_kernel_compiles_for_convolution_models(lib):
    """Energy shifts become literals of the leaf calls; a flux normalisation adds a pre-pass over
    its own grid in `norms` (value accumulator in ModelV and ModelG, one more per tangent column
    in ModelG) and one multiply in `bin`."""
    import conv_cases

    buf = C.create_string_buffer(1 << 16)
    for name, (m, _theta) in conv_cases.cases().items():
        cm = m.compile()
        md, _keep = cm.to_desc()
        n = lib.elisa_b200_specialise_check(C.byref(md), max(cm.nparam, 1), buf, len(buf))
        assert n > 0, (name, lib.elisa_b200_last_error().decode()[:500])
        src = buf.value.decode()
        assert src.count('acc_v = warp_sum(acc_v)') == 2 * len(cm.norms)  # ModelV + ModelG
        assert src.count('const GridView gv = a.ngv[') == 2 * len(cm.norms)
    # a malformed program: the normalisation opcode without a descriptor
    cm = conv_cases.cases()['phflux_pl'][0].compile()
    md, _keep = cm.to_desc()
    md.n_norms = 0
    assert lib.elisa_b200_specialise_check(C.byref(md), 2, buf, len(buf)) < 0


# ------------------------------------------------------------------ sharding of the N axis
def test_shard_bounds():
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for n in (1, 7, 64, 1_000_003):
        b = shard_bounds(n, 8)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(7))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    from elisa_b200.parallel import ShardedLikelihood, sharded_map

    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    theta = torch.arange(n * 3, dtype=torch.float64).reshape(n, 3)
    counts = torch.arange(n * 2, dtype=torch.float64).reshape(n, 2)

    class Fake:  # stands in for elisa_b200.Likelihood: same call signature, CPU arithmetic
        calls = []

        def loglike_and_grad(self, t, on=None, off=None):
            Fake.calls.append(len(t))
            return (t.sum(1, keepdim=True) + on.sum(1, keepdim=True)), t.unsqueeze(1) * 2.0

        def loglike(self, t, on=None, off=None):
            return t.sum(1, keepdim=True)

        nparam, device = 3, 0

        def fit(self, t, counts_on=None, counts_off=None, **kw):
            if len(t) == 1 and counts_on is not None:  # one start vector for every replica (Likelihood.fit)
                t = t.expand(len(counts_on), -1)
            return {'theta': t + 1.0, 'deviance': t.sum(1), 'grad_norm': t[:, 0] * 0.0,
                    'n_iter': torch.full((len(t),), 3), 'valid': t[:, 0] >= 0, 'loglike': -0.5 * t.sum(1)}

    sl = ShardedLikelihood(Fake())
    fit = sl.fit(theta, counts)
    fit_ok = (torch.equal(fit['theta'], theta + 1.0) and torch.equal(fit['deviance'], theta.sum(1))
              and fit['n_iter'].dtype == torch.int64 and bool((fit['n_iter'] == 3).all()) and fit['valid'].dtype == torch.bool
              and bool(fit['valid'].all()) and fit['theta'].shape[0] == n)
    # one start vector broadcast over the replicas: the REPLICAS are sharded, not the vector
    bfit = sl.fit(theta[:1], counts)
    fit_ok = fit_ok and bfit['theta'].shape == (n, 3) and torch.equal(bfit['theta'], (theta[:1] + 1.0).expand(n, 3))
    bfit1 = sl.fit(theta[0], counts)  # 1-D theta = ONE parameter vector, not P rows
    fit_ok = fit_ok and torch.equal(bfit1['theta'], bfit['theta'])
    ll, g = sl.loglike_and_grad(theta, counts)
    local = sl.loglike(theta, gather=False)
    same = sharded_map(lambda t: t * 1.0, (theta,), n)
    ok = (torch.equal(ll, theta.sum(1, keepdim=True) + counts.sum(1, keepdim=True)) and torch.equal(g, theta.unsqueeze(1) * 2.0)
          and torch.equal(same, theta) and fit_ok)
    q.put((rank, bool(ok), Fake.calls[0], local.shape[0]))
    dist.destroy_process_group()


def _peer_setup_worker(rank, world, port, q):
    import torch.distributed as dist

    from elisa_b200 import _lib
    from elisa_b200.parallel import PeerGather

    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)

    class Fake:  # the attributes PeerGather reads of a Likelihood
        _lib, _S, _P, device, _h = _lib.load(), 1, 3, 0, None

    try:
        PeerGather(Fake(), 10)
        out = 'no error'
    except RuntimeError as exc:
        out = str(exc)
    dist.barrier()  # the ranks are still in step after the failure
    q.put((rank, out))
    dist.destroy_process_group()


def test_peer_gather_setup_fails_on_every_rank_together():
    """Without a GPU the CUDA-IPC allocation fails: every rank must raise the same RuntimeError (naming the ranks
    that failed) and stay in step -- a rank that raised alone would leave the others in a collective for ever."""
    import socket

    import torch.multiprocessing as mp

    if torch.cuda.is_available():
        pytest.skip('needs a machine WITHOUT a GPU: the allocation has to fail')
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_setup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] and 'peer gather could not be set up' in res[0][1] and 'rank 0' in res[0][1] and 'rank 1' in res[0][1], res


@pytest.mark.parametrize('n,world', [(7, 2), (64, 2), (7, 3)])
def test_sharded_map_gloo(n, world):
    """world 2: even and uneven blocks; world 3 with 7 rows: blocks of 3, 2, 2 -- the in-place
    compaction of the padded gather moves the last block."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    b = shard_bounds(n, world)
    for rank, ok, first_call, local_rows in res:
        assert ok
        assert first_call == b[rank][1] - b[rank][0] == local_rows


# ------------------------------------------------------------------ table-driven exp / log
def test_fastmath_against_libm(tmp_path):
    """csrc/fastmath.cuh compiles for the host too: its exp / log / expm1 against glibc's long-double
    versions over 2e6 arguments -- <= 1.5 ulp (4.5 for expm1 = exp - 1), specials handed to libm."""
    import subprocess

    exe = tmp_path / 'fastmath_check'
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(['g++', '-O2', '-std=c++17', '-I', os.path.join(root, 'elisa_b200', 'csrc'),
                    os.path.join(root, 'tools', 'fastmath_check.cc'), '-o', str(exe)], check=True)
    res = subprocess.run([str(exe), '2000000'], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout
    assert 'exp(-800)=0 exp(800)=inf exp(nan)=nan log(0)=-inf log(-1)=-nan log(inf)=inf' in res.stdout


def test_peer_gather_entry_points_reject_bad_arguments(lib):
    """No GPU here: the argument checks of the peer-gather entry points (include/elisa_b200.h) come before any CUDA call."""
    ptr, h = C.c_void_p(), C.create_string_buffer(64)
    assert lib.elisa_b200_peer_alloc(0, C.byref(ptr), h) == -1
    assert lib.elisa_b200_peer_alloc(64, None, h) == -1
    assert lib.elisa_b200_peer_open(None, C.byref(ptr)) == -1
    assert lib.elisa_b200_peer_close(None) == 0 and lib.elisa_b200_peer_free(None) == 0
    assert lib.elisa_b200_plan_set_gather(None, 2, 0, None, None, None, 0) == -1
    assert b'null plan' in lib.elisa_b200_last_error()
