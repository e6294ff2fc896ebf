"""The multi-GPU path on real NCCL hardware (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_nccl.py -m gpu`): every rank owns a plan on its GPU, evaluates its block of the
batch, and one all_gather_into_tensor per result puts the full arrays on every rank -- equal to
the single-GPU evaluation to the fold engine's accuracy (each rank's plan balances its integer fold
around a pilot taken from its OWN block of rows, so the last bits differ between plans; rank 0's own
block is reproduced bit for bit)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    import torch.distributed as dist

    from elisa_b200 import Likelihood, synthetic
    from elisa_b200.parallel import ShardedLikelihood

    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world,
                            device_id=torch.device(f'cuda:{rank}'))
    cfg = synthetic.config2(n=n, E=512, C=256)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], device=rank)
    theta = torch.as_tensor(cfg['theta'], dtype=torch.float64, device=f'cuda:{rank}')
    sl = ShardedLikelihood(lk)
    ll, g = sl.loglike_and_grad(theta)
    ll_local, _ = sl.loglike_and_grad(theta, gather=False)
    ok = ll.shape == (n, 1) and g.shape == (n, 1, 5)
    if rank == 0:
        ll1, g1 = lk.loglike_and_grad(theta)  # the whole batch on one GPU
        lo = ll_local.shape[0]
        close = lambda a, b: bool(((a - b).abs() <= 1e-11 * b.abs().amax(dim=0, keepdim=True)).all())
        ok = ok and close(ll, ll1) and close(g, g1) and torch.equal(ll[:lo], ll1[:lo]) and torch.equal(g[:lo], g1[:lo])
    # cfg 3: 64 chains, sharded, lazily guarded (no host synchronisation per call)
    cfg3 = synthetic.config3(n=64)
    lk3 = Likelihood(cfg3['data'], cfg3['model'], cfg3['stat'], device=rank, guard='lazy')
    th3 = torch.as_tensor(cfg3['theta'], dtype=torch.float64, device=f'cuda:{rank}')
    s3 = ShardedLikelihood(lk3)
    for _ in range(3):
        ll3, g3 = s3.loglike_and_grad(th3)
    if rank == 0:
        ref = Likelihood(cfg3['data'], cfg3['model'], cfg3['stat'], device=rank)
        l3, gg3 = ref.loglike_and_grad(th3)
        ok = ok and close(ll3, l3) and close(g3, gg3)
        ref.close()
    # the gather fused into the evaluation (peer-to-peer pushes of the finished chunks, copy engines): the same bits as
    # the NCCL gather of the same plan's results, call after call (two sets of arrays used in turn), for a batch of
    # several chunks, for unequal blocks, value only, and on the one-kernel small-batch path
    sp = ShardedLikelihood(lk, peer=True)
    for rep in range(5):
        th = theta * (1.0 + 0.01 * rep)
        ll_n, g_n = sl.loglike_and_grad(th)
        ll_p, g_p = sp.loglike_and_grad(th)
        ok = ok and torch.equal(ll_p, ll_n) and torch.equal(g_p, g_n)
        if rep == 0:
            ok = ok and torch.equal(sp.loglike(th), sl.loglike(th))
    lk_c = Likelihood(cfg['data'], cfg['model'], cfg['stat'], device=rank, chunk=128)  # four chunks per rank: two pushes
    sc, scn = ShardedLikelihood(lk_c, peer=True), ShardedLikelihood(lk_c)
    ll_p, g_p = sc.loglike_and_grad(theta)
    ll_n, g_n = scn.loglike_and_grad(theta)
    ok = ok and torch.equal(ll_p, ll_n) and torch.equal(g_p, g_n)
    s3p = ShardedLikelihood(lk3, peer=True)
    for rep in range(4):
        th = th3 * (1.0 + 0.002 * rep)
        a, b = s3.loglike_and_grad(th)
        c, d = s3p.loglike_and_grad(th)
        ok = ok and torch.equal(a, c) and torch.equal(b, d)
    torch.cuda.synchronize()
    for x in (sp, sc, s3p):
        x.close()
    lk_c.close()
    torch.cuda.synchronize()
    q.put((rank, bool(ok), int(ll_local.shape[0])))
    lk.close()
    lk3.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [1000, 1001])
def test_sharded_likelihood_nccl_world2(lib, n):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run under gpurun --gpus 2)')
    import socket

    import torch.multiprocessing as mp

    from elisa_b200.parallel import shard_bounds

    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    b = shard_bounds(n, 2)
    for rank, ok, rows in res:
        assert ok and rows == b[rank][1] - b[rank][0]
