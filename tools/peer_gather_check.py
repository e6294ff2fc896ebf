"""torchrun --nproc-per-node N tools/peer_gather_check.py: the fused gather of the small-batch kernel (cfg 3, 64 chains
over N GPUs) against the NCCL gather of the same plan -- bit-identical -- and the time per gathered leapfrog of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from elisa_b200 import Likelihood, synthetic
from elisa_b200.parallel import ShardedLikelihood

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
cfg = synthetic.config3(n=64)
lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], device=local)
th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device=f'cuda:{local}')
res = {}
for name, sh in (('nccl', ShardedLikelihood(lk)), ('peer', ShardedLikelihood(lk, peer=True))):
    out = []
    for rep in range(20):
        ll, g = sh.loglike_and_grad(th * (1.0 + 1e-3 * rep))
        out.append((ll.clone(), g.clone()))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(300):
        sh.loglike_and_grad(th)
    torch.cuda.synchronize()
    us = torch.tensor([(time.perf_counter() - t0) / 300 * 1e6], dtype=torch.float64, device=f'cuda:{local}')
    dist.all_reduce(us, op=dist.ReduceOp.MAX)
    res[name] = (out, float(us.item()))
    sh.close()
same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(res['nccl'][0], res['peer'][0]))
flag = torch.tensor([int(same)], device=f'cuda:{local}')
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f'world {world}: cfg3 64 chains, gathered leapfrog: NCCL {res["nccl"][1]:.1f} us, fused peer gather {res["peer"][1]:.1f} us, '
          f'bit-identical on every rank: {bool(flag.item())}', flush=True)
lk.close()
dist.destroy_process_group()
