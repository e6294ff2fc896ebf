"""A few direct (not graph-replayed) small-batch calls of cfg 3 for `ncu -k regex:tiny_eval`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['ELISA_B200_GRAPHS'] = '0'
import torch
from elisa_b200 import Likelihood, synthetic
cfg = synthetic.config3(n=64)
lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
for _ in range(4):
    ll, g = lk.loglike_and_grad(th)
torch.cuda.synchronize()
print(float(ll.sum()), lk.small_batch_kernel)
lk.close()
