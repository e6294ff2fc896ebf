"""cfg4 (band-sparse 4096 x 4096, chi2, per-replica counts) with the default engine: two calls of one
chunk each, for `ncu -s <kernels of call 1> -c <kernels of call 2>`.  Prints the kernels per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from elisa_b200 import Likelihood, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
cfg = synthetic.config4(n=n)
lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], **({'fold': sys.argv[2]} if len(sys.argv) > 2 else {}))
th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
on = torch.as_tensor(cfg['counts_on'], dtype=torch.float64, device='cuda')
c0 = lk.launch_count
lk.loglike_and_grad(th, counts_on=on, guard='off'); torch.cuda.synchronize()
c1 = lk.launch_count
lk.loglike_and_grad(th, counts_on=on, guard='off'); torch.cuda.synchronize()
print('engine slices', lk.fold_slices, 'launches first call', c1 - c0, 'second call', lk.launch_count - c1)
