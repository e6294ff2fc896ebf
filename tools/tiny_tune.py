"""Launch shape of the one-kernel small-batch path (cfg3): per-call time through the Python API and the kernel's
own device time (the plan's profiling spans), for ELISA_B200_TINY_TUNE = threads,split,unroll given in argv[1]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] != 'default':
    os.environ['ELISA_B200_TINY_TUNE'] = sys.argv[1]
import torch
from elisa_b200 import Likelihood, synthetic
out = []
for n in (64, 8):
    cfg = synthetic.config3(n=n)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], guard='lazy')
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    for _ in range(20):
        lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(500):
        ll, g = lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    call_us = (time.perf_counter() - t0) / 500 * 1e6
    lk.set_profiling(True)
    for _ in range(100):
        lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    prof = lk.get_profile()
    lk.set_profiling(False)
    out.append(f"n={n}: {call_us:.1f} us/call, kernel {prof['tiny'][0] / max(prof['tiny'][1], 1) * 1e3:.2f} us")
    lk.close()
print(os.environ.get('ELISA_B200_TINY_TUNE', 'default (256 threads, 4 channels per thread for <= 4 parameters else 2, unroll 4)'), '|', ' | '.join(out), flush=True)
