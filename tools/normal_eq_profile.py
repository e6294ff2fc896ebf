"""Per-kernel-class device time of one normal-equations call (the inner step of the batched
re-fit) on the bench shape.  usage: python tools/normal_eq_profile.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from elisa_b200 import Likelihood, synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
cfg = synthetic.config2(n=n)
for bal in (True, False):
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], balance=bal)
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    for _ in range(2):
        lk.normal_equations(th)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        lk.normal_equations(th)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    lk.set_profiling(True)
    for _ in range(5):
        lk.normal_equations(th)
    torch.cuda.synchronize()
    print(f'balance={bal}: {ms:.3f} ms per call of {n} rows;', {k: round(v[0] / 5, 3) for k, v in lk.get_profile().items()})
    lk.close()
