"""Probe: two plans on two CUDA streams (kernels of different chunks free to overlap) vs one plan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, torch
from elisa_b200 import Likelihood, synthetic, _lib
n = 18944 * 12
cfg = synthetic.config2(n=n)
dev = 'cuda'
def mk():
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], fold='i8')
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device=dev)
    ll = torch.empty((n, 1), dtype=torch.float64, device=dev); g = torch.empty((n, 1, lk.nparam), dtype=torch.float64, device=dev)
    return lk, th, ll, g
def run(lk, th, ll, g, stream):
    _lib.check(lk._lib.elisa_b200_loglike_grad_dev(lk._h, C.c_void_p(th.data_ptr()), None, None, n, C.c_void_p(ll.data_ptr()), C.c_void_p(g.data_ptr()), C.c_void_p(stream.cuda_stream)))
a, b = mk(), mk()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for _ in range(2):
    run(*a, s1); run(*b, s2)
torch.cuda.synchronize()
t0 = time.perf_counter(); run(*a, s1); run(*b, s1); torch.cuda.synchronize(); t_seq = time.perf_counter() - t0
t0 = time.perf_counter(); run(*a, s1); run(*b, s2); torch.cuda.synchronize(); t_par = time.perf_counter() - t0
print(f'2 x {n} draws: one stream {t_seq*1e3:.1f} ms ({2*n/t_seq:.3e} evals/s), two streams {t_par*1e3:.1f} ms ({2*n/t_par:.3e} evals/s)')
