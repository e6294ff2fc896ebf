#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tiny.py tests/test_gpu_autograd.py tests/test_gpu_free_shift.py -m gpu -q -x 2>&1 | tail -5
for rep in 1 2; do
for t in default 256,4,2 256,4,8 256,2,8 512,4,4 128,4,4 256,8,2; do
  timeout 120 python tools/tiny_tune.py $t 2>&1 | tail -1
done; done | tee $OUT/r02ae_tiny_tune.txt
ELISA_B200_TINY_CLOCKS=1 python - <<'PY' 2>&1 | tail -4 | tee gpurun_out/r02ae_tiny_clocks.txt
import torch
from elisa_b200 import Likelihood, synthetic
for n in (64, 8):
    cfg = synthetic.config3(n=n)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'])
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    for _ in range(3):
        lk.loglike_and_grad(th)
    torch.cuda.synchronize()
    lk.close()
PY
