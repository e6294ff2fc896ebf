"""Power, clock and time of each kernel class of the bench step in ISOLATION, and of two chunks
overlapped on two streams (the experiment DESIGN.md section 2 refers to).

`elisa_b200_plan_set_debug` skips kernel classes (results are garbage, timing is not), so each
class can be looped on its own for ~2 s while nvidia-smi samples power and SM clock.  Answers:
which kernels sit at the 1 kW cap on their own, what the FP64-pipe kernels draw, and whether
running an FP64-pipe kernel next to a fold on a second stream gains anything.

usage: python tools/power_probe.py [draws]   -> markdown table on stdout
"""
import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from elisa_b200 import Likelihood, synthetic

Q = 'clocks.sm,power.draw,clocks_event_reasons.sw_power_cap'


class Sampler:
    def __enter__(self):
        self.f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
        self.p = subprocess.Popen(['nvidia-smi', f'--query-gpu={Q}', '--format=csv,noheader,nounits', '-lms', '50', '-i', '0'],
                                  stdout=self.f, stderr=subprocess.DEVNULL)
        return self

    def __exit__(self, *a):
        self.p.terminate(); self.p.wait(timeout=5)
        sm, pw, cap = [], [], 0
        for ln in open(self.f.name):
            x = [t.strip() for t in ln.split(',')]
            try:
                sm.append(float(x[0])); pw.append(float(x[1])); cap += x[2].lower().startswith('active')
            except (ValueError, IndexError):
                pass
        os.unlink(self.f.name)
        k = len(sm) // 4  # drop the ramp
        self.sm = float(np.median(sm[k:])) if sm else float('nan')
        self.pw = float(np.median(pw[k:])) if pw else float('nan')
        self.cap = cap / max(len(sm), 1)


def loop(fn, seconds):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        fn(); n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 18944 * 8
    cfg = synthetic.config2(n=n)
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], guard='off')
    th = torch.as_tensor(cfg['theta'], dtype=torch.float64, device='cuda')
    chunks = (n + lk.chunk - 1) // lk.chunk
    lk.loglike_and_grad(th); torch.cuda.synchronize()
    print(f'| kernels run (n = {n}, {chunks} chunks per call) | ms per chunk | median W | median SM MHz | share of samples at sw_power_cap |')
    print('|---|---|---|---|---|')
    ALL = 31
    cases = [('all five classes (the bench step)', 0), ('unfold_grad only', ALL & ~1), ('forward fold only', ALL & ~2),
             ('statistic + W digits only', ALL & ~4), ('transposed fold + gradient epilogue only', ALL & ~8),
             ('unfold_grad only, dU/dtheta stores dropped', (ALL & ~1) | 32), ('both folds only', ALL & ~(2 | 8)), ('unfold + statistic only (FP64-pipe kernels)', ALL & ~(1 | 4))]
    for label, skip in cases:
        lk.set_debug(skip)
        with Sampler() as s:
            ms = loop(lambda: lk.loglike_and_grad(th), 2.5)
        print(f'| {label} | {ms / chunks:.3f} | {s.pw:.0f} | {s.sm:.0f} | {s.cap:.2f} |', flush=True)
    lk.set_debug(0)
    # two plans, two streams: plan B's FP64-pipe kernels next to plan A's folds
    lk2 = Likelihood(cfg['data'], cfg['model'], cfg['stat'], guard='off')
    lk2.loglike_and_grad(th); torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def both(skip_a, skip_b):
        lk.set_debug(skip_a); lk2.set_debug(skip_b)
        def fn():
            with torch.cuda.stream(s1):
                lk.loglike_and_grad(th)
            with torch.cuda.stream(s2):
                lk2.loglike_and_grad(th)
        return fn
    for label, a, b in (('two full steps on two streams (per chunk PAIR)', 0, 0),
                        ('stream 1: both folds, stream 2: unfold + statistic (per chunk pair)', ALL & ~(2 | 8), ALL & ~(1 | 4))):
        with Sampler() as s:
            ms = loop(both(a, b), 2.5)
        print(f'| {label} | {ms / chunks:.3f} | {s.pw:.0f} | {s.sm:.0f} | {s.cap:.2f} |', flush=True)
    lk.close(); lk2.close()


if __name__ == '__main__':
    main()
