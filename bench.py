#!/usr/bin/env python
"""bench.py -- likelihood+gradient evaluations per second of the forward-folding path.

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
CutoffPL + Blackbody with background, wstat, 1024-channel x 2048-energy dense response,
1e6 posterior parameter draws per GPU and per step (synthetic data).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--draws D] [--impl reference]

One "step" = one pass of the hot path over the whole batch of draws: value AND gradient
of the per-spectrum log-likelihood for every draw.  One eval = one (draw x spectrum).
Prints ONE JSON line on rank 0 (see DESIGN.md "Measurement" for the fields).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'likelihood+grad evals/sec (samples x spectra)'
UNIT = 'evals/s'
WORKLOAD = 'cfg2: CutoffPL+Blackbody, wstat, 1024ch x 2048E dense RSP, 1e6 draws/GPU'
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the ncu --set full
# captures under profiles/ (same command, --draws 37888): forward + transposed fold, averaged
TRAFFIC = {
    # profiles/r01d_kernels.md: gemm_f64_kernel<Stat> 0.406 + 0.139 GB, <Grad> 1.944 + 0.027 GB
    'dmma': 0.5 * (0.406204 + 0.139177 + 1.943642 + 0.026987) * 1e9,
    # profiles/r01i_kernels.md: fold_i8_kernel<Store> 0.286 + 0.129 GB, <Grad> 1.793 + 0.064 GB
    # (the gradient epilogue streams dU/dtheta: 18944 x 2048 x 5 x 8 B = 1.55 GB per launch)
    'i8': 0.5 * (0.285639 + 0.128670 + 1.792700 + 0.064248) * 1e9,
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--draws', type=int, default=1_000_000, help='parameter draws per GPU and step')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--cpu-sample', type=int, default=16384, help='draws of the bounded CPU-baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--fold', default=None, choices=['i8', 'dmma'], help='response-fold engine (default: library default)')
    ap.add_argument('--slices', type=int, default=0, help='digits of the sliced-integer fold (0 = default 6)')
    return ap.parse_args()


# --------------------------------------------------------------------------- CPU baseline
def cpu_oracle_rate(cfg, n_sample, repeats=1, warmup=0):
    """Time the CPU oracle (torch CPU float64, all host threads; batched-GEMM formulation of
    the reference arithmetic) on the first n_sample draws.  Returns (evals/s, seconds, threads)."""
    import torch

    import oracle

    models = [m.compile() for m in cfg['model']]
    specs = [m.to_spec() for m in models]
    datas = [d.as_mapping() for d in cfg['data']]
    theta = torch.as_tensor(cfg['theta'][:n_sample], dtype=torch.float64)
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1, which would
    # leave the CPU arm single-threaded at N > 1, where rank 0 alone runs it)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    if torch.get_num_threads() < cores:
        torch.set_num_threads(cores)
    threads = torch.get_num_threads()
    for _ in range(warmup):
        oracle.loglike_and_grad(cfg['stat'], datas, specs, theta)
    t0 = time.perf_counter()
    for _ in range(repeats):
        oracle.loglike_and_grad(cfg['stat'], datas, specs, theta)
    dt = (time.perf_counter() - t0) / repeats
    return theta.shape[0] * len(datas) / dt, dt, threads


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  JAX is not
    installable here (no jax/jaxlib/numpyro wheels, no network), so the timed code is the
    oracle port of the reference arithmetic (kind = "port") on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from elisa_b200 import synthetic

    n = args.cpu_sample
    cfg = synthetic.config2(n=n)
    rate, dt, threads = cpu_oracle_rate(cfg, n, repeats=max(args.steps, 1), warmup=max(args.warmup, 1))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': f'{n} draws per step'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{n} of the 1e6 draws per step, torch CPU float64 oracle (JAX unavailable offline)'},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.index)],
                stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for ln in open(self.path):
                p = [x.strip() for x in ln.split(',')]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); mx.append(float(p[2])); power.append(float(p[3]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), p[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                   'power_w_max': float(max(power)), 'samples': len(sm)}
        return out


# --------------------------------------------------------------------------- B200 arm
def measure_dgemm_peak(torch):
    """cuBLAS DGEMM 8192^3, best of 6 (burst): the FP64 tensor-pipe denominator.
    MEASURED_PEAKS.json carries no FP64 entry; same recipe as its bf16 figure."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    torch.cuda.empty_cache()
    return 2.0 * n**3 / best / 1e9


def run_b200(args):
    import torch
    import torch.distributed as dist

    from elisa_b200 import Likelihood, synthetic

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: elisa_b200 has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))

    n = args.draws
    cfg = synthetic.config2(n=n)
    if world > 1:  # every rank evaluates its own block of draws (weak scaling over the N axis)
        cfg['theta'] = synthetic._theta_batch(cfg['truth'], n, 5 + 1000 * rank)
    d = cfg['data'][0]
    E, C, S = d.n_energies, d.n_channels, len(cfg['data'])
    lk = Likelihood(cfg['data'], cfg['model'], cfg['stat'], device=local, fold=args.fold, slices=args.slices)
    ks = lk.fold_slices[0]
    P = lk.nparam
    theta = torch.as_tensor(cfg['theta'], dtype=torch.float64, device=f'cuda:{local}')
    ll = torch.empty((n, S), dtype=torch.float64, device=theta.device)
    grad = torch.empty((n, S, P), dtype=torch.float64, device=theta.device)

    import ctypes as Ct

    from elisa_b200 import _lib

    stream = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        _lib.check(lk._lib.elisa_b200_loglike_grad_dev(lk._h, Ct.c_void_p(theta.data_ptr()), None, None, n,
                                                       Ct.c_void_p(ll.data_ptr()), Ct.c_void_p(grad.data_ptr()), stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = measure_dgemm_peak(torch) if rank == 0 else None

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lk.set_profiling(True)
    launches0 = lk.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = lk.launch_count - launches0
    prof = lk.get_profile()
    lk.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=theta.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * S * args.steps / (ms_max * 1e-3)

    # ---- end to end: host buffers through the C ABI, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        th_host = torch.empty((n, P), dtype=torch.float64).pin_memory()
        th_host.copy_(theta.cpu())
        ll_host = torch.empty((n, S), dtype=torch.float64).pin_memory()
        g_host = torch.empty((n, S, P), dtype=torch.float64).pin_memory()

        def step_host():
            _lib.check(lk._lib.elisa_b200_loglike_grad_host(
                lk._h, Ct.c_void_p(th_host.data_ptr()), None, None, n, Ct.c_void_p(ll_host.data_ptr()),
                Ct.c_void_p(g_host.data_ptr())))

        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=theta.device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {'value': world * n * S * args.steps / float(tt.item()), 'unit': UNIT,
               'h2d_bytes_per_step': int(th_host.numel() * 8), 'd2h_bytes_per_step': int((ll_host.numel() + g_host.numel()) * 8)}
        assert torch.equal(ll_host, ll.cpu()), 'host and device paths disagree'

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the two response folds (forward + transposed)
    fold_ms, fold_n = prof['fold']
    tfold_ms, tfold_n = prof['tfold']
    gemm_ms = fold_ms + tfold_ms
    gemm_launches = fold_n + tfold_n
    flops_total = 4.0 * E * C * n * S * args.steps  # 2 E C per eval and fold, both folds
    achieved = flops_total / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    roofline = {
        'bound': 'tensor',
        'kernel': ('fold_i8_kernel (forward + transposed fold; int8 tcgen05.mma, %d digits per operand, %d digit '
                   'products per FP64 product, int32 accumulation in TMEM)' % (ks, ks * (ks + 1) // 2)) if ks else
                  'gemm_f64_kernel (forward + transposed fold, FP64 DMMA)',
        'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
        'frac': achieved / fp64_peak if achieved and fp64_peak else None,
        'peak_source': 'cuBLAS DGEMM 8192^3 best-of-6 measured in this run (MEASURED_PEAKS.json has no FP64 entry); '
                       'achieved = ALGORITHMIC FP64 flops (4 E C per eval) / device time of the fold kernels',
        'traffic': TRAFFIC.get('i8' if ks else 'dmma'),
        'avg_launch_ms': gemm_ms / gemm_launches if gemm_launches else None,
        'flop_per_launch': flops_total / gemm_launches if gemm_launches else None,
        'share_of_step': gemm_ms / ms if ms > 0 else None,
        'classes_ms_per_step': {k: v[0] / args.steps for k, v in prof.items()},
        'whole_path_frac': (4.0 * E * C * n * S * args.steps / (ms * 1e-3) / 1e12) / fp64_peak if fp64_peak else None,
    }
    if ks:
        # the integer engine beats the FP64 pipe roofline by construction: also report the int8
        # work it executes against the int8 tensor peak (nominal dense 4500 TOP/s; bf16 burst in
        # MEASURED_PEAKS.json x 2 as the measured stand-in)
        int8_ops = flops_total * (ks * (ks + 1) // 2)
        int8_tops = int8_ops / (gemm_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        int8_peak = 2.0 * peaks.get('bf16_tflops', 1590.0)
        roofline['int8'] = {'executed_tops': int8_tops, 'peak_tops': int8_peak,
                            'peak_source': '2 x measured bf16 burst of MEASURED_PEAKS.json (of measured)' if peaks else
                                           '2 x 1590 fallback bf16 (of fallback)',
                            'frac': int8_tops / int8_peak, 'nominal_peak_tops': 4500.0}
        roofline['note'] = ('frac > 1: the fold is FP64-accurate (parity 1e-10) but runs on the int8 tensor pipe, '
                            'so the FP64 DMMA roofline does not bound it; int8.frac is the fraction of the pipe it '
                            'does run on')

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        ns = min(args.cpu_sample, n)
        rate, dt, threads = cpu_oracle_rate(cfg, ns, repeats=2, warmup=1)
        cpu = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': f'first {ns} of the {n} draws, 2 timed repeats ({dt:.2f} s each), torch CPU float64 oracle; '
                         'the reference JAX path cannot be installed offline'}

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'draws_per_gpu': n, 'spectra': S, 'n_energies': E, 'n_channels': C,
                   'n_params': P, 'chunk': lk.chunk, 'parallelism': f'dp{world} over draws',
                   'fold_engine': f'int8 tcgen05 sliced, {ks} digits' if ks else 'FP64 DMMA',
                   'l2': 'inputs larger than L2 (per-chunk workspace 2 GB, theta 40 MB)'},
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner
    # under NCCL_DEBUG=VERSION) are sent to stderr, and print() writes to the saved descriptor
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, 'w', buffering=1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
