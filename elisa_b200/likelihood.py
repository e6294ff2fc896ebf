"""Host-side mirror of the reference's likelihood seam for the hot path.

Reference: ``likelihood_wrapper[stat](data, model.eval)`` closures
(infer/helper.py:313-323; infer/likelihood.py:184-450) and the reductions
``get_loglike`` / ``deviance`` / ``deviance_total`` (infer/helper.py:494-512, 571-586),
differentiated by ``jax.grad`` (infer/fit.py:487) and batched by ``jax.vmap`` over
parameter vectors.  Here one :class:`Likelihood` owns a CUDA plan (replicated constants
+ workspace) and evaluates a whole batch per call through the C ABI
(include/elisa_b200.h).  torch supplies device memory and streams only.
"""

from __future__ import annotations

import ctypes as C
import weakref
from typing import Sequence

import numpy as np

from . import _lib
from .data import FixedData
from .models import CompiledModel, Model

import contextlib

_NULL_CTX = contextlib.nullcontext()

# statistic -> needs (infer/likelihood.py:184-450; validity rules of infer/fit.py:_parse_input)
_NEEDS_BACK = ('pstat', 'pgstat', 'wstat')


def _ptr(a, ctype=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


def _dev_ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _raw_stream(torch, device):
    """cudaStream_t of torch's current stream on `device` as an integer (no Stream object)."""
    try:
        return torch._C._cuda_getCurrentRawStream(device)
    except AttributeError:  # an older / newer torch without the private accessor
        return torch.cuda.current_stream(device).cuda_stream


_SMALL_BATCH_ROWS = 1024  # TINY_MAX_N of the library: calls up to this size take the one-kernel path where a spectrum qualifies


class GuardTripped(_lib.ElisaB200Error):
    """guard='lazy': the accuracy guard of the integer fold flagged rows of earlier calls."""


class Likelihood:
    """Batched log-likelihood (+ gradient) of S datasets for N parameter vectors.

    Parameters
    ----------
    data : sequence of FixedData
    model : CompiledModel / Model, or one per dataset.  Free parameters shared between
        models (the same UniformParameter instance) are one column of ``theta``.
    stat : sequence of {'chi2', 'cstat', 'pstat', 'pgstat', 'wstat'}, one per dataset
    device : CUDA device ordinal (default: torch's current device)
    chunk : parameter vectors per internal pass (0 = library default)
    specialise : True = require the NVRTC-specialised component kernel, False = use the
        ahead-of-time interpreter kernel, None = specialise when NVRTC is available
    fold : response-fold engine.  None = library default (the sliced-integer tcgen05 fold,
        FP64 DMMA for spectra it cannot take), 'i8' = require the sliced-integer fold,
        'dmma' = native FP64 DMMA
    slices : digits per operand of the sliced-integer fold (4..7, 0 = default 6)
    slices_bwd : digits of the transposed fold (gradient side; 0 = library default)
    balance : balance the integer fold's contraction indices around the first batch evaluated
        (power-of-two column envelopes); False = plain normwise accuracy
    guard : how evaluations consult the accuracy guard of the integer fold (``fold=None`` only).
        'sync' (default): every call goes through ``elisa_b200_loglike_grad_guarded_dev``, which
        synchronises the stream, and rows the guard flags are re-evaluated with the FP64 DMMA
        fold before the call returns.  'lazy': calls only enqueue (no host synchronisation, the
        small-batch CUDA graph is replayed as is); the guard words are snapshotted into pinned
        memory every ``guard_every`` calls and looked at on the NEXT call -- a flagged snapshot
        raises :class:`GuardTripped` there (callers of a leapfrog loop catch it and repeat the
        trajectory with guard='sync').  'off': never looked at.
    small_batch : calls of <= 1024 parameter vectors on a small spectrum (E * C * (1 + free parameters)
        <= 2^18: NUTS chains on GBM-like detectors) are ONE kernel per spectrum in plain FP64 instead of
        the chunk pipeline (``small_batch_kernel`` tells which spectra qualify); False, or an explicit
        ``fold``, keeps every call on the pipeline.

    ``params_name`` lists the columns of ``theta`` (constrained space; priors and
    transforms stay with the caller, infer/helper.py:409-426).
    """

    def __init__(self, data: Sequence[FixedData], model, stat: Sequence[str], device: int | None = None, chunk: int = 0,
                 specialise: bool | None = None, fold: str | None = None, slices: int = 0, balance: bool = True,
                 slices_bwd: int = 0, guard: str = 'sync', guard_every: int = 16, small_batch: bool = True):
        import torch

        if not torch.cuda.is_available():
            raise _lib.ElisaB200Error('elisa_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        lib = _lib.load()
        data = list(data)
        if isinstance(model, (CompiledModel, Model)):
            model = [model] * len(data)
        model = [m if isinstance(m, CompiledModel) else m.compile() for m in model]
        stat = [stat] * len(data) if isinstance(stat, str) else list(stat)
        if not (len(data) == len(model) == len(stat)) or not data:
            raise ValueError('data, model and stat must have the same non-zero length')
        for d, s in zip(data, stat):
            if s not in _lib.STAT:
                raise ValueError(f'unknown statistic {s!r}')
            if s in _NEEDS_BACK and not d.has_back:
                raise ValueError(f'{s} requires background data ({d.name})')  # likelihood.py:280,331,397
            if s == 'chi2' and (d.net_counts is None or d.net_errors is None):
                raise ValueError(f'chi2 requires net_counts / net_errors ({d.name})')
            if s == 'pgstat' and d.back_errors is None:
                raise ValueError(f'pgstat requires back_errors ({d.name})')
        for m in model:
            if m.type != 'add':
                raise TypeError('the model folded through a response must be additive')
        self.data, self.model, self.stat = data, model, stat
        self.names = [str(d.name) for d in data]
        # global free-parameter list (columns of theta), first appearance order
        self.free, self.params_name = [], []
        for m in model:
            for p, n in zip(m.free, m.params_name):
                if all(p is not q for q in self.free):
                    self.free.append(p)
                    self.params_name.append(n)
        self.nparam = len(self.free)
        if self.nparam > _lib.MAX_PARAMS:
            raise NotImplementedError(f'more than {_lib.MAX_PARAMS} free parameters')
        self.params_default = np.array([p.default for p in self.free], dtype=np.float64)
        self._P = max(self.nparam, 1)
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._torch = torch
        self._lib = lib

        keep = []
        specs = (_lib.SpectrumDesc * len(data))()
        for i, (d, m, s) in enumerate(zip(data, model, stat)):
            sd = specs[i]
            sd.n_energies, sd.n_channels = d.n_energies, d.n_channels
            sd.photon_egrid = _ptr(d.photon_egrid)
            if d.response_sparse:
                ip, ix, v = d.sparse_matrix
                sd.csr_indptr, sd.csr_indices, sd.csr_values = _ptr(ip, C.c_int32), _ptr(ix, C.c_int32), _ptr(v)
            else:
                sd.response_dense = _ptr(d.response_matrix)
            sd.area_scale = _ptr(d.area_scale)
            sd.exposure = d.spec_exposure
            sd.stat = _lib.STAT[s]
            if s == 'chi2':
                sd.on_counts, sd.on_error = _ptr(d.net_counts), _ptr(d.net_errors)
            else:
                sd.on_counts = _ptr(d.spec_counts)
            if s in _NEEDS_BACK:
                sd.off_counts, sd.back_ratio = _ptr(d.back_counts), _ptr(d.back_ratio)
            if s == 'pgstat':
                sd.off_error = _ptr(d.back_errors)
            sd.model, k = m.to_desc(self.free)
            keep.append(k)
        flags = 0 if specialise is None else (_lib.FLAG_REQUIRE_JIT if specialise else _lib.FLAG_NO_JIT)
        if fold not in (None, 'i8', 'dmma'):
            raise ValueError("fold must be None, 'i8' or 'dmma'")
        if slices and not 4 <= int(slices) <= 7:
            raise ValueError('slices must be 0 or 4..7')
        flags |= {None: 0, 'i8': _lib.FLAG_FOLD_I8, 'dmma': _lib.FLAG_FOLD_DMMA}[fold]
        flags |= int(slices) << _lib.FLAG_FOLD_SLICES_SHIFT
        if slices_bwd and not 4 <= int(slices_bwd) <= 7:
            raise ValueError('slices_bwd must be 0 or 4..7')
        flags |= int(slices_bwd) << _lib.FLAG_FOLD_SLICES_BWD_SHIFT
        if guard not in ('sync', 'lazy', 'off'):
            raise ValueError("guard must be 'sync', 'lazy' or 'off'")
        if not balance:
            flags |= _lib.FLAG_NO_BALANCE
        if not small_batch:  # keep calls of <= 1024 rows on the chunk pipeline (see small_batch_kernel)
            flags |= _lib.FLAG_NO_SMALL_BATCH_KERNEL
        desc = _lib.PlanDesc(_lib.ABI_VERSION, self.device, self._P, len(data), specs, int(chunk), flags)
        handle = C.c_void_p()
        _lib.check(lib.elisa_b200_plan_create(C.byref(desc), C.byref(handle)))
        self._h = handle
        # the plan owns gigabytes of device memory: release it when the object is collected, too
        self._finalizer = weakref.finalize(self, lib.elisa_b200_plan_destroy, C.c_void_p(handle.value))
        for i, d in enumerate(data):
            _lib.check(lib.elisa_b200_plan_set_channel_width(self._h, i, _ptr(d.channel_width)))
        self.n_channels = [d.n_channels for d in data]
        self.sum_channels = int(sum(self.n_channels))
        # fold=None: the integer fold runs guarded (see the `guard` parameter)
        self._auto_guard = fold is None and any(self.fold_slices)
        self.guard_mode = guard if self._auto_guard else 'off'
        self.guard_every = max(1, int(guard_every))
        self._calls_since_snapshot = 0
        self._snapshot_pending = False
        self.guard_fallbacks = 0   # calls in which the DMMA fold answered for some rows
        self.fallback_rows = 0     # ... and how many rows in total
        self.rebalances = 0        # envelope renewals the guard triggered
        self._dev_str = f'cuda:{self.device}'
        self._fb = C.c_int64(0)
        self._all_small_batch = all(self.small_batch_kernel)
        self._S = len(self.data)
        self._f_grad, self._f_val = self._lib.elisa_b200_loglike_grad_dev, self._lib.elisa_b200_loglike_dev

    # -- lifetime ------------------------------------------------------------------
    def close(self):
        """Destroy the plan now (idempotent; also runs when the object is garbage-collected)."""
        if getattr(self, '_h', None) is not None and self._h.value:
            self._finalizer()  # calls elisa_b200_plan_destroy exactly once
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def fold_guard(self, reset: bool = True):
        """(rows flagged, largest relative error estimate) of the integer fold's accuracy guard
        since the last reset (elisa_b200_plan_fold_guard); synchronises the device."""
        self._torch.cuda.synchronize(self.device)
        n, w = C.c_int64(0), C.c_double(0.0)
        _lib.check(self._lib.elisa_b200_plan_fold_guard(self._h, C.byref(n), C.byref(w), 1 if reset else 0))
        return int(n.value), float(w.value)

    def guard_poll(self, wait: bool = False):
        """Lazy mode: (ready, rows flagged, largest estimate) of the last asynchronous snapshot of
        the guard words (``elisa_b200_plan_guard_poll``); never synchronises unless ``wait``."""
        n, w = C.c_int64(0), C.c_double(0.0)
        ready = self._lib.elisa_b200_plan_guard_poll(self._h, 1 if wait else 0, C.byref(n), C.byref(w))
        if ready < 0:
            _lib.check(ready)
        return bool(ready), int(n.value), float(w.value)

    def _lazy_guard(self, stream):
        """Called after every lazily guarded evaluation: look at the previous snapshot (if it
        has landed), take a new one every ``guard_every`` calls."""
        if self._snapshot_pending:
            ready, flagged, worst = self.guard_poll(wait=False)
            if ready:
                self._snapshot_pending = False
                if flagged:
                    self.fold_guard(reset=True)
                    raise GuardTripped(f'the accuracy guard of the integer fold flagged {flagged} rows '
                                       f'(estimate {worst:.1e}) in the last {self.guard_every} lazily guarded calls: '
                                       "repeat them with guard='sync' or fold='dmma'")
        self._calls_since_snapshot += 1
        if self._calls_since_snapshot >= self.guard_every and not self._snapshot_pending:
            _lib.check(self._lib.elisa_b200_plan_guard_snapshot(self._h, stream))
            self._snapshot_pending = True
            self._calls_since_snapshot = 0

    def _sync_counters(self):
        out = (C.c_int64 * 3)()
        _lib.check(self._lib.elisa_b200_plan_guard_stats(self._h, out))
        self.guard_fallbacks, self.fallback_rows, self.rebalances = int(out[0]), int(out[1]), int(out[2])

    def rebalance(self):
        """Mark the balancing envelopes of the integer fold stale: the next evaluation derives
        them from its own first rows (``elisa_b200_plan_rebalance``)."""
        _lib.check(self._lib.elisa_b200_plan_rebalance(self._h))

    @property
    def specialised(self) -> list[bool]:
        """Per dataset: is the component kernel the runtime-specialised one?"""
        return [bool(self._lib.elisa_b200_plan_is_specialised(self._h, k)) for k in range(len(self.data))]

    @property
    def fold_slices(self) -> list[int]:
        """Per dataset: digits of the sliced-integer fold in use, 0 = FP64 DMMA fold."""
        return [int(self._lib.elisa_b200_plan_fold_slices(self._h, k)) for k in range(len(self.data))]

    @property
    def fold_slices_bwd(self) -> list[int]:
        """Per dataset: digits of the TRANSPOSED fold (gradient side), 0 = FP64 DMMA fold."""
        return [int(self._lib.elisa_b200_plan_fold_slices_bwd(self._h, k)) for k in range(len(self.data))]

    @property
    def small_batch_kernel(self) -> list[bool]:
        """Per dataset: calls of <= 1024 parameter vectors evaluate it in ONE kernel (small spectra, the
        NUTS shape) instead of the chunk pipeline."""
        return [self._lib.elisa_b200_plan_small_batch_kernel(self._h, k) == 1 for k in range(len(self.data))]

    def fold_products(self, k: int = 0):
        """(forward, transposed): int8 digit products the integer fold executes per FP64
        multiply-add of dataset k's dense contraction (``elisa_b200_plan_fold_products``); 21 for 6
        digits on a response without structure, fewer where digit planes are zero.  Synchronises."""
        out = (C.c_double * 2)()
        _lib.check(self._lib.elisa_b200_plan_fold_products(self._h, int(k), out))
        return float(out[0]), float(out[1])

    def set_debug(self, skip_mask: int):
        """Timing experiments only (``elisa_b200_plan_set_debug``)."""
        _lib.check(self._lib.elisa_b200_plan_set_debug(self._h, int(skip_mask)))

    @property
    def chunk(self) -> int:
        return int(self._lib.elisa_b200_plan_chunk(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(self._lib.elisa_b200_plan_workspace_bytes(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.elisa_b200_plan_launch_count(self._h))

    PROFILE_CLASSES = ('unfold', 'fold', 'tfold', 'reduce', 'slice', 'stat', 'contract', 'tiny', 'gather')

    def set_profiling(self, enable: bool):
        _lib.check(self._lib.elisa_b200_plan_set_profiling(self._h, int(bool(enable))))

    def get_profile(self) -> dict:
        """{class: (milliseconds, launches)} since set_profiling(True) (CUDA events)."""
        ms = (C.c_double * len(self.PROFILE_CLASSES))()
        n = (C.c_int64 * len(self.PROFILE_CLASSES))()
        _lib.check(self._lib.elisa_b200_plan_get_profile(self._h, ms, n))
        return {k: (ms[i], n[i]) for i, k in enumerate(self.PROFILE_CLASSES)}

    # -- argument plumbing -----------------------------------------------------------
    def _theta(self, theta):
        torch = self._torch
        t = torch.as_tensor(theta, dtype=torch.float64)
        if t.dim() == 1:
            t = t.unsqueeze(0)
        if t.dim() != 2 or t.shape[1] != self.nparam:
            raise ValueError(f'expected theta of shape (N, {self.nparam}), got {tuple(t.shape)}')
        if self.nparam == 0:
            t = torch.zeros((t.shape[0], 1), dtype=torch.float64)
        return t.to(f'cuda:{self.device}').contiguous()

    def _counts(self, c, n):
        if c is None:
            return None
        torch = self._torch
        t = torch.as_tensor(c, dtype=torch.float64)
        if tuple(t.shape) != (n, self.sum_channels):
            raise ValueError(f'expected per-replica counts of shape ({n}, {self.sum_channels}), got {tuple(t.shape)}')
        return t.to(f'cuda:{self.device}').contiguous()

    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.device).cuda_stream)

    def _device_ctx(self):
        """torch.cuda.device(self.device), or nothing when it is the current device already (the
        context manager costs ~10 us per call: a tenth of a replayed small-batch graph)."""
        if self._torch.cuda.current_device() == self.device:
            return _NULL_CTX
        return self._torch.cuda.device(self.device)

    # -- device API (torch tensors in, torch tensors out) ----------------------------
    def _evaluate_small(self, t, want_grad: bool, out=None):
        """The leapfrog path: a small batch already on the device, every spectrum on the one-kernel path (plain FP64, no
        guard to consult).  The kernel takes ~12 us, so what this method spends is the rate of a sampler that calls it
        back to back: no tensor conversion, no device context, no stream object, plain integers through ctypes."""
        torch = self._torch
        n = t.shape[0]
        out, grad = out if out is not None else (torch.empty((n, self._S), dtype=torch.float64, device=t.device), None)
        stream = _raw_stream(torch, self.device)
        if want_grad:
            if grad is None:
                grad = torch.empty((n, self._S, self._P), dtype=torch.float64, device=t.device)
            rc = self._f_grad(self._h, t.data_ptr(), None, None, n, out.data_ptr(), grad.data_ptr(), stream)
        else:
            grad = None
            rc = self._f_val(self._h, t.data_ptr(), None, None, n, out.data_ptr(), stream)
        if rc:
            _lib.check(rc)
        return out, grad

    def _check_out(self, out, n, want_grad):
        """Caller-provided result arrays (ll (n, S), grad (n, S, P_plan) or None): contiguous float64 on this device."""
        ll, g = out
        shapes = [(ll, (n, self._S))] + ([(g, (n, self._S, self._P))] if want_grad else [])
        for t, shape in shapes:
            if (t is None or tuple(t.shape) != shape or t.dtype is not self._torch.float64 or not t.is_cuda
                    or t.device.index != self.device or not t.is_contiguous()):
                raise ValueError(f'out: expected a contiguous float64 tensor of shape {shape} on cuda:{self.device}')
        return ll, (g if want_grad else None)

    def _evaluate(self, theta, counts_on, counts_off, want_grad: bool, guard, out=None):
        torch = self._torch
        if (self._all_small_batch and counts_on is None and counts_off is None and type(theta) is torch.Tensor
                and theta.is_cuda and theta.dtype is torch.float64 and theta.dim() == 2 and theta.shape[1] == self.nparam
                and 0 < theta.shape[0] <= _SMALL_BATCH_ROWS and self.nparam > 0 and theta.is_contiguous()
                and theta.device.index == self.device and torch.cuda.current_device() == self.device):
            return self._evaluate_small(theta, want_grad, None if out is None else self._check_out(out, theta.shape[0], want_grad))
        mode = self.guard_mode if guard is None else (guard if self._auto_guard else 'off')
        with self._device_ctx():
            t = self._theta(theta)
            n = t.shape[0]
            on, off = self._counts(counts_on, n), self._counts(counts_off, n)
            S = len(self.data)
            if out is not None:
                out, grad = self._check_out(out, n, want_grad)
            else:
                out = torch.empty((n, S), dtype=torch.float64, device=t.device)
                grad = torch.empty((n, S, self._P), dtype=torch.float64, device=t.device) if want_grad else None
            stream = self._stream()
            if n <= _SMALL_BATCH_ROWS and self._all_small_batch:
                mode = 'off'  # the one-kernel path is plain FP64: nothing for the guard to watch, no synchronisation
            if mode == 'sync':
                _lib.check(self._lib.elisa_b200_loglike_grad_guarded_dev(
                    self._h, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n, _dev_ptr(out), _dev_ptr(grad), stream,
                    C.byref(self._fb)))
                self._sync_counters()
            else:
                if want_grad:
                    _lib.check(self._lib.elisa_b200_loglike_grad_dev(self._h, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n,
                                                                      _dev_ptr(out), _dev_ptr(grad), stream))
                else:
                    _lib.check(self._lib.elisa_b200_loglike_dev(self._h, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n,
                                                                 _dev_ptr(out), stream))
                if mode == 'lazy':
                    self._lazy_guard(stream)
        return out, grad

    def loglike(self, theta, counts_on=None, counts_off=None, guard=None, out=None):
        """(N, S) log-likelihood per dataset: get_loglike()['group'], infer/helper.py:494-512.
        ``guard`` overrides the object's guard mode for this call; ``out``: a (N, S) tensor to write into."""
        return self._evaluate(theta, counts_on, counts_off, False, guard, None if out is None else (out, None))[0]

    def loglike_and_grad(self, theta, counts_on=None, counts_off=None, guard=None, out=None):
        """((N, S), (N, S, P)): value and d loglike[n, s] / d theta[n, :] -- what
        jax.value_and_grad of each dataset's log-likelihood returns upstream.  ``out``: (ll (N, S), grad (N, S,
        plan_params)) tensors to write into (``plan_params`` = max(P, 1))."""
        out, grad = self._evaluate(theta, counts_on, counts_off, True, guard, out)
        return out, (grad if self._P == self.nparam else grad[:, :, : self.nparam])

    def normal_equations(self, theta, counts_on=None, counts_off=None, guard: bool = True):
        """(loglike (N,), grad (N, P), jtj (N, P, P)) of the deviance residuals
        r = sqrt(-2 * point log-likelihood) over all datasets (infer/helper.py:588-594):
        loglike = -|r|^2 / 2, grad = -J^T r, jtj = J^T J -- what one Levenberg-Marquardt step of
        the reference's re-fit (helper.py:601-625) needs (``elisa_b200_normal_eq_dev``).
        ``guard``: consult the integer fold's accuracy guard (the kernel accumulates the same
        first-order estimate of the log-likelihood error as the statistic kernel) and repeat a
        flagged call on the FP64 DMMA fold; synchronises."""
        torch = self._torch
        if self.nparam > 16:
            raise NotImplementedError('normal equations / re-fit with more than 16 free parameters')
        guard = bool(guard) and self._auto_guard
        with torch.cuda.device(self.device):
            t = self._theta(theta)
            n = t.shape[0]
            on, off = self._counts(counts_on, n), self._counts(counts_off, n)
            ll = torch.empty((n,), dtype=torch.float64, device=t.device)
            g = torch.empty((n, self._P), dtype=torch.float64, device=t.device)
            h = torch.empty((n, self._P, self._P), dtype=torch.float64, device=t.device)

            def run():
                _lib.check(self._lib.elisa_b200_normal_eq_dev(self._h, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n,
                                                              _dev_ptr(ll), _dev_ptr(g), _dev_ptr(h), self._stream()))
            if guard:
                self.fold_guard(reset=True)
            run()
            if guard and self.fold_guard(reset=True)[0]:
                _lib.check(self._lib.elisa_b200_plan_set_fold(self._h, 1))
                try:
                    run()
                    torch.cuda.synchronize(self.device)
                finally:
                    _lib.check(self._lib.elisa_b200_plan_set_fold(self._h, 0))
                self.normal_eq_fallbacks = getattr(self, 'normal_eq_fallbacks', 0) + 1
        return ll, g[:, : self.nparam], h[:, : self.nparam, : self.nparam]

    def jacobian(self, k: int, theta, counts_on=None, counts_off=None):
        """Forward-mode per-channel Jacobian of dataset k (``elisa_b200_jacobian_dev``): returns
        ``(x, ds, dl_dx)`` with x (N, C) the clipped model counts the statistic sees, ds (N, P, C)
        = d x / d theta and dl_dx (N, C) = d loglike_c / d x_c.  The LM residual Jacobian of the
        reference (infer/helper.py:588-594, 616-622: jacfwd of ``residual``) is
        ``-dl_dx[:, None, :] * ds / residual[:, None, :]``."""
        torch = self._torch
        if self.nparam > 16:
            raise NotImplementedError('per-channel Jacobian with more than 16 free parameters')
        with torch.cuda.device(self.device):
            t = self._theta(theta)
            n = t.shape[0]
            on, off = self._counts(counts_on, n), self._counts(counts_off, n)
            Ck = self.n_channels[k]
            x = torch.empty((n, Ck), dtype=torch.float64, device=t.device)
            ds = torch.empty((n, self._P, Ck), dtype=torch.float64, device=t.device)
            dl = torch.empty((n, Ck), dtype=torch.float64, device=t.device)
            _lib.check(self._lib.elisa_b200_jacobian_dev(self._h, k, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n, _dev_ptr(x),
                                                         _dev_ptr(ds), _dev_ptr(dl), self._stream()))
        return x, ds[:, : self.nparam], dl

    def hessian(self, theta, counts_on=None, counts_off=None, rel_step: float = 1e-5):
        """(N, P, P) Hessian of the joint log-likelihood sum_s loglike[n, s] by central differences
        of the analytic gradient (2 P gradient evaluations) -- the consumer upstream is
        ``jax.hessian(deviance_total)`` for the covariance at the optimum (infer/helper.py:533-540),
        called O(1) times per fit; agreement with autodiff is ~1e-7 relative (tested against torch
        autograd of the oracle), well inside the rtol 1e-3 the reference's own test asks
        (tests/test_results.py:70-90)."""
        torch = self._torch
        t = self._theta(theta)[:, : self.nparam]
        n, P = t.shape
        H = torch.empty((n, P, P), dtype=torch.float64, device=t.device)
        for j in range(P):
            h = rel_step * torch.maximum(t[:, j].abs(), torch.full_like(t[:, j], 1e-3))
            tp, tm = t.clone(), t.clone()
            tp[:, j] += h
            tm[:, j] -= h
            gp = self.loglike_and_grad(tp, counts_on, counts_off)[1].sum(1)
            gm = self.loglike_and_grad(tm, counts_on, counts_off)[1].sum(1)
            H[:, j, :] = (gp - gm) / (tp[:, j] - tm[:, j])[:, None]
        return 0.5 * (H + H.transpose(1, 2))

    def fit(self, theta0, counts_on=None, counts_off=None, lower=None, upper=None, max_iter: int = 200,
            xtol: float = 1e-9, gtol: float = 1e-7):
        """Batched re-fit: one Levenberg-Marquardt minimisation of the deviance residuals per
        row of ``theta0`` / per replica of ``counts_on`` (``counts_off``), all rows advancing
        together on the GPU -- the reference's ``fit_once`` / ``batch_fit`` (infer/helper.py:
        601-806: optimistix LM on ``residual`` in unconstrained space, sequential per device).

        Parameters live in the unconstrained space of the reference's interval bijection
        (theta = lower + (upper - lower) * sigmoid(u), numpyro ``biject_to(interval)``);
        ``lower`` / ``upper`` default to the parameters' prior bounds.  Returns a dict with
        'theta' (N, P), 'loglike' / 'deviance' (N,) totals, 'grad_norm' (N,) = |J^T r| in
        unconstrained space, 'n_iter' and 'valid' (finite deviance and grad_norm <= 1e-3, the
        reference's validity rule, helper.py:652-657)."""
        torch = self._torch
        dev = f'cuda:{self.device}'
        P = self.nparam
        t0 = self._theta(theta0)[:, :P]
        if counts_on is not None and t0.shape[0] == 1:
            t0 = t0.expand(self._counts(counts_on, len(counts_on)).shape[0], P).contiguous()
        n = t0.shape[0]
        on, off = self._counts(counts_on, n), self._counts(counts_off, n)
        lo = torch.as_tensor(np.array([p.min for p in self.free]) if lower is None else lower, dtype=torch.float64, device=dev)
        hi = torch.as_tensor(np.array([p.max for p in self.free]) if upper is None else upper, dtype=torch.float64, device=dev)
        boxed = torch.isfinite(lo) & torch.isfinite(hi)
        span = torch.where(boxed, hi - lo, torch.ones_like(lo))
        base = torch.where(boxed, lo, torch.zeros_like(lo))

        def to_theta(u):
            sg = torch.sigmoid(u)
            theta = torch.where(boxed, base + span * sg, u)
            d = torch.where(boxed, span * sg * (1.0 - sg), torch.ones_like(u))
            return theta, d

        z = ((t0 - base) / span).clamp(1e-12, 1.0 - 1e-12)
        u = torch.where(boxed, torch.log(z) - torch.log1p(-z), t0)

        def evaluate(u):
            theta, d = to_theta(u)
            ll, g, h = self.normal_equations(theta, on, off, guard=False)
            return ll, g * d, h * d[:, :, None] * d[:, None, :]

        ll, g, h = evaluate(u)
        lam = torch.full((n,), 1e-3, dtype=torch.float64, device=dev)
        done = ~torch.isfinite(ll)
        n_iter = torch.zeros((n,), dtype=torch.int64, device=dev)
        eye = torch.eye(P, dtype=torch.float64, device=dev)
        for _ in range(max_iter):
            diag = torch.diagonal(h, dim1=1, dim2=2)
            a = h + (lam[:, None] * (diag + 1e-12 * diag.amax(1, keepdim=True) + 1e-300))[:, :, None] * eye
            a = torch.where(torch.isfinite(a), a, eye.expand_as(a))
            delta = torch.linalg.solve(a, torch.nan_to_num(g).unsqueeze(-1)).squeeze(-1)
            delta = torch.where(done[:, None], torch.zeros_like(delta), delta)
            u_t = u + delta
            ll_t, g_t, h_t = evaluate(u_t)
            better = torch.isfinite(ll_t) & (ll_t >= ll) & ~done
            step = delta.abs().amax(1)
            small = (step <= xtol * (1.0 + u.abs().amax(1))) | (g_t.abs().amax(1) <= gtol)
            u = torch.where(better[:, None], u_t, u)
            g = torch.where(better[:, None], g_t, g)
            h = torch.where(better[:, None, None], h_t, h)
            ll = torch.where(better, ll_t, ll)
            n_iter += (~done).to(torch.int64)
            lam = torch.where(better, (lam * 0.2).clamp_min(1e-15), (lam * 4.0).clamp_max(1e15))
            done = done | (better & small) | (~better & (lam >= 1e15))
            if bool(done.all()):
                break
        theta, dth = to_theta(u)
        if self._auto_guard:
            # the LM iterations run on the unguarded integer fold (the optimum does not care about
            # 1e-10); what is REPORTED -- log-likelihood, deviance, gradient norm, validity -- comes
            # from one guarded evaluation of the final parameters (1e-10 contract, DMMA for flagged rows)
            ll_s, g_s = self.loglike_and_grad(theta, on, off, guard='sync')
            ll = ll_s.sum(-1)
            g = g_s.sum(1) * dth
        grad_norm = torch.linalg.norm(g, dim=1)
        valid = torch.isfinite(ll) & torch.isfinite(grad_norm) & (grad_norm <= 1e-3)
        return {'theta': theta, 'loglike': ll, 'deviance': -2.0 * ll, 'grad_norm': grad_norm, 'n_iter': n_iter,
                'valid': valid, 'converged': done}

    def simulate(self, theta, n: int = 1, seed: int = 0):
        """Simulated data sets drawn on the device from the model at ``theta`` (infer/helper.py:
        221-278): ``{name}_Non`` ~ Poisson(``{name}_Non_model``) -- Normal(model, net_errors) for
        chi2 -- and, for pgstat / wstat, ``{name}_Noff`` ~ Normal(``{name}_Noff_model``,
        back_errors) / Poisson(``{name}_Noff_model``).  ``theta`` is (P,) with ``n`` replicas or
        (N, P) with one replica per row (the reference's posterior-predictive case).  Returns
        ``(counts_on, counts_off)``, (rows, sum of channels) CUDA tensors in the layout the
        evaluation entry points take (``counts_off`` is None when no dataset has a sampled
        background).  The draws come from the library's own sampling kernel
        (``elisa_b200_sample_dev``: Philox4x32-10, one stream per (seed, dataset, array, replica,
        channel); Poisson by multiplication below a mean of 12 and by transformed rejection above,
        Normal by Box-Muller): the reference's NumPy stream cannot be reproduced on a GPU, so
        parity here is distributional (tests check moments, the pmf at small means and the fit
        recovering the truth)."""
        torch = self._torch
        t = self._theta(theta)
        if t.shape[0] != 1 and n != 1:
            raise ValueError('n > 1 needs a single parameter vector (helper.py:851-857)')
        rows = t.shape[0] if t.shape[0] != 1 else int(n)
        W = self.sum_channels
        any_off = any(s in ('pgstat', 'wstat') for s in self.stat)
        with self._device_ctx():
            on = torch.empty((rows, W), dtype=torch.float64, device=t.device)
            off = torch.empty((rows, W), dtype=torch.float64, device=t.device) if any_off else None
            col = 0
            for k, (d, stat) in enumerate(zip(self.data, self.stat)):
                sites = self.sites(k, t)
                name = self.names[k]
                Ck = self.n_channels[k]

                def draw(which, mean, out):
                    mean = mean.contiguous()
                    view = out[:, col: col + Ck]  # element [row, c] at out + col + row * W + c
                    _lib.check(self._lib.elisa_b200_sample_dev(
                        self._h, k, which, _dev_ptr(mean), mean.shape[0], rows, int(seed) & (2**64 - 1),
                        C.c_void_p(view.data_ptr()), W, self._stream()))

                draw(0, sites[f'{name}_Non_model'], on)
                if stat in ('pgstat', 'wstat'):
                    draw(1, sites[f'{name}_Noff_model'], off)
                elif off is not None:  # not sampled: the observed background (if any) stays in place
                    b = d.back_counts if getattr(d, 'back_counts', None) is not None else np.zeros(Ck)
                    off[:, col: col + Ck] = torch.as_tensor(b, dtype=torch.float64, device=t.device)
                col += Ck
        return on, off

    def simulate_and_fit(self, theta, n: int = 1, seed: int = 0, **fit_kw):
        """``simulate_and_fit`` of the reference (infer/helper.py:808-876), the core of ``boot()``
        and ``ppc()``: simulate ``n`` data sets at ``theta`` (or one per row of ``theta``), then
        re-fit every one of them starting from the generating parameters -- entirely on the
        device.  Returns the ``fit`` dict plus the simulated 'counts_on' / 'counts_off'."""
        on, off = self.simulate(theta, n, seed)
        t = self._theta(theta)[:, : self.nparam]
        start = t.expand(on.shape[0], -1).contiguous() if t.shape[0] == 1 else t
        out = self.fit(start, counts_on=on, counts_off=off, **fit_kw)
        out['counts_on'], out['counts_off'] = on, off
        return out

    def deviance(self, theta, **kw):
        """-2 * loglike per dataset (infer/helper.py:571-580)."""
        return -2.0 * self.loglike(theta, **kw)

    def deviance_total(self, theta, **kw):
        """(N,) joint deviance (infer/helper.py:582-586)."""
        return -2.0 * self.loglike(theta, **kw).sum(-1)

    def deviance_total_and_grad(self, theta, **kw):
        """(N,), (N, P): what Minuit's FCN / GRAD pair evaluates (infer/fit.py:474-500)."""
        ll, g = self.loglike_and_grad(theta, **kw)
        return -2.0 * ll.sum(-1), -2.0 * g.sum(1)

    def sites(self, k: int, theta, counts_on=None, counts_off=None):
        """Per-channel sites of dataset k with the reference's names
        (infer/likelihood.py:206-225, 350-385): '{name}', '{name}_Non_model',
        '{name}_Non_loglike', ['{name}_Noff_model', '{name}_Noff_loglike'], '{name}_loglike'."""
        torch = self._torch
        name, stat = self.names[k], self.stat[k]
        with torch.cuda.device(self.device):
            t = self._theta(theta)
            n = t.shape[0]
            on, off = self._counts(counts_on, n), self._counts(counts_off, n)
            mk = lambda: torch.empty((n, self.n_channels[k]), dtype=torch.float64, device=t.device)
            rate, m_on, l_on = mk(), mk(), mk()
            with_off = stat in ('pgstat', 'wstat')
            m_off, l_off = (mk(), mk()) if with_off else (None, None)
            _lib.check(self._lib.elisa_b200_sites_dev(self._h, k, _dev_ptr(t), _dev_ptr(on), _dev_ptr(off), n, _dev_ptr(rate), _dev_ptr(m_on), _dev_ptr(l_on), _dev_ptr(m_off), _dev_ptr(l_off), self._stream()))
        out = {name: rate, f'{name}_Non_model': m_on, f'{name}_Non_loglike': l_on}
        if with_off:
            out[f'{name}_Noff_model'] = m_off
            out[f'{name}_Noff_loglike'] = l_off
            out[f'{name}_loglike'] = l_on + l_off
        else:
            out[f'{name}_loglike'] = l_on
        return out

    def residual(self, theta, **kw):
        """sqrt(-2 * point log-likelihood) over all channels (infer/helper.py:588-594)."""
        torch = self._torch
        pts = [self.sites(k, theta, **kw)[f'{self.names[k]}_loglike'] for k in range(len(self.data))]
        return torch.sqrt(-2.0 * torch.cat(pts, dim=-1))

    def unfold(self, k: int, theta):
        """(N, E_k) unfolded model of dataset k (CompiledModel.eval, before the clip)."""
        torch = self._torch
        with torch.cuda.device(self.device):
            t = self._theta(theta)
            out = torch.empty((t.shape[0], self.data[k].n_energies), dtype=torch.float64, device=t.device)
            _lib.check(self._lib.elisa_b200_unfold_dev(self._h, k, _dev_ptr(t), t.shape[0], _dev_ptr(out), self._stream()))
        return out

    def csr_fold(self, k: int, unfold):
        """(N, C_k) = unfold (N, E_k) folded through the CSR-by-channel-row response of dataset k with the
        row-wise sparse kernel (``elisa_b200_csr_fold_dev``): the plain BCSR product ``resp_matrix @
        unfold`` of infer/likelihood.py:205, without area / exposure factors."""
        torch = self._torch
        with self._device_ctx():
            u = torch.as_tensor(unfold, dtype=torch.float64).to(self._dev_str).contiguous()
            out = torch.empty((u.shape[0], self.n_channels[k]), dtype=torch.float64, device=u.device)
            _lib.check(self._lib.elisa_b200_csr_fold_dev(self._h, int(k), _dev_ptr(u), u.shape[1], u.shape[0], _dev_ptr(out),
                                                         out.shape[1], self._stream()))
        return out

    # -- host API (numpy in, numpy out; copies inside) -------------------------------
    def loglike_and_grad_host(self, theta: np.ndarray, counts_on=None, counts_off=None, grad: bool = True):
        """Same through ``elisa_b200_loglike_grad_host``: host buffers, H2D / D2H inside."""
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        if theta.shape[1] != self.nparam:
            raise ValueError(f'expected theta of shape (N, {self.nparam}), got {theta.shape}')
        if self.nparam == 0:
            theta = np.zeros((theta.shape[0], 1))
        n = theta.shape[0]
        on = None if counts_on is None else np.ascontiguousarray(counts_on, dtype=np.float64)
        off = None if counts_off is None else np.ascontiguousarray(counts_off, dtype=np.float64)
        for c in (on, off):
            if c is not None and c.shape != (n, self.sum_channels):
                raise ValueError(f'expected per-replica counts of shape ({n}, {self.sum_channels})')
        ll = np.empty((n, len(self.data)))
        g = np.empty((n, len(self.data), self._P)) if grad else None
        _lib.check(self._lib.elisa_b200_loglike_grad_host(self._h, _ptr(theta), _ptr(on), _ptr(off), n, _ptr(ll), _ptr(g)))
        self._sync_counters()
        return (ll, g[:, :, : self.nparam]) if grad else ll


_unfold_cache: dict = {}


def unfold_only(model: CompiledModel, egrid: np.ndarray, theta: np.ndarray, device=None) -> np.ndarray:
    """CompiledModel.eval backend: a plan with a one-channel dummy response around the
    given grid, cached per (model, grid)."""
    key = (id(model), egrid.tobytes(), device)
    lk = _unfold_cache.get(key)
    if lk is None:
        E = egrid.size - 1
        d = FixedData(name='eval', photon_egrid=egrid, spec_counts=np.zeros(1), area_scale=np.ones(1),
                      spec_exposure=1.0, response_matrix=np.zeros((E, 1)))
        lk = _EvalOnly([d], [model], ['cstat'], device=device, chunk=128)
        for old in _unfold_cache.values():
            old.close()  # one cached plan at a time: release the evicted one's device memory now
        _unfold_cache.clear()
        _unfold_cache[key] = lk
    return lk.unfold(0, theta).cpu().numpy()


class _EvalOnly(Likelihood):
    """Plan used by CompiledModel.eval: multiplicative models are allowed here."""

    def __init__(self, data, model, stat, device=None, chunk=0, specialise=None):
        types = [m.type for m in model]
        for m in model:
            m.type = 'add'
        try:
            super().__init__(data, model, stat, device=device, chunk=chunk, specialise=specialise, fold='dmma')
        finally:
            for m, t in zip(model, types):
                m.type = t


def sliced_gemm(a, b, slices: int = 6):
    """``a @ b.T`` for CUDA float64 tensors a (M, K), b (N, K) on the sliced-integer fold
    engine (int8 tcgen05 tensor cores, FP64-class accuracy; csrc/fold_i8.cuh).  Returns
    ``(d, ms)``: the (M, N) product and the device time of the contraction kernel."""
    import torch

    lib = _lib.load()
    if not (a.is_cuda and b.is_cuda and a.dtype == torch.float64 and b.dtype == torch.float64):
        raise TypeError('sliced_gemm needs CUDA float64 tensors')
    a, b = a.contiguous(), b.contiguous()
    (m, k), (n, k2) = a.shape, b.shape
    if k != k2:
        raise ValueError('inner dimensions differ')
    d = torch.empty((m, n), dtype=torch.float64, device=a.device)
    ms = C.c_double(0.0)
    with torch.cuda.device(a.device):
        _lib.check(lib.elisa_b200_sliced_gemm_dev(_dev_ptr(a), k, _dev_ptr(b), k, m, n, k, int(slices), _dev_ptr(d), n,
                                                  C.byref(ms), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return d, ms.value
