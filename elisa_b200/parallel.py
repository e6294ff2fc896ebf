"""Multi-GPU layout of the path: the N axis (parameter vectors / replicas / chains) is cut
into contiguous blocks, one per rank; every plan constant (responses, grids, observed
data) is replicated; there is NO collective on the data path.  One optional all-gather
of the per-sample results afterwards -- tens of MB at most over NVSwitch.

The reference's counterpart is ``jax.pmap`` / ``shard_map`` over the same axis
(infer/helper.py:690-718, infer/results.py:557-569): embarrassingly parallel, results
concatenated by the runtime.  One process per GPU here, ``torch.distributed`` (nccl on
GPUs, gloo in the CPU tests) for the plumbing.
"""

from __future__ import annotations

from typing import Callable, Sequence


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced blocks: the first n % world ranks get one extra row."""
    if n < 0 or world < 1:
        raise ValueError('n >= 0 and world >= 1 required')
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _dist():
    import torch.distributed as dist

    return dist


def world_info(group=None) -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def all_gather_rows(local, n_total: int, group=None):
    """Concatenate the ranks' row blocks (shard_bounds order) into the full (n_total, ...)
    tensor on every rank with ONE ``all_gather_into_tensor``.  Equal blocks (n_total divisible
    by the world size -- the bench shapes, 64 chains over 8 GPUs) are gathered straight into the
    result, no staging copy; blocks that differ by one row are gathered into a padded buffer
    whose layout matches the result except for one spare row per rank, and compacted in place."""
    import torch

    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    local = local.contiguous()
    bounds = shard_bounds(n_total, world)
    width = bounds[0][1] - bounds[0][0]  # the largest block (the first ranks carry the extra row)
    tail = tuple(local.shape[1:])
    if n_total % world == 0:
        out = torch.empty((n_total, *tail), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    send = local
    if local.shape[0] != width:
        send = torch.empty((width, *tail), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
        send[local.shape[0]:] = 0
    buf = torch.empty((world * width, *tail), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send, group=group)
    # ranks [0, extra) are full; from rank `extra` on each block is one row short: close the gaps
    extra = n_total % world
    out = buf[:n_total]
    for r in range(extra + 1, world):  # block r moves up by (r - extra) rows; regions may overlap -> clone
        lo, hi = bounds[r]
        out[lo:hi] = buf[r * width: r * width + (hi - lo)].clone()
    return out


def sharded_map(fn: Callable, args: Sequence, n_total: int, group=None, gather: bool = True):
    """Evaluate ``fn(*[a[lo:hi] for a in args])`` on this rank's block of rows and, if
    ``gather``, return the results of all ranks concatenated in row order (a tensor or a
    tuple of tensors, as ``fn`` returns).  ``None`` entries of ``args`` pass through."""
    rank, world = world_info(group)
    lo, hi = shard_bounds(n_total, world)[rank]
    local = fn(*[None if a is None else a[lo:hi] for a in args])
    if not gather or world == 1:
        return local
    if isinstance(local, (tuple, list)):
        return tuple(all_gather_rows(t, n_total, group) for t in local)
    return all_gather_rows(local, n_total, group)


class _DeviceArray:
    """A raw device allocation as something ``torch.as_tensor`` takes without copying."""

    def __init__(self, ptr: int, shape: tuple, typestr: str = '<f8'):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


class PeerGather:
    """The gather fused into the evaluation (``elisa_b200_plan_set_gather``, include/elisa_b200.h): every rank owns
    FULL result arrays ``ll [n_total, S]`` / ``grad [n_total, S, P]`` in CUDA-IPC memory which every peer maps, and a
    plan with the gather set pushes the rows of its finished chunks into all of them over NVLink with the copy
    engines while it computes the next chunks; the call ends with a flag to every peer and a wait for theirs.  No
    NCCL on the data path: ``torch.distributed`` only carries the 64-byte IPC handles, once.  ``sets`` copies of the
    arrays are used in turn (a rank may be one call ahead of a peer whose consumer still reads the previous
    results)."""

    def __init__(self, likelihood, n_total: int, group=None, sets: int = 2):
        import ctypes as C

        import torch

        from . import _lib

        dist = _dist()
        self.lk, self.group, self.n_total, self.sets = likelihood, group, int(n_total), int(sets)
        self.rank, self.world = world_info(group)
        if self.world < 2:
            raise ValueError('PeerGather needs an initialised process group of at least 2 ranks')
        self._lib = lib = likelihood._lib
        S, P = likelihood._S, likelihood._P
        self.bounds = shard_bounds(self.n_total, self.world)
        sizes = [self.world * 8] + [self.n_total * S * 8, self.n_total * S * P * 8] * self.sets  # flags, (ll, grad) per set
        # Setting up is collective and must fail on EVERY rank or on none (a rank that raised while the others wait in
        # a collective would hang the job): errors are collected, exchanged, and raised together.
        self._own, handles, self._maps, err = [], [], None, None
        import contextlib

        stack = contextlib.ExitStack()
        try:
            stack.enter_context(torch.cuda.device(likelihood.device))
        except Exception as exc:
            err = f'rank {self.rank}: {exc}'
        with stack:
            try:
                if err is not None:
                    raise RuntimeError(err.split(': ', 1)[1])
                for b in sizes:
                    ptr, h = C.c_void_p(), C.create_string_buffer(64)
                    _lib.check(lib.elisa_b200_peer_alloc(max(int(b), 8), C.byref(ptr), h))
                    self._own.append(ptr.value)
                    handles.append(h.raw)
            except Exception as exc:
                err = f'rank {self.rank}: {exc}'
            everyone = [None] * self.world
            dist.all_gather_object(everyone, None if err else handles, group=group)
            maps = []  # [rank][array] device pointers as mapped here
            if err is None and all(h is not None for h in everyone):
                try:
                    for r in range(self.world):
                        if r == self.rank:
                            maps.append(list(self._own))
                            continue
                        row = []
                        maps.append(row)
                        for h in everyone[r]:
                            ptr = C.c_void_p()
                            _lib.check(lib.elisa_b200_peer_open(h, C.byref(ptr)))
                            row.append(ptr.value)
                except Exception as exc:
                    err = f'rank {self.rank}: {exc}'
            errors = [None] * self.world
            dist.all_gather_object(errors, err, group=group)  # also: every rank has mapped every array before anyone pushes
            if any(e is not None for e in errors) or any(h is None for h in everyone):
                for r, row in enumerate(maps):
                    if r != self.rank:
                        for p in row:
                            lib.elisa_b200_peer_close(p)
                for p in self._own:
                    lib.elisa_b200_peer_free(p)
                self._own = []
                raise RuntimeError('peer gather could not be set up: ' + '; '.join(e for e in errors if e))
            self._maps = maps
        dev = f'cuda:{likelihood.device}'
        self._views = []
        for k in range(self.sets):
            ll = torch.as_tensor(_DeviceArray(self._own[1 + 2 * k], (self.n_total, S)), device=dev)
            g = torch.as_tensor(_DeviceArray(self._own[2 + 2 * k], (self.n_total, S, P)), device=dev)
            self._views.append((ll, g))
        self._next = 0
        ptrs = lambda index: (C.c_void_p * self.world)(*[self._maps[r][index] for r in range(self.world)])
        self._flag_ptrs = ptrs(0)
        self._set_ptrs = [(ptrs(1 + 2 * k), ptrs(2 + 2 * k)) for k in range(self.sets)]
        self._set_gather = lib.elisa_b200_plan_set_gather

    def begin(self):
        """Set the plan's gather to the next set of arrays; returns (ll_full, grad_full, lo, hi): the full arrays
        of this rank and its block of rows -- evaluate with ``out=(ll_full[lo:hi], grad_full[lo:hi])``."""
        k = self._next
        self._next = (k + 1) % self.sets
        lo, hi = self.bounds[self.rank]
        rc = self._set_gather(self.lk._h, self.world, self.rank, self._set_ptrs[k][0], self._set_ptrs[k][1], self._flag_ptrs, lo)
        if rc:
            from . import _lib

            _lib.check(rc)
        ll, g = self._views[k]
        return ll, g, lo, hi

    def end(self):
        rc = self._set_gather(self.lk._h, 0, 0, None, None, None, 0)
        if rc:
            from . import _lib

            _lib.check(rc)

    def close(self):
        """Unmap the peers' arrays and free this rank's (collective: the peers must have stopped pushing)."""
        if self._maps is None:
            return
        import torch

        torch.cuda.synchronize(self.lk.device)
        _dist().barrier(group=self.group)
        with torch.cuda.device(self.lk.device):
            for r in range(self.world):
                if r != self.rank:
                    for p in self._maps[r]:
                        self._lib.elisa_b200_peer_close(p)
            _dist().barrier(group=self.group)
            self._views = []
            for p in self._own:
                self._lib.elisa_b200_peer_free(p)
        self._maps = None


class ShardedLikelihood:
    """A :class:`elisa_b200.Likelihood` per rank (constants replicated on its GPU) evaluating
    this rank's block of the batch.  ``gather=False`` keeps results sharded (e.g. NUTS with
    the chains resident per GPU: no collective at all).  ``peer=True`` (GPUs of one box, one process each): the
    gather of ``loglike`` / ``loglike_and_grad`` is fused into the evaluation (:class:`PeerGather`) instead of
    NCCL all-gathers after it; the arrays returned are then reused two gathered calls later."""

    def __init__(self, likelihood, group=None, peer: bool = False):
        self.lk = likelihood
        self.group = group
        self.peer = bool(peer)
        self._gathers = {}

    def _peer_gather(self, n):
        if n not in self._gathers:
            self._gathers[n] = PeerGather(self.lk, n, self.group)
        return self._gathers[n]

    def close(self):
        for g in self._gathers.values():
            g.close()
        self._gathers = {}

    def _peer_eval(self, theta, counts_on, counts_off, n, want_grad):
        pg = self._peer_gather(n)
        ll, g, lo, hi = pg.begin()
        try:
            sl = lambda a: None if a is None else a[lo:hi]
            if want_grad:
                self.lk.loglike_and_grad(theta[lo:hi], sl(counts_on), sl(counts_off), out=(ll[lo:hi], g[lo:hi]))
            else:
                self.lk.loglike(theta[lo:hi], sl(counts_on), sl(counts_off), out=ll[lo:hi])
        finally:
            pg.end()
        return (ll, g[:, :, : self.lk.nparam]) if want_grad else ll

    @staticmethod
    def _rows(theta, counts_on=None, counts_off=None):
        """Number of rows of the batch and theta as a 2-D array: a 1-D theta is ONE parameter
        vector, and a single row next to per-replica counts is broadcast over the replicas."""
        import numpy as np
        import torch

        if not isinstance(theta, torch.Tensor):
            theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        elif theta.dim() == 1:
            theta = theta.unsqueeze(0)
        n = theta.shape[0]
        for c in (counts_on, counts_off):
            if c is not None and n == 1:
                n = len(c)
        return theta, n

    def _use_peer(self, gather, n):
        return self.peer and gather and world_info(self.group)[1] > 1 and n >= world_info(self.group)[1]

    def loglike(self, theta, counts_on=None, counts_off=None, gather=True):
        theta, n = self._rows(theta)
        if self._use_peer(gather, n):
            return self._peer_eval(theta, counts_on, counts_off, n, False)
        return sharded_map(self.lk.loglike, (theta, counts_on, counts_off), n, self.group, gather)

    def loglike_and_grad(self, theta, counts_on=None, counts_off=None, gather=True):
        theta, n = self._rows(theta)
        if self._use_peer(gather, n):
            return self._peer_eval(theta, counts_on, counts_off, n, True)
        return sharded_map(self.lk.loglike_and_grad, (theta, counts_on, counts_off), n, self.group, gather)

    def fit(self, theta0, counts_on=None, counts_off=None, gather=True, **fit_kw):
        """Batched re-fit of this rank's block of replicas (``Likelihood.fit``); with ``gather``
        every rank receives the full ('theta', 'deviance', 'grad_norm', 'n_iter', 'valid')
        -- the reference's ``fit_in_parallel`` (infer/helper.py:690-718) concatenates the
        per-device results the same way.  ``theta0`` may be one row (or 1-D) broadcast over the
        replicas of ``counts_on``, as in ``Likelihood.fit``."""
        keys = ('theta', 'deviance', 'grad_norm', 'n_iter', 'valid')
        theta0, n = self._rows(theta0, counts_on, counts_off)
        broadcast = theta0.shape[0] == 1 and n > 1
        rank, world = world_info(self.group)
        lo, hi = shard_bounds(n, world)[rank]

        def local(t, on, off):
            if hi == lo:  # more ranks than replicas: this rank has nothing to fit
                import torch

                P = self.lk.nparam
                dev = f'cuda:{self.lk.device}'
                z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)
                return z(0, P), z(0), z(0), z(0), z(0)
            out = self.lk.fit(t, counts_on=on, counts_off=off, **fit_kw)
            return tuple(out[k].double() for k in keys)  # one dtype for the gather

        if broadcast:  # slice the replicas, not the single start vector
            res = local(theta0, None if counts_on is None else counts_on[lo:hi], None if counts_off is None else counts_off[lo:hi])
            if gather and world > 1:
                res = tuple(all_gather_rows(t, n, self.group) for t in res)
        else:
            res = sharded_map(local, (theta0, counts_on, counts_off), n, self.group, gather)
        out = dict(zip(keys, res))
        out['n_iter'] = out['n_iter'].long()
        out['valid'] = out['valid'] > 0.5
        return out
