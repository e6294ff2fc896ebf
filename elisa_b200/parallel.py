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


class ShardedLikelihood:
    """A :class:`elisa_b200.Likelihood` per rank (constants replicated on its GPU) evaluating
    this rank's block of the batch.  ``gather=False`` keeps results sharded (e.g. NUTS with
    the chains resident per GPU: no collective at all)."""

    def __init__(self, likelihood, group=None):
        self.lk = likelihood
        self.group = group

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

    def loglike(self, theta, counts_on=None, counts_off=None, gather=True):
        theta, n = self._rows(theta)
        return sharded_map(self.lk.loglike, (theta, counts_on, counts_off), n, self.group, gather)

    def loglike_and_grad(self, theta, counts_on=None, counts_off=None, gather=True):
        theta, n = self._rows(theta)
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
