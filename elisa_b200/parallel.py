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
    tensor on every rank.  Blocks may differ by one row: pad to the largest, gather, trim."""
    import torch

    dist = _dist()
    rank, world = world_info(group)
    if world == 1:
        return local
    bounds = shard_bounds(n_total, world)
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((width, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * width: r * width + (hi - lo)] for r, (lo, hi) in enumerate(bounds)]
    return torch.cat(parts, dim=0)


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

    def loglike(self, theta, counts_on=None, counts_off=None, gather=True):
        return sharded_map(self.lk.loglike, (theta, counts_on, counts_off), len(theta), self.group, gather)

    def loglike_and_grad(self, theta, counts_on=None, counts_off=None, gather=True):
        return sharded_map(self.lk.loglike_and_grad, (theta, counts_on, counts_off), len(theta), self.group, gather)

    def fit(self, theta0, counts_on=None, counts_off=None, gather=True, **fit_kw):
        """Batched re-fit of this rank's block of replicas (``Likelihood.fit``); with ``gather``
        every rank receives the full ('theta', 'deviance', 'grad_norm', 'n_iter', 'valid')
        -- the reference's ``fit_in_parallel`` (infer/helper.py:690-718) concatenates the
        per-device results the same way."""
        keys = ('theta', 'deviance', 'grad_norm', 'n_iter', 'valid')

        def local(t, on, off):
            out = self.lk.fit(t, counts_on=on, counts_off=off, **fit_kw)
            return tuple(out[k].double() for k in keys)  # one dtype for the gather

        res = sharded_map(local, (theta0, counts_on, counts_off), len(theta0), self.group, gather)
        out = dict(zip(keys, res))
        out['n_iter'] = out['n_iter'].long()
        out['valid'] = out['valid'] > 0.5
        return out
