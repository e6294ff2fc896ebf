"""Autodiff binding of the C ABI: the custom-JVP contract of INTEGRATION.md section 4, executed.

Upstream the likelihood sits inside ``numpyro_model`` (infer/helper.py:332-367) and is
differentiated by JAX: reverse mode for NUTS (``value_and_grad(potential_fn)``, set up at
infer/fit.py:833-843) and for Minuit's gradient (infer/fit.py:487).  The binding a maintainer adds
is a ``jax.custom_jvp`` around the FFI call whose primal returns ``(value, J)`` and whose tangent
is ``J . t`` in plain JAX -- linear in ``t``, hence transposable, hence reverse mode works.  JAX is
not installable in this image (``xla/ffi/api/ffi.h`` is absent too, csrc/elisa_b200_xla.cc is
compiled only where it exists), so the SAME contract is wired into the autodiff framework that is
here: a ``torch.autograd.Function`` whose forward evaluates value and Jacobian in one call of the
library and caches ``J``, and whose backward is the transposed product ``J^T . gbar``.  Everything
upstream of the boundary (transforms to unconstrained space, priors, log-Jacobians, the sampler)
differentiates through it with ordinary autograd, as numpyro does with the JAX version.

``hmc`` is a small batched Hamiltonian Monte Carlo sampler (fixed step, fixed trajectory length,
all chains advancing together).  It is a CONSUMER of the binding for tests and the cfg-3 latency
bench -- the production samplers (numpyro NUTS, blackjax, emcee ...) stay in the reference.
"""

from __future__ import annotations

import math
from typing import Callable

import torch

__all__ = ['LogLikeFunction', 'loglike', 'jvp', 'hmc', 'interval_bijection']


class LogLikeFunction(torch.autograd.Function):
    """``ll[n, s] = loglike(theta[n, :])`` with the Jacobian ``J[n, s, p]`` from the library.

    forward : one ``loglike_and_grad`` call of :class:`elisa_b200.Likelihood` (value + J);
    backward: ``grad_theta[n, p] = sum_s gbar[n, s] J[n, s, p]``.
    """

    @staticmethod
    def forward(ctx, theta, lk, counts_on, counts_off, guard):
        ll, jac = lk.loglike_and_grad(theta.detach(), counts_on, counts_off, guard=guard)
        ctx.save_for_backward(jac)
        return ll

    @staticmethod
    def backward(ctx, gbar):
        (jac,) = ctx.saved_tensors
        return torch.einsum('ns,nsp->np', gbar, jac), None, None, None, None


def loglike(lk, theta, counts_on=None, counts_off=None, guard=None):
    """Differentiable (N, S) log-likelihood of ``lk`` (an ``elisa_b200.Likelihood``) at the CUDA
    tensor ``theta`` (N, P): use it inside any torch expression and call ``.backward()``."""
    return LogLikeFunction.apply(theta, lk, counts_on, counts_off, guard)


def jvp(lk, theta, tangent, counts_on=None, counts_off=None, guard=None):
    """``(value, J . t)``: the forward-mode rule of the custom-JVP contract (tangent (N, P))."""
    ll, jac = lk.loglike_and_grad(theta, counts_on, counts_off, guard=guard)
    return ll, torch.einsum('nsp,np->ns', jac, tangent.to(jac.device))


def interval_bijection(lower, upper):
    """numpyro's ``biject_to(interval(lower, upper))``: theta = lower + (upper - lower) sigmoid(u)
    and log |d theta / d u|, as torch functions of u (priors uniform on the interval, the
    reference's default, infer/helper.py:409-426)."""
    def forward(u):
        lo, hi = lower.to(u), upper.to(u)
        sg = torch.sigmoid(u)
        theta = lo + (hi - lo) * sg
        log_jac = (torch.log(hi - lo) + torch.nn.functional.logsigmoid(u) + torch.nn.functional.logsigmoid(-u)).sum(-1)
        return theta, log_jac

    return forward


def hmc(potential: Callable, q0, n_iter: int, n_leapfrog: int, step: float, seed: int = 0, mass=None):
    """Batched HMC on ``potential(q) -> U (N,)`` (a differentiable torch function), all N chains
    advancing together.  Momenta and acceptance draws come from a CPU generator seeded with
    ``seed`` (so that two back-ends see the same random numbers).  Returns a dict with 'q'
    (n_iter, N, D) positions after each iteration, 'accept' (n_iter, N) and 'n_grad' (gradient
    evaluations = leapfrog steps + 1 per iteration)."""
    dev, dt = q0.device, q0.dtype
    gen = torch.Generator(device='cpu').manual_seed(int(seed))
    n, d = q0.shape
    inv_mass = torch.ones(d, dtype=dt, device=dev) if mass is None else (1.0 / mass.to(q0))

    def value_and_grad(q):
        q = q.detach().requires_grad_(True)
        u = potential(q)
        (g,) = torch.autograd.grad(u.sum(), q)
        return u.detach(), g

    q = q0.detach().clone()
    u, g = value_and_grad(q)
    out_q, out_acc, n_grad = [], [], 1
    for _ in range(n_iter):
        p = (torch.randn((n, d), dtype=dt, generator=gen) / torch.sqrt(inv_mass.cpu())).to(dev)
        logu = torch.log(torch.rand((n,), dtype=dt, generator=gen)).to(dev)
        h0 = u + 0.5 * (p * p * inv_mass).sum(-1)
        qn, pn, un, gn = q, p, u, g
        for _ in range(n_leapfrog):
            pn = pn - 0.5 * step * gn
            qn = qn + step * pn * inv_mass
            un, gn = value_and_grad(qn)
            pn = pn - 0.5 * step * gn
            n_grad += 1
        h1 = un + 0.5 * (pn * pn * inv_mass).sum(-1)
        acc = (logu < (h0 - h1)) & torch.isfinite(h1)
        q = torch.where(acc[:, None], qn, q)
        u = torch.where(acc, un, u)
        g = torch.where(acc[:, None], gn, g)
        out_q.append(q.clone())
        out_acc.append(acc.clone())
    return {'q': torch.stack(out_q), 'accept': torch.stack(out_acc), 'n_grad': n_grad,
            'potential': u, 'log_step': math.log(step)}
