"""CPU oracle for the ELISA forward-folding likelihood path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU in float64, the arithmetic of the reference's
hot path: component photon-flux integrals, model composition,
response fold, background handling, the five fit statistics and their reduction,
with gradients obtained by torch's reverse-mode autograd (the stand-in for
``jax.grad`` -- both differentiate the same primitive graph, ``where`` included).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or
as the timed CPU baseline.  The product (``elisa_b200``) never imports it.

PARITY STATUS: pinned.  The reference itself (pure Python on JAX/numpyro) cannot be
imported in the build container (no jax, numpyro, ... and no network) and its test-suite
holds no golden vectors for statistic values or gradients, so the
fixtures under tests/golden/ were produced by EXECUTING THE REFERENCE'S OWN SOURCE for
this path on NumPy stand-ins for jax.numpy / numpyro (tools/make_golden.py:
infer/likelihood.py imported unmodified; component bodies, _powerlaw_integral and the
trapz / Simpson closures extracted from models/add.py, mul.py, model.py with ast).
tests/test_cpu_oracle_golden.py holds this oracle to those outputs (<= 2e-11 on the bin
values of all 23 components, 1e-12 on every numpyro site of the five statistics,
finite-difference gradients of the reference functions at 5e-6) and to the closed forms
the reference's tests pin (tests/conftest.py:23-31, tests/test_fit.py:24-42,
tests/test_model.py:178-195).  What stays unpinned: XLA's own libm versus NumPy's (a
few ulp), and reverse-mode gradients beyond the finite-difference check -- torch
autograd differentiates the same primitive graph as jax.grad.
"""

from .components import ADDITIVE, MULTIPLICATIVE, eval_component  # noqa: F401
from .likelihood import (  # noqa: F401
    STATISTICS,
    eval_model,
    loglike,
    loglike_and_grad,
    pgstat_background,
    sites,
    wstat_background,
)
