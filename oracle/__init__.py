"""CPU oracle for the ELISA forward-folding likelihood path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU in float64, the arithmetic of the reference's
hot path: component photon-flux integrals, model composition,
response fold, background handling, the five fit statistics and their reduction,
with gradients obtained by torch's reverse-mode autograd (the stand-in for
``jax.grad`` -- both differentiate the same primitive graph, ``where`` included).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or
as the timed CPU baseline.  The product (``elisa_b200``) never imports it.

PARITY STATUS: the reference itself (pure Python on JAX/numpyro) cannot be imported
in the build container (no jax, numpyro, ... and no network), and its test-suite
holds no golden vectors for statistic values or gradients.  The
oracle is therefore pinned only to what the reference's tests do pin -- the
closed-form power-law integral (tests/conftest.py:23-31), the closed-form cstat MLE
and error for PowerLaw(alpha=0) on an identity response (tests/test_fit.py:24-42),
finiteness of every built-in component on geomspace(1e-4, 1e9, 1301)
(tests/test_model.py:178-195) -- and to the reference's own independent
restatements of the deviances (plot/residuals.py:406-427, plot/errorbars.py:20-34).
Beyond those rows: **parity unpinned** (authoritative by line-by-line restatement).
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
