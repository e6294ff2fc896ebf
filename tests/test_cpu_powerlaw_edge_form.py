"""The power law's tangent in the edge-wise gradient contraction (Comp<ELISA_COMP_POWERLAW>::basis,
csrc/components.cuh), restated in NumPy and held against extended precision.

grad_alpha = sum_e G_e d/dalpha [K (F(E_e+1) - F(E_e))],  F = E^(1-alpha) / (1-alpha)  (models/add.py:40-50).
Summed over EDGES the weights w_k = G_k-1 - G_k add up to zero, so a constant may be taken off F: the kernel
uses F~ = expm1((1-alpha) ln E) / (1-alpha) and dF~/dalpha = (F~ - E^(1-alpha) ln E) / (1-alpha), which stay of
order ln E near alpha = 1 (the reference's DEFAULT is alpha = 1.01), where F is 1 / (1-alpha) plus that.  The
per-bin form -- what forward duals, and JAX on the reference formula, evaluate -- loses (1-alpha)^-2 digits."""

import numpy as np
import pytest

E = np.geomspace(0.3, 80.0, 2049)  # the bench's photon grid
K = 10.0


def _weights(G):
    w = np.zeros(E.size)
    w[:-1] -= G
    w[1:] += G
    return w


def per_bin(G, alpha, dt):
    e, a = E.astype(dt), dt(alpha)
    ln = np.log(e)
    t = np.exp((1 - a) * ln) / (1 - a) * (1 / (1 - a) - ln)  # dF/dalpha on the edges
    return (G.astype(dt) * K * (t[1:] - t[:-1])).sum()


def kernel_form(G, alpha):
    ln = np.log(E)
    c0 = 1.0 - alpha
    c1 = 1.0 / c0
    m = np.expm1(c0 * ln)
    b0 = m * c1
    b1 = (b0 - (m + 1.0) * ln) * c1
    return K * (_weights(G) * b1).sum()


@pytest.mark.parametrize('alpha', [1.7, 1.05, 1.01, 1.001, 0.999, 0.3, 2.5])
def test_kernel_form_against_extended_precision(alpha):
    if np.finfo(np.longdouble).eps > 1e-18:
        pytest.skip('no extended precision on this platform')
    rng = np.random.default_rng(0)
    G = rng.standard_normal(2048) * np.exp(-0.5 * np.log(E[:-1] / 3.0) ** 2)  # weights of both signs
    ref = per_bin(G, alpha, np.longdouble)
    err_kernel = abs(kernel_form(G, alpha) - ref) / abs(ref)
    err_per_bin = abs(per_bin(G, alpha, np.float64) - ref) / abs(ref)
    assert err_kernel <= 2e-10, (alpha, float(err_kernel))
    # never worse than the reference's own form by more than rounding, and far better near alpha = 1
    assert err_kernel <= max(4.0 * err_per_bin, 1e-12), (alpha, float(err_kernel), float(err_per_bin))
    if abs(alpha - 1.0) <= 0.0011:
        assert err_kernel <= 1e-2 * err_per_bin
