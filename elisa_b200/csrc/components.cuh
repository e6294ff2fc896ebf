// Built-in spectral components of the reference, written once and generic in the
// scalar type (double -> value, Dual<NP> -> value + partials; see dual.cuh).
//
// Each component restates the body cited beside it (reference file:line).  A component
// is described by
//   pre()   per-sample scalars hoisted out of the energy loop,
//   edge()  the function evaluated on a grid EDGE: the continuum N(E) for
//           NumericalIntegral components (models/model.py:1707-1761) or the
//           antiderivative for AnalyticalIntegral ones (models/model.py:1684-1698),
//   bin()   how two neighbouring edge values become the BIN value.
// Powers of an energy are evaluated as exp(expo * ln E) with ln E tabulated per grid at
// plan creation (relative deviation from jnp.power <= |expo ln E| * 2^-53).
#pragma once
#include "elisa_b200.h"
#include "dual.cuh"

namespace elisa {

enum Rule { RULE_CONTINUUM_ADD = 0, RULE_CONTINUUM_MUL = 1, RULE_CUSTOM = 2 };

template <typename T>
struct Cst {
    __device__ __forceinline__ static T make(double c);
};
template <>
struct Cst<double> {
    __device__ __forceinline__ static double make(double c) { return c; }
};
template <int N>
struct Cst<Dual<N>> {
    __device__ __forceinline__ static Dual<N> make(double c) { return make_const<N>(c); }
};

// a grid value as the leaf's scalar type: plain energies are constants, energies under a FREE shift
// (models/conv.py:294-432 with z / v a fit parameter) already are of that type
template <typename T>
__device__ __forceinline__ T lift(double x) {
    return Cst<T>::make(x);
}
template <typename T, int N>
__device__ __forceinline__ T lift(const Dual<N>& x) {
    return x;
}

#define ELISA_PI 3.141592653589793238462643383279502884

// F(x; alpha) of _powerlaw_integral (models/add.py:40-50) on one edge, given ln x.
// oma = where(alpha != 1, 1 - alpha, 1); F = alpha != 1 ? x^oma / oma : ln x.
template <typename T>
__device__ __forceinline__ T powerlaw_antiderivative(bool cond, const T& oma, double lnx) {
    if (cond) return pw_e(lnx, oma) / oma;
    return Cst<T>::make(lnx);
}
template <int N>
__device__ __forceinline__ Dual<N> powerlaw_antiderivative(bool cond, const Dual<N>& oma, const Dual<N>& lnx) {
    if (cond) return ex(oma * lnx) / oma;  // the energy itself carries a tangent (free energy shift)
    return lnx;
}
template <typename T>
__device__ __forceinline__ T powerlaw_antiderivative_t(bool cond, const T& oma, const T& lnx) {
    if (cond) return ex(oma * lnx) / oma;
    return lnx;
}

template <int KIND>
struct Comp;

// ------------------------------------------------------------------ additive, continuum
template <>
struct Comp<ELISA_COMP_BAND> {  // Band.continuum, models/add.py:100-121
    static constexpr int NP = 4, NG = 1, NC = 3, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        const T &alpha = p[0], &beta = p[1], &Ec = p[2];
        const T amb_ = alpha - beta;
        const T inv_Ec = 1.0 / Ec;
        const T amb = sel(value(amb_) < value(inv_Ec), inv_Ec, amb_);
        c[0] = inv_Ec;
        c[1] = Ec * amb;                                // Ebreak
        c[2] = amb * lg(amb * Ec * (1.0 / 100.0)) - amb;  // amb log(amb Ec / e0) - amb
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE lnE, T* g) {
        const TE le = lnE - 4.605170185988091368;  // log(E / e0), e0 = 100
        const T lg_ = sel(value(E) < value(c[1]), p[0] * le - E * c[0], c[2] + p[1] * le);
        g[0] = p[3] * ex(lg_);
    }
};

template <>
struct Comp<ELISA_COMP_BANDEP> {  // BandEp.continuum, models/add.py:177-200
    static constexpr int NP = 4, NG = 1, NC = 3, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        const T &alpha = p[0], &beta = p[1], &Ep = p[2];
        const T Ec = Ep / (2.0 + alpha);
        const T amb_ = alpha - beta;
        const T inv_Ec = 1.0 / Ec;
        const T amb = sel(value(amb_) < value(inv_Ec), inv_Ec, amb_);
        c[0] = inv_Ec;
        c[1] = amb_ * Ec;  // Ebreak uses (alpha - beta), not the guarded amb (add.py:185)
        c[2] = amb * lg(amb * Ec * (1.0 / 100.0)) - amb;
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE lnE, T* g) {
        const TE le = lnE - 4.605170185988091368;
        const T lg_ = sel(value(E) < value(c[1]), p[0] * le - E * c[0], c[2] + p[1] * le);
        g[0] = p[3] * ex(lg_);
    }
};

template <>
struct Comp<ELISA_COMP_BENTPL> {  // BentPL.continuum, models/add.py:231-236
    static constexpr int NP = 3, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = lg(p[1]);
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE lnE, T* g) {
        g[0] = p[2] * 2.0 / (1.0 + ex(p[0] * (lnE - c[0])));
    }
};

template <>
struct Comp<ELISA_COMP_BLACKBODY> {  // Blackbody.continuum, models/add.py:266-286
    static constexpr int NP = 2, NG = 1, NC = 2, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[0];
        c[1] = 8.0525 * p[1] / (p[0] * p[0] * p[0]);
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        const T x = E * c[0];
        const T tmp = c[1] * E;
        const double xv = value(x);
        if (xv <= 1e-4) {
            g[0] = tmp;
        } else if (xv >= 50.0) {
            g[0] = Cst<T>::make(0.0);
        } else {
            g[0] = tmp * x * rinv(em1(x));  // x in (1e-4, 50): expm1(x) in (1e-4, 5.2e21)
        }
    }
    // closed-form edge tangents (see EdgeLeaf): b0 = g / K, b1 = (g / K) (x (1 + 1 / expm1 x) - 4), x = E / kT
    static constexpr int NB = 2;
    __device__ static void basis(const double* p, const double* c, double E, double, double& v, double* b) {
        const double x = E * c[0];
        const double t = (8.0525 * c[0] * c[0] * c[0]) * E;
        if (x <= 1e-4) {
            b[0] = t;
            b[1] = -3.0 * t;
        } else if (x >= 50.0) {
            b[0] = 0.0;
            b[1] = 0.0;
        } else {
            const double r = rinv(em1(x));
            b[0] = t * x * r;
            b[1] = b[0] * (fma(x, r, x) - 4.0);
        }
        v = p[1] * b[0];
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        d[0] = p[1] * c[0] * S[1];
        d[1] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_BLACKBODYRAD> {  // BlackbodyRad.continuum, models/add.py:317-338
    static constexpr int NP = 2, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[0];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        const T x = E * c[0];
        const T tmp = (1.0344e-3 * E) * p[1];
        const double xv = value(x);
        if (xv <= 1e-4) {
            g[0] = tmp * p[0];
        } else if (xv >= 50.0) {
            g[0] = Cst<T>::make(0.0);
        } else {
            g[0] = tmp * E * rinv(em1(x));
        }
    }
    // closed-form edge tangents: b0 = g / K, b1 = kT / K dg/dkT
    static constexpr int NB = 2;
    __device__ static void basis(const double* p, const double* c, double E, double, double& v, double* b) {
        const double x = E * c[0];
        const double t = 1.0344e-3 * E;
        if (x <= 1e-4) {
            b[0] = t * p[0];
            b[1] = b[0];
        } else if (x >= 50.0) {
            b[0] = 0.0;
            b[1] = 0.0;
        } else {
            const double r = rinv(em1(x));
            b[0] = t * E * r;
            b[1] = b[0] * fma(x, r, x);
        }
        v = p[1] * b[0];
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        d[0] = p[1] * c[0] * S[1];
        d[1] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_COMPT> {  // Compt.continuum, models/add.py:473-479
    static constexpr int NP = 3, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = (p[0] - 2.0) / p[1];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE lnE, T* g) {
        g[0] = p[2] * ex(E * c[0] - p[0] * lnE);
    }
    // closed-form edge tangents: b = e (1, E, ln E), e = exp(E (a - 2) / b - a ln E)
    static constexpr int NB = 3;
    __device__ static void basis(const double* p, const double* c, double E, double lnE, double& v, double* b) {
        const double e = ex(E * c[0] - p[0] * lnE);
        v = p[2] * e;
        b[0] = e;
        b[1] = e * E;
        b[2] = e * lnE;
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        const double inv = 1.0 / p[1];
        d[0] = p[2] * (S[1] * inv - S[2]);
        d[1] = -(p[2] * c[0] * inv * S[1]);
        d[2] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_CUTOFFPL> {  // CutoffPL.continuum, models/add.py:434-439
    static constexpr int NP = 3, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[1];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE lnE, T* g) {
        g[0] = p[2] * ex(-(p[0] * lnE) - E * c[0]);
    }
    // closed-form edge tangents: b = e (1, ln E, E), e = exp(-alpha ln E - E / Ec)
    static constexpr int NB = 3;
    __device__ static void basis(const double* p, const double* c, double E, double lnE, double& v, double* b) {
        const double e = ex(-(p[0] * lnE) - E * c[0]);
        v = p[2] * e;
        b[0] = e;
        b[1] = e * lnE;
        b[2] = e * E;
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        d[0] = -(p[2] * S[1]);
        d[1] = p[2] * c[0] * c[0] * S[2];
        d[2] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_GAUSS> {  // Gauss.continuum, models/add.py:511-516 (jax norm.pdf)
    static constexpr int NP = 3, NG = 1, NC = 2, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        const T s2 = p[1] * p[1];
        c[0] = lg((2.0 * ELISA_PI) * s2);
        c[1] = 1.0 / s2;
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        const T dx = E - p[0];
        g[0] = p[2] * ex((c[0] + dx * dx * c[1]) * -0.5);
    }
    // closed-form edge tangents: b = e (1, dx, dx^2), e = pdf(E; El, sigma), dx = E - El
    static constexpr int NB = 3;
    __device__ static void basis(const double* p, const double* c, double E, double, double& v, double* b) {
        const double dx = E - p[0];
        const double e = ex((c[0] + dx * dx * c[1]) * -0.5);
        v = p[2] * e;
        b[0] = e;
        b[1] = e * dx;
        b[2] = b[1] * dx;
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        d[0] = p[2] * c[1] * S[1];
        d[1] = p[2] * (c[1] * S[2] - S[0]) / p[1];
        d[2] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_OTTB> {  // OTTB.continuum, models/add.py:590-594
    static constexpr int NP = 2, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[0];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        g[0] = p[1] * ex((1.0 - E) * c[0]) * (1.0 / E);
    }
    // closed-form edge tangents: b0 = g / K, b1 = (g / K) (E - 1)
    static constexpr int NB = 2;
    __device__ static void basis(const double* p, const double* c, double E, double, double& v, double* b) {
        b[0] = ex((1.0 - E) * c[0]) * (1.0 / E);
        b[1] = b[0] * (E - 1.0);
        v = p[1] * b[0];
    }
    __device__ static void finish(const double* p, const double* c, const double* S, double* d) {
        d[0] = p[1] * c[0] * c[0] * S[1];
        d[1] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_OTTS> {  // OTTS.continuum, models/add.py:625-629
    static constexpr int NP = 2, NG = 1, NC = 1, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = lg(p[0]);
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE, TE lnE, T* g) {
        g[0] = p[1] * ex(-ex((lnE - c[0]) * (1.0 / 3.0)));
    }
};

template <>
struct Comp<ELISA_COMP_SMOOTHLYBROKENPL> {  // SmoothlyBrokenPL.continuum, models/add.py:846-861
    static constexpr int NP = 5, NG = 1, NC = 2, RULE = RULE_CONTINUUM_ADD;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = lg(p[1]);
        c[1] = (p[2] - p[0]) / p[3];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE, TE lnE, T* g) {
        const T le = lnE - c[0];  // log(E / Eb)
        const T x = c[1] * le;
        const double xv = value(x);
        const T alpha = sel(xv > 30.0, p[2], p[0]);
        T r;
        if (fabs(xv) > 30.0) {
            r = Cst<T>::make(0.5);
        } else {
            r = 0.5 * (1.0 + ex(x));
        }
        g[0] = p[4] * ex(-(alpha * le)) * ex(-(p[3] * lg(r)));
    }
};

// ------------------------------------------------------------------ additive, analytic
template <>
struct Comp<ELISA_COMP_POWERLAW> {  // PowerLaw.integral, models/add.py:40-50,655-657
    static constexpr int NP = 2, NG = 1, NC = 2, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = sel(value(p[0]) != 1.0, 1.0 - p[0], Cst<T>::make(1.0));
        c[1] = 1.0 / c[0];  // the division by (1 - alpha) once per parameter vector, not once per edge
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE, TE lnE, T* g) {
        if (value(p[0]) != 1.0)
            g[0] = pw_e(lnE, c[0]) * c[1];
        else
            g[0] = lift<T>(lnE);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T* p, const T*, const T* g0, const T* g1, TE, TE) {
        return p[1] * (g1[0] - g0[0]);
    }
    // closed-form edge tangents of the antiderivative: the bin is K (F(E1) - F(E0)), a DIFFERENCE of
    // edge values (EDGE_DIFF).  The edge weights of a difference form add up to zero, so any constant may be taken
    // off F: the contraction uses F~ = (E^(1-alpha) - 1) / (1 - alpha) = expm1((1 - alpha) ln E) / (1 - alpha), which
    // stays of order ln E near alpha = 1, where F itself is 1 / (1 - alpha) plus that -- summed over two thousand
    // edges with weights of both signs, F and F ln E lose (1 - alpha)^-2 digits of the gradient before
    // S0 / (1 - alpha) - S1 is formed (1.5e-4 at alpha = 1.00001 on the bench grid), F~ and its own derivative
    // dF~/dalpha = (F~ - E^(1-alpha) ln E) / (1 - alpha) do not.  b0 = F~, b1 = dF~/dalpha.
    static constexpr int NB = 2;
    static constexpr bool EDGE_DIFF = true;
    __device__ static void basis(const double* p, const double* c, double, double lnE, double& v, double* b) {
        const bool cond = p[0] != 1.0;
        const double x = c[0] * lnE;
        const double m = em1(x);  // E^(1-alpha) - 1
        b[0] = cond ? m * c[1] : lnE;
        b[1] = cond ? (b[0] - (m + 1.0) * lnE) * c[1] : 0.0;
        v = p[1] * b[0];
    }
    __device__ static void finish(const double* p, const double*, const double* S, double* d) {
        d[0] = p[0] != 1.0 ? p[1] * S[1] : 0.0;
        d[1] = S[0];
    }
};

template <>
struct Comp<ELISA_COMP_BROKENPL> {  // BrokenPL.integral, models/add.py:378-393
    // line 392 (`f.at[idx].set(...)`) discards its result: no break-bin fix-up, and the
    // integral is in E/Eb units (lines 385-387) -- reproduced as the code does it.
    static constexpr int NP = 4, NG = 2, NC = 3, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = lg(p[1]);
        c[1] = sel(value(p[0]) != 1.0, 1.0 - p[0], Cst<T>::make(1.0));
        c[2] = sel(value(p[2]) != 1.0, 1.0 - p[2], Cst<T>::make(1.0));
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE, TE lnE, T* g) {
        const T lx = lnE - c[0];  // log(E / Eb)
        g[0] = powerlaw_antiderivative_t(value(p[0]) != 1.0, c[1], lx);
        g[1] = powerlaw_antiderivative_t(value(p[2]) != 1.0, c[2], lx);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T* p, const T*, const T* g0, const T* g1, TE E0, TE) {
        const bool mask = value(E0) <= value(p[1]);
        return p[3] * sel(mask, g1[0] - g0[0], g1[1] - g0[1]);
    }
};

template <>
struct Comp<ELISA_COMP_LORENTZ> {  // Lorentz.integral, models/add.py:553-561
    static constexpr int NP = 3, NG = 1, NC = 2, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 2.0 / p[1];
        c[1] = 2.0 * p[2] / (ELISA_PI + 2.0 * at(2.0 * p[0] / p[1]));
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        g[0] = at((E - p[0]) * c[0]);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T*, const T* c, const T* g0, const T* g1, TE, TE) {
        return c[1] * (g1[0] - g0[0]);
    }
};

template <bool ENERGY>
struct CompPLFlux {  // PLFluxNorm.eval, models/add.py:682-708
    static constexpr int NP = 2, NG = 1, NC = 2, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T* p, const double* aux, T* c) {
        c[0] = sel(value(p[0]) != 1.0, 1.0 - p[0], Cst<T>::make(1.0));
        const double lmin = log(aux[0]), lmax = log(aux[1]);
        if (ENERGY) {
            const T a1 = p[0] - 1.0;
            const bool cond = value(a1) != 1.0;
            const T oma = sel(cond, 1.0 - a1, Cst<T>::make(1.0));
            const T f = powerlaw_antiderivative(cond, oma, lmax) - powerlaw_antiderivative(cond, oma, lmin);
            c[1] = p[1] / (f * 1.602176634e-9);
        } else {
            const bool cond = value(p[0]) != 1.0;
            const T f = powerlaw_antiderivative(cond, c[0], lmax) - powerlaw_antiderivative(cond, c[0], lmin);
            c[1] = p[1] / f;
        }
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE, TE lnE, T* g) {
        g[0] = powerlaw_antiderivative(value(p[0]) != 1.0, c[0], lnE);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T*, const T* c, const T* g0, const T* g1, TE, TE) {
        return c[1] * (g1[0] - g0[0]);
    }
};
template <>
struct Comp<ELISA_COMP_PLENFLUX> : CompPLFlux<true> {};
template <>
struct Comp<ELISA_COMP_PLPHFLUX> : CompPLFlux<false> {};

template <>
struct Comp<ELISA_COMP_PHOTONABS> {  // PhotonAbsorption.continuum, models/mul.py:360-388
    // the "ln E" slot of edge() carries sigma(E) of the leaf's table (leaf_bin_pc)
    static constexpr int NP = 1, NG = 1, NC = 1, RULE = RULE_CONTINUUM_MUL;
    static constexpr bool TABLE = true;
    template <typename T>
    __device__ static void pre(const T*, const double*, T*) {}
    template <typename T>
    __device__ static void edge(const T* p, const T*, double, double sigma, T* g) {
        g[0] = ex(-(p[0] * sigma));
    }
};

// ------------------------------------------------------------------ multiplicative
template <>
struct Comp<ELISA_COMP_CONSTANT> {  // Constant.integral, models/mul.py:54-56
    static constexpr int NP = 1, NG = 1, NC = 1, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T*, const double*, T*) {}
    template <typename T, typename TE>
    __device__ static void edge(const T*, const T*, TE, TE, T* g) {
        g[0] = Cst<T>::make(0.0);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T* p, const T*, const T*, const T*, TE, TE) {
        return p[0];
    }
};

template <>
struct Comp<ELISA_COMP_EDGE> {  // Edge.continuum, models/mul.py:88-94
    static constexpr int NP = 2, NG = 1, NC = 1, RULE = RULE_CONTINUUM_MUL;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[0];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        if (value(E) >= value(p[0])) {
            const T t = E * c[0];
            g[0] = ex(-(p[1] * (t * t * t)));
        } else {
            g[0] = Cst<T>::make(1.0);
        }
    }
};

template <>
struct Comp<ELISA_COMP_EXPABS> {  // ExpAbs.continuum, models/mul.py:116-118
    static constexpr int NP = 1, NG = 1, NC = 1, RULE = RULE_CONTINUUM_MUL;
    template <typename T>
    __device__ static void pre(const T*, const double*, T*) {}
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T*, TE E, TE, T* g) {
        g[0] = ex(p[0] * (-1.0 / E));
    }
};

template <>
struct Comp<ELISA_COMP_EXPFAC> {  // ExpFac.continuum, models/mul.py:154-159
    static constexpr int NP = 3, NG = 1, NC = 1, RULE = RULE_CONTINUUM_MUL;
    template <typename T>
    __device__ static void pre(const T*, const double*, T*) {}
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T*, TE E, TE, T* g) {
        if (value(E) >= value(p[2])) {
            g[0] = 1.0 + p[0] * ex(p[1] * (-E));
        } else {
            g[0] = Cst<T>::make(1.0);
        }
    }
};

template <>
struct Comp<ELISA_COMP_GABS> {  // GAbs.continuum, models/mul.py:195-204
    static constexpr int NP = 3, NG = 1, NC = 2, RULE = RULE_CONTINUUM_MUL;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[1];
        c[1] = -(p[2] / (2.50662827463100050242 * p[1]));  // -tau / (sqrt(2 pi) sigma)
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        const T z = (E - p[0]) * c[0];
        g[0] = ex(c[1] * ex(z * z * -0.5));
    }
};

template <>
struct Comp<ELISA_COMP_HIGHECUT> {  // HighECut.continuum, models/mul.py:236-240
    static constexpr int NP = 2, NG = 1, NC = 1, RULE = RULE_CONTINUUM_MUL;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 / p[1];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T* p, const T* c, TE E, TE, T* g) {
        if (value(E) >= value(p[0])) {
            g[0] = ex((p[0] - E) * c[0]);
        } else {
            g[0] = Cst<T>::make(1.0);
        }
    }
};

template <>
struct Comp<ELISA_COMP_PLABS> {  // PLAbs.integral, models/mul.py:266-271 (alpha = 1 ignored)
    static constexpr int NP = 2, NG = 1, NC = 2, RULE = RULE_CUSTOM;
    template <typename T>
    __device__ static void pre(const T* p, const double*, T* c) {
        c[0] = 1.0 - p[0];
        c[1] = p[1] / c[0];
    }
    template <typename T, typename TE>
    __device__ static void edge(const T*, const T* c, TE, TE lnE, T* g) {
        g[0] = c[1] * pw_e(lnE, c[0]);
    }
    template <typename T, typename TE>
    __device__ static T bin(const T*, const T*, const T* g0, const T* g1, TE E0, TE E1) {
        return (g1[0] - g0[0]) * (1.0 / (E1 - E0));
    }
};

// number of parameters per kind (host + device)
__host__ __device__ inline int comp_num_params(int kind) {
    switch (kind) {
        case ELISA_COMP_BAND: case ELISA_COMP_BANDEP: case ELISA_COMP_BROKENPL: return 4;
        case ELISA_COMP_BENTPL: case ELISA_COMP_COMPT: case ELISA_COMP_CUTOFFPL: case ELISA_COMP_GAUSS:
        case ELISA_COMP_LORENTZ: case ELISA_COMP_EXPFAC: case ELISA_COMP_GABS: return 3;
        case ELISA_COMP_BLACKBODY: case ELISA_COMP_BLACKBODYRAD: case ELISA_COMP_OTTB: case ELISA_COMP_OTTS:
        case ELISA_COMP_PLENFLUX: case ELISA_COMP_PLPHFLUX: case ELISA_COMP_POWERLAW: case ELISA_COMP_EDGE:
        case ELISA_COMP_HIGHECUT: case ELISA_COMP_PLABS: return 2;
        case ELISA_COMP_SMOOTHLYBROKENPL: return 5;
        case ELISA_COMP_CONSTANT: case ELISA_COMP_EXPABS: case ELISA_COMP_PHOTONABS: return 1;
        default: return -1;
    }
}
__host__ __device__ inline bool comp_is_additive(int kind) { return kind >= 0 && kind <= ELISA_COMP_SMOOTHLYBROKENPL; }

// ---------------------------------------------------------------------------------------
// Warp-cooperative leaf evaluation.  One warp owns one parameter vector; lane l evaluates
// the edge function on edge e (= first edge of the warp's pass + l) and receives the next
// edge's value from lane l + 1, so every edge is evaluated once per pass and lanes 0..30
// hold 31 bin values.  Simpson's midpoint (models/model.py:1748-1757) is evaluated by the
// lane itself.  The per-parameter-vector scalars (pre()) are computed once per warp and
// parked, flattened to doubles, in a shared-memory scratch slot.
struct GridView {
    const double* __restrict__ egrid;   // [E + 1]
    const double* __restrict__ lngrid;  // [E + 1]  ln(egrid)
    const double* __restrict__ emid;    // [E]      0.5 (E_i + E_{i+1})
    const double* __restrict__ lnmid;   // [E]      ln(emid)
    int E;
    // grid values of the calling lane's edge, loaded one pass ahead by the unfold kernel
    // (has_point != 0): E, ln E of edge min(e, E) and of the midpoint of bin min(e, E - 1)
    int has_point;
    double pE, pl, pEm, plm;
    // table components (ELISA_COMP_PHOTONABS): sigma at the edges [E + 1] / bin midpoints [E]
    // of leaf i in tab[2 i] / tab[2 i + 1]
    const double* tab[2 * ELISA_MAX_COMPONENTS];
};

template <class C, class = void>
struct IsTable {
    static constexpr bool value = false;
};
template <class C>
struct IsTable<C, decltype((void)C::TABLE)> {
    static constexpr bool value = true;
};

constexpr int LEAF_MAX_NC = 3;
// scratch doubles per leaf: parameter values, then NC flattened constants
constexpr int LEAF_SCRATCH = ELISA_MAX_COMP_PARAMS + LEAF_MAX_NC * (1 + ELISA_MAX_COMP_PARAMS);

template <typename T>
struct Flat;
template <>
struct Flat<double> {
    static constexpr int SIZE = 1;
    __device__ __forceinline__ static void put(double* s, double x) { s[0] = x; }
    __device__ __forceinline__ static double get(const double* s) { return s[0]; }
    __device__ __forceinline__ static double seed(double v, int) { return v; }
};
template <int N>
struct Flat<Dual<N>> {
    static constexpr int SIZE = 1 + N;
    __device__ __forceinline__ static void put(double* s, const Dual<N>& x) {
        s[0] = x.v;
#pragma unroll
        for (int i = 0; i < N; ++i) s[1 + i] = x.d[i];
    }
    __device__ __forceinline__ static Dual<N> get(const double* s) {
        Dual<N> r;
        r.v = s[0];
#pragma unroll
        for (int i = 0; i < N; ++i) r.d[i] = s[1 + i];
        return r;
    }
    __device__ __forceinline__ static Dual<N> seed(double v, int k) {
        Dual<N> r;
        r.v = v;
#pragma unroll
        for (int i = 0; i < N; ++i) r.d[i] = (i == k) ? 1.0 : 0.0;
        return r;
    }
};

// scratch[0 .. NP) must already hold the parameter values; writes the constants after them.
template <int KIND, typename T>
__device__ __forceinline__ void leaf_pre(double* scratch, const double* aux) {
    using C = Comp<KIND>;
    T p[C::NP], c[C::NC];
#pragma unroll
    for (int k = 0; k < C::NP; ++k) p[k] = Flat<T>::seed(scratch[k], k);
#pragma unroll
    for (int k = 0; k < C::NC; ++k) c[k] = Flat<T>::seed(0.0, -1);
    C::pre(p, aux, c);
#pragma unroll
    for (int k = 0; k < C::NC; ++k) Flat<T>::put(scratch + ELISA_MAX_COMP_PARAMS + k * Flat<T>::SIZE, c[k]);
}

// Bin value of leaf KIND for bin e (valid on lanes 0..30 of the pass, for e < E), given the
// parameters p and the hoisted constants c.
// gs / lgs: the leaf sees the grid scaled by gs (lgs = ln gs) -- energy shifts with a fixed
// factor (models/conv.py:294-432); the tables of table components are built on the scaled grid.
template <int KIND, typename T>
__device__ __forceinline__ T leaf_bin_pc(const T* p, const T* c, int method, const GridView& gv, int e, int leaf = 0,
                                         double gs = 1.0, double lgs = 0.0) {
    using C = Comp<KIND>;
    const int ec = e <= gv.E ? e : gv.E;  // lanes beyond the grid re-evaluate the last edge
    const double E0 = gs * (gv.has_point ? gv.pE : gv.egrid[ec]);
    double l0;
    if constexpr (IsTable<C>::value) {
        l0 = gv.tab[2 * leaf][ec];
    } else {
        l0 = gv.has_point ? gv.pl : gv.lngrid[ec];
        if (lgs != 0.0) l0 += lgs;
    }
    T g0[C::NG], g1[C::NG];
    C::edge(p, c, E0, l0, g0);
#pragma unroll
    for (int i = 0; i < C::NG; ++i) g1[i] = shfl_down1(g0[i]);
    const double E1 = __shfl_down_sync(0xffffffffu, E0, 1);
    if constexpr (C::RULE == RULE_CUSTOM) {
        return C::bin(p, c, g0, g1, E0, E1);
    } else {
        const bool add = C::RULE == RULE_CONTINUUM_ADD;
        if (method == ELISA_METHOD_SIMPSON) {
            const int em = e < gv.E ? e : gv.E - 1;
            T gm[C::NG];
            double lm;
            if constexpr (IsTable<C>::value) {
                lm = gv.tab[2 * leaf + 1][em];
            } else {
                lm = gv.has_point ? gv.plm : gv.lnmid[em];
                if (lgs != 0.0) lm += lgs;
            }
            C::edge(p, c, gs * (gv.has_point ? gv.pEm : gv.emid[em]), lm, gm);
            const double factor = add ? (E1 - E0) / 6.0 : 1.0 / 6.0;
            return factor * (g0[0] + 4.0 * gm[0] + g1[0]);
        }
        const double factor = add ? 0.5 * (E1 - E0) : 0.5;
        return factor * (g0[0] + g1[0]);
    }
}

// The same under FREE energy shifts (ZAShift / ZMShift / VAShift / VMShift with z or v a fit parameter,
// models/conv.py:294-432): the leaf sees E' = F gs E, where F -- the product of the enclosing factors 1 + z,
// 1 - v / c -- is of the leaf's dual type and carries the tangent of the shift parameter.  Every component body
// is generic in the type of its energy argument, so d(bin)/dz comes out of the same dual arithmetic that gives
// the parameter derivatives (the reference gets it from JAX differentiating model_fn(egrid * factor)).
// lnF = ln F.  T = double (values only) works as well.
template <int KIND, typename T>
__device__ __forceinline__ T leaf_bin_shift(const T* p, const T* c, int method, const GridView& gv, int e, double gs,
                                            double lgs, const T& F, const T& lnF) {
    using C = Comp<KIND>;
    static_assert(!IsTable<C>::value, "free energy shifts around table components are not supported");
    const int ec = e <= gv.E ? e : gv.E;
    const double E0d = gs * (gv.has_point ? gv.pE : gv.egrid[ec]);
    const double l0d = (gv.has_point ? gv.pl : gv.lngrid[ec]) + lgs;
    const double E1d = __shfl_down_sync(0xffffffffu, E0d, 1);
    const T E0 = F * E0d, E1 = F * E1d, l0 = lnF + l0d;
    T g0[C::NG], g1[C::NG];
    C::edge(p, c, E0, l0, g0);
#pragma unroll
    for (int i = 0; i < C::NG; ++i) g1[i] = shfl_down1(g0[i]);
    if constexpr (C::RULE == RULE_CUSTOM) {
        return C::bin(p, c, g0, g1, E0, E1);
    } else {
        const bool add = C::RULE == RULE_CONTINUUM_ADD;
        if (method == ELISA_METHOD_SIMPSON) {
            const int em = e < gv.E ? e : gv.E - 1;
            T gm[C::NG];
            const T Em = F * (gs * (gv.has_point ? gv.pEm : gv.emid[em]));
            const T lm = lnF + ((gv.has_point ? gv.plm : gv.lnmid[em]) + lgs);
            C::edge(p, c, Em, lm, gm);
            const T s = g0[0] + 4.0 * gm[0] + g1[0];
            if (add) return (E1 - E0) * (1.0 / 6.0) * s;
            return s * (1.0 / 6.0);
        }
        const T s = g0[0] + g1[0];
        if (add) return (E1 - E0) * 0.5 * s;
        return s * 0.5;
    }
}

// Same, with p and c decoded from the interpreter's shared-memory scratch slot.
template <int KIND, typename T>
__device__ __forceinline__ T leaf_bin(const double* scratch, int method, const GridView& gv, int e, int leaf = 0,
                                      double gs = 1.0, double lgs = 0.0) {
    using C = Comp<KIND>;
    T p[C::NP], c[C::NC];
#pragma unroll
    for (int k = 0; k < C::NP; ++k) p[k] = Flat<T>::seed(scratch[k], k);
#pragma unroll
    for (int k = 0; k < C::NC; ++k) c[k] = Flat<T>::get(scratch + ELISA_MAX_COMP_PARAMS + k * Flat<T>::SIZE);
    return leaf_bin_pc<KIND, T>(p, c, method, gv, e, leaf, gs, lgs);
}

}  // namespace elisa
