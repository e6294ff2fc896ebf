// Forward-mode dual numbers for the component derivatives.
//
// The reference gets d(unfold)/d(theta) from JAX autodiff of the component bodies
// (models/add.py, models/mul.py).  Here every component is written once, generic in
// its scalar type; instantiated with `double` it gives the photon-flux value, with
// `Dual<NP>` the value and its NP partial derivatives in one pass (transcendentals
// are evaluated once and shared by value and tangent).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <math.h>
#endif
#include "fastmath.cuh"

namespace elisa {

template <int N>
struct Dual {
    double v;
    double d[N];
};

template <typename T>
struct Scalar;  // traits
template <>
struct Scalar<double> {
    static constexpr int NP = 0;
};
template <int N>
struct Scalar<Dual<N>> {
    static constexpr int NP = N;
};

// ---- construction -------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T constant(double c);
template <>
__device__ __forceinline__ double constant<double>(double c) {
    return c;
}
template <int N>
__device__ __forceinline__ Dual<N> make_const(double c) {
    Dual<N> r;
    r.v = c;
#pragma unroll
    for (int i = 0; i < N; ++i) r.d[i] = 0.0;
    return r;
}

__device__ __forceinline__ double value(double x) { return x; }
template <int N>
__device__ __forceinline__ double value(const Dual<N>& x) {
    return x.v;
}

// ---- arithmetic ---------------------------------------------------------------------
#define ELISA_DUAL_UNARY_LOOP(expr)          \
    _Pragma("unroll") for (int i = 0; i < N; ++i) { expr; }

template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v + b.v;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = a.d[i] + b.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N>& a, double b) {
    Dual<N> r = a;
    r.v += b;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator+(double a, const Dual<N>& b) {
    return b + a;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a) {
    Dual<N> r;
    r.v = -a.v;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = -a.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v - b.v;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = a.d[i] - b.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N>& a, double b) {
    Dual<N> r = a;
    r.v -= b;
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator-(double a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a - b.v;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = -b.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    r.v = a.v * b.v;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = a.d[i] * b.v + a.v * b.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N>& a, double b) {
    Dual<N> r;
    r.v = a.v * b;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = a.d[i] * b);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator*(double a, const Dual<N>& b) {
    return b * a;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, const Dual<N>& b) {
    Dual<N> r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = (a.d[i] - r.v * b.d[i]) * inv);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N>& a, double b) {
    return a * (1.0 / b);
}
template <int N>
__device__ __forceinline__ Dual<N> operator/(double a, const Dual<N>& b) {
    Dual<N> r;
    const double inv = 1.0 / b.v;
    r.v = a * inv;
    const double s = -r.v * inv;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = s * b.d[i]);
    return r;
}

// ---- elementary functions (chain rule: r = f(a.v), r' = f'(a.v) a') -----------------
template <int N>
__device__ __forceinline__ Dual<N> chain(const Dual<N>& a, double f, double df) {
    Dual<N> r;
    r.v = f;
    ELISA_DUAL_UNARY_LOOP(r.d[i] = df * a.d[i]);
    return r;
}

// exp / log / expm1: the table-driven versions of fastmath.cuh (<= 1 ulp; 4 ulp for expm1)
__device__ __forceinline__ double ex(double a) { return fm::exp(a); }
template <int N>
__device__ __forceinline__ Dual<N> ex(const Dual<N>& a) {
    const double e = fm::exp(a.v);
    return chain(a, e, e);
}
__device__ __forceinline__ double lg(double a) { return fm::log(a); }
template <int N>
__device__ __forceinline__ Dual<N> lg(const Dual<N>& a) {
    return chain(a, fm::log(a.v), 1.0 / a.v);
}
__device__ __forceinline__ double sq(double a) { return sqrt(a); }
template <int N>
__device__ __forceinline__ Dual<N> sq(const Dual<N>& a) {
    const double s = sqrt(a.v);
    return chain(a, s, 0.5 / s);
}
__device__ __forceinline__ double at(double a) { return atan(a); }
template <int N>
__device__ __forceinline__ Dual<N> at(const Dual<N>& a) {
    return chain(a, atan(a.v), 1.0 / (1.0 + a.v * a.v));
}
// x / expm1(x) helper pieces
__device__ __forceinline__ double em1(double a) { return fm::expm1(a); }
template <int N>
__device__ __forceinline__ Dual<N> em1(const Dual<N>& a) {
    const double e = fm::expm1(a.v);
    return chain(a, e, e + 1.0);
}
// 1 / a for a normal a well inside the exponent range: fm::rcp (fastmath.cuh)
__device__ __forceinline__ double rinv(double a) { return fm::rcp(a); }
template <int N>
__device__ __forceinline__ Dual<N> rinv(const Dual<N>& a) {
    const double r = rinv(a.v);
    return chain(a, r, -(r * r));
}
// base^expo with a plain-double base whose log is known (energies): exp(expo * ln base)
__device__ __forceinline__ double pw_e(double lnbase, double expo) { return fm::exp(expo * lnbase); }
template <int N>
__device__ __forceinline__ Dual<N> pw_e(double lnbase, const Dual<N>& expo) {
    const double r = fm::exp(expo.v * lnbase);
    return chain(expo, r, r * lnbase);
}
template <int N>
__device__ __forceinline__ Dual<N> pw_e(const Dual<N>& lnbase, const Dual<N>& expo) {  // energies under a free shift
    return ex(expo * lnbase);
}
// general pow for positive base
__device__ __forceinline__ double pw(double base, double expo) { return pow(base, expo); }
template <int N>
__device__ __forceinline__ Dual<N> pw(const Dual<N>& base, const Dual<N>& expo) {
    Dual<N> r;
    r.v = pow(base.v, expo.v);
    const double db = expo.v * pow(base.v, expo.v - 1.0);  // jax: y * x^(y-1)
    const double de = (base.v == 0.0) ? 0.0 : r.v * log(base.v);
    ELISA_DUAL_UNARY_LOOP(r.d[i] = db * base.d[i] + de * expo.d[i]);
    return r;
}
template <int N>
__device__ __forceinline__ Dual<N> pw(const Dual<N>& base, double expo) {
    const double r = pow(base.v, expo);
    return chain(base, r, expo * pow(base.v, expo - 1.0));
}

// select (jnp.where): value and tangent of the taken branch
__device__ __forceinline__ double sel(bool c, double a, double b) { return c ? a : b; }
template <int N>
__device__ __forceinline__ Dual<N> sel(bool c, const Dual<N>& a, const Dual<N>& b) {
    return c ? a : b;
}

// ---- warp shuffle of a scalar ------------------------------------------------------
__device__ __forceinline__ double shfl_down1(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }
template <int N>
__device__ __forceinline__ Dual<N> shfl_down1(const Dual<N>& x) {
    Dual<N> r;
    r.v = __shfl_down_sync(0xffffffffu, x.v, 1);
    ELISA_DUAL_UNARY_LOOP(r.d[i] = __shfl_down_sync(0xffffffffu, x.d[i], 1));
    return r;
}

}  // namespace elisa
