// Per-channel fit statistics and their derivative with respect to the folded source
// counts.  Restates infer/likelihood.py of the reference:
//   BetterNormal.log_prob  :136-140      BetterPoisson.log_prob (dense branch) :171-174
//   pgstat_background      :41-86        wstat_background                      :89-133
//   chi2 :184-227  cstat :230-272  pstat :275-321  pgstat :324-387  wstat :390-450
// The reference differentiates these with JAX autodiff; the closed forms below are the
// same derivatives written out, verified against torch autograd
// of the oracle in tests/.
#pragma once
#include "elisa_b200.h"
#include "fastmath.cuh"

namespace elisa {

struct ChannelData {
    double on;       // observed "on" data (net_counts for chi2, spec_counts otherwise)
    double off;      // observed background counts (pstat / pgstat / wstat)
    double sigma;    // net_errors (chi2) or back_errors (pgstat)
    double ratio;    // back_ratio
    double gof_on;   // xlogy(on, on) - on     (saturated-model offset, likelihood.py:172)
    double gof_off;  // xlogy(off, off) - off
    double den;      // stat_den(ratio): 1 / (2 a (a + 1)) for wstat, 1 / (2 a) for pgstat -- per channel, not per row
};

// per-channel constants of one spectrum as the kernels see them (arrays of C_pad doubles)
struct ChanArrays {
    const double* scale;  // area_scale * exposure
    const double* on;
    const double* off;
    const double* sigma;
    const double* ratio;
    const double* gof_on;
    const double* gof_off;
    const double* den;  // stat_den(ratio)
};

template <int STAT>
__device__ __forceinline__ double stat_den(double a) {
    if (STAT == ELISA_STAT_WSTAT) return 1.0 / (2.0 * a * (a + 1.0));
    if (STAT == ELISA_STAT_PGSTAT) return 1.0 / (2.0 * a);
    return 0.0;
}

// sqrt(q) and 1 / sqrt(q) together for normal positive q: one MUFU.RSQ64H seed and two Newton steps
// instead of a square root AND a division (the profile-background formulas need both,
// likelihood.py:76-86, 118-132, and their derivatives).  Both within 1 ulp.
__device__ __forceinline__ void sqrt_and_rsqrt(double q, double& root, double& inv) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q));
    const double h = 0.5 * q;
    y = fma(y, fma(-h * y, y, 0.5), y);  // y <- y (1.5 - h y^2)
    y = fma(y, fma(-h * y, y, 0.5), y);
    double r = q * y;
    r = fma(fma(-r, r, q), 0.5 * y, r);  // one correction of the root
    root = r;
    inv = y;
}

__device__ __forceinline__ double xlogy(double x, double y) { return x == 0.0 ? 0.0 : x * fm::log(y); }
__device__ __forceinline__ double poisson_gof(double n) { return xlogy(n, n) - n; }

// BetterPoisson.log_prob(n; mu) and d/dmu.
// FAST: n / mu as n * fm::rcp(mu) when mu lies in [1e-290, 1e290] (the clipped model counts always do).
template <bool GRAD, bool FAST = false>
__device__ __forceinline__ double poisson_ll(double n, double mu, double gof, double& dmu) {
    const double v = xlogy(n, mu) - mu - gof;
    if (GRAD) {
        double q;
        if (FAST && mu > 1e-290 && mu < 1e290)
            q = fma(n, fm::rcp(mu), -1.0);
        else
            q = n / mu - 1.0;
        dmu = (v > 0.0) ? 0.0 : (n == 0.0 ? -1.0 : q);
    }
    return v > 0.0 ? 0.0 : v;  // jnp.minimum(v, 0) semantics: NaN propagates (fmin would drop it)
}

// clip(x, 1e-30, 1e15) of likelihood.py:208 etc.; pass = d clip / d x
__device__ __forceinline__ double clip_counts(double x, double& pass) {
    pass = 1.0;
    if (!(x >= 1e-30 && x <= 1e15)) {  // rare: one predictable branch instead of min / max / NaN selects per channel
        pass = 0.0;
        x = x < 1e-30 ? 1e-30 : (x > 1e15 ? 1e15 : x);  // jnp.clip propagates NaN
    }
    return x;
}

// One channel of one statistic.  `x` = source_rate * exposure BEFORE the clip.
// Returns the channel log-likelihood (on + off terms); `dx` receives d loglike / d x;
// `m_on`, `m_off`, `l_on`, `l_off` the per-channel sites when SITES.
template <int STAT, bool GRAD, bool SITES = false, bool FAST = false>
__device__ __forceinline__ double stat_channel(double x, const ChannelData& d, double& dx, double* site = nullptr) {
    double pass, ll;
    if constexpr (STAT == ELISA_STAT_CHI2) {
        const double s = clip_counts(x, pass);
        const double inv = 1.0 / d.sigma;
        const double z = (d.on - s) * inv;
        ll = -0.5 * z * z;
        if (GRAD) dx = z * inv * pass;
        if (SITES) { site[0] = s; site[1] = ll; }
    } else if constexpr (STAT == ELISA_STAT_CSTAT) {
        const double s = clip_counts(x, pass);
        double dmu;
        ll = poisson_ll<GRAD, FAST>(d.on, s, d.gof_on, dmu);
        if (GRAD) dx = dmu * pass;
        if (SITES) { site[0] = s; site[1] = ll; }
    } else if constexpr (STAT == ELISA_STAT_PSTAT) {
        const double mu = clip_counts(x + d.ratio * d.off, pass);  // likelihood.py:301-302
        double dmu;
        ll = poisson_ll<GRAD, FAST>(d.on, mu, d.gof_on, dmu);
        if (GRAD) dx = dmu * pass;
        if (SITES) { site[0] = mu; site[1] = ll; }
    } else if constexpr (STAT == ELISA_STAT_PGSTAT) {
        const double s = clip_counts(x, pass);
        const double a = d.ratio, n = d.on;
        const double var = d.sigma * d.sigma;
        const double e = d.off - a * var;
        const double f = a * var * n + e * s;
        const double c = a * e - s;
        const double q = c * c + 4.0 * a * f;
        double dd, inv_dd;
        if (q > 1e-280 && q < 1e280) {
            sqrt_and_rsqrt(q, dd, inv_dd);
        } else {  // zero, negative (NaN as in the reference), denormal or huge: the plain operations
            dd = sqrt(q);
            inv_dd = 1.0 / dd;
        }
        double b, db = 0.0;
        if (e >= 0.0 || f >= 0.0) {
            if (n > 0.0) {
                b = (c + dd) * d.den;
                if (GRAD) db = (-1.0 + (-c + 2.0 * a * e) * inv_dd) * d.den;
            } else {
                b = e;
            }
        } else {
            b = 0.0;
        }
        const double mu = s + a * b;
        double dmu;
        const double l_on = poisson_ll<GRAD, FAST>(n, mu, d.gof_on, dmu);
        const double inv = 1.0 / d.sigma;
        const double z = (d.off - b) * inv;
        const double l_off = -0.5 * z * z;
        ll = l_on + l_off;
        if (GRAD) dx = (dmu * (1.0 + a * db) + z * inv * db) * pass;
        if (SITES) { site[0] = mu; site[1] = l_on; site[2] = b; site[3] = l_off; }
    } else {  // ELISA_STAT_WSTAT
        const double s = clip_counts(x, pass);
        const double a = d.ratio, n_on = d.on, n_off = d.off;
        double b, db = 0.0;
        if (n_on == 0.0) {
            b = n_off / (1.0 + a);
        } else if (n_off == 0.0) {
            if (s <= a / (a + 1.0) * n_on) {
                b = n_on / (1.0 + a) - s / a;
                if (GRAD) db = -1.0 / a;
            } else {
                b = 0.0;
            }
        } else {
            const double c = a * (n_on + n_off) - (a + 1.0) * s;
            const double q = c * c + 4.0 * a * (a + 1.0) * n_off * s;
            double dd, inv_dd;
            if (q > 1e-280 && q < 1e280) {
                sqrt_and_rsqrt(q, dd, inv_dd);
            } else {
                dd = sqrt(q);
                inv_dd = 1.0 / dd;
            }
            const double den = d.den;
            b = (c + dd) * den;
            if (GRAD) db = (-(a + 1.0) + (-(a + 1.0) * c + 2.0 * a * (a + 1.0) * n_off) * inv_dd) * den;
        }
        const double mu = s + a * b;
        double dmu, dmb;
        const double l_on = poisson_ll<GRAD, FAST>(n_on, mu, d.gof_on, dmu);
        const double l_off = poisson_ll<GRAD, FAST>(n_off, b, d.gof_off, dmb);
        ll = l_on + l_off;
        if (GRAD) dx = (dmu * (1.0 + a * db) + dmb * db) * pass;
        if (SITES) { site[0] = mu; site[1] = l_on; site[2] = b; site[3] = l_off; }
    }
    return ll;
}

}  // namespace elisa
