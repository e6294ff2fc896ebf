// Device-side simulator: Poisson / Normal draws of the data sets the reference simulates on the host
// with NumPy (infer/helper.py:221-278: rng.poisson(model) for counting data, rng.normal(model, error)
// for Gaussian data) -- here a counter-based generator, one independent stream per (replica, channel),
// so that a draw depends on (seed, replica, channel) only: reproducible, independent of the launch
// shape, nothing stored between calls.  The NumPy stream itself cannot be reproduced on a GPU; parity
// is distributional (tests: moments, the exact pmf at small means, reproducibility per seed).
//   Philox4x32-10 (Salmon et al. 2011) -> 53-bit uniforms
//   Poisson: mean < 12 multiplication (Knuth); mean >= 12 transformed rejection PTRS (Hoermann 1993),
//            the algorithm NumPy itself uses for large means
//   Normal : Box-Muller
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elisa {

struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t out[4];
    int have;
    __device__ __forceinline__ Philox(uint64_t seed, uint64_t stream) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0;
        ctr[1] = 0;
        ctr[2] = (uint32_t)stream;
        ctr[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    __device__ __forceinline__ void round(uint32_t k0, uint32_t k1, uint32_t* c) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    __device__ __forceinline__ void refill() {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            round(k0, k1, c);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
        if (++ctr[0] == 0) ++ctr[1];
        have = 4;
    }
    // uniform in (0, 1): 53 random bits, never 0 or 1
    __device__ __forceinline__ double uniform() {
        if (have < 2) refill();
        const uint32_t a = out[have - 1], b = out[have - 2];
        have -= 2;
        const uint64_t bits = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);  // 53 bits
        return ((double)bits + 0.5) * 1.1102230246251565e-16;             // 2^-53
    }
};

__device__ __forceinline__ double sample_normal(Philox& g) {
    const double u = g.uniform(), v = g.uniform();
    return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
}

__device__ __forceinline__ double sample_poisson(Philox& g, double lam) {
    if (!(lam > 0.0)) return lam == 0.0 ? 0.0 : lam != lam ? lam : 0.0;  // NaN propagates, negative -> 0 (clipped upstream)
    if (lam < 12.0) {
        const double limit = exp(-lam);
        double prod = g.uniform();
        int k = 0;
        while (prod > limit) {
            prod *= g.uniform();
            ++k;
        }
        return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (int it = 0; it < 1000; ++it) {
        const double U = g.uniform() - 0.5, V = g.uniform();
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return floor(lam + 0.5);  // not reached in practice (acceptance > 0.9 per trial)
}

struct SampleArgs {
    const double* mean;   // [mean_rows, C] model counts ('{name}_Non_model' / '{name}_Noff_model')
    int64_t mean_rows;    // 1: every replica draws from the same row; else one row per replica
    const double* sigma;  // [C] or nullptr: Normal(mean, sigma) instead of Poisson(mean)
    int64_t n;            // replicas
    int32_t C;
    double* out;          // element [row * ld + col]
    int64_t ld;
    uint64_t seed;
    uint64_t stream0;     // first stream index of this (call, array): stream = stream0 + row * C + col
};

__global__ void __launch_bounds__(256) sample_kernel(const SampleArgs a) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n * a.C) return;
    const int64_t row = i / a.C;
    const int col = (int)(i % a.C);
    const double m = a.mean[(a.mean_rows == 1 ? 0 : row) * a.C + col];
    Philox g(a.seed, a.stream0 + (uint64_t)i);
    a.out[row * a.ld + col] = a.sigma != nullptr ? m + a.sigma[col] * sample_normal(g) : sample_poisson(g, m);
}

}  // namespace elisa
