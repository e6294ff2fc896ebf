// Digit slicing of FP64 rows for the sliced-integer fold (fold_i8.cuh): device functions shared
// by the slicing kernels and by the runtime-compiled unfold kernel (NVRTC: no host headers).
#pragma once
#include "elisa_b200.h"

namespace elisa {

constexpr int I8_BM = 128;        // rows of A per tile = TMEM lanes
constexpr int I8_BK = 128;        // contraction bytes per k-block = one 128-byte swizzle atom
constexpr int I8_TOP_DIGIT = 120; // |most significant digit| <= 120 (+1 carry)

// Scale factors of a row with largest magnitude amax: q = rint(x * f1 * f2), |q| <= TOP * 256^(KS-1).
// f1 * f2 = TOP * 256^(KS-1) / amax, split so that neither factor nor the intermediate leaves the
// normal range (amax may sit anywhere in [1e-300, 1e300]).  Returns the row scale
// (D = sc_a * sc_b * sum_o acc_o 256^-o); NaN for a row with a non-finite element.
template <int KS>
__device__ __forceinline__ double slice_scale(double amax, bool bad, double& f1, double& f2) {
    f1 = f2 = 0.0;
    if (bad) return __longlong_as_double(0x7ff8000000000000LL);
    if (!(amax > 0.0)) return 0.0;
    int ex;
    const double m = frexp(amax, &ex);  // amax = m 2^ex, m in [0.5, 1)
    const int h = ex / 2;
    f1 = ldexp(1.0, -h);
    f2 = ldexp((double)I8_TOP_DIGIT / m, 8 * (KS - 1) - (ex - h));
    return amax / (double)I8_TOP_DIGIT;
}

// Digits of one row (one warp): balanced base-256 digits of q = rint(x f1 f2), four consecutive
// k per lane packed into one 32-bit store per slice, k-blocks [kb0, kb1) into the images
// starting at img0.  kscale (optional): x[k] is multiplied by kscale[k] first.
template <int KS>
__device__ __forceinline__ void slice_row_digits(const double* x, bool live, bool aligned, int cols, double f1,
                                                 double f2, int rr, int tile_rows, int kb0, int kb1, int8_t* img,
                                                 int64_t img0, int lane, const double* kscale = nullptr) {
    const int sub = (rr >> 3) * 1024 + (rr & 7) * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + ((lane & 3) << 2);
    const size_t slice_bytes = (size_t)tile_rows * I8_BK;
    for (int kb = kb0; kb < kb1; ++kb) {
        const int k = kb * I8_BK + 4 * lane;
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        if (live && f2 != 0.0) {
            if (k + 3 < cols && aligned) {
                const double2 p0 = *reinterpret_cast<const double2*>(x + k);
                const double2 p1 = *reinterpret_cast<const double2*>(x + k + 2);
                v[0] = p0.x; v[1] = p0.y; v[2] = p1.x; v[3] = p1.y;
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (k + t < cols) v[t] = x[k + t];
            }
            if (kscale != nullptr) {  // balancing of the contraction index: exact powers of two;
                                      // the array is padded with ones to a multiple of 128 entries
                const double2 s0 = *reinterpret_cast<const double2*>(kscale + k);
                const double2 s1 = *reinterpret_cast<const double2*>(kscale + k + 2);
                v[0] *= s0.x; v[1] *= s0.y; v[2] *= s1.x; v[3] *= s1.y;
            }
        }
        // q + bias (bias = 0x80 in each of the KS digit bytes) has the plain base-256 digits
        // u_k = d_k + 128 of the balanced digits d_k; xor with the bias turns each byte back into
        // d_k as an int8.  Byte k of the four elements is then gathered with PRMT.
        constexpr unsigned long long bias = 0x8080808080808080ull >> (8 * (8 - KS));
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const unsigned long long u = ((unsigned long long)__double2ll_rn(v[t] * f1 * f2) + bias) ^ bias;
            lo[t] = (uint32_t)u;
            hi[t] = (uint32_t)(u >> 32);
        }
        uint32_t w[KS];  // w[i]: digit of significance KS-1-i (slice 0 = most significant)
#pragma unroll
        for (int i = 0; i < KS; ++i) {
            const int k = KS - 1 - i;  // byte index in u
            const uint32_t* src = k < 4 ? lo : hi;
            const uint32_t sel = 0x0040u + 0x0011u * (uint32_t)(k & 3);  // bytes k of (a, b) -> low half
            const uint32_t p01 = __byte_perm(src[0], src[1], sel);
            const uint32_t p23 = __byte_perm(src[2], src[3], sel);
            w[i] = __byte_perm(p01, p23, 0x5410);
        }
        int8_t* base = img + (size_t)(img0 + (kb - kb0)) * KS * slice_bytes + sub;
#pragma unroll
        for (int i = 0; i < KS; ++i) *reinterpret_cast<uint32_t*>(base + i * slice_bytes) = w[i];
    }
}

}  // namespace elisa
