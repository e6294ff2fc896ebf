// Sliced-integer response fold on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   D[M, N] = A[M, K] . B[N, K]^T      to FP64 accuracy, computed by int8 tcgen05.mma
//
// forward fold   : A = unfold U [samples, E],  B = R^T [C, E]   (infer/likelihood.py:205)
// transposed fold: A = W        [samples, C],  B = R   [E, C]   (its reverse-mode transpose)
//
// tcgen05.mma has no FP64 kind and DMMA tops out at the FP64 pipe rate (~37 TFLOP/s,
// gemm_f64.cuh).  This path gets FP64-class accuracy from the int8 pipe instead: every
// operand row x[k] is written as   x[k] = s_row * sum_{i<KS} d_i[k] * 256^(KS-1-i),
// d_i in [-128, 127] (balanced base-256 digits of q = rint(x / s_row), |q| < 2^(8 KS - 1)),
// so that
//       sum_k x[k] y[k] = s_x s_y * sum_{i,j} 256^(2KS-2-i-j) * (d_i . e_j)
// with every digit product d_i . e_j an EXACT int32 dot product of the tensor core.
// Products of order i + j > KS - 1 lie below the digits already dropped and are skipped:
// KS (KS + 1) / 2 int8 GEMMs replace one FP64 GEMM (21 for KS = 6).  The error is bounded
// NORMWISE like any floating-point GEMM's,   |dD[m, n]| <~ 2^(-8 KS + 3) * max_k|A[m, k]| *
// sum_k|B[n, k]| + (A <-> B),  i.e. ~2^-45 for KS = 6; tests/ hold the path to the
// north-star tolerance (1e-10 on statistic and gradient) on every BASELINE configuration.
//
// Layout.  Both operands live in HBM as ready-made shared-memory images: per (row tile,
// k-block of 128 bytes) one image of KS slices, each slice [rows, 128 B] in the canonical
// K-major SWIZZLE_128B layout of tcgen05 (8-row atoms of 1024 B, 16-byte chunk index
// XOR-ed with the row index mod 8).  A tile load is therefore ONE linear bulk copy
// (cp.async.bulk, the TMA engine in 1-D mode) -- no tensor map, no per-thread addressing.
// The digits of B are produced once per plan; those of A by slice_rows_kernel per chunk.
//
// Kernel.  One persistent CTA per SM, 19 warps:
//   warps 0-15 epilogue: tcgen05.ld (16x256b: four threads share 64 contiguous bytes of a
//              row) the KS int32 accumulators of their 32 rows x 16 columns, recombine in FP64
//              (Horner in base 256), hand TMEM back, then run the epilogue (plain store of
//              R u, or the contraction with dU/dtheta) while the next tile accumulates;
//   warp  16   producer: bulk copies into a ring of A slices (16 KB) and a ring of B
//              blocks (KS x 8 KB), each with full/empty mbarriers;
//   warps 17-18 MMA issuers (slices i even / odd): for slice i of A one tcgen05.mma of N = (KS - i) * 64 columns
//              against the STACKED B slices 0 .. KS-1-i, landing on the TMEM accumulators
//              of orders i .. KS-1 (column offset 64 i): A is read once per slice and the
//              instruction shapes stay wide (N up to 256).
// TMEM: KS accumulators of 128 lanes x 64 columns (384 of 512 columns for KS = 6).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "slice_i8.cuh"
#include "stats.cuh"

namespace elisa {

constexpr int I8_BN = 64;        // output columns per tile (one accumulator per order)
constexpr int I8_A_SLICE = I8_BM * I8_BK;  // 16 KB
constexpr int I8_B_SLICE = I8_BN * I8_BK;  // 8 KB
constexpr int I8_NB = 2;         // B-block ring depth
constexpr int I8_EPI_WARPS = 16;               // 4 per TMEM lane quarter (8 also works: 32 columns per thread)
constexpr int I8_CG = I8_EPI_WARPS / 4;        // column groups per tile
constexpr int I8_TC = I8_BN / I8_CG;           // columns per epilogue thread
constexpr int I8_ISSUERS = 2;                  // MMA issuer warps (slices i = w mod 2)
constexpr int I8_THREADS = (I8_EPI_WARPS + 1 + I8_ISSUERS) * 32;
constexpr int I8_SMEM_BUDGET = 212 * 1024;  // rings; + 1 KB alignment slack + barriers below 227 KB
constexpr int I8_MAX_K = 16384;             // int32 headroom: KS * K * 2^14 < 2^31 for KS <= 7

__host__ __device__ constexpr int i8_na(int ks) { return (I8_SMEM_BUDGET - I8_NB * ks * I8_B_SLICE) / I8_A_SLICE; }
__host__ __device__ constexpr int i8_smem_bytes(int ks) {
    return 1024 + I8_NB * ks * I8_B_SLICE + i8_na(ks) * I8_A_SLICE + 256;
}

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol error traps (launch failure) instead of hanging the device.
// try_wait carries a suspend-time hint: the warp sleeps in hardware until the phase completes
// (or the hint expires) instead of spinning -- in the first version more than half of all
// instructions the kernel executed were these polls, stealing issue slots and power from the
// epilogue warps sharing the scheduler.
// `what` names the barrier for the post-mortem record (host-mapped memory, FoldI8Args::trap_info):
// 1 a_full 2 a_empty 3 b_full 4 b_empty 5 t_full 6 t_empty 7 t_init, + 16 * ring slot
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* trap_info = nullptr, int what = 0,
                                          int tile = 0) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 12); ++spin) {  // ~4 s of 1 ms sleeps
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(1000000u)
            : "memory");
        if (done) return;
        if (spin == 1024 && trap_info != nullptr && (threadIdx.x & 31) == 0) {
            // a wait this long is a stall: every stuck warp of one block leaves what it waits for
            const int me = (int)blockIdx.x + 1;
            const int owner = atomicCAS(trap_info + 6, 0, me);
            if (owner == 0 || owner == me) {
                int32_t* st = trap_info + 8 + 3 * (threadIdx.x >> 5);
                st[0] = what;
                st[1] = (int32_t)parity;
                st[2] = tile;
                __threadfence_system();
            }
        }
    }
    if (trap_info != nullptr && (threadIdx.x & 31) == 0 && atomicCAS(trap_info, 0, 1) == 0) {
        trap_info[1] = (int32_t)blockIdx.x;
        trap_info[2] = (int32_t)(threadIdx.x >> 5);
        trap_info[3] = what;
        trap_info[4] = (int32_t)parity;
        trap_info[5] = tile;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 8 consecutive 32-bit columns: thread = lane (row), v[j] = column j
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset (between 8-row atoms)
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}
// instruction descriptor: s8 x s8 -> s32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_i8(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(I8_BM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------
// slicing: FP64 rows -> digit images + row scales
// ---------------------------------------------------------------------------------------
struct SliceArgs {
    const double* X;   // [rows, ld] row-major
    int64_t ld;
    int32_t rows;      // rows holding data (further rows of the last tile are zero-filled)
    int32_t cols;      // columns holding data (k >= cols is zero)
    int32_t row_tiles; // tiles of `tile_rows` rows written
    int32_t tile_rows; // 128 (A operand) or 64 (B operand)
    int32_t kb_total;  // k-blocks per row tile when kb_lo == nullptr
    const int64_t* tile_img;  // [row_tiles] first image of the tile (nullptr: tile * kb_total)
    const int32_t* kb_lo;     // [row_tiles] stored k-block range (nullptr: [0, kb_total))
    const int32_t* kb_hi;
    int8_t* img;       // images of KS * tile_rows * 128 bytes
    double* scale;     // [row_tiles * tile_rows]  x ~= scale * 256^-(KS-1) * q  (see kernel)
    // balancing of the contraction index (powers of two, see col_envelope_kernel): the digits are
    // those of x[k] * kscale[k]; nullptr = 1.  abs_sum (optional): sum_k |x[k] kscale[k]| per row.
    const double* kscale;
    double* abs_sum;
};

// One warp per row.  Pass 1: amax (and a non-finite flag).  Pass 2: q = rint(x * f) as
// int64, balanced digits from the low end (d = sign-extended low byte, q = (q - d) >> 8),
// four consecutive k per lane packed into one 32-bit store per slice.
// One warp per row.  Pass 1: amax (and a non-finite flag).  Pass 2: the digits.
template <int KS>
__global__ void __launch_bounds__(256) slice_rows_kernel(const SliceArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.row_tiles * a.tile_rows) return;
    const int tile = row / a.tile_rows, rr = row % a.tile_rows;
    const int kb0 = a.kb_lo ? a.kb_lo[tile] : 0;
    const int kb1 = a.kb_hi ? a.kb_hi[tile] : a.kb_total;
    const int64_t img0 = a.tile_img ? a.tile_img[tile] : (int64_t)tile * a.kb_total;
    const bool live = row < a.rows;
    const double* x = a.X + (size_t)row * a.ld;

    double amax = 0.0, asum = 0.0;
    bool bad = false;
    if (live) {
        for (int k = 2 * lane; k < a.cols; k += 64) {
            double v0 = x[k], v1 = (k + 1 < a.cols) ? x[k + 1] : 0.0;
            if (a.kscale != nullptr) {  // k is even and the array is padded with ones: one 16-byte load
                const double2 s = *reinterpret_cast<const double2*>(a.kscale + k);
                v0 *= s.x;
                v1 *= s.y;
            }
            bad |= !(fabs(v0) <= 1.79e308) | !(fabs(v1) <= 1.79e308);
            amax = fmax(amax, fmax(fabs(v0), fabs(v1)));
            asum += fabs(v0) + fabs(v1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmax(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        asum += __shfl_xor_sync(0xffffffffu, asum, o);
    }
    bad = __any_sync(0xffffffffu, bad);
    // the same row maximum as the kernels that slice their own output (unfold_generic, stat_slice_kernel):
    // the upper end of the interval of doubles sharing amax's high word -- a row must get the same digits
    // whichever kernel slices it (results are bit-identical under re-chunking, which changes that)
    if (amax > 0.0 && !bad) amax = __hiloint2double(__double2hiint(amax), (int)0xffffffff);
    double f1, f2;
    const double sc = slice_scale<KS>(amax, bad, f1, f2);
    if (lane == 0) {
        a.scale[row] = sc;
        if (a.abs_sum != nullptr) a.abs_sum[row] = live ? asum : 0.0;
    }
    slice_row_digits<KS>(x, live, (a.ld & 1) == 0, a.cols, f1, f2, rr, a.tile_rows, kb0, kb1, a.img, img0, lane,
                         a.kscale);
}

// Column envelopes for balancing the contraction index of a fold.  The engine's error is
// relative to (largest element of the A row) x (sum over the B row); when the operand's columns
// differ by orders of magnitude (a steep photon spectrum; a few starved channels in W) the small
// columns lose digits.  Dividing column k of A by t_k and multiplying column k of B by t_k leaves
// the product unchanged -- exactly, t_k being a power of two -- and makes every column of A of
// order one.  t_k = 2^round(log2(max over the pilot rows and over columns k - w .. k + w of
// |X|)): an UPPER envelope (a too-small t_k would let one column tower over its row; the window
// bridges accidental zeros such as sign changes of a residual), 1 where the envelope is zero or
// not finite.  One thread per column.
// The envelope is floored at 2^-I8_BALANCE_BITS of its largest column: columns the pilot saw as
// negligible (an exponential cut-off far above its folding energy: 1e-300 against 1) stay
// negligible for every row instead of being blown up to order one, where a row with a slightly
// larger cut-off energy would tower over its own significant columns.  28 bits: columns within
// 4e-9 of the peak keep full relative precision, a count fed only by columns at 1e-11 of the peak
// still keeps 1e-12; a deeper floor buys nothing physical and makes tails that vary by many
// orders between rows (exp(-E/Ec) at E = 20 Ec) dangerous.  Two passes: `peak` (one double,
// zeroed by the caller) receives the largest column maximum first.
constexpr int I8_BALANCE_BITS = 28;

__global__ void __launch_bounds__(128) col_peak_kernel(const double* X, int64_t ld, int rows, int cols,
                                                       unsigned long long* peak) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    double m = 0.0;
    if (k < cols)
        for (int r = 0; r < rows; ++r) {
            const double v = fabs(X[(size_t)r * ld + k]);
            if (v <= 1.79e308) m = fmax(m, v);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(peak, (unsigned long long)__double_as_longlong(m));
}

__global__ void __launch_bounds__(128) col_envelope_kernel(const double* X, int64_t ld, int rows, int cols, int cols_pad,
                                                           int window, const unsigned long long* peak, double* scale,
                                                           double* inv, uint8_t* inv_exp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= cols_pad) return;
    double m = 0.0;
    if (k < cols) {
        const int lo = max(k - window, 0), hi = min(k + window, cols - 1);
        for (int r = 0; r < rows; ++r)
            for (int j = lo; j <= hi; ++j) {
                const double v = fabs(X[(size_t)r * ld + j]);
                if (v <= 1.79e308) m = fmax(m, v);
            }
    }
    // Only the ratios between columns matter (the row scales absorb any common factor), so the
    // envelope is expressed relative to its peak: t in [2^-I8_BALANCE_BITS, 1], 1 / t = 2^inv_exp.
    const double top = __longlong_as_double((long long)*peak);  // bits of a non-negative double order like integers
    int down = 0;
    if (top > 0.0 && k < cols) {
        m = fmax(m, ldexp(top, -I8_BALANCE_BITS));
        int ex, ex_top;
        const double f = frexp(m, &ex);  // m = f 2^ex, f in [0.5, 1)
        const double f_top = frexp(top, &ex_top);
        ex = (f > 0.70710678118654752 ? ex : ex - 1);
        ex_top = (f_top > 0.70710678118654752 ? ex_top : ex_top - 1);
        down = max(0, min(I8_BALANCE_BITS, ex_top - ex));
    }
    scale[k] = ldexp(1.0, -down);
    inv[k] = ldexp(1.0, down);
    inv_exp[k] = (uint8_t)down;
}

// Leading all-zero slices of every image of a static operand: lead[img] in [0, KS].  One block
// per image.
template <int KS>
__global__ void __launch_bounds__(256) lead_zero_kernel(const int8_t* img, int64_t n_images, int slice_bytes, uint8_t* lead) {
    const int64_t im = blockIdx.x;
    if (im >= n_images) return;
    const uint4* base = reinterpret_cast<const uint4*>(img + (size_t)im * KS * slice_bytes);
    const int per_slice = slice_bytes / 16;
    int z = 0;
    for (; z < KS; ++z) {
        unsigned any = 0;
        for (int t = threadIdx.x; t < per_slice; t += blockDim.x) {
            const uint4 v = base[(size_t)z * per_slice + t];
            any |= v.x | v.y | v.z | v.w;
        }
        if (__syncthreads_or(any != 0)) break;
    }
    if (threadIdx.x == 0) lead[im] = (uint8_t)z;
}

// ---------------------------------------------------------------------------------------
// the fold kernel
// ---------------------------------------------------------------------------------------
struct FoldI8Args {
    const int8_t* A_img;        // image (mt, kb) at ((mt * kb_a + kb) * KS) * 16 KB
    const double* a_scale;      // [m_tiles * 128]
    int32_t kb_a;               // k-blocks stored per row tile of A
    const int8_t* B_img;        // image tile_img[nt] + (kb - kb_lo[nt]), KS * 8 KB each
    const double* b_scale;      // [n_tiles * 64]
    const int64_t* b_tile_img;  // [n_tiles]
    const int32_t* kb_lo;       // [n_tiles] k-block range of column tile nt
    const int32_t* kb_hi;
    // Leading all-zero digit slices of each stored B image (lead_zero_kernel), same indexing as the
    // images; nullptr: none.  A response's redistribution core towers over its tails by 4-6 decades,
    // so outside the core the most significant one or two digit planes of a (64-row, 128-byte)
    // block are zero throughout: the products with those planes -- 6 + 5 of the 21 for two leading
    // zero slices -- and the A slices only they would meet are neither loaded nor issued.  Exact:
    // only multiplications by zero digits are skipped, results are bit-identical.
    const uint8_t* b_lead;
    int32_t m_tiles, n_tiles;
    int32_t n_groups;  // ceil(n_tiles / I8_GROUP)
    int32_t* trap_info;  // host-mapped int32[128] or nullptr: which waits stalled (see mbar_wait)
    int32_t debug;  // experiments only: 1 = no A copies, 2 = no B copies, 4 = no MMAs, 8 = no epilogue, 16 = no TMEM drain, 256 = every epilogue warp polls t_full itself (as before round 2)
};

// Exact int32 -> double without the quarter-rate I2F.F64: the bits of 2^52 + (v + 2^31) are
// assembled with one XOR, and one full-rate DADD removes the offset.
__device__ __forceinline__ double i32_to_f64(int32_t v) {
    return __hiloint2double(0x43300000, v ^ (int32_t)0x80000000) - 4503601774854144.0;  // 2^52 + 2^31
}

// Two neighbouring orders are first merged in integer arithmetic, p = 256 acc_2j + acc_2j+1
// (one IMAD.WIDE; |p| < 2^40 under the int32 headroom rule), and the int64 is converted by the
// same trick: bits(1.5 * 2^52) + p is the double 1.5 * 2^52 + p.  The _m8 variant uses the
// exponent of 2^44, whose unit in the last place is 2^-8, and returns p / 256.  Halves the FP64
// instructions of the recombination (the gradient epilogue stalls on the FP64 pipe).
__device__ __forceinline__ double i64_to_f64(long long p) {
    return __longlong_as_double(0x4338000000000000LL + p) - 6755399441055744.0;  // 1.5 * 2^52
}
__device__ __forceinline__ double i64_to_f64_m8(long long p) {
    return __longlong_as_double(0x42B8000000000000LL + p) - 26388279066624.0;  // 1.5 * 2^44
}

// x = sum_o acc_o 256^-o for one element, acc(o) = accumulator of order o
template <int KS, class Acc>
__device__ __forceinline__ double i8_recombine(const Acc& acc) {
    static_assert(KS >= 4, "at least two pairs of orders");
    constexpr int NP = KS / 2;  // pairs (2j, 2j + 1); an odd KS leaves order KS - 1 on its own
    auto pair = [&](int j) { return (long long)acc(2 * j) * 256 + (long long)acc(2 * j + 1); };
    // T = sum_{j >= 1} 65536^-(j-1) p_j (+ the odd order), Horner from the least significant end
    double T;
    if (KS & 1) {
        T = fma(i32_to_f64(acc(KS - 1)), 0.00390625, i64_to_f64(pair(NP - 1)));
    } else {
        T = i64_to_f64(pair(NP - 1));
    }
#pragma unroll
    for (int j = NP - 2; j >= 1; --j) T = fma(T, 1.52587890625e-05, i64_to_f64(pair(j)));  // 2^-16
    return fma(T, 5.9604644775390625e-08, i64_to_f64_m8(pair(0)));  // 2^-24 T + p_0 / 256
}

// Epilogue register layout ("quad layout").  A warp owns 32 rows (its TMEM lane quarter) x 16
// columns (its column group) of the tile and reads them with tcgen05.ld.16x256b.x2: thread
// (r = lane / 4, q = lane % 4) receives rows r + 8 k (k = 0..3) and columns
// {2q, 2q + 1, 8 + 2q, 9 + 2q}: the four threads of a quad cover 64 contiguous bytes of an
// FP64 row, so the epilogue's global loads and stores touch 8 rows per warp instruction with
// every 32-byte sector used in full (a thread-per-row layout touches 32 lines for 16 bytes each).
// x[k][m]: row r + 8 k, column i8_col(m, q) (a second 16-column block follows for I8_TC = 32).
__device__ __forceinline__ void tc_ld_16x256b_x2(uint32_t taddr, int32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

constexpr int I8_NM = I8_TC / 4;  // elements per row and thread
// column (within the warp's group) of element m of thread q: 16-column blocks of the layout above
__host__ __device__ constexpr int i8_col(int m, int q) { return 16 * (m >> 2) + 8 * ((m >> 1) & 1) + 2 * q + (m & 1); }

// TMEM -> FP64: x = sum_o acc_o * 256^-o (Horner from the least significant order)
template <int KS>
__device__ __forceinline__ void i8_combine_quad(uint32_t taddr, double (&x)[4][I8_NM]) {
#pragma unroll
    for (int b = 0; b < I8_TC / 16; ++b) {  // 16-column blocks of the warp's group
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // TMEM lanes [16 h, 16 h + 16) of the warp's quarter
            int32_t v[KS][8];
#pragma unroll
            for (int d = 0; d < KS; ++d)
                tc_ld_16x256b_x2(taddr + ((uint32_t)(16 * h) << 16) + d * I8_BN + 16 * b, v[d]);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                // registers: {0,1} row r cols 2q,2q+1 | {2,3} row r+8 | {4,5} row r cols 8+2q.. | {6,7} row r+8
                x[2 * h + ((e >> 1) & 1)][4 * b + (e & 1) + 2 * (e >> 2)] =
                    i8_recombine<KS>([&](int d) { return v[d][e]; });
            }
        }
    }
}

// Work decomposition.  The columns of the output are cut into GROUPS of I8_GROUP tiles (fixed,
// so that the summation tree of the fused reductions never depends on the batch size); a work
// ITEM is one row tile x one group, items are numbered row-tile major and dealt round-robin
// to the persistent CTAs: neighbouring CTAs work on the same row tile at the same time, so
// its A images are fetched from HBM once and shared through L2 (a whole-row sweep per CTA
// would put 148 different 1.5 MB A sets in flight and thrash the 126 MB L2).
//
// Epilogue interface (quad layout above; row0 = first row of the thread, c0 = first column of
// the warp's group in the tile's matrix, q = lane % 4):
//   State st; epi.begin(st)                      once per group
//   epi.prefetch(row_acc, c0)                    before the tile's accumulators are awaited
//   epi.tile(st, x, row0, c0, q, sb)             x[k][m] recombined (unscaled) dot products,
//                                                sb = column scales of the warp's 16 columns
//   epi.end(st, group, cg, row_acc, sa)          once per group: the thread's row partial
constexpr int I8_GROUP = 4;

// the row (k of row0 + 8 k) whose partial thread q of a quad keeps after the transposing butterfly
// of the fused reductions: q = 0, 1, 2, 3 -> k = 0, 2, 1, 3
__device__ __forceinline__ int i8_quad_row(int q) { return 2 * (q & 1) + (q >> 1); }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// item -> (row tile, first column tile, one past the last column tile)
struct I8Item {
    int mt, g;
};
__device__ __forceinline__ I8Item i8_item(const FoldI8Args& a, int item) {
    I8Item it;
    it.mt = item / a.n_groups;
    it.g = item % a.n_groups;
    return it;
}
#define I8_FOR_EACH_TILE(it, a, nt) \
    for (int nt = (it).g * I8_GROUP, ne_ = min(nt + I8_GROUP, (a).n_tiles); nt < ne_; ++nt)

// Registers per thread of the fold kernel.  One persistent CTA per SM owns the shared memory whatever it
// is compiled to; what its REGISTER footprint decides is whether blocks of the FP64-pipe kernels of the
// neighbouring chunk (component evaluation, statistic, contraction: no shared memory to speak of) can be
// resident next to it.  -DI8_MAXNREG=n caps it ("Two chunks in flight").
#ifdef I8_MAXNREG
#define I8_KERNEL_BOUNDS __maxnreg__(I8_MAXNREG)  // (nvcc refuses __launch_bounds__ next to it)
#else
#define I8_KERNEL_BOUNDS __launch_bounds__(I8_THREADS, 1)
#endif

template <int V>
struct I8Int {
    static constexpr int value = V;
};

template <int KS, class Epi, int NI = Epi::ISSUERS>
__global__ void I8_KERNEL_BOUNDS
fold_i8_kernel(const __grid_constant__ FoldI8Args a, const __grid_constant__ Epi epi) {
    constexpr int NA = i8_na(KS);
    constexpr int B_BLOCK = KS * I8_B_SLICE;
    static_assert(NI >= 1 && NI <= I8_ISSUERS, "issuer warps");  // MMA issuer warps in use
    extern __shared__ uint8_t smem_raw[];
    const uint32_t s_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_B = s_base;
    const uint32_t s_A = s_base + I8_NB * B_BLOCK;
    const uint32_t s_bar = s_A + NA * I8_A_SLICE;
    // barriers: a_full[NA] a_empty[NA] b_full[NB] b_empty[NB] t_full t_empty ; then the TMEM slot
    auto a_full = [&](int s) { return s_bar + 8u * s; };
    auto a_empty = [&](int s) { return s_bar + 8u * (NA + s); };
    auto b_full = [&](int s) { return s_bar + 8u * (2 * NA + s); };
    auto b_empty = [&](int s) { return s_bar + 8u * (2 * NA + I8_NB + s); };
    // ring r (of NI) owns slots [ring_base(r), ring_base(r) + ring_size(r)) of the NA
    auto ring_size = [](int r) { return NA / NI + (r < NA % NI ? 1 : 0); };
    auto ring_base = [](int r) { return r * (NA / NI) + (r < NA % NI ? r : NA % NI); };
    const uint32_t t_full = s_bar + 8u * (2 * NA + 2 * I8_NB);
    const uint32_t t_empty = t_full + 8u;
    const uint32_t t_init = t_empty + 8u;  // accumulators of the tile initialised (issuer 0 -> the others)
    const uint32_t s_slot = t_init + 8u;
    uint8_t* gen_base = smem_raw + (s_base - smem_u32(smem_raw));
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + (s_slot - s_base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == I8_EPI_WARPS && lane == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < I8_NB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), NI); }
        mbar_init(t_full, NI);
        mbar_init(t_empty, I8_EPI_WARPS);
        mbar_init(t_init, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == I8_EPI_WARPS + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_slot), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot_ptr;

    const int n_items = a.m_tiles * a.n_groups;

    if (warp == I8_EPI_WARPS) {
        // ================= producer =================
        // One ring of A slots PER ISSUER (slices i = r mod NI go to ring r): a single shared ring
        // would let an issuer test a slot's full barrier a whole phase early -- bulk copies
        // complete out of order, so the other issuer's slice of the previous round can still be in
        // flight when the phase parity is read -- and multiply stale data.
        int as[NI], bs = 0;
        uint32_t aph[NI], bph = 0;
#pragma unroll
        for (int r = 0; r < NI; ++r) { as[r] = 0; aph[r] = 0; }
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const I8Item it = i8_item(a, item);
            I8_FOR_EACH_TILE(it, a, nt) {
                const int lo = a.kb_lo[nt], hi = a.kb_hi[nt];
                const int8_t* bsrc = a.B_img + (size_t)a.b_tile_img[nt] * B_BLOCK;
                const int8_t* asrc = a.A_img + ((size_t)it.mt * a.kb_a + lo) * KS * I8_A_SLICE;
                for (int kb = lo; kb < hi; ++kb) {
                    // leading zero slices of this B block; the tile's first block is taken whole (its
                    // slice-0 products initialise every accumulator)
                    const int z = (a.b_lead != nullptr && kb != lo) ? min((int)a.b_lead[a.b_tile_img[nt] + (kb - lo)], KS) : 0;
                    mbar_wait(b_empty(bs), bph ^ 1u, a.trap_info, 4 + 16 * bs, nt);
                    if (lane == 0) {
                        if ((a.debug & 2) || z == KS) {
                            mbar_arrive(b_full(bs));
                        } else {
                            const uint32_t bytes = (uint32_t)(KS - z) * I8_B_SLICE;
                            mbar_expect_tx(b_full(bs), bytes);
                            bulk_g2s(s_B + bs * B_BLOCK + z * I8_B_SLICE, bsrc + z * I8_B_SLICE, bytes, b_full(bs));
                        }
                    }
                    bsrc += B_BLOCK;
                    if (++bs == I8_NB) { bs = 0; bph ^= 1u; }
#pragma unroll
                    for (int i = 0; i < KS; ++i) {
                        if (i + z > KS - 1) {  // this A slice would only meet zero digits
                            asrc += I8_A_SLICE;
                            continue;
                        }
                        const int r = i % NI;
                        const int slot = ring_base(r) + as[r];
                        mbar_wait(a_empty(slot), aph[r] ^ 1u, a.trap_info, 2 + 16 * slot, nt);
                        if (lane == 0) {
                            if (a.debug & 1) {
                                mbar_arrive(a_full(slot));
                            } else {
                                mbar_expect_tx(a_full(slot), I8_A_SLICE);
                                bulk_g2s(s_A + slot * I8_A_SLICE, asrc, I8_A_SLICE, a_full(slot));
                            }
                        }
                        asrc += I8_A_SLICE;
                        if (++as[r] == ring_size(r)) { as[r] = 0; aph[r] ^= 1u; }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp > I8_EPI_WARPS + NI) {
        // spare issuer warp of this instantiation
    } else if (warp > I8_EPI_WARPS) {
        // ================= MMA issuers =================
        // Issuer w owns the A slices i = w (mod NI): the issue path (barrier wait, four to eight
        // UTCIMMA, commit) of a single thread cannot keep the tensor pipe busy with 64..256-column
        // int8 instructions, two threads can (0.60 -> 0.54 ms on the bare engine).  Digit products
        // accumulate into shared TMEM columns from both threads; integer accumulation is
        // order-independent, and the one ordering that matters -- the overwriting (accumulate = 0)
        // MMAs of slice 0 in the tile's first k-block -- is published through t_init.
        // Each issuer consumes its own ring of A slots (see the producer).  Bit-exact against the
        // single-issuer kernel and row-partition invariant (tools/i8_issuer_crosscheck.py,
        // tools/i8_determinism.py).
        const int w = warp - (I8_EPI_WARPS + 1);
        const int base = ring_base(w), size = ring_size(w);
        int as = 0, bs = 0;  // as: position in this issuer's own ring
        uint32_t aph = 0, bph = 0, tph = 0, iph = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const I8Item it = i8_item(a, item);
            I8_FOR_EACH_TILE(it, a, nt) {
                const int lo = a.kb_lo[nt], hi = a.kb_hi[nt];
                mbar_wait(t_empty, tph ^ 1u, a.trap_info, 6, nt);  // epilogue has drained the accumulators
                tph ^= 1u;
                tc_fence_after();
                if (w != 0 && lo < hi) {
                    mbar_wait(t_init, iph, a.trap_info, 7, nt);
                    tc_fence_after();
                }
                if (lo < hi) iph ^= 1u;
                // The MMAs of one k-block whose B image has Z leading zero slices, with every instruction
                // shape a compile-time constant (a run-time column loop costs the single issuing
                // thread -- the bottleneck of this kernel -- more than the skipped MMAs give back).
                auto issue_block = [&](auto zc, int kb) {
                    constexpr int Z = decltype(zc)::value;
                    const uint32_t sb = s_B + bs * B_BLOCK + Z * I8_B_SLICE;  // first non-zero slice
#pragma unroll
                    for (int i = 0; i + Z <= KS - 1; ++i) {
                        if (i % NI == w) {
                            const int slot = base + as;
                            mbar_wait(a_full(slot), aph, a.trap_info, 1 + 16 * slot, nt);
                            tc_fence_after();
                            if (lane == 0) {
                                const uint32_t sa = s_A + slot * I8_A_SLICE;
#pragma unroll
                                for (int ks = 0; ks < ((a.debug & 4) ? 0 : I8_BK / 32); ++ks) {
                                    const uint64_t ad = umma_desc_sw128(sa + 32 * ks);
                                    const uint32_t acc = (i > 0 || ks > 0 || kb > lo) ? 1u : 0u;
                                    // (KS - i - Z) * 64 columns in equal parts of <= 256 (multiples of 32)
                                    const int width = (KS - i - Z) * I8_BN;
                                    const int parts = (width + 255) / 256;
                                    const int n = width / parts;  // 192, 160, 224 ...: multiples of 32 for KS <= 8
#pragma unroll
                                    for (int c0 = 0; c0 < width; c0 += n) {
                                        const uint64_t bd = umma_desc_sw128(sb + c0 * I8_BK + 32 * ks);
                                        tc_mma_i8(tmem + (i + Z) * I8_BN + c0, ad, bd, umma_idesc_i8(n), acc);
                                    }
                                }
                                tc_commit(a_empty(slot));
                                if (i == 0 && kb == lo) tc_commit(t_init);
                            }
                            __syncwarp();
                            if (++as == size) { as = 0; aph ^= 1u; }
                        }
                    }
                };
                for (int kb = lo; kb < hi; ++kb) {
                    const int z = (a.b_lead != nullptr && kb != lo) ? min((int)a.b_lead[a.b_tile_img[nt] + (kb - lo)], KS) : 0;
                    mbar_wait(b_full(bs), bph, a.trap_info, 3 + 16 * bs, nt);
                    switch (z) {
                        case 0: issue_block(I8Int<0>{}, kb); break;
                        case 1: issue_block(I8Int<1>{}, kb); break;
                        case 2: issue_block(I8Int<2>{}, kb); break;
                        case 3: issue_block(I8Int<3>{}, kb); break;
                        case 4: issue_block(I8Int<(KS > 4 ? 4 : KS)>{}, kb); break;
                        case 5: issue_block(I8Int<(KS > 5 ? 5 : KS)>{}, kb); break;
                        case 6: issue_block(I8Int<(KS > 6 ? 6 : KS)>{}, kb); break;
                        default: break;  // every slice zero: nothing to multiply
                    }
                    if (lane == 0) tc_commit(b_empty(bs));
                    __syncwarp();
                    if (++bs == I8_NB) { bs = 0; bph ^= 1u; }
                }
                if (lane == 0) tc_commit(t_full);
                __syncwarp();
            }
        }
    } else {
        // ================= epilogue =================
        const int quarter = warp & 3, cg = warp >> 2;
        const int r = lane >> 2, q = lane & 3;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + cg * I8_TC;
        uint32_t tph = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const I8Item it = i8_item(a, item);
            const int row0 = it.mt * I8_BM + quarter * 32 + r;  // the thread's rows: row0 + 8 k
            const int row_acc = row0 + 8 * i8_quad_row(q);       // the row whose partial it keeps
            const double sa = a.a_scale[row_acc];
            typename Epi::State st;
            I8_FOR_EACH_TILE(it, a, nt) {
                const int c0 = nt * I8_BN + cg * I8_TC;
                if (nt % I8_GROUP == 0) epi.begin(st);
                epi.prefetch(row_acc, c0);
                const bool empty = a.kb_lo[nt] >= a.kb_hi[nt];  // no k-blocks: accumulators are stale
                // One warp waits on the mbarrier, the other fifteen on a named barrier behind it: a warp
                // sleeping in mbarrier.try_wait is woken by EVERY barrier event of the CTA (~500 per tile:
                // each bulk-copy completion and each MMA commit) only to go back to sleep, and sixteen of
                // them polling were a quarter of all instructions this kernel issued (19 M NANOSLEEP /
                // TRYWAIT pairs per launch, profiles/r02_stalls.md) -- on a chip at its power cap.
                if ((a.debug & 256) || warp == 0) mbar_wait(t_full, tph, a.trap_info, 5, nt);
                if (!(a.debug & 256)) asm volatile("bar.sync 1, %0;" ::"n"(I8_EPI_WARPS * 32) : "memory");
                tph ^= 1u;
                tc_fence_after();
                // drain: recombine this thread's elements into registers, hand TMEM back to the
                // MMA warps at once, then run the epilogue math while the next tile accumulates
                double x[4][I8_NM];
                if ((a.debug & 16) || empty) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int m = 0; m < I8_NM; ++m) x[k][m] = 0.0;
                } else {
                    i8_combine_quad<KS>(taddr, x);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty);
                if (!(a.debug & 8)) epi.tile(st, x, row0, c0, q, a.b_scale + c0);
                if (nt + 1 == ne_) epi.end(st, nt / I8_GROUP, cg, row_acc, sa);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == I8_EPI_WARPS + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------------------
// plain store: D[row, col] = x * scale_row * scale_col
struct I8EpilogueStore {
    static constexpr int ISSUERS = 2;
    double* D;
    int64_t ldd;
    int32_t rows, cols;  // bounds of D
    struct State {};
    __device__ __forceinline__ void begin(State&) const {}
    __device__ __forceinline__ void prefetch(int, int) const {}
    __device__ __forceinline__ void end(State&, int, int, int, double) const {}
    __device__ __forceinline__ void tile(State&, const double (&x)[4][I8_NM], int row0, int c0, int q,
                                         const double* sb) const {
        // every row of the thread needs its own scale here (the fused epilogues only that of the
        // row whose partial they keep)
        double2 sc[I8_NM / 2];
#pragma unroll
        for (int m = 0; m < I8_NM; m += 2) sc[m / 2] = *reinterpret_cast<const double2*>(sb + i8_col(m, q));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int row = row0 + 8 * k;
            if (row >= rows) continue;
            const double sa = row_scale[row];
            double* dst = D + (size_t)row * ldd + c0;
#pragma unroll
            for (int m = 0; m < I8_NM; m += 2) {
                const int c = i8_col(m, q);
                const double v0 = x[k][m] * sa * sc[m / 2].x, v1 = x[k][m + 1] * sa * sc[m / 2].y;
                if (c0 + I8_TC <= cols && (ldd & 1) == 0) {
                    *reinterpret_cast<double2*>(dst + c) = make_double2(v0, v1);
                } else {
                    if (c0 + c < cols) dst[c] = v0;
                    if (c0 + c + 1 < cols) dst[c + 1] = v1;
                }
            }
        }
    }
    const double* row_scale;  // a_scale of the fold (set by the launcher)
};

// per-channel observed data, padded to C_pad with benign values (shared with the DMMA path)
// transposed fold: G[n, e] contracted with dU/dtheta_j[n, e] over the tile's energies.
// Per tile and parameter a thread multiplies its 4 rows x 4 columns, the quad adds up each row
// (a transposing butterfly), and thread q keeps the sum of row row0 + 8 i8_quad_row(q).
// gpart[(j * n_part + I8_CG * group + cg) * ld_part + row],  n_part = I8_CG * n_groups
constexpr int I8_GRAD_PMAX = 16;  // parameters per launch of the gradient epilogue (register budget)
struct I8EpilogueGrad {
    static constexpr int ISSUERS = 2;
    const double* dU;  // [P, rows, ldu], already offset to plane p0
    int64_t plane;
    int32_t ldu;
    int32_t P;           // parameters of this launch: theta columns p0 .. p0 + P - 1, P <= I8_GRAD_PMAX
    uint32_t used_mask;  // bit j: column p0 + j is used by the model
    int32_t n_part;
    double* gpart;       // already offset to plane p0
    int32_t ld_part;
    int32_t dbg;  // timing experiments: 32 = no dU loads, 64 = L2 prefetch of the tile's dU lines

    struct State {
        double rs[I8_GRAD_PMAX];
    };
    __device__ __forceinline__ void begin(State& st) const {
#pragma unroll
        for (int j = 0; j < I8_GRAD_PMAX; ++j) st.rs[j] = 0.0;
    }
    // Optional (dbg & 64): request the tile's dU lines (one 128-byte line per parameter and row;
    // thread q of a quad takes row row0 + 8 q) from HBM while the tile is still accumulating.
    // It paid while the epilogue was a thread-per-row loop; with the quad layout the loads of a
    // plane overlap the arithmetic of the previous one and the extra L2 traffic only costs
    // (ablation on the bench shape: transposed fold 16.0 ms with, 15.6 ms without), so it is off.
    __device__ __forceinline__ void prefetch(int row, int c0) const {
        if (!(dbg & 64)) return;
        const double* base = dU + (size_t)row * ldu + c0;
#pragma unroll
        for (int p = 0; p < I8_GRAD_PMAX; ++p)
            if (p < P && ((used_mask >> p) & 1u)) {
#pragma unroll
                for (int b = 0; b < I8_TC / 16; ++b) prefetch_l2(base + (size_t)p * plane + 16 * b);
            }
    }
    __device__ __forceinline__ void end(State& st, int group, int cg, int row, double sa) const {
#pragma unroll
        for (int p = 0; p < I8_GRAD_PMAX; ++p)
            if (p < P && ((used_mask >> p) & 1u))
                gpart[((size_t)p * n_part + I8_CG * group + cg) * ld_part + row] = st.rs[p] * sa;
    }
    __device__ __forceinline__ void tile(State& st, const double (&x)[4][I8_NM], int row0, int c0, int q,
                                         const double* sb) const {
        double g[4][I8_NM];
#pragma unroll
        for (int m = 0; m < I8_NM; m += 2) {
            const double2 sc = *reinterpret_cast<const double2*>(sb + i8_col(m, q));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                g[k][m] = x[k][m] * sc.x;
                g[k][m + 1] = x[k][m + 1] * sc.y;
            }
        }
        const double* base = dU + (size_t)row0 * ldu + c0;
#pragma unroll
        for (int p = 0; p < I8_GRAD_PMAX; ++p) {
            if (p < P && ((used_mask >> p) & 1u)) {
                const double* src = base + (size_t)p * plane;
                double s[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double t = 0.0;
#pragma unroll
                    for (int m = 0; m < I8_NM; m += 2) {
                        const double2 v = (dbg & 32) ? make_double2(g[k][m], g[k][m + 1])
                                                     : *reinterpret_cast<const double2*>(src + (size_t)(8 * k) * ldu + i8_col(m, q));
                        t = fma(g[k][m + 1], v.y, fma(g[k][m], v.x, t));
                    }
                    s[k] = t;
                }
                // transposing butterfly over the quad: 3 shuffles + 3 adds instead of 8 + 8 (the four
                // threads hold partials of the same four rows; each ends up with ONE row's total,
                // thread q with row i8_quad_row(q))
                const bool b0 = q & 1, b1 = q & 2;
                const double r0 = __shfl_xor_sync(0xffffffffu, b0 ? s[0] : s[2], 1);
                const double r1 = __shfl_xor_sync(0xffffffffu, b0 ? s[1] : s[3], 1);
                const double k0 = (b0 ? s[2] : s[0]) + r0, k1 = (b0 ? s[3] : s[1]) + r1;
                const double r2 = __shfl_xor_sync(0xffffffffu, b1 ? k0 : k1, 2);
                st.rs[p] += (b1 ? k1 : k0) + r2;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// statistic + digits of W in one pass over the folded rates (one warp per row)
// ---------------------------------------------------------------------------------------
// The forward fold stores X = R u with the plain epilogue; this kernel turns each row into
// its log-likelihood (fixed summation order: lane-serial over channels lane, lane + 32, ...,
// then an xor tree -- bit-reproducible and independent of the batch size), overwrites the
// row with W = d loglike / d (R u), and slices W into the A-operand images of the transposed
// fold while the row is still in L1/L2.  At 64 warps per SM the long FP64 chains of the
// statistics (log, sqrt, divisions) hide each other; inside the fold kernel's 16 epilogue
// warps they were the critical path.
struct StatSliceArgs {
    double* X;  // [rows_pad, ldx]  in: R u   out: W
    int32_t ldx;
    ChanArrays ch;
    const double* counts_on;
    const double* counts_off;
    int64_t ld_counts;
    int32_t n_valid, rows_pad, C;
    double* loglike;  // element [row * ll_stride]
    int64_t ll_stride;
    // digits of W (GRAD only); on entry scale[] holds the row scales of U's digits
    int8_t* img;
    double* scale;
    int32_t kb_total;
    // accuracy guard of the integer fold: colsum[c] = sum_e |R[e, c]|; guard[0] counts the rows
    // whose estimated log-likelihood error exceeds guard_tol * max(|loglike|, 1) or whose
    // gradient-side estimate does (each row once), guard[1] keeps the largest value-side
    // estimate / max(|loglike|, 1) seen (bits of a non-negative double), guard[2] the largest
    // gradient-side estimate.  The first flag_cap flagged rows are also listed (global row
    // index row_base + row, in arrival order) for the per-row DMMA fallback.
    const double* colsum;
    unsigned long long* guard;
    int32_t* flag_list;
    int32_t flag_cap;
    int64_t row_base;
    double guard_eta;    // typical normwise error of the forward fold, 2^(-8 KS_forward)
    double guard_eta_g;  // of the transposed fold, 2^(-8 KS) -- KS is this kernel's template parameter
    double guard_grad;   // calibration of the gradient-side estimate (see the kernel)
    double guard_tol;
    const double* kinv;  // balancing of the transposed fold: the digits are those of W[c] * kinv[c] (nullptr: 1)
};

__device__ __forceinline__ void guard_flag_row(const StatSliceArgs& a, int row) {
    const unsigned long long idx = atomicAdd(a.guard, 1ull);
    if (a.flag_list != nullptr && idx < (unsigned long long)a.flag_cap) a.flag_list[idx] = (int32_t)(a.row_base + row);
}

// KS: slices of W (the transposed fold's A operand).
// 64 registers = 4 blocks per SM: the kernel lives on occupancy (its FP64 chains -- two logarithms, a
// square root, two reciprocals per channel -- hide behind other warps); measured on the bench shape
// (profiles/r02b_ab_stat.txt): 66 registers = 3 blocks 6.47 ms per 400k rows, 64 registers 5.6, forced
// below that (48 / 40 registers, spills) 6.5; channel constants as 64-byte records (4 x LDG.128 instead
// of 9 x LDG.64, but 4x the L1 wavefronts of the coalesced column arrays) 6.4 - 7.2.
template <int KS, int STAT, bool GRAD>
__global__ void __launch_bounds__(256, 4) stat_slice_kernel(const StatSliceArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.rows_pad) return;
    const bool live = row < a.n_valid;
    double* x = a.X + (size_t)row * a.ldx;
    double rs = 0.0, eb = 0.0, wsum = 0.0;
    int amax_hi = 0;  // largest |W'| of the row, tracked on the high word (see unfold_generic)
    bool flagged = false;
    if (live) {
        const double u_max = a.scale[row] * (double)I8_TOP_DIGIT;  // largest |U| of the row
        const double* con = a.counts_on ? a.counts_on + (size_t)row * a.ld_counts : nullptr;
        const double* coff = a.counts_off ? a.counts_off + (size_t)row * a.ld_counts : nullptr;
        constexpr bool HAS_OFF = STAT == ELISA_STAT_PSTAT || STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT;
        const bool use_off = (STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) && coff != nullptr;
        // the folded rate (and the row's own counts) of the next pass are loaded during this one: they
        // come from L2 / HBM, the per-channel constants below from L1 (every row of the SM reads the same)
        double xn = 0.0, con_n = 0.0, coff_n = 0.0;
        if (lane < a.C) {
            xn = x[lane];
            if (con != nullptr) con_n = con[lane];
            if (use_off) coff_n = coff[lane];
        }
        for (int col = lane; col < a.C; col += 32) {
            const double xv = xn, con_v = con_n, coff_v = coff_n;
            if (col + 32 < a.C) {
                xn = x[col + 32];
                if (con != nullptr) con_n = con[col + 32];
                if (use_off) coff_n = coff[col + 32];
            }
            ChannelData d;
            const double sc = a.ch.scale[col];
            d.on = a.ch.on[col];
            d.gof_on = a.ch.gof_on[col];
            d.off = d.sigma = d.ratio = d.gof_off = d.den = 0.0;
            if (STAT == ELISA_STAT_CHI2) d.sigma = a.ch.sigma[col];
            if (HAS_OFF) {
                d.off = a.ch.off[col];
                d.ratio = a.ch.ratio[col];
            }
            if (STAT == ELISA_STAT_PGSTAT) d.sigma = a.ch.sigma[col];
            if (STAT == ELISA_STAT_WSTAT) d.gof_off = a.ch.gof_off[col];
            if (STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) d.den = a.ch.den[col];
            if (con != nullptr) {
                d.on = con_v;
                if (STAT != ELISA_STAT_CHI2) d.gof_on = poisson_gof(d.on);
            }
            if (use_off) {
                d.off = coff_v;
                if (STAT == ELISA_STAT_WSTAT) d.gof_off = poisson_gof(d.off);
            }
            double dx = 0.0;
            rs += stat_channel<STAT, true, false, true>(xv * sc, d, dx);  // FAST: n * (1 / mu) for the two divisions
            eb = fma(fabs(dx * sc), a.colsum[col], eb);
            if (GRAD) {
                const double w = dx * sc;
                x[col] = w;
                const double wb = a.kinv != nullptr ? fabs(w) * a.kinv[col] : fabs(w);  // as the digits see it
                amax_hi = max(amax_hi, __double2hiint(wb) & 0x7fffffff);  // NaN / Inf: >= 0x7ff00000
                wsum += wb;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rs += __shfl_xor_sync(0xffffffffu, rs, o);
            eb += __shfl_xor_sync(0xffffffffu, eb, o);
        }
        if (lane == 0) {
            a.loglike[(size_t)row * a.ll_stride] = rs;
            // first-order estimate of the log-likelihood error the sliced fold may have caused:
            // sum_c |dl/dx_c| * (typical normwise error of x_c).  A NaN row (NaN parameters) is
            // NaN under either fold engine and is not flagged.
            const double est = a.guard_eta * u_max * eb / fmax(fabs(rs), 1.0);
            flagged = est > a.guard_tol;
            if (est > 0.0 && est < 1.79e308) atomicMax(a.guard + 1, (unsigned long long)__double_as_longlong(est));
            if (!GRAD && flagged) guard_flag_row(a, row);
        }
    }
    if (!GRAD) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax_hi = max(amax_hi, __shfl_xor_sync(0xffffffffu, amax_hi, o));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    const bool bad = amax_hi >= 0x7ff00000;
    const double amax = amax_hi == 0 ? 0.0 : __hiloint2double(amax_hi, (int)0xffffffff);
    if (live && lane == 0) {
        if (wsum > 0.0 && !bad) {
            // gradient side of the guard.  The transposed fold's error is relative to the LARGEST
            // |W| of the row; when a few starved channels (n >> mu) tower over the rest, the bulk of
            // the row -- and with it G and the gradient -- keeps fewer digits.  peak / mean of |W|
            // times the engine's normwise error estimates the relative error of G.
            const double est_g = a.guard_eta_g * a.guard_grad * amax * (double)a.C / wsum;
            flagged |= est_g > a.guard_tol;
            atomicMax(a.guard + 2, (unsigned long long)__double_as_longlong(est_g));
        }
        if (flagged) guard_flag_row(a, row);
    }
    double f1, f2;
    const double sc = slice_scale<KS>(amax, bad, f1, f2);
    if (lane == 0) a.scale[row] = sc;
    __syncwarp();  // the row of W written above is read back by other lanes
    slice_row_digits<KS>(x, live, (a.ldx & 1) == 0, a.C, f1, f2, row % I8_BM, I8_BM, 0, a.kb_total, a.img,
                         (int64_t)(row / I8_BM) * a.kb_total, lane, a.kinv);
}

}  // namespace elisa
