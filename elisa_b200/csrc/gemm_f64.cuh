// FP64 tensor-core GEMM for the response fold on sm_100a.
//
//   D[M, N] = A[M, K] . B[K, N]      FP64; the epilogue functor consumes D in registers
//
// forward fold   : A = unfold U [samples, E],  B = R   [E, C]  (FixedData.response_matrix)
// transposed fold: A = W        [samples, C],  B = R^T [C, E]
// (the contraction of infer/likelihood.py:205 `resp_matrix @ unfold` batched over the
// parameter axis, and its reverse-mode transpose.)
//
// Blackwell has no FP64 kind of tcgen05.mma; the native FP64 tensor op on sm_100a is
// DMMA.8x8x4 (mma.sync.m8n8k4.f64), measured here at the same ~37 TFLOP/s as the DFMA
// pipe (tools/fp64_microbench.cu), so the kernel is built to keep that one pipe busy:
//   CTA tile 128 x 64 x 16, 4 warps (2 x 2), warp tile 64 x 32 -> 32 independent DMMA
//   accumulators per k-step; 3-stage cp.async ring (16-byte, L1-bypassing); 2 CTAs/SM so
//   one CTA's epilogue overlaps the other's main loop; smem rows padded so that the
//   fragment loads (LDS.64 for A, LDS.128 for B) are bank-conflict free.
//
// B is PACKED by the plan, tile-major: column tile j (64 columns) is a contiguous
// [k_hi[j] - k_lo[j], 64] block holding only the rows between the first and last
// non-zero of those columns (rounded to 16).  A dense response has k range [0, K); a
// band-sparse one (CCD-like RMF) keeps ~(64 + band) rows per tile, which turns the
// sparse fold into a block-banded dense DMMA contraction with the same kernel.
// M is padded to 128 and N to 64 by the plan, so the main loop carries no bounds checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace elisa {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 64;
constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 128;
constexpr int GEMM_LDA_S = GEMM_BK + 4;  // 20 doubles: row stride = 32 B mod 128 B
constexpr int GEMM_LDB_S = GEMM_BN + 4;  // 68 doubles: row stride = 32 B mod 128 B
constexpr int GEMM_A_STAGE = GEMM_BM * GEMM_LDA_S;
constexpr int GEMM_B_STAGE = GEMM_BK * GEMM_LDB_S;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * (GEMM_A_STAGE + GEMM_B_STAGE) * 8;

// Packed B operand (device pointers; owned by the plan).
struct PackedB {
    const double* __restrict__ data;      // concatenated tiles
    const int64_t* __restrict__ tile_off;  // [n_tiles] offset of tile j in `data` (doubles)
    const int32_t* __restrict__ k_lo;      // [n_tiles] first k row (multiple of 16)
    const int32_t* __restrict__ k_hi;      // [n_tiles] one past the last k row (multiple of 16)
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Accumulator fragment owned by one thread after the main loop.
//   v[i][t][r]:  row = m0 + warp_m*64 + 8*i + g          (g = lane >> 2, q = lane & 3)
//                col = n0 + warp_n*32 + 16*(t>>1) + 4*q + 2*r + (t&1)
// so for a fixed (i, l = t>>1) the thread owns 4 CONSECUTIVE columns 4q .. 4q+3, in the
// order (t&1, r) = (0,0), (1,0), (0,1), (1,1).
struct AccFrag {
    double v[8][4][2];
    __device__ __forceinline__ double get(int i, int l, int cc) const { return v[i][2 * l + (cc & 1)][cc >> 1]; }
    __device__ __forceinline__ void set(int i, int l, int cc, double x) { v[i][2 * l + (cc & 1)][cc >> 1] = x; }
};

// Epilogue signature:
//   epi(acc, smem, tile_m, tile_n, row0, col0, warp_m, warp_n, lane)
// row0 = m0 + warp_m*64 + g is the thread's first row (rows row0 + 8 i), col0 =
// n0 + warp_n*32 + 4 q its first column (columns col0 + 16 l + cc).
template <class Epilogue>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_f64_kernel(const double* __restrict__ A, int lda, PackedB Bp, Epilogue epi) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + GEMM_STAGES * GEMM_A_STAGE;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int warp_m = warp >> 1, warp_n = warp & 1;
    const int g = lane >> 2, q = lane & 3;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;

    const int k_lo = Bp.k_lo[blockIdx.x];
    const int KB = (Bp.k_hi[blockIdx.x] - k_lo) / GEMM_BK;
    const double* Ag = A + (size_t)m0 * lda + k_lo;
    const double* Bg = Bp.data + Bp.tile_off[blockIdx.x];

    auto load_stage = [&](int stage, int kb) {
        double* as = As + stage * GEMM_A_STAGE;
        double* bs = Bs + stage * GEMM_B_STAGE;
        const double* ag = Ag + kb * GEMM_BK;
        const double* bg = Bg + (size_t)kb * (GEMM_BK * GEMM_BN);
#pragma unroll
        for (int r = 0; r < 8; ++r) {  // A tile: 128 rows x 8 chunks of 16 B
            const int c = tid + GEMM_THREADS * r;
            const int row = c >> 3, kc = c & 7;
            cp_async16(as + row * GEMM_LDA_S + 2 * kc, ag + (size_t)row * lda + 2 * kc);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // B tile: 16 rows x 32 chunks of 16 B, contiguous 8 KB
            const int c = tid + GEMM_THREADS * r;
            const int row = c >> 5, nc = c & 31;
            cp_async16(bs + row * GEMM_LDB_S + 2 * nc, bg + 2 * c);
        }
    };

    AccFrag acc;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) acc.v[i][t][0] = acc.v[i][t][1] = 0.0;

#pragma unroll
    for (int s = 0; s < GEMM_STAGES - 1; ++s) {
        if (s < KB) load_stage(s, s);
        cp_async_commit();
    }

    const int a_off = (warp_m * 64 + g) * GEMM_LDA_S + q;
    const int b_off = q * GEMM_LDB_S + warp_n * 32 + 2 * g;

    for (int kb = 0; kb < KB; ++kb) {
        cp_async_wait<GEMM_STAGES - 2>();
        __syncthreads();
        {
            const int nk = kb + GEMM_STAGES - 1;
            if (nk < KB) load_stage(nk % GEMM_STAGES, nk);
            cp_async_commit();
        }
        const double* as = As + (kb % GEMM_STAGES) * GEMM_A_STAGE + a_off;
        const double* bs = Bs + (kb % GEMM_STAGES) * GEMM_B_STAGE + b_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * GEMM_LDA_S + 4 * j];
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                const double2 bb = *reinterpret_cast<const double2*>(bs + 4 * j * GEMM_LDB_S + 16 * l);
                b[2 * l] = bb.x;
                b[2 * l + 1] = bb.y;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int t = 0; t < 4; ++t) dmma_8x8x4(acc.v[i][t][0], acc.v[i][t][1], a[i], b[t]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // smem is free for the epilogue from here on

    epi(acc, smem, (int)blockIdx.y, (int)blockIdx.x, m0 + warp_m * 64 + g, n0 + warp_n * 32 + 4 * q, warp_m, warp_n,
        lane);
}

// Plain store epilogue: D[row, col] = acc (two 16-byte stores per 4-column run).
struct EpilogueStore {
    double* __restrict__ D;
    int ldd;
    __device__ __forceinline__ void operator()(AccFrag& acc, double*, int, int, int row0, int col0, int, int,
                                               int) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            double* drow = D + (size_t)(row0 + 8 * i) * ldd + col0;
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                double2 lo, hi;
                lo.x = acc.get(i, l, 0);
                lo.y = acc.get(i, l, 1);
                hi.x = acc.get(i, l, 2);
                hi.y = acc.get(i, l, 3);
                *reinterpret_cast<double2*>(drow + 16 * l) = lo;
                *reinterpret_cast<double2*>(drow + 16 * l + 2) = hi;
            }
        }
    }
};

}  // namespace elisa
