// elisa_b200 -- plan, kernels and C ABI of the B200-native forward-folding likelihood path.
//
// Per chunk of parameter vectors and per spectrum the path is four launches:
//   1. unfold_kernel      component photon-flux integrals + model composition -> U [n, E]
//                         (and, with gradients, dU/dtheta_j [j, n, E] by forward duals)
//   2. gemm_f64_kernel<EpilogueStat>   fold U . R with DMMA; the epilogue scales by
//                         area * exposure, evaluates the fit statistic per channel, reduces
//                         it over the tile's channels and emits W = d loglike / d(R u)
//   3. gemm_f64_kernel<EpilogueGrad>   transposed fold W . R^T with DMMA; the epilogue
//                         contracts G[n, e] with dU/dtheta_j[n, e] over the tile's energies
//   4. reduce_kernel      fixed-order sum of the per-tile partials -> loglike[n, s], grad[n, s, :]
// Folded counts, G and the per-channel log-likelihoods never reach HBM.
//
// Reference lines restated by each piece are cited where the arithmetic lives
// (components.cuh, stats.cuh) and in include/elisa_b200.h.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nvrtc.h>  // types only: libnvrtc is opened lazily with dlopen

#include <map>
#include <mutex>

#include "elisa_b200.h"
#include "components.cuh"
#include "dual.cuh"
#include "fold_i8.cuh"
#include "gemm_f64.cuh"
#include "sample.cuh"
#include "stats.cuh"
#include "unfold_skel.cuh"

namespace elisa {

// ======================================================================================
// device-side descriptors
// ======================================================================================
struct LeafDev {
    int32_t kind, method, np, pad_;
    int32_t src[ELISA_MAX_COMP_PARAMS];
    int32_t pad2_;
    double cst[ELISA_MAX_COMP_PARAMS];
    double aux[2];
    // energy shifts around the leaf: grid scale, its log, output scale -- on the photon grid
    // and on the grid of the enclosing flux normalisation
    double gs, lgs, os;
    double ngs, lngs, nos;
    // FREE energy shifts around the leaf (ELISA_SHIFT_*, theta column); n_shift of them
    int32_t n_shift, pad3_;
    int32_t shift_kind[ELISA_MAX_SHIFTS];
    int32_t shift_src[ELISA_MAX_SHIFTS];
};

constexpr int MAX_OPS = 2 * ELISA_MAX_COMPONENTS + ELISA_MAX_NORMS;

struct NormDev {
    int32_t kind, f_src;
    int32_t op_lo, op_hi;  // the normalised sub-expression is ops[op_lo, op_hi); the opcode sits at op_hi
    double f_cst;
};

struct ModelDev {
    int32_t n_leaf, n_ops;
    int32_t ops[MAX_OPS];
    uint32_t used_mask;  // theta columns this model depends on
    int32_t n_norms;
    LeafDev leaf[ELISA_MAX_COMPONENTS];
    NormDev norm[ELISA_MAX_NORMS];
};

struct UnfoldArgs {
    ModelDev model;
    UnfoldCore core;
};

// ======================================================================================
// 1. unfold: component evaluation + composition
// ======================================================================================
template <int PU>
struct StackT {
    using type = Dual<PU>;
};
template <>
struct StackT<0> {
    using type = double;
};

template <int PU, int NPK>
__device__ __forceinline__ typename StackT<PU>::type scatter_leaf(const Dual<NPK>& x, const int32_t* src) {
    Dual<PU> r;
    r.v = x.v;
#pragma unroll
    for (int j = 0; j < PU; ++j) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < NPK; ++k) s += (src[k] == j) ? x.d[k] : 0.0;
        r.d[j] = s;
    }
    return r;
}

#define ELISA_FOR_EACH_KIND(X)                                                                                     \
    X(ELISA_COMP_BAND) X(ELISA_COMP_BANDEP) X(ELISA_COMP_BENTPL) X(ELISA_COMP_BLACKBODY)                           \
    X(ELISA_COMP_BLACKBODYRAD) X(ELISA_COMP_BROKENPL) X(ELISA_COMP_COMPT) X(ELISA_COMP_CUTOFFPL)                   \
    X(ELISA_COMP_GAUSS) X(ELISA_COMP_LORENTZ) X(ELISA_COMP_OTTB) X(ELISA_COMP_OTTS) X(ELISA_COMP_PLENFLUX)         \
    X(ELISA_COMP_PLPHFLUX) X(ELISA_COMP_POWERLAW) X(ELISA_COMP_SMOOTHLYBROKENPL) X(ELISA_COMP_CONSTANT)            \
    X(ELISA_COMP_EDGE) X(ELISA_COMP_EXPABS) X(ELISA_COMP_EXPFAC) X(ELISA_COMP_GABS) X(ELISA_COMP_HIGHECUT)         \
    X(ELISA_COMP_PLABS) X(ELISA_COMP_PHOTONABS)

// Stack machine over ops[lo, hi) for one bin per lane.  on_norm_grid: `gv` is the grid of a flux
// normalisation, leaves take their scales on that grid; nrm: finished normalisation factors.
template <int PU>
__device__ __noinline__ typename StackT<PU>::type interp_range(const ModelDev& model,
                                                               const double (*scr)[LEAF_SCRATCH], int lo, int hi,
                                                               const GridView& gv, int e, bool on_norm_grid,
                                                               const typename StackT<PU>::type* nrm) {
    using TT = typename StackT<PU>::type;
    constexpr bool GRAD = PU > 0;
    TT stk[ELISA_MAX_COMPONENTS];
    int sp = 0;
    for (int o = lo; o < hi; ++o) {
        const int op = model.ops[o];
        if (op >= 0) {
            const LeafDev& L = model.leaf[op];
            const double gs = on_norm_grid ? L.ngs : L.gs, lgs = on_norm_grid ? L.lngs : L.lgs;
            const double os = on_norm_grid ? L.nos : L.os;
            TT val;
            switch (L.kind) {
#define X(KIND)                                                                                            \
    case KIND:                                                                                             \
        if constexpr (GRAD)                                                                                \
            val = scatter_leaf<PU, Comp<KIND>::NP>(                                                        \
                leaf_bin<KIND, Dual<Comp<KIND>::NP>>(scr[op], L.method, gv, e, op, gs, lgs), L.src);       \
        else                                                                                               \
            val = leaf_bin<KIND, double>(scr[op], L.method, gv, e, op, gs, lgs);                           \
        break;
                ELISA_FOR_EACH_KIND(X)
#undef X
                default:
                    val = Cst<TT>::make(0.0);
            }
            if (os != 1.0) val = os * val;
            stk[sp++] = val;
        } else if (op <= ELISA_OP_NORM0) {
            stk[sp - 1] = nrm[ELISA_OP_NORM0 - op] * stk[sp - 1];
        } else {
            const TT rhs = stk[--sp];
            const TT lhs = stk[sp - 1];
            stk[sp - 1] = (op == ELISA_OP_ADD) ? lhs + rhs : lhs * rhs;
        }
    }
    return stk[0];
}

// Ahead-of-time interpreter: runs any model program without runtime compilation (the
// fallback when NVRTC is unavailable, and the cross-check of the specialised kernels).
template <int PU>
__global__ void __launch_bounds__(UNFOLD_WARPS * 32) unfold_kernel(const __grid_constant__ UnfoldArgs args) {
    using TT = typename StackT<PU>::type;
    constexpr bool GRAD = PU > 0;
    __shared__ double s_scratch[UNFOLD_WARPS][ELISA_MAX_COMPONENTS][LEAF_SCRATCH];
    const UnfoldCore& a = args.core;
    const ModelDev& model = args.model;

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * UNFOLD_WARPS + wib;
    const int row = (int)(wid / a.n_seg), seg = (int)(wid % a.n_seg);
    if (row >= a.n_rows) return;
    const int E = a.gv.E;
    const int e_lo = seg * a.seg_len;
    const int e_hi = min(e_lo + a.seg_len, E);
    const bool last_seg = seg == a.n_seg - 1;
    const size_t row_off = (size_t)row * a.ld;

    if (row >= a.n_valid) {  // padding rows: zeros, so that the fold sees finite operands
        unfold_zero(a, row_off, e_lo, last_seg ? a.e_pad : e_hi, lane, GRAD);
        return;
    }

    // ---- per-parameter-vector scalars, once per warp -------------------------------
    double(*scr)[LEAF_SCRATCH] = s_scratch[wib];
    const double* th = a.theta + (size_t)row * a.P;
    for (int i = 0; i < model.n_leaf; ++i) {
        const LeafDev& L = model.leaf[i];
        if (lane < ELISA_MAX_COMP_PARAMS)
            scr[i][lane] = lane < L.np ? (L.src[lane] >= 0 ? th[L.src[lane]] : L.cst[lane]) : 0.0;
    }
    __syncwarp();
    for (int i = 0; i < model.n_leaf; ++i) {
        const LeafDev& L = model.leaf[i];
        double tmp[LEAF_SCRATCH];
#pragma unroll
        for (int k = 0; k < ELISA_MAX_COMP_PARAMS; ++k) tmp[k] = scr[i][k];
        switch (L.kind) {
#define X(KIND)                                                                               \
    case KIND:                                                                                \
        if (GRAD) leaf_pre<KIND, Dual<Comp<KIND>::NP>>(tmp, L.aux);                           \
        else leaf_pre<KIND, double>(tmp, L.aux);                                              \
        break;
            ELISA_FOR_EACH_KIND(X)
#undef X
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = ELISA_MAX_COMP_PARAMS; k < LEAF_SCRATCH; ++k) scr[i][k] = tmp[k];
        }
    }
    __syncwarp();

    // ---- flux normalisations: F / sum over the normalisation's own grid ----------------
    TT nrm[ELISA_MAX_NORMS];
    for (int k = 0; k < model.n_norms; ++k) {
        const NormDev& nd = model.norm[k];
        const GridView& gv = a.ngv[k];
        TT acc = Cst<TT>::make(0.0);
        for (int e0 = 0; e0 < gv.E; e0 += 31) {
            const int e = e0 + lane;
            const TT v = interp_range<PU>(model, scr, nd.op_lo, nd.op_hi, gv, e, true, nrm);
            if (lane < 31 && e < gv.E) acc = acc + (a.nw[k] != nullptr ? a.nw[k][e] : 1.0) * v;
        }
        acc = warp_sum(acc);
        TT F = Cst<TT>::make(nd.f_src >= 0 ? th[nd.f_src] : nd.f_cst);
        if constexpr (GRAD) {
            if (nd.f_src >= 0) F.d[nd.f_src] = 1.0;
        }
        nrm[k] = F / acc;
    }

    // ---- passes of 31 bins ------------------------------------------------------------
    for (int e0 = e_lo; e0 < e_hi; e0 += 31) {
        const int e = e0 + lane;
        const TT res = interp_range<PU>(model, scr, 0, model.n_ops, a.gv, e, false, nrm);
        if (lane < 31 && e < e_hi) unfold_store<TT>(a, row_off + e, res);
    }
    if (last_seg) unfold_zero(a, row_off, E, a.e_pad, lane, GRAD);
}

// Interpreter version of unfold_contract_generic (unfold_skel.cuh): gradient by contraction of
// G = d loglike / d u with the tangents, which never reach memory.  One warp per row.
template <int PU>
__global__ void __launch_bounds__(UNFOLD_WARPS * 32) unfold_contract_kernel(const __grid_constant__ UnfoldArgs args) {
    using TT = Dual<PU>;
    __shared__ double s_scratch[UNFOLD_WARPS][ELISA_MAX_COMPONENTS][LEAF_SCRATCH];
    const UnfoldCore& a = args.core;
    const ModelDev& model = args.model;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = (int)((int64_t)blockIdx.x * UNFOLD_WARPS + wib);
    if (row >= a.n_valid) return;
    const int E = a.gv.E;
    double(*scr)[LEAF_SCRATCH] = s_scratch[wib];
    const double* th = a.theta + (size_t)row * a.P;
    for (int i = 0; i < model.n_leaf; ++i) {
        const LeafDev& L = model.leaf[i];
        if (lane < ELISA_MAX_COMP_PARAMS)
            scr[i][lane] = lane < L.np ? (L.src[lane] >= 0 ? th[L.src[lane]] : L.cst[lane]) : 0.0;
    }
    __syncwarp();
    for (int i = 0; i < model.n_leaf; ++i) {
        const LeafDev& L = model.leaf[i];
        double tmp[LEAF_SCRATCH];
#pragma unroll
        for (int k = 0; k < ELISA_MAX_COMP_PARAMS; ++k) tmp[k] = scr[i][k];
        switch (L.kind) {
#define X(KIND)                                                   \
    case KIND:                                                    \
        leaf_pre<KIND, Dual<Comp<KIND>::NP>>(tmp, L.aux);         \
        break;
            ELISA_FOR_EACH_KIND(X)
#undef X
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int k = ELISA_MAX_COMP_PARAMS; k < LEAF_SCRATCH; ++k) scr[i][k] = tmp[k];
        }
    }
    __syncwarp();
    TT nrm[ELISA_MAX_NORMS];
    for (int k = 0; k < model.n_norms; ++k) {
        const NormDev& nd = model.norm[k];
        const GridView& gv = a.ngv[k];
        TT acc = Cst<TT>::make(0.0);
        for (int e0 = 0; e0 < gv.E; e0 += 31) {
            const int e = e0 + lane;
            const TT v = interp_range<PU>(model, scr, nd.op_lo, nd.op_hi, gv, e, true, nrm);
            if (lane < 31 && e < gv.E) acc = acc + (a.nw[k] != nullptr ? a.nw[k][e] : 1.0) * v;
        }
        acc = warp_sum(acc);
        TT F = Cst<TT>::make(nd.f_src >= 0 ? th[nd.f_src] : nd.f_cst);
        if (nd.f_src >= 0) F.d[nd.f_src] = 1.0;
        nrm[k] = F / acc;
    }
    const double* g_row = a.G + (size_t)row * a.ldg;
    double acc[PU];
#pragma unroll
    for (int j = 0; j < PU; ++j) acc[j] = 0.0;
    for (int e0 = 0; e0 < E; e0 += 31) {
        const int e = e0 + lane;
        const TT res = interp_range<PU>(model, scr, 0, model.n_ops, a.gv, e, false, nrm);
        if (lane < 31 && e < E) {
            const double g = g_row[e];
            const bool pass = !(a.clip && !(res.v >= 1e-300 && res.v <= 1e300));
#pragma unroll
            for (int j = 0; j < PU; ++j) acc[j] = fma(g, pass ? res.d[j] : 0.0, acc[j]);
        }
    }
    double* out = a.grad_out + (size_t)row * a.grad_stride;
#pragma unroll
    for (int j = 0; j < PU; ++j) {
        if (j < a.P) {
            const double t = ((a.used_mask >> j) & 1u) ? warp_sum(acc[j]) : 0.0;
            if (lane == 0) out[j] = t;
        }
    }
}

// ======================================================================================
// 2. statistic epilogue of the forward fold
// ======================================================================================
// The accumulator tile is parked in shared memory (the pipeline buffers are free by then)
// and re-read ROW-WISE: thread t owns row t of the 128 x 64 tile and walks its 64 channels,
// so the row sum is a private register (no shuffles, fixed order), the code is one compact
// loop (the fully unrolled register version overflowed the instruction cache) and W
// leaves through coalesced 512-byte row segments.
// partial log-likelihoods: part[tile_n * ld_part + row]
constexpr int EPI_LD = GEMM_BN + 1;  // 65 doubles: odd stride -> conflict-free row-wise access
static_assert(GEMM_BM * EPI_LD * 8 <= GEMM_SMEM_BYTES, "epilogue tile must fit the pipeline buffers");

__device__ __forceinline__ void park_tile(const AccFrag& acc, double* tile, int warp_m, int warp_n, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        double* row = tile + (warp_m * 64 + 8 * i + g) * EPI_LD + warp_n * 32 + 4 * q;
#pragma unroll
        for (int l = 0; l < 2; ++l)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) row[16 * l + cc] = acc.get(i, l, cc);
    }
}

template <int STAT, bool GRAD>
struct EpilogueStat {
    ChanArrays ch;
    const double* counts_on;   // per-replica observed data [n, ld_counts] or nullptr
    const double* counts_off;  // idem
    int64_t ld_counts;
    int32_t n_valid;  // rows with real parameter vectors
    int32_t C;        // real channels
    double* W;        // [rows, ldw]  d loglike / d (R u) (GRAD only)
    int32_t ldw;
    double* part;
    int32_t ld_part;

    __device__ __forceinline__ void operator()(AccFrag& acc, double* smem, int tile_m, int tile_n, int, int,
                                               int warp_m, int warp_n, int lane) const {
        double* tile = smem;
        park_tile(acc, tile, warp_m, warp_n, lane);
        __syncthreads();
        const int t = threadIdx.x;
        const int row = tile_m * GEMM_BM + t;
        const int c0 = tile_n * GEMM_BN;
        const int n_cols = min(GEMM_BN, C - c0);  // real channels of this tile
        double* xr = tile + t * EPI_LD;
        const bool rvalid = row < n_valid;
        const double* con = (counts_on != nullptr && rvalid) ? counts_on + (size_t)row * ld_counts + c0 : nullptr;
        const double* coff = (counts_off != nullptr && rvalid) ? counts_off + (size_t)row * ld_counts + c0 : nullptr;
        double rs = 0.0;
#pragma unroll 4
        for (int j = 0; j < GEMM_BN; ++j) {
            const int col = c0 + j;
            ChannelData d;
            const double sc = ch.scale[col];
            d.on = ch.on[col];
            d.gof_on = ch.gof_on[col];
            d.off = d.sigma = d.ratio = d.gof_off = 0.0;
            if (STAT == ELISA_STAT_CHI2) d.sigma = ch.sigma[col];
            if (STAT == ELISA_STAT_PSTAT || STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) {
                d.off = ch.off[col];
                d.ratio = ch.ratio[col];
            }
            if (STAT == ELISA_STAT_PGSTAT) d.sigma = ch.sigma[col];
            if (STAT == ELISA_STAT_WSTAT) d.gof_off = ch.gof_off[col];
            d.den = ch.den[col];
            const bool cvalid = j < n_cols;
            if (con != nullptr && cvalid) {
                d.on = con[j];
                if (STAT != ELISA_STAT_CHI2) d.gof_on = poisson_gof(d.on);
            }
            if ((STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) && coff != nullptr && cvalid) {
                d.off = coff[j];
                if (STAT == ELISA_STAT_WSTAT) d.gof_off = poisson_gof(d.off);
            }
            const double x = xr[j] * sc;
            double dx = 0.0;
            const double ll = stat_channel<STAT, GRAD>(x, d, dx);
            rs += cvalid ? ll : 0.0;
            if (GRAD) xr[j] = cvalid ? dx * sc : 0.0;
        }
        part[(size_t)tile_n * ld_part + row] = rs;
        if (GRAD) {
            __syncthreads();
            // coalesced copy of the W tile: each warp streams whole 512-byte row segments
            const int warp = t >> 5;
            for (int r = warp; r < GEMM_BM; r += GEMM_THREADS / 32) {
                const double* src = tile + r * EPI_LD;
                double* dst = W + (size_t)(tile_m * GEMM_BM + r) * ldw + c0;
                dst[lane] = src[lane];
                dst[lane + 32] = src[lane + 32];
            }
        }
    }
};

// ======================================================================================
// 3. gradient epilogue of the transposed fold
// ======================================================================================
// gpart[(j * n_tiles + tile_n) * ld_part + row]
struct EpilogueGrad {
    const double* dU;  // [P, rows, ldu]
    int64_t plane;
    int32_t ldu;
    int32_t P;
    uint32_t used_mask;
    int32_t n_tiles;
    double* gpart;
    int32_t ld_part;

    __device__ __forceinline__ void operator()(AccFrag& acc, double* smem, int, int tile_n, int row0, int col0,
                                               int warp_m, int warp_n, int lane) const {
        double* red = smem;
        const int m0 = row0 - warp_m * 64 - (lane >> 2);
        for (int j = 0; j < P; ++j) {
            if (!((used_mask >> j) & 1u)) continue;
            const double* base = dU + j * plane + (size_t)row0 * ldu + col0;
            double rs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double* p = base + (size_t)(8 * i) * ldu;
                double s = 0.0;
#pragma unroll
                for (int l = 0; l < 2; ++l) {
                    const double2 lo = *reinterpret_cast<const double2*>(p + 16 * l);
                    const double2 hi = *reinterpret_cast<const double2*>(p + 16 * l + 2);
                    s = fma(acc.get(i, l, 0), lo.x, s);
                    s = fma(acc.get(i, l, 1), lo.y, s);
                    s = fma(acc.get(i, l, 2), hi.x, s);
                    s = fma(acc.get(i, l, 3), hi.y, s);
                }
                rs[i] = s;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 1);
                rs[i] += __shfl_xor_sync(0xffffffffu, rs[i], 2);
            }
            __syncthreads();
            if ((lane & 3) == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) red[warp_n * GEMM_BM + warp_m * 64 + 8 * i + (lane >> 2)] = rs[i];
            }
            __syncthreads();
            gpart[((size_t)j * n_tiles + tile_n) * ld_part + m0 + threadIdx.x] =
                red[threadIdx.x] + red[GEMM_BM + threadIdx.x];
        }
    }
};

// ======================================================================================
// 4. fixed-order reduction of the tile partials
// ======================================================================================
struct ReduceArgs {
    const double* part;  // [tiles_c, ld]
    const double* gpart;  // [P, tiles_e, ld]
    int32_t tiles_c, tiles_e, ld;
    int32_t n;
    int32_t P, S, s;
    uint32_t used_mask;
    double* loglike;  // [n, S]   (already offset to the chunk's first row)
    double* grad;     // [n, S, P] or nullptr
};

__global__ void reduce_kernel(const ReduceArgs a) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= a.n) return;
    if (a.part != nullptr) {
        double s = 0.0;
        for (int t = 0; t < a.tiles_c; ++t) s += a.part[(size_t)t * a.ld + row];
        a.loglike[(size_t)row * a.S + a.s] = s;
    }
    if (a.grad != nullptr) {
        for (int j = 0; j < a.P; ++j) {
            double g = 0.0;
            if ((a.used_mask >> j) & 1u)
                for (int t = 0; t < a.tiles_e; ++t) g += a.gpart[((size_t)j * a.tiles_e + t) * a.ld + row];
            a.grad[((size_t)row * a.S + a.s) * a.P + j] = g;
        }
    }
}

// ======================================================================================
// per-channel sites (likelihood.py:206-225 etc.): elementwise on the folded counts
// ======================================================================================
struct SitesArgs {
    const double* X;  // [rows, ldx] folded R u
    int32_t ldx;
    ChanArrays ch;
    const double* inv_width;  // 1 / channel_width (or nullptr: rate = source_rate)
    const double* area;       // area_scale
    const double* counts_on;
    const double* counts_off;
    int64_t ld_counts;
    int32_t n, C, stat;
    double *rate, *model_on, *ll_on, *model_off, *ll_off;  // [n, C] each, nullable
};

template <int STAT>
__device__ __forceinline__ void sites_one(const SitesArgs& a, int row, int col) {
    ChannelData d;
    d.on = a.ch.on[col];
    d.off = a.ch.off[col];
    d.sigma = a.ch.sigma[col];
    d.ratio = a.ch.ratio[col];
    if (a.counts_on != nullptr) d.on = a.counts_on[(size_t)row * a.ld_counts + col];
    if (a.counts_off != nullptr) d.off = a.counts_off[(size_t)row * a.ld_counts + col];
    d.gof_on = poisson_gof(d.on);
    d.gof_off = poisson_gof(d.off);
    d.den = a.ch.den[col];
    const double fold = a.X[(size_t)row * a.ldx + col];
    double dx, site[4] = {0.0, 0.0, 0.0, 0.0};
    stat_channel<STAT, false, true>(fold * a.ch.scale[col], d, dx, site);
    const size_t o = (size_t)row * a.C + col;
    if (a.rate) a.rate[o] = fold * a.area[col] * (a.inv_width ? a.inv_width[col] : 1.0);
    if (a.model_on) a.model_on[o] = site[0];
    if (a.ll_on) a.ll_on[o] = site[1];
    if (a.model_off) a.model_off[o] = site[2];
    if (a.ll_off) a.ll_off[o] = site[3];
}

__global__ void sites_kernel(const SitesArgs a) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= a.C || row >= a.n) return;
    switch (a.stat) {
        case ELISA_STAT_CHI2: sites_one<ELISA_STAT_CHI2>(a, row, col); break;
        case ELISA_STAT_CSTAT: sites_one<ELISA_STAT_CSTAT>(a, row, col); break;
        case ELISA_STAT_PSTAT: sites_one<ELISA_STAT_PSTAT>(a, row, col); break;
        case ELISA_STAT_PGSTAT: sites_one<ELISA_STAT_PGSTAT>(a, row, col); break;
        default: sites_one<ELISA_STAT_WSTAT>(a, row, col); break;
    }
}

// ======================================================================================
// Gauss-Newton normal equations of the deviance residuals (batched re-fit, infer/helper.py:588-660)
// ======================================================================================
// The reference re-fits simulated data sets with Levenberg-Marquardt on the residual vector
// r_c = sqrt(-2 l_c) (helper.py:588-594, 601-625).  One LM step needs, per replica,
//   loglike = sum_c l_c,   grad_p = d loglike / d theta_p = -(J^T r)_p,
//   (J^T J)_pq = sum_c (dr_c/dtheta_p)(dr_c/dtheta_q) = sum_c (dl_c/dx)^2 / (-2 l_c) s_pc s_qc,
// with s_pc = d x_c / d theta_p = scale_c (R dU/dtheta_p)_c the FORWARD-folded tangent spectra
// (forward mode: no transposed fold).  One warp per replica, lanes stride the channels.
struct NormalEqArgs {
    const double* X;  // [1 + P planes][rows_pad, ldx]: plane 0 = R u, plane 1 + j = R du/dtheta_j
    int64_t plane;
    int32_t ldx;
    ChanArrays ch;
    const double* counts_on;
    const double* counts_off;
    int64_t ld_counts;
    int32_t n, C, P;
    uint32_t used_mask;
    int32_t accumulate;  // 0: overwrite the outputs (first spectrum), 1: add to them
    double* loglike;     // [n]
    double* grad;        // [n, P]
    double* jtj;         // [n, P, P]
    // accuracy guard of the integer fold (value side, as in stat_slice_kernel); guard == nullptr: off
    const double* u_scale;  // [rows] row scales of U's digits
    const double* colsum;   // [C] sum_e |t_e R[e, c]|
    unsigned long long* guard;
    double guard_eta, guard_tol;
};

template <int STAT, int PT>
__global__ void __launch_bounds__(256) normal_eq_kernel(const NormalEqArgs a) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= a.n) return;
    const double* x0 = a.X + (size_t)row * a.ldx;
    const double* con = a.counts_on ? a.counts_on + (size_t)row * a.ld_counts : nullptr;
    const double* coff = a.counts_off ? a.counts_off + (size_t)row * a.ld_counts : nullptr;
    double ll = 0.0, eb = 0.0, g[PT], h[PT * (PT + 1) / 2];
#pragma unroll
    for (int p = 0; p < PT; ++p) g[p] = 0.0;
#pragma unroll
    for (int k = 0; k < PT * (PT + 1) / 2; ++k) h[k] = 0.0;
    for (int col = lane; col < a.C; col += 32) {
        ChannelData d;
        const double sc = a.ch.scale[col];
        d.on = con ? con[col] : a.ch.on[col];
        d.off = coff ? coff[col] : a.ch.off[col];
        d.sigma = a.ch.sigma[col];
        d.ratio = a.ch.ratio[col];
        d.gof_on = con ? poisson_gof(d.on) : a.ch.gof_on[col];
        d.gof_off = coff ? poisson_gof(d.off) : a.ch.gof_off[col];
        d.den = a.ch.den[col];
        double dx = 0.0;
        const double l = stat_channel<STAT, true>(x0[col] * sc, d, dx);
        ll += l;
        if (a.guard != nullptr) eb = fma(fabs(dx * sc), a.colsum[col], eb);
        const double r2 = -2.0 * l;
        double w;  // (dr/dx)^2
        if (r2 > 1e-9) {
            w = dx * dx / r2;
        } else {
            // r -> 0 (model on the data): dx^2 / r^2 is 0/0 and l itself is rounding noise there;
            // its limit is the curvature -d2l/dx2, taken from a secant of dl/dx
            const double xs = x0[col] * sc, step = 1e-4 * fabs(xs);
            double dx2 = 0.0;
            if (step > 0.0) stat_channel<STAT, true>(xs + step, d, dx2);
            w = step > 0.0 ? fmax((dx - dx2) / step, 0.0) : 0.0;
        }
        double sp[PT];
#pragma unroll
        for (int p = 0; p < PT; ++p)
            sp[p] = (p < a.P && ((a.used_mask >> p) & 1u)) ? sc * x0[(size_t)(1 + p) * a.plane + col] : 0.0;
        int k = 0;
#pragma unroll
        for (int p = 0; p < PT; ++p) {
            g[p] = fma(dx, sp[p], g[p]);
            const double wp = w * sp[p];
#pragma unroll
            for (int q = p; q < PT; ++q, ++k) h[k] = fma(wp, sp[q], h[k]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ll += __shfl_xor_sync(0xffffffffu, ll, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
#pragma unroll
        for (int p = 0; p < PT; ++p) g[p] += __shfl_xor_sync(0xffffffffu, g[p], o);
#pragma unroll
        for (int k = 0; k < PT * (PT + 1) / 2; ++k) h[k] += __shfl_xor_sync(0xffffffffu, h[k], o);
    }
    if (lane != 0) return;
    if (a.guard != nullptr) {
        const double est = a.guard_eta * a.u_scale[row] * (double)I8_TOP_DIGIT * eb / fmax(fabs(ll), 1.0);
        if (est > a.guard_tol) atomicAdd(a.guard, 1ull);
        if (est > 0.0 && est < 1.79e308) atomicMax(a.guard + 1, (unsigned long long)__double_as_longlong(est));
    }
    a.loglike[row] = (a.accumulate ? a.loglike[row] : 0.0) + ll;
    int k = 0;
#pragma unroll
    for (int p = 0; p < PT; ++p) {
        if (p < a.P) {
            double* gp = a.grad + (size_t)row * a.P + p;
            *gp = (a.accumulate ? *gp : 0.0) + g[p];
        }
#pragma unroll
        for (int q = p; q < PT; ++q, ++k) {
            if (q < a.P) {
                double* hpq = a.jtj + ((size_t)row * a.P + p) * a.P + q;
                double* hqp = a.jtj + ((size_t)row * a.P + q) * a.P + p;
                const double v = (a.accumulate ? *hpq : 0.0) + h[k];
                *hpq = v;
                *hqp = v;
            }
        }
    }
}

// ======================================================================================
// row-wise sparse fold: X[n, c] = sum_k val[c, k] * U[n, idx[c, k]]   (CSR by channel row)
// ======================================================================================
// The north star's "warp-per-channel-row" kernel for band-sparse responses, in the mapping that
// keeps its loads coalesced with U stored [n, E]: a block owns 128 consecutive channel rows and
// CSR_RT replicas, one THREAD per channel row (a warp = 32 neighbouring rows, whose column indices
// are neighbours too in a banded response, so the 32 loads of U[n, idx] fall into 2-3 sectors), every
// value / index loaded once and used for CSR_RT replicas.  4 nnz useful flop per replica, FP64 FMA
// pipe.  Raced against the block-banded DMMA fold in tools/csr_race.py (profiles/r02_cfg4.md).
constexpr int CSR_RT = 8;
__global__ void __launch_bounds__(128) csr_fold_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                                                       const double* __restrict__ val, const double* __restrict__ U,
                                                       int64_t ldu, int n, int C, double* __restrict__ X, int64_t ldx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = blockIdx.y * CSR_RT;
    if (c >= C) return;
    const int k0 = ptr[c], k1 = ptr[c + 1];
    double acc[CSR_RT];
#pragma unroll
    for (int r = 0; r < CSR_RT; ++r) acc[r] = 0.0;
    const double* u = U + (size_t)n0 * ldu;
    const int rows = min(CSR_RT, n - n0);
    if (rows == CSR_RT) {
        for (int k = k0; k < k1; ++k) {
            const double v = val[k];
            const double* ue = u + idx[k];
#pragma unroll
            for (int r = 0; r < CSR_RT; ++r) acc[r] = fma(v, ue[(size_t)r * ldu], acc[r]);
        }
    } else {
        for (int k = k0; k < k1; ++k) {
            const double v = val[k];
            const double* ue = u + idx[k];
            for (int r = 0; r < rows; ++r) acc[r] = fma(v, ue[(size_t)r * ldu], acc[r]);
        }
    }
    for (int r = 0; r < rows; ++r) X[(size_t)(n0 + r) * ldx + c] = acc[r];
}

// ======================================================================================
// host side: plan
// ======================================================================================
static thread_local std::string g_last_error = "";

// Post-mortem record of the fold kernel's bounded barrier waits: host-mapped, so it survives the
// launch failure a protocol error ends in (fold_i8.cuh: mbar_wait).
static int32_t* g_trap_host = nullptr;
static int32_t* g_trap_dev = nullptr;
static int32_t* trap_info_dev() {
    static std::once_flag once;
    std::call_once(once, [] {
        if (cudaHostAlloc((void**)&g_trap_host, 512, cudaHostAllocMapped) == cudaSuccess) {
            memset(g_trap_host, 0, 512);
            if (cudaHostGetDevicePointer((void**)&g_trap_dev, g_trap_host, 0) != cudaSuccess) g_trap_dev = nullptr;
        }
    });
    return g_trap_dev;
}
static std::string trap_report() {
    if (g_trap_host == nullptr || g_trap_host[0] == 0) return "";
    char buf[160];
    snprintf(buf, sizeof buf, " [fold kernel gave up waiting: block %d warp %d barrier code %d parity %d tile %d;",
             g_trap_host[1], g_trap_host[2], g_trap_host[3], g_trap_host[4], g_trap_host[5]);
    std::string out = buf;
    snprintf(buf, sizeof buf, " stalled warps of block %d (warp:code/parity/tile):", g_trap_host[6] - 1);
    out += buf;
    for (int w = 0; w < 24; ++w) {
        const int32_t* st = g_trap_host + 8 + 3 * w;
        if (st[0] == 0) continue;
        snprintf(buf, sizeof buf, " %d:%d/%d/%d", w, st[0], st[1], st[2]);
        out += buf;
    }
    return out + "]";
}

static int fail(int code, const std::string& msg) {
    g_last_error = code == ELISA_ERR_CUDA ? msg + trap_report() : msg;
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(ELISA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));            \
    } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct PackedHost {
    std::vector<double> data;
    std::vector<int64_t> off;
    std::vector<int32_t> k_lo, k_hi;
};

struct PackedDevice {
    double* data = nullptr;
    int64_t* off = nullptr;
    int32_t* k_lo = nullptr;
    int32_t* k_hi = nullptr;
    int n_tiles = 0;
    int64_t n_doubles = 0;
    PackedB view() const { return PackedB{data, off, k_lo, k_hi}; }
};

// Static (B) operand of the sliced-integer fold: digit images per (64-row tile, k-block).
struct I8Operand {
    double* dense = nullptr;      // [rows, K] FP64 copy the digits are (re)sliced from (balancing)
    int rows = 0, K = 0;
    int8_t* img = nullptr;
    double* abs_sum = nullptr;    // [n_tiles * 64] sum_k |row| (accuracy guard)
    double* scale = nullptr;      // [n_tiles * 64]
    int64_t* tile_img = nullptr;  // [n_tiles] first image of the tile
    int32_t* kb_lo = nullptr;     // [n_tiles] stored k-block range
    int32_t* kb_hi = nullptr;
    int n_tiles = 0;
    int64_t n_images = 0;
    int ks = 0;  // slices per element of this operand
    uint8_t* lead = nullptr;  // [n_images] leading all-zero slices of each image (lead_zero_kernel)
    std::vector<int32_t> h_kb_lo, h_kb_hi;  // host copies of the stored k-block ranges (work accounting)
};

struct SpectrumPlan {
    int E = 0, C = 0, E_pad = 0, C_pad = 0;
    int stat = 0;
    ModelDev model;
    // device arrays
    double *egrid = nullptr, *lngrid = nullptr, *emid = nullptr, *lnmid = nullptr;
    double* tab[2 * ELISA_MAX_COMPONENTS] = {};  // sigma(E) of table components at edges / midpoints
    // flux normalisations: their grids (egrid, lngrid, emid, lnmid), EnFlux weights, tables
    struct NormGrid {
        double *egrid = nullptr, *lngrid = nullptr, *emid = nullptr, *lnmid = nullptr, *w = nullptr;
        double* tab[2 * ELISA_MAX_COMPONENTS] = {};
        int E = 0;
    } norm_grid[ELISA_MAX_NORMS];
    double* chan = nullptr;  // 10 arrays of C_pad: scale on off sigma ratio gof_on gof_off area inv_width den
    PackedDevice fwd, bwd;   // R [E, C] tiled over C ; R^T [C, E] tiled over E
    I8Operand i8_fwd, i8_bwd;  // digit images of R^T rows (channels) / R rows (energies); img == nullptr: DMMA fold
    bool use_i8() const { return i8_fwd.img != nullptr; }
    int64_t counts_off = 0;  // column offset of this spectrum in the per-replica counts arrays
    int64_t nnz = 0;
    // CSR by channel row of R^T (what BCSR.from_scipy_sparse(sparse.T) holds, likelihood.py:177-181),
    // kept for the row-wise sparse fold (csr_fold_kernel); built from the non-zeros of a dense response too
    int32_t *csr_ptr = nullptr, *csr_idx = nullptr;
    double* csr_val = nullptr;
    double fill = 1.0;  // stored fraction of the packed response operands
    // small spectra: the dense response [E, C] for the one-kernel small-batch path (tiny_eval, unfold_skel.cuh);
    // nullptr: the spectrum does not qualify (see tiny_shape_ok)
    double* tiny_R = nullptr;
    // balancing of the integer folds' contraction indices (col_envelope_kernel): power-of-two
    // envelopes of U over the energies / of W over the channels, and their reciprocals
    double *t_scale = nullptr, *t_inv = nullptr, *r_scale = nullptr, *r_inv = nullptr;
    uint8_t *t_exp = nullptr, *r_exp = nullptr;  // log2 of t_inv / r_inv
    bool balanced_fwd = false, balanced_bwd = false;
    bool stale = false;  // elisa_b200_plan_rebalance(): derive the envelopes afresh at the next call
    const struct JitKernels* jit = nullptr;  // runtime-specialised unfold (nullptr: interpreter)
    GridView grid_view() const {
        GridView g{egrid, lngrid, emid, lnmid, E};
        for (int i = 0; i < 2 * ELISA_MAX_COMPONENTS; ++i) g.tab[i] = tab[i];
        return g;
    }
    GridView norm_view(int k) const {
        const NormGrid& n = norm_grid[k];
        GridView g{n.egrid, n.lngrid, n.emid, n.lnmid, n.E};
        for (int i = 0; i < 2 * ELISA_MAX_COMPONENTS; ++i) g.tab[i] = n.tab[i];
        return g;
    }
    ChanArrays chan_view() const {
        return ChanArrays{chan, chan + C_pad, chan + 2 * C_pad, chan + 3 * C_pad,
                          chan + 4 * C_pad, chan + 5 * C_pad, chan + 6 * C_pad, chan + 9 * C_pad};
    }
};

}  // namespace elisa

using namespace elisa;

struct ElisaPlan {
    // Calls on one plan share its workspace: they are serialised here, so that a plan may be used
    // from several host threads (numpyro chain_method='parallel'); for calls that actually overlap
    // on the device create one plan per thread / stream.
    std::mutex mu;
    int device = 0;
    int P = 0, S = 0;
    int64_t chunk = 0;
    int64_t sum_C = 0;
    int max_E_pad = 0, max_C_pad = 0;
    std::vector<SpectrumPlan> spec;
    // workspace
    double *U = nullptr, *dU = nullptr, *W = nullptr, *part = nullptr, *gpart = nullptr;
    double* G = nullptr;  // [chunk, max_E_pad] d loglike / d u of the transposed fold (gradient by contraction)
    bool grad_contract = true;  // integer fold: gradient by contraction instead of tangent planes (ELISA_B200_GRAD=planes: off)
    // sliced-integer fold: digit images of the per-chunk A operand (U, then W) and its row scales
    int ks = 0;      // slices per operand of the forward fold (0: every spectrum folds with DMMA)
    int ks_bwd = 0;  // slices per operand of the transposed fold
    int8_t* A_img = nullptr;
    double* a_scale = nullptr;
    unsigned long long* guard = nullptr;  // [4] accuracy guard of the integer fold (stat_slice_kernel)
    int32_t* flag_list = nullptr;         // [flag_cap] rows flagged by the guard (global row indices)
    int flag_cap = 0;
    int64_t row_base = 0;                 // global index of the first row of the chunk being evaluated
    unsigned long long* guard_host = nullptr;  // pinned [4]: asynchronous snapshots of `guard`
    cudaEvent_t guard_event = nullptr;
    bool guard_snapshot_pending = false;
    int64_t fallback_rows = 0;            // rows answered by the DMMA fold (per-row fallback)
    int64_t renewals = 0;                 // envelope renewals triggered by the guard
    int debug_skip = 0;                   // timing experiments: kernel classes to skip
    double* a_scale_u = nullptr;          // [chunk] row scales of U kept across the tangent folds (normal equations guard)
    double* Xp = nullptr;  // [1 + P][chunk, max_C_pad] forward-folded model and tangents (normal equations; lazily allocated)
    // Small batches (NUTS chains, single fits) are launch-bound: ~5 kernels per spectrum of a few
    // microseconds each.  The launch sequence of a (n, with gradient) shape is captured into a CUDA
    // graph on its second use and replayed afterwards; the caller's buffers enter and leave
    // through memcpy nodes whose addresses are patched per call, so the graph never goes stale.
    struct GraphEntry {
        int64_t n = 0;
        bool grad = false;
        int uses = 0;
        int64_t kernels = 0;  // kernel launches one replay stands for
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        cudaGraphNode_t in_theta = nullptr, out_ll = nullptr, out_grad = nullptr;
    };
    std::vector<GraphEntry> graphs;
    double *g_theta = nullptr, *g_ll = nullptr, *g_grad = nullptr;  // staging, GRAPH_MAX_N rows
    bool graphs_enabled = true;
    bool guard_off = false;   // ELISA_FLAG_FOLD_I8: the caller asked for the integer fold, unguarded
    int64_t guard_fallbacks = 0;
    bool force_dmma = false;  // evaluate with the FP64 DMMA fold although the integer operands exist (guard fallback)
    bool balance = true;      // balance the integer folds around the first batch seen (ELISA_B200_BALANCE=0: off)
    bool renew_envelopes = true;  // host entry point: answer a flagged call by renewing the envelopes first
    double* cal_ll = nullptr; // [128 * S] log-likelihoods of the pilot rows (scratch)
    double *cal_theta = nullptr, *cal_on = nullptr, *cal_off = nullptr;  // the gathered pilot rows
    int ld_part = 0;  // leading dimension (rows) of the part / gpart arrays in use
    // Joint fits: inside the captured graph every spectrum runs on its own branch (own stream
    // during capture, own small workspace), so the S kernel chains of a small batch overlap
    // instead of queueing behind one another.
    struct SmallWs {
        double *U = nullptr, *dU = nullptr, *W = nullptr, *part = nullptr, *gpart = nullptr, *a_scale = nullptr, *G = nullptr;
        int8_t* A_img = nullptr;
    };
    std::vector<SmallWs> small_ws;
    std::vector<cudaStream_t> small_stream;
    std::vector<cudaEvent_t> small_done;
    cudaEvent_t small_fork = nullptr;
    int sm_count = 148;
    // one-kernel path for small batches of small spectra (tiny_eval): the decision is taken per CALL
    // (tiny_call_n >= 0: rows of the whole call when it is evaluated in pieces), so that re-chunking a
    // batch never changes which kernels compute a row
    bool tiny_enabled = true, tiny_now = false;
    int64_t tiny_call_n = -1;
    TinyArgs* tiny_args = nullptr;  // [S] device array (entries of spectra that do not qualify are unused)
    int64_t workspace_bytes = 0;
    int64_t launches = 0;
    // host-path staging (device side), grown on demand: theta / loglike / grad for the whole call,
    // per-replica counts double-buffered per piece of `chunk` rows (H2D of piece k + 1 under the
    // evaluation of piece k)
    double *h_theta = nullptr, *h_ll = nullptr, *h_grad = nullptr;
    int64_t h_n = 0;
    double *pipe_on[2] = {nullptr, nullptr}, *pipe_off[2] = {nullptr, nullptr};
    cudaEvent_t pipe_free[2] = {nullptr, nullptr}, pipe_ready[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaStream_t stream = nullptr;
    // peer gather (elisa_b200_plan_set_gather): full result arrays of every rank as mapped here
    struct PeerGather {
        int world = 0, rank = 0;
        std::vector<double*> ll, grad;
        std::vector<long long*> flags;
        int64_t row_offset = 0;
        long long epoch = 0;
        cudaStream_t copy = nullptr;
        cudaEvent_t ready = nullptr, pushed = nullptr;
        int64_t pend_lo = 0, pend_hi = 0;  // rows computed and not pushed yet
        int pend_chunks = 0;
        int every = 4;                     // chunks per push
        unsigned int* done = nullptr;      // device counter of the small-batch kernel's fused signal
        bool signalled = false;            // this call's flags were raised by the evaluation kernel itself
        int32_t* timeout = nullptr;        // host-mapped: a peer's flag did not arrive
        unsigned long long limit_ns = 60ull * 1000000000ull;  // how long a call waits for a peer (ranks may be skewed by host work)
    } gather;
    std::vector<void*> owned;
    // optional per-kernel-class device timing (CUDA events on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Span { int cls; cudaEvent_t a, b; };
    std::vector<Span> spans;
    double prof_ms[ELISA_PROFILE_CLASSES] = {};
    int64_t prof_n[ELISA_PROFILE_CLASSES] = {};
};

namespace elisa {

template <typename T>
static int upload(ElisaPlan* plan, const std::vector<T>& h, T** out) {
    *out = nullptr;
    const size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
    CUDA_TRY(cudaMalloc((void**)out, bytes));
    plan->owned.push_back(*out);
    if (!h.empty()) CUDA_TRY(cudaMemcpy(*out, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ELISA_OK;
}

static int upload_packed(ElisaPlan* plan, const PackedHost& h, PackedDevice* d) {
    int rc;
    if ((rc = upload(plan, h.data, &d->data))) return rc;
    if ((rc = upload(plan, h.off, &d->off))) return rc;
    if ((rc = upload(plan, h.k_lo, &d->k_lo))) return rc;
    if ((rc = upload(plan, h.k_hi, &d->k_hi))) return rc;
    d->n_tiles = (int)h.off.size();
    d->n_doubles = (int64_t)h.data.size();
    return ELISA_OK;
}

// Pack a K x N matrix given by an element accessor into 64-column tiles holding rows
// [k_lo, k_hi) between the first and last row with a non-zero (16-aligned).
// get_col_range(n) -> (first k, one past last k) of the non-zeros of column n; fill(k, n).
template <class Range, class Get>
static void pack_tiles(int K, int N, int K_pad, Range col_range, Get get, PackedHost* out) {
    const int n_tiles = round_up(N, GEMM_BN) / GEMM_BN;
    out->off.resize(n_tiles);
    out->k_lo.resize(n_tiles);
    out->k_hi.resize(n_tiles);
    int64_t total = 0;
    for (int j = 0; j < n_tiles; ++j) {
        int lo = K, hi = 0;
        for (int n = j * GEMM_BN; n < std::min(N, (j + 1) * GEMM_BN); ++n) {
            int a, b;
            col_range(n, &a, &b);
            if (a < b) {
                lo = std::min(lo, a);
                hi = std::max(hi, b);
            }
        }
        if (lo >= hi) lo = hi = 0;
        lo = lo / GEMM_BK * GEMM_BK;
        hi = std::min(round_up(hi, GEMM_BK), K_pad);
        out->k_lo[j] = lo;
        out->k_hi[j] = hi;
        out->off[j] = total;
        total += (int64_t)(hi - lo) * GEMM_BN;
    }
    out->data.assign((size_t)total, 0.0);
    for (int j = 0; j < n_tiles; ++j) {
        double* dst = out->data.data() + out->off[j];
        const int lo = out->k_lo[j], hi = std::min(out->k_hi[j], K);
        for (int n = j * GEMM_BN; n < std::min(N, (j + 1) * GEMM_BN); ++n)
            get(n, lo, hi, dst + (n - j * GEMM_BN));  // writes dst[(k - lo) * 64] for k in [lo, hi)
    }
}

static int build_model(const ElisaModelDesc& md, int P, ModelDev* out) {
    memset(out, 0, sizeof(*out));
    if (md.n_components < 1 || md.n_components > ELISA_MAX_COMPONENTS || md.components == nullptr)
        return fail(ELISA_ERR_INVALID, "model: n_components out of range");
    out->n_leaf = md.n_components;
    for (int i = 0; i < md.n_components; ++i) {
        const ElisaComponentDesc& c = md.components[i];
        const int np = comp_num_params(c.kind);
        if (np < 0) return fail(ELISA_ERR_INVALID, "model: unknown component kind");
        if (c.method != ELISA_METHOD_TRAPZ && c.method != ELISA_METHOD_SIMPSON)
            return fail(ELISA_ERR_INVALID, "model: unknown integration method");
        LeafDev& L = out->leaf[i];
        L.kind = c.kind;
        L.method = c.method;
        L.np = np;
        for (int k = 0; k < ELISA_MAX_COMP_PARAMS; ++k) {
            L.src[k] = -1;
            L.cst[k] = 0.0;
        }
        for (int k = 0; k < np; ++k) {
            if (c.param_src[k] >= P) return fail(ELISA_ERR_INVALID, "model: param_src >= n_params");
            L.src[k] = c.param_src[k] < 0 ? -1 : c.param_src[k];
            L.cst[k] = c.param_const[k];
            if (L.src[k] >= 0) out->used_mask |= 1u << L.src[k];
        }
        L.aux[0] = c.aux[0];
        L.aux[1] = c.aux[1];
        if ((c.kind == ELISA_COMP_PLENFLUX || c.kind == ELISA_COMP_PLPHFLUX) && !(c.aux[0] > 0.0 && c.aux[1] > c.aux[0]))
            return fail(ELISA_ERR_INVALID, "model: PLPhFlux/PLEnFlux need 0 < emin < emax in aux");
        auto one = [](double v) { return v == 0.0 ? 1.0 : v; };
        L.gs = one(c.grid_scale);
        L.os = one(c.out_scale);
        L.ngs = one(c.norm_grid_scale);
        L.nos = one(c.norm_out_scale);
        if (!(L.gs > 0.0 && L.gs < HUGE_VAL) || !(L.ngs > 0.0 && L.ngs < HUGE_VAL) || !(fabs(L.os) < HUGE_VAL) ||
            !(fabs(L.nos) < HUGE_VAL))
            return fail(ELISA_ERR_INVALID, "model: grid_scale must be positive and finite, out_scale finite");
        L.lgs = L.gs == 1.0 ? 0.0 : log(L.gs);
        L.lngs = L.ngs == 1.0 ? 0.0 : log(L.ngs);
        L.n_shift = 0;
        for (int k = 0; k < ELISA_MAX_SHIFTS; ++k) {
            if (c.shift_kind[k] == 0) continue;
            if (c.shift_kind[k] != ELISA_SHIFT_Z && c.shift_kind[k] != ELISA_SHIFT_ZA && c.shift_kind[k] != ELISA_SHIFT_V)
                return fail(ELISA_ERR_INVALID, "model: unknown shift_kind");
            if (c.shift_src[k] < 0 || c.shift_src[k] >= P) return fail(ELISA_ERR_INVALID, "model: shift_src out of range");
            if (c.kind == ELISA_COMP_PHOTONABS)
                return fail(ELISA_ERR_UNSUPPORTED, "model: a free energy shift around a table component");
            if (md.n_norms > 0)
                return fail(ELISA_ERR_UNSUPPORTED, "model: free energy shifts together with flux normalisations");
            L.shift_kind[L.n_shift] = c.shift_kind[k];
            L.shift_src[L.n_shift] = c.shift_src[k];
            ++L.n_shift;
            out->used_mask |= 1u << c.shift_src[k];
        }
        if (c.kind == ELISA_COMP_PHOTONABS) {
            if (c.table_energy == nullptr || c.table_xsect == nullptr || c.table_len < 2)
                return fail(ELISA_ERR_INVALID, "model: PhotonAbsorption needs a cross-section table (>= 2 points)");
            for (int k = 0; k < c.table_len; ++k) {
                const bool ok = c.table_log ? (std::isfinite(c.table_energy[k]) && std::isfinite(c.table_xsect[k]))
                                            : (c.table_energy[k] > 0.0 && c.table_xsect[k] > 0.0);
                if (!ok || (k > 0 && !(c.table_energy[k] >= c.table_energy[k - 1])))
                    return fail(ELISA_ERR_INVALID,
                                "model: cross-section table must be positive with non-decreasing energies");
            }
            if (!(c.table_energy[c.table_len - 1] > c.table_energy[0]))
                return fail(ELISA_ERR_INVALID, "model: cross-section table spans no energy range");
        }
    }
    // postfix program; a single component may omit it
    if (md.n_ops == 0 && md.n_components == 1 && md.n_norms == 0) {
        out->n_ops = 1;
        out->ops[0] = 0;
        return ELISA_OK;
    }
    if (md.n_ops < 1 || md.n_ops > MAX_OPS || md.ops == nullptr)
        return fail(ELISA_ERR_INVALID, "model: n_ops out of range");
    if (md.n_norms < 0 || md.n_norms > ELISA_MAX_NORMS || (md.n_norms > 0 && md.norms == nullptr))
        return fail(ELISA_ERR_INVALID, "model: n_norms out of range");
    out->n_norms = md.n_norms;
    for (int k = 0; k < md.n_norms; ++k) {
        const ElisaNormDesc& nd = md.norms[k];
        if (nd.kind != ELISA_NORM_PHFLUX && nd.kind != ELISA_NORM_ENFLUX)
            return fail(ELISA_ERR_INVALID, "model: unknown flux normalisation");
        if (nd.f_src >= P) return fail(ELISA_ERR_INVALID, "model: norm f_src >= n_params");
        if (nd.grid == nullptr || nd.n_grid < 2) return fail(ELISA_ERR_INVALID, "model: flux grid needs >= 2 points");
        for (int j = 0; j < nd.n_grid; ++j)
            if (!(nd.grid[j] > 0.0) || (j > 0 && !(nd.grid[j] > nd.grid[j - 1])))
                return fail(ELISA_ERR_INVALID, "model: flux grid must be positive and increasing");
        out->norm[k] = NormDev{nd.kind, nd.f_src < 0 ? -1 : nd.f_src, -1, -1, nd.f_const};
        if (nd.f_src >= 0) out->used_mask |= 1u << nd.f_src;
    }
    int depth = 0;
    // type check as models/model.py:1228-1258: '+' needs add, add; '*' forbids add * add
    std::vector<int> types;   // 1 = additive, 0 = multiplicative
    std::vector<int> starts;  // first op of the sub-expression on each stack level
    std::vector<int> normed;  // that sub-expression contains a flux normalisation
    for (int o = 0; o < md.n_ops; ++o) {
        const int op = md.ops[o];
        if (op >= 0) {
            if (op >= md.n_components) return fail(ELISA_ERR_INVALID, "model: op references unknown component");
            types.push_back(comp_is_additive(md.components[op].kind) ? 1 : 0);
            starts.push_back(o);
            normed.push_back(0);
            ++depth;
        } else if (op <= ELISA_OP_NORM0 && op > ELISA_OP_NORM0 - ELISA_MAX_NORMS) {
            const int k = ELISA_OP_NORM0 - op;
            if (k >= md.n_norms) return fail(ELISA_ERR_INVALID, "model: op references unknown flux normalisation");
            if (depth < 1) return fail(ELISA_ERR_INVALID, "model: malformed postfix program");
            if (types.back() != 1)  // conv.py:24 _supported = {'add'}
                return fail(ELISA_ERR_INVALID, "model: PhFlux / EnFlux need an additive operand");
            if (normed.back()) return fail(ELISA_ERR_UNSUPPORTED, "model: nested flux normalisations");
            if (out->norm[k].op_hi >= 0) return fail(ELISA_ERR_INVALID, "model: flux normalisation used twice");
            out->norm[k].op_lo = starts.back();
            out->norm[k].op_hi = o;
            normed.back() = 1;
        } else if (op == ELISA_OP_ADD || op == ELISA_OP_MUL) {
            if (depth < 2) return fail(ELISA_ERR_INVALID, "model: malformed postfix program");
            const int r = types.back();
            types.pop_back();
            const int l = types.back();
            types.pop_back();
            if (op == ELISA_OP_ADD && !(l == 1 && r == 1))
                return fail(ELISA_ERR_INVALID, "model: '+' needs two additive operands");
            if (op == ELISA_OP_MUL && l == 1 && r == 1)
                return fail(ELISA_ERR_INVALID, "model: '*' of two additive operands");
            types.push_back(op == ELISA_OP_ADD ? 1 : (l | r));
            const int nr = normed.back();
            normed.pop_back();
            starts.pop_back();
            normed.back() |= nr;
            --depth;
        } else {
            return fail(ELISA_ERR_INVALID, "model: unknown opcode");
        }
        if (depth > ELISA_MAX_COMPONENTS) return fail(ELISA_ERR_INVALID, "model: expression too deep");
        out->ops[o] = op;
    }
    if (depth != 1) return fail(ELISA_ERR_INVALID, "model: malformed postfix program");
    for (int k = 0; k < md.n_norms; ++k)
        if (out->norm[k].op_hi < 0) return fail(ELISA_ERR_INVALID, "model: flux normalisation without its opcode");
    out->n_ops = md.n_ops;
    return ELISA_OK;
}

// Shapes the one-kernel small-batch path takes: every thread folds whole response columns in FP64 with
// 1 + (free parameters) accumulators, u and du/dtheta of one parameter vector sit in shared memory.
constexpr int64_t TINY_MAX_N = 1024;          // rows per call
constexpr int64_t TINY_MAX_WORK = 1 << 18;    // E * C * (1 + used parameters) per row
static int popcount32(uint32_t v) {
    int n = 0;
    for (; v; v &= v - 1) ++n;
    return n;
}
// launch shape of the small-batch kernel: compiled defaults, or ELISA_B200_TINY_TUNE="threads,channels per thread,unroll"
// (threads a multiple of 32 and >= 32 * channels per thread; 0 channels = by the number of parameters; the same numbers
// are written into the generated translation unit)
struct TinyTune {
    int threads = TINY_THREADS, ch = 0, unroll = TINY_UNROLL;
    bool custom = false;
};
static const TinyTune& tiny_tune() {
    static const TinyTune t = [] {
        TinyTune v;
        if (const char* env = getenv("ELISA_B200_TINY_TUNE")) {
            int a = 0, b = 0, c = 0;
            if (sscanf(env, "%d,%d,%d", &a, &b, &c) == 3 && b >= 0 && b <= 8 && a >= 32 * std::max(b, 4) && a <= 1024 && a % 32 == 0 &&
                c >= 1 && c <= 16) {
                v.threads = a;
                v.ch = b;
                v.unroll = c;
                v.custom = true;
            }
        }
        return v;
    }();
    return t;
}
constexpr size_t TINY_MAX_SMEM = 96 * 1024;  // opt-in dynamic shared memory of the kernel (two CTAs per SM still fit)
static size_t tiny_smem_bytes(int E, int P) { return tiny_smem_doubles(E, P, tiny_tune().threads, tiny_tune().ch) * 8; }
static bool tiny_shape_ok(int E, int C, int P, uint32_t used_mask) {
    return P <= 8 && (int64_t)E * C * (1 + popcount32(used_mask)) <= TINY_MAX_WORK && tiny_smem_bytes(E, P) <= TINY_MAX_SMEM;
}

static int build_spectrum(ElisaPlan* plan, const ElisaSpectrumDesc& sd, SpectrumPlan* sp) {
    const int E = sd.n_energies, C = sd.n_channels;
    if (E < 1 || C < 1) return fail(ELISA_ERR_INVALID, "spectrum: empty energy grid or channel set");
    if (sd.photon_egrid == nullptr || sd.area_scale == nullptr || sd.on_counts == nullptr)
        return fail(ELISA_ERR_INVALID, "spectrum: photon_egrid, area_scale and on_counts are required");
    const bool dense = sd.response_dense != nullptr;
    const bool sparse = sd.csr_indptr != nullptr && sd.csr_indices != nullptr && sd.csr_values != nullptr;
    if (dense == sparse) return fail(ELISA_ERR_INVALID, "spectrum: give exactly one of response_dense / csr_*");
    if (sd.stat < ELISA_STAT_CHI2 || sd.stat > ELISA_STAT_WSTAT) return fail(ELISA_ERR_INVALID, "spectrum: unknown stat");
    if (sd.stat == ELISA_STAT_CHI2 && sd.on_error == nullptr)
        return fail(ELISA_ERR_INVALID, "chi2 needs on_error (net_errors)");
    if (sd.stat >= ELISA_STAT_PSTAT && (sd.off_counts == nullptr || sd.back_ratio == nullptr))
        return fail(ELISA_ERR_INVALID, "pstat/pgstat/wstat need off_counts and back_ratio");
    if (sd.stat == ELISA_STAT_PGSTAT && sd.off_error == nullptr)
        return fail(ELISA_ERR_INVALID, "pgstat needs off_error (back_errors)");
    for (int e = 0; e <= E; ++e)
        if (!(sd.photon_egrid[e] > 0.0) || (e > 0 && !(sd.photon_egrid[e] > sd.photon_egrid[e - 1])))
            return fail(ELISA_ERR_INVALID, "spectrum: photon_egrid must be positive and increasing");

    sp->E = E;
    sp->C = C;
    sp->E_pad = round_up(E, GEMM_BN);
    sp->C_pad = round_up(C, GEMM_BN);
    sp->stat = sd.stat;
    int rc;
    if ((rc = build_model(sd.model, plan->P, &sp->model))) return rc;

    // grids
    std::vector<double> eg(sd.photon_egrid, sd.photon_egrid + E + 1), lg(E + 1), em(E), lm(E);
    for (int e = 0; e <= E; ++e) lg[e] = log(eg[e]);
    for (int e = 0; e < E; ++e) {
        em[e] = 0.5 * (eg[e] + eg[e + 1]);
        lm[e] = log(em[e]);
    }
    if ((rc = upload(plan, eg, &sp->egrid))) return rc;
    if ((rc = upload(plan, lg, &sp->lngrid))) return rc;
    if ((rc = upload(plan, em, &sp->emid))) return rc;
    if ((rc = upload(plan, lm, &sp->lnmid))) return rc;
    // table components: sigma on the photon grid, _make_interp of models/mul.py:274-283 --
    // exp(interp(log e, log energy, log xsect, left='extrapolate', right='extrapolate'))
    for (int i = 0; i < sd.model.n_components; ++i) {
        const ElisaComponentDesc& c = sd.model.components[i];
        if (c.kind != ELISA_COMP_PHOTONABS) continue;
        const int n = c.table_len;
        std::vector<double> xp(n), fp(n);
        for (int k = 0; k < n; ++k) {
            xp[k] = c.table_log ? c.table_energy[k] : log(c.table_energy[k]);
            fp[k] = c.table_log ? c.table_xsect[k] : log(c.table_xsect[k]);
        }
        auto sigma = [&](double lx) {
            int j = (int)(std::upper_bound(xp.begin(), xp.end(), lx) - xp.begin()) - 1;  // xp[j] <= lx < xp[j+1]
            j = std::max(0, std::min(j, n - 2));                                          // end segments extrapolate
            const double dx = xp[j + 1] - xp[j];
            // repeated table energies: only a clipped end segment can have zero width; jnp.interp
            // returns the left value there (jax/_src/numpy/lax_numpy.py, the dx0 branch)
            if (fabs(dx) <= 4.9303806576313238e-32) return exp(fp[j]);
            return exp(fp[j] + ((lx - xp[j]) / dx) * (fp[j + 1] - fp[j]));
        };
        // the component sees the grid scaled by the energy shifts around it
        auto tables = [&](const std::vector<double>& g, double gs, double** edges, double** mids) -> int {
            const int n_e = (int)g.size() - 1;
            std::vector<double> se(n_e + 1), sm(n_e);
            for (int e = 0; e <= n_e; ++e) se[e] = sigma(log(gs * g[e]));
            for (int e = 0; e < n_e; ++e) sm[e] = sigma(log(gs * (0.5 * (g[e] + g[e + 1]))));
            int r;
            if ((r = upload(plan, se, edges))) return r;
            return upload(plan, sm, mids);
        };
        if ((rc = tables(eg, sp->model.leaf[i].gs, &sp->tab[2 * i], &sp->tab[2 * i + 1]))) return rc;
        for (int k = 0; k < sd.model.n_norms; ++k) {
            const NormDev& nd = sp->model.norm[k];
            bool inside = false;
            for (int o = nd.op_lo; o < nd.op_hi; ++o) inside |= sp->model.ops[o] == i;
            if (!inside) continue;
            const std::vector<double> ng(sd.model.norms[k].grid, sd.model.norms[k].grid + sd.model.norms[k].n_grid);
            if ((rc = tables(ng, sp->model.leaf[i].ngs, &sp->norm_grid[k].tab[2 * i], &sp->norm_grid[k].tab[2 * i + 1])))
                return rc;
        }
    }
    // flux normalisations: their own grids; EnFlux weights keV_to_erg * sqrt(e_j e_j+1) (conv.py:251-254)
    for (int k = 0; k < sd.model.n_norms; ++k) {
        const ElisaNormDesc& nd = sd.model.norms[k];
        const int n_e = nd.n_grid - 1;
        std::vector<double> g(nd.grid, nd.grid + nd.n_grid), l(n_e + 1), m(n_e), ml(n_e), w(n_e);
        for (int e = 0; e <= n_e; ++e) l[e] = log(g[e]);
        for (int e = 0; e < n_e; ++e) {
            m[e] = 0.5 * (g[e] + g[e + 1]);
            ml[e] = log(m[e]);
            w[e] = 1.602176634e-9 * sqrt(g[e] * g[e + 1]);
        }
        SpectrumPlan::NormGrid& ngd = sp->norm_grid[k];
        ngd.E = n_e;
        if ((rc = upload(plan, g, &ngd.egrid))) return rc;
        if ((rc = upload(plan, l, &ngd.lngrid))) return rc;
        if ((rc = upload(plan, m, &ngd.emid))) return rc;
        if ((rc = upload(plan, ml, &ngd.lnmid))) return rc;
        if (nd.kind == ELISA_NORM_ENFLUX && (rc = upload(plan, w, &ngd.w))) return rc;
    }

    // per-channel data, padded with benign values
    const int Cp = sp->C_pad;
    std::vector<double> ch((size_t)10 * Cp, 0.0);
    auto xlogy_h = [](double x, double y) { return x == 0.0 ? 0.0 : x * log(y); };
    for (int c = 0; c < Cp; ++c) {
        const bool v = c < C;
        const double on = v ? sd.on_counts[c] : 0.0;
        const double off = v && sd.off_counts ? sd.off_counts[c] : 0.0;
        double sigma = 1.0;
        if (v && sd.stat == ELISA_STAT_CHI2) sigma = sd.on_error[c];
        if (v && sd.stat == ELISA_STAT_PGSTAT) sigma = sd.off_error[c];
        ch[0 * Cp + c] = v ? sd.area_scale[c] * sd.exposure : 0.0;
        ch[1 * Cp + c] = on;
        ch[2 * Cp + c] = off;
        ch[3 * Cp + c] = sigma;
        ch[4 * Cp + c] = v && sd.back_ratio ? sd.back_ratio[c] : 1.0;
        ch[5 * Cp + c] = xlogy_h(on, on) - on;
        ch[6 * Cp + c] = xlogy_h(off, off) - off;
        ch[7 * Cp + c] = v ? sd.area_scale[c] : 0.0;
        ch[8 * Cp + c] = 1.0;  // 1 / channel_width: set by elisa_b200_plan_set_channel_width
        {  // 1 / (2 a (a + 1)) (wstat) or 1 / (2 a) (pgstat): likelihood.py:79, 127 -- a division per channel, not per row
            const double a = ch[4 * Cp + c];
            ch[9 * Cp + c] = sd.stat == ELISA_STAT_WSTAT ? 1.0 / (2.0 * a * (a + 1.0)) : sd.stat == ELISA_STAT_PGSTAT ? 1.0 / (2.0 * a) : 0.0;
        }
    }
    if ((rc = upload(plan, ch, &sp->chan))) return rc;

    // response -> two packed operands
    PackedHost fwd, bwd;
    if (dense) {
        const double* R = sd.response_dense;  // [E, C]
        int64_t nnz = 0;
        std::vector<int> clo(C, E), chi(C, 0), elo(E, C), ehi(E, 0);
        for (int e = 0; e < E; ++e)
            for (int c = 0; c < C; ++c)
                if (R[(size_t)e * C + c] != 0.0) {
                    ++nnz;
                    clo[c] = std::min(clo[c], e);
                    chi[c] = std::max(chi[c], e + 1);
                    elo[e] = std::min(elo[e], c);
                    ehi[e] = std::max(ehi[e], c + 1);
                }
        sp->nnz = nnz;
        pack_tiles(
            E, C, sp->E_pad, [&](int n, int* a, int* b) { *a = clo[n]; *b = chi[n]; },
            [&](int n, int lo, int hi, double* dst) {
                for (int k = lo; k < hi; ++k) dst[(size_t)(k - lo) * GEMM_BN] = R[(size_t)k * C + n];
            },
            &fwd);
        pack_tiles(
            C, E, sp->C_pad, [&](int n, int* a, int* b) { *a = elo[n]; *b = ehi[n]; },
            [&](int n, int lo, int hi, double* dst) {
                for (int k = lo; k < hi; ++k) dst[(size_t)(k - lo) * GEMM_BN] = R[(size_t)n * C + k];
            },
            &bwd);
    } else {
        const int32_t* ip = sd.csr_indptr;
        if (ip[0] != 0) return fail(ELISA_ERR_INVALID, "csr_indptr[0] != 0");
        for (int c = 0; c < C; ++c)
            if (ip[c + 1] < ip[c]) return fail(ELISA_ERR_INVALID, "csr_indptr not monotone");
        const int64_t nnz = ip[C];
        sp->nnz = nnz;
        for (int64_t i = 0; i < nnz; ++i)
            if (sd.csr_indices[i] < 0 || sd.csr_indices[i] >= E) return fail(ELISA_ERR_INVALID, "csr_indices out of range");
        std::vector<int> clo(C, E), chi(C, 0), elo(E, C), ehi(E, 0);
        for (int c = 0; c < C; ++c)
            for (int i = ip[c]; i < ip[c + 1]; ++i) {
                const int e = sd.csr_indices[i];
                clo[c] = std::min(clo[c], e);
                chi[c] = std::max(chi[c], e + 1);
                elo[e] = std::min(elo[e], c);
                ehi[e] = std::max(ehi[e], c + 1);
            }
        // forward: columns = channels; duplicates accumulate like scipy's tocsr would
        pack_tiles(
            E, C, sp->E_pad, [&](int n, int* a, int* b) { *a = clo[n]; *b = chi[n]; },
            [&](int n, int lo, int, double* dst) {
                for (int i = ip[n]; i < ip[n + 1]; ++i)
                    dst[(size_t)(sd.csr_indices[i] - lo) * GEMM_BN] += sd.csr_values[i];
            },
            &fwd);
        // transposed: columns = energies; build a CSC view (by energy) first
        std::vector<int64_t> cp(E + 1, 0);
        for (int64_t i = 0; i < nnz; ++i) ++cp[sd.csr_indices[i] + 1];
        for (int e = 0; e < E; ++e) cp[e + 1] += cp[e];
        std::vector<int32_t> crow((size_t)nnz);
        std::vector<double> cval((size_t)nnz);
        {
            std::vector<int64_t> pos(cp.begin(), cp.end() - 1);
            for (int c = 0; c < C; ++c)
                for (int i = ip[c]; i < ip[c + 1]; ++i) {
                    const int64_t p = pos[sd.csr_indices[i]]++;
                    crow[p] = c;
                    cval[p] = sd.csr_values[i];
                }
        }
        pack_tiles(
            C, E, sp->C_pad, [&](int n, int* a, int* b) { *a = elo[n]; *b = ehi[n]; },
            [&](int n, int lo, int, double* dst) {
                for (int64_t i = cp[n]; i < cp[n + 1]; ++i) dst[(size_t)(crow[i] - lo) * GEMM_BN] += cval[i];
            },
            &bwd);
    }
    if ((rc = upload_packed(plan, fwd, &sp->fwd))) return rc;
    if ((rc = upload_packed(plan, bwd, &sp->bwd))) return rc;
    if (plan->tiny_enabled && tiny_shape_ok(E, C, plan->P, sp->model.used_mask)) {
        std::vector<double> Rd((size_t)E * C, 0.0);
        if (dense) {
            std::copy(sd.response_dense, sd.response_dense + (size_t)E * C, Rd.begin());
        } else {
            for (int c = 0; c < C; ++c)
                for (int i = sd.csr_indptr[c]; i < sd.csr_indptr[c + 1]; ++i) Rd[(size_t)sd.csr_indices[i] * C + c] += sd.csr_values[i];
        }
        if ((rc = upload(plan, Rd, &sp->tiny_R))) return rc;
    }
    {  // CSR by channel row for the row-wise sparse fold
        std::vector<int32_t> ptr(C + 1, 0), idx;
        std::vector<double> val;
        if (dense) {
            const double* R = sd.response_dense;
            if (sp->nnz <= (int64_t)E * C / 4) {  // worth keeping only when the response is actually sparse
                idx.reserve((size_t)sp->nnz);
                val.reserve((size_t)sp->nnz);
                for (int c = 0; c < C; ++c) {
                    for (int e = 0; e < E; ++e)
                        if (R[(size_t)e * C + c] != 0.0) {
                            idx.push_back(e);
                            val.push_back(R[(size_t)e * C + c]);
                        }
                    ptr[c + 1] = (int32_t)idx.size();
                }
            }
        } else {
            ptr.assign(sd.csr_indptr, sd.csr_indptr + C + 1);
            idx.assign(sd.csr_indices, sd.csr_indices + sp->nnz);
            val.assign(sd.csr_values, sd.csr_values + sp->nnz);
        }
        if (!idx.empty()) {
            if ((rc = upload(plan, ptr, &sp->csr_ptr))) return rc;
            if ((rc = upload(plan, idx, &sp->csr_idx))) return rc;
            if ((rc = upload(plan, val, &sp->csr_val))) return rc;
        }
    }
    {  // fraction of the two operands that is stored (1 for a dense response)
        double stored = 0.0;
        for (size_t j = 0; j < fwd.k_lo.size(); ++j) stored += std::max(0, fwd.k_hi[j] - fwd.k_lo[j]);
        for (size_t j = 0; j < bwd.k_lo.size(); ++j) stored += std::max(0, bwd.k_hi[j] - bwd.k_lo[j]);
        sp->fill = stored / ((double)fwd.k_lo.size() * E + (double)bwd.k_lo.size() * C);
    }
    return ELISA_OK;
}

// ======================================================================================
// runtime specialisation of the unfold kernel (NVRTC)
// ======================================================================================
struct JitKernels {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t val = nullptr, grad = nullptr, contract = nullptr;
    cudaKernel_t tiny = nullptr;  // one-kernel small-batch evaluation (only in translation units generated with a statistic)
};

static const char* const k_src_elisa_h =
#include "embedded_elisa_b200_h.inc"
    ;
static const char* const k_src_dual =
#include "embedded_dual_cuh.inc"
    ;
static const char* const k_src_components =
#include "embedded_components_cuh.inc"
    ;
static const char* const k_src_skel =
#include "embedded_unfold_skel_cuh.inc"
    ;
static const char* const k_src_slice =
#include "embedded_slice_i8_cuh.inc"
    ;
static const char* const k_src_fastmath =
#include "embedded_fastmath_cuh.inc"
    ;
static const char* const k_src_stats =
#include "embedded_stats_cuh.inc"
    ;

struct NvrtcApi {
    void* handle = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    bool ok = false;
};

static NvrtcApi& nvrtc_api() {
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
#define ELISA_SYM(field, sym) api.field = (decltype(api.field))dlsym(api.handle, #sym)
        ELISA_SYM(create, nvrtcCreateProgram);
        ELISA_SYM(compile, nvrtcCompileProgram);
        ELISA_SYM(destroy, nvrtcDestroyProgram);
        ELISA_SYM(cubin_size, nvrtcGetCUBINSize);
        ELISA_SYM(cubin, nvrtcGetCUBIN);
        ELISA_SYM(log_size, nvrtcGetProgramLogSize);
        ELISA_SYM(log, nvrtcGetProgramLog);
#undef ELISA_SYM
        api.ok = api.create && api.compile && api.destroy && api.cubin_size && api.cubin && api.log_size && api.log;
    });
    return api;
}

static std::string hexd(double v) {
    char buf[64];
    if (v != v) return "(0.0/0.0)";
    if (v == HUGE_VAL) return "(1.0/0.0)";
    if (v == -HUGE_VAL) return "(-1.0/0.0)";
    snprintf(buf, sizeof buf, "%a", v);
    return std::string("(") + buf + ")";
}

// The translation unit for one model: see unfold_skel.cuh.
// ks > 0: the kernels also write the digits of U for the sliced-integer fold (fold_i8.cuh).
//
// Two policy structs are generated, ModelV (values: every scalar a double) and ModelG (values +
// d/dtheta).  In ModelG every leaf is evaluated in its OWN dual type, Dual<n_i> with n_i the number
// of distinct free theta columns among the leaf's parameters -- not Dual<P>: a tangent that is
// structurally zero costs an IEEE multiply-add all the same (0 * x is not foldable), and for
// CutoffPL + Blackbody that is 12 scalars per edge carried instead of 7.  The +/* tree above the
// leaves is emitted as straight-line scalar code over (value, {column: tangent}) pairs, so sums
// of leaves with disjoint parameters cost nothing beyond the value.
namespace {
struct SV {  // a scalar expression and its non-zero tangents by theta column
    std::string v;
    std::map<int, std::string> d;
};
}  // namespace

static bool model_has_free_shift(const ModelDev& m) {
    for (int i = 0; i < m.n_leaf; ++i)
        if (m.leaf[i].n_shift > 0) return true;
    return false;
}

static std::string generate_model_struct(const ModelDev& m, int P, bool grad, const std::string& name) {
    std::string s;
    auto str = [](int v) { return std::to_string(v); };
    // distinct free columns per leaf (local tangent index = position in this list)
    std::vector<std::vector<int>> cols(m.n_leaf);
    std::vector<std::string> ltype(m.n_leaf);
    for (int i = 0; i < m.n_leaf; ++i) {
        if (grad) {
            for (int q = 0; q < m.leaf[i].np; ++q) {
                const int c = m.leaf[i].src[q];
                if (c >= 0 && std::find(cols[i].begin(), cols[i].end(), c) == cols[i].end()) cols[i].push_back(c);
            }
            for (int k = 0; k < m.leaf[i].n_shift; ++k) {  // a free shift's z / v is one more tangent of the leaf
                const int c = m.leaf[i].shift_src[k];
                if (std::find(cols[i].begin(), cols[i].end(), c) == cols[i].end()) cols[i].push_back(c);
            }
        }
        std::sort(cols[i].begin(), cols[i].end());
        ltype[i] = cols[i].empty() ? "double" : "Dual<" + str((int)cols[i].size()) + ">";
    }
    // columns each flux normalisation factor F / flux depends on
    std::vector<std::vector<int>> ncols(m.n_norms);
    for (int k = 0; k < m.n_norms; ++k) {
        if (!grad) continue;
        std::vector<int>& nc = ncols[k];
        for (int o = m.norm[k].op_lo; o < m.norm[k].op_hi; ++o)
            if (m.ops[o] >= 0)
                for (int c : cols[m.ops[o]])
                    if (std::find(nc.begin(), nc.end(), c) == nc.end()) nc.push_back(c);
        if (m.norm[k].f_src >= 0 && std::find(nc.begin(), nc.end(), m.norm[k].f_src) == nc.end()) nc.push_back(m.norm[k].f_src);
        std::sort(nc.begin(), nc.end());
    }
    s += "struct " + name + " {\n  struct Pre {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const std::string k = str(m.leaf[i].kind), id = str(i);
        s += "    " + ltype[i] + " p" + id + "[Comp<" + k + ">::NP]; " + ltype[i] + " c" + id + "[Comp<" + k + ">::NC];\n";
        if (m.leaf[i].n_shift > 0)  // free energy shifts: grid factor, its logarithm, output factor
            s += "    " + ltype[i] + " F" + id + ", lnF" + id + ", oF" + id + ";\n";
    }
    for (int k = 0; k < m.n_norms; ++k) {
        s += "    double n" + str(k) + "_v;\n";
        for (int c : ncols[k]) s += "    double n" + str(k) + "_d" + str(c) + ";\n";
    }
    s += "  };\n  __device__ __forceinline__ static void pre(const double* th, Pre& s) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        const std::string k = str(L.kind), id = str(i), T = ltype[i];
        for (int q = 0; q < L.np; ++q) {
            s += "    s.p" + id + "[" + str(q) + "] = ";
            if (L.src[q] >= 0) {
                const int local = grad ? (int)(std::find(cols[i].begin(), cols[i].end(), L.src[q]) - cols[i].begin()) : -1;
                s += "Flat<" + T + ">::seed(th[" + str(L.src[q]) + "], " + str(local) + ");\n";
            } else {
                s += "Flat<" + T + ">::seed(" + hexd(L.cst[q]) + ", -1);\n";
            }
        }
        s += "    { const double aux[2] = {" + hexd(L.aux[0]) + ", " + hexd(L.aux[1]) + "};\n";
        s += "      for (int k = 0; k < Comp<" + k + ">::NC; ++k) s.c" + id + "[k] = Flat<" + T + ">::seed(0.0, -1);\n";
        s += "      Comp<" + k + ">::pre(s.p" + id + ", aux, s.c" + id + "); }\n";
        if (L.n_shift > 0) {
            // F = prod of (1 + z) / (1 - v / c) over the free shifts around the leaf, as a dual number seeded in
            // the shift parameter's tangent; oF = prod of 1 / (1 + z) over the ZAShifts among them (conv.py:299-300)
            s += "    s.F" + id + " = Flat<" + T + ">::seed(1.0, -1); s.oF" + id + " = Flat<" + T + ">::seed(1.0, -1);\n";
            for (int q = 0; q < L.n_shift; ++q) {
                const int local = grad ? (int)(std::find(cols[i].begin(), cols[i].end(), L.shift_src[q]) - cols[i].begin()) : -1;
                const std::string z = "Flat<" + T + ">::seed(th[" + str(L.shift_src[q]) + "], " + str(local) + ")";
                if (L.shift_kind[q] == ELISA_SHIFT_V)
                    s += "    { const " + T + " f = 1.0 - " + z + " * (1.0 / 299792.458); s.F" + id + " = s.F" + id + " * f; }\n";
                else if (L.shift_kind[q] == ELISA_SHIFT_ZA)
                    s += "    { const " + T + " f = 1.0 + " + z + "; s.F" + id + " = s.F" + id + " * f; s.oF" + id + " = s.oF" + id + " / f; }\n";
                else
                    s += "    { const " + T + " f = 1.0 + " + z + "; s.F" + id + " = s.F" + id + " * f; }\n";
            }
            s += "    s.lnF" + id + " = lg(s.F" + id + ");\n";
        }
    }
    s += "  }\n";
    // straight-line code of ops[lo, hi) on grid view `gv`; on_norm_grid selects the leaf's scales
    // on a flux-normalisation grid (shifts inside the normalised sub-expression only)
    auto emit_range = [&](int lo, int hi, bool on_norm_grid, const std::string& pfx, const std::string& ind) -> SV {
        std::vector<SV> stack;
        std::vector<bool> emitted(m.n_leaf, false);
        int tmp = 0;
        auto fresh = [&](const SV& val) {  // name the pieces of a computed value
            SV out;
            const std::string t = pfx + "t" + str(tmp++);
            s += ind + "const double " + t + "_v = " + val.v + ";\n";
            out.v = t + "_v";
            for (const auto& kv : val.d) {
                s += ind + "const double " + t + "_d" + str(kv.first) + " = " + kv.second + ";\n";
                out.d[kv.first] = t + "_d" + str(kv.first);
            }
            return out;
        };
        auto mul = [&](const SV& a, const SV& b) {
            SV r;
            r.v = a.v + " * " + b.v;
            for (const auto& kv : a.d) r.d[kv.first] = kv.second + " * " + b.v;
            for (const auto& kv : b.d) {
                auto it = r.d.find(kv.first);
                if (it == r.d.end())
                    r.d[kv.first] = a.v + " * " + kv.second;
                else
                    it->second = "fma(" + a.v + ", " + kv.second + ", " + it->second + ")";
            }
            return fresh(r);
        };
        for (int o = lo; o < hi; ++o) {
            const int op = m.ops[o];
            if (op >= 0) {
                const LeafDev& L = m.leaf[op];
                const std::string id = str(op), T = ltype[op];
                if (!emitted[op]) {
                    const double gs = on_norm_grid ? L.ngs : L.gs, lgs = on_norm_grid ? L.lngs : L.lgs;
                    const double os = on_norm_grid ? L.nos : L.os;
                    if (L.n_shift > 0)
                        s += ind + "const " + T + " " + pfx + "v" + id + " = " + (os != 1.0 ? hexd(os) + " * " : std::string()) +
                             "(s.oF" + id + " * leaf_bin_shift<" + str(L.kind) + ", " + T + ">(s.p" + id + ", s.c" + id + ", " +
                             str(L.method) + ", gv, e, " + hexd(gs) + ", " + hexd(lgs) + ", s.F" + id + ", s.lnF" + id + "));\n";
                    else
                    s += ind + "const " + T + " " + pfx + "v" + id + " = " + (os != 1.0 ? hexd(os) + " * " : std::string()) +
                         "leaf_bin_pc<" + str(L.kind) + ", " + T + ">(s.p" + id + ", s.c" + id + ", " + str(L.method) +
                         ", gv, e, " + id + ", " + hexd(gs) + ", " + hexd(lgs) + ");\n";
                    emitted[op] = true;
                }
                SV leaf;
                if (cols[op].empty()) {
                    leaf.v = pfx + "v" + id;
                } else {
                    leaf.v = pfx + "v" + id + ".v";
                    for (size_t k = 0; k < cols[op].size(); ++k) leaf.d[cols[op][k]] = pfx + "v" + id + ".d[" + str((int)k) + "]";
                }
                stack.push_back(leaf);
            } else if (op <= ELISA_OP_NORM0) {
                const int k = ELISA_OP_NORM0 - op;
                SV nrm;
                nrm.v = "s.n" + str(k) + "_v";
                for (int c : ncols[k]) nrm.d[c] = "s.n" + str(k) + "_d" + str(c);
                stack.back() = mul(nrm, stack.back());
            } else {
                const SV r = stack.back();
                stack.pop_back();
                const SV l = stack.back();
                stack.pop_back();
                if (op == ELISA_OP_ADD) {
                    SV t;
                    t.v = l.v + " + " + r.v;
                    t.d = l.d;
                    for (const auto& kv : r.d) {
                        auto it = t.d.find(kv.first);
                        if (it == t.d.end())
                            t.d[kv.first] = kv.second;
                        else
                            it->second = it->second + " + " + kv.second;
                    }
                    stack.push_back(fresh(t));
                } else {
                    stack.push_back(mul(l, r));
                }
            }
        }
        return stack.back();
    };
    s += "  __device__ __forceinline__ static void norms(Pre& s, const UnfoldCore& a, const double* th, int lane) {\n";
    for (int k = 0; k < m.n_norms; ++k) {
        const NormDev& nd = m.norm[k];
        const std::string id = str(k);
        s += "    {\n      const GridView gv = a.ngv[" + id + "];\n      double acc_v = 0.0;\n";
        for (int c : ncols[k]) s += "      double acc_d" + str(c) + " = 0.0;\n";
        s += "      for (int e0 = 0; e0 < gv.E; e0 += 31) {\n        const int e = e0 + lane;\n";
        const SV sub = emit_range(nd.op_lo, nd.op_hi, true, "n", "        ");
        const std::string w = nd.kind == ELISA_NORM_ENFLUX ? "a.nw[" + id + "][e]" : "";
        s += "        if (lane < 31 && e < gv.E) {\n";
        if (!w.empty()) s += "          const double w = " + w + ";\n";
        s += "          acc_v += " + (w.empty() ? sub.v : "w * (" + sub.v + ")") + ";\n";
        for (int c : ncols[k]) {
            auto it = sub.d.find(c);
            if (it != sub.d.end())
                s += "          acc_d" + str(c) + " += " + (w.empty() ? it->second : "w * (" + it->second + ")") + ";\n";
        }
        s += "        }\n      }\n      acc_v = warp_sum(acc_v);\n";
        for (int c : ncols[k]) s += "      acc_d" + str(c) + " = warp_sum(acc_d" + str(c) + ");\n";
        s += "      const double F = " + (nd.f_src >= 0 ? "th[" + str(nd.f_src) + "]" : hexd(nd.f_cst)) + ";\n";
        s += "      const double inv = 1.0 / acc_v;\n      s.n" + id + "_v = F * inv;\n";
        for (int c : ncols[k])  // d(F / acc) = (dF - (F / acc) d acc) / acc
            s += "      s.n" + id + "_d" + str(c) + " = (" + (c == nd.f_src ? "1.0" : "0.0") + " - s.n" + id + "_v * acc_d" + str(c) +
                 ") * inv;\n";
        s += "    }\n";
    }
    s += "  }\n";
    const std::string RT = grad ? "Dual<" + str(P) + ">" : "double";
    s += "  __device__ __forceinline__ static " + RT + " bin(const Pre& s, const GridView& gv, int e) {\n";
    {
        const SV r = emit_range(0, m.n_ops, false, "", "    ");
        if (!grad) {
            s += "    return " + r.v + ";\n";
        } else {
            s += "    " + RT + " r;\n    r.v = " + r.v + ";\n";
            for (int j = 0; j < P; ++j) {
                auto it = r.d.find(j);
                s += "    r.d[" + str(j) + "] = " + (it != r.d.end() ? it->second : std::string("0.0")) + ";\n";
            }
            s += "    return r;\n";
        }
    }
    s += "  }\n};\n";
    return s;
}

// Gradient by contraction on edges (unfold_contract_edges, unfold_skel.cuh): possible when the model is a
// plain sum of trapezoid continuum leaves and power laws -- the bin value is then linear in the edge
// values of each leaf.  Returns the generated policy struct ModelE, or an empty string when the model
// does not qualify (products, Simpson's rule, flux normalisations, analytic leaves other than the power
// law) or ELISA_B200_CONTRACT=generic asks for the dual-number contraction.
static std::string generate_edge_model(const ModelDev& m, int P) {
    if (const char* env = getenv("ELISA_B200_CONTRACT"))
        if (strcmp(env, "generic") == 0) return std::string();
    if (m.n_norms != 0) return std::string();
    for (int i = 0; i < m.n_leaf; ++i)
        if (m.leaf[i].n_shift > 0) return std::string();  // a free shift differentiates the edge ENERGIES as well
    auto str = [](int v) { return std::to_string(v); };
    std::vector<int> mult(m.n_leaf, 0);
    int n_acc = 0;
    for (int o = 0; o < m.n_ops; ++o) {
        const int op = m.ops[o];
        if (op >= 0)
            ++mult[op];
        else if (op != ELISA_OP_ADD)
            return std::string();
    }
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        const bool trapz_continuum = comp_is_additive(L.kind) && L.kind != ELISA_COMP_POWERLAW && L.kind != ELISA_COMP_BROKENPL &&
                                     L.kind != ELISA_COMP_LORENTZ && L.kind != ELISA_COMP_PLENFLUX && L.kind != ELISA_COMP_PLPHFLUX &&
                                     L.method == ELISA_METHOD_TRAPZ;
        if (!trapz_continuum && L.kind != ELISA_COMP_POWERLAW) return std::string();
        n_acc += L.np;
    }
    if (n_acc > 12) return std::string();  // accumulators and basis values live in registers
    std::string s = "struct ModelE {\n  static constexpr int NPAR = " + str(P) + ";\n  static constexpr int O0 = 0;\n";
    for (int i = 0; i < m.n_leaf; ++i)
        s += "  using L" + str(i) + " = EdgeLeaf<" + str(m.leaf[i].kind) + ">;\n  static constexpr int O" + str(i + 1) + " = O" + str(i) +
             " + L" + str(i) + "::NB;\n";
    s += "  static constexpr int NACC = O" + str(m.n_leaf) + ";\n  static constexpr unsigned DMASK = 0u";
    for (int i = 0; i < m.n_leaf; ++i) s += " | (L" + str(i) + "::DIFF ? ((1u << L" + str(i) + "::NB) - 1u) << O" + str(i) + " : 0u)";
    s += ";\n  static constexpr bool HAS_T = false";
    for (int i = 0; i < m.n_leaf; ++i) s += " || !L" + str(i) + "::DIFF";
    s += ";\n  static constexpr bool HAS_D = false";
    for (int i = 0; i < m.n_leaf; ++i) s += " || L" + str(i) + "::DIFF";
    s += ";\n  struct Pre {\n";
    for (int i = 0; i < m.n_leaf; ++i) s += "    L" + str(i) + "::Pre l" + str(i) + ";\n";
    s += "  };\n  __device__ __forceinline__ static void pre(const double* th, Pre& s) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        s += "    { const double val[" + str(L.np) + "] = {";
        for (int q = 0; q < L.np; ++q) s += (q ? ", " : "") + (L.src[q] >= 0 ? "th[" + str(L.src[q]) + "]" : hexd(L.cst[q]));
        s += "}; const double aux[2] = {" + hexd(L.aux[0]) + ", " + hexd(L.aux[1]) + "}; L" + str(i) + "::pre(val, aux, s.l" + str(i) + "); }\n";
    }
    s += "  }\n  __device__ __forceinline__ static void edge(const Pre& s, const GridView&, double E, double lnE, double& vT, double& vD, double* b) {\n"
         "    double v;\n";
    // scale of a leaf's contribution: out_scale (x grid_scale for the trapezoid's bin width) x multiplicity
    std::vector<double> f(m.n_leaf);
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        const bool diff = L.kind == ELISA_COMP_POWERLAW;
        f[i] = L.os * (diff ? 1.0 : L.gs) * (double)mult[i];
        s += "    L" + str(i) + "::edge(s.l" + str(i) + ", " + (L.gs != 1.0 ? hexd(L.gs) + " * E" : std::string("E")) + ", " +
             (L.lgs != 0.0 ? "lnE + " + hexd(L.lgs) : std::string("lnE")) + ", v, b + O" + str(i) + "); " + (diff ? "vD" : "vT") + " += " +
             (f[i] != 1.0 ? hexd(f[i]) + " * v" : std::string("v")) + ";\n";
    }
    s += "  }\n  __device__ __forceinline__ static void finish(const Pre& s, const double* S, double* grad) {\n"
         "    for (int j = 0; j < NPAR; ++j) grad[j] = 0.0;\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        s += "    { double d[" + str(L.np) + "]; L" + str(i) + "::finish(s.l" + str(i) + ", S + O" + str(i) + ", d);";
        for (int q = 0; q < L.np; ++q)
            if (L.src[q] >= 0)
                s += " grad[" + str(L.src[q]) + "] += " + (f[i] != 1.0 ? hexd(f[i]) + " * " : std::string()) + "d[" + str(q) + "];";
        s += " }\n";
    }
    s += "  }\n};\n";
    return s;
}

// Edge-wise contraction for a TREE of sums and products (unfold_contract_tree, csrc/unfold_skel.cuh): every leaf an
// additive trapezoid continuum, a power law, a multiplicative trapezoid continuum (incl. the table absorption) or a
// Constant; no Simpson's rule, no flux normalisation, no free shift.  Empty string: does not apply.
static std::string generate_tree_model(const ModelDev& m, int P) {
    if (const char* env = getenv("ELISA_B200_CONTRACT"))
        if (strcmp(env, "generic") == 0) return std::string();
    if (m.n_norms != 0 || m.n_leaf > 8) return std::string();
    auto str = [](int v) { return std::to_string(v); };
    enum { T_ADD = 0, T_DIFF = 1, T_MUL = 2 };
    std::vector<int> rule(m.n_leaf);
    int n_acc = 0;
    bool any_mul_op = false;
    for (int o = 0; o < m.n_ops; ++o) {
        if (m.ops[o] >= 0) continue;
        if (m.ops[o] == ELISA_OP_MUL)
            any_mul_op = true;
        else if (m.ops[o] != ELISA_OP_ADD)
            return std::string();
    }
    if (!any_mul_op) return std::string();  // plain sums have their own, leaner form (generate_edge_model)
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        if (L.n_shift > 0) return std::string();
        if (L.kind == ELISA_COMP_POWERLAW) {
            rule[i] = T_DIFF;
        } else if (comp_is_additive(L.kind)) {
            if (L.kind == ELISA_COMP_BROKENPL || L.kind == ELISA_COMP_LORENTZ || L.kind == ELISA_COMP_PLENFLUX ||
                L.kind == ELISA_COMP_PLPHFLUX || L.method != ELISA_METHOD_TRAPZ)
                return std::string();
            rule[i] = T_ADD;
        } else if (L.kind == ELISA_COMP_CONSTANT) {
            rule[i] = T_MUL;
        } else {
            if (L.kind == ELISA_COMP_PLABS || L.method != ELISA_METHOD_TRAPZ) return std::string();
            rule[i] = T_MUL;
        }
        n_acc += L.np;
    }
    if (n_acc > 12) return std::string();  // accumulators and basis values live in registers
    std::string s = "struct ModelT {\n  static constexpr int NPAR = " + str(P) + ", NLEAF = " + str(m.n_leaf) + ";\n  static constexpr int O0 = 0;\n";
    for (int i = 0; i < m.n_leaf; ++i)
        s += "  using L" + str(i) + " = EdgeLeaf<" + str(m.leaf[i].kind) + ">;\n  static constexpr int O" + str(i + 1) + " = O" + str(i) +
             " + L" + str(i) + "::NB;\n";
    s += "  static constexpr int NACC = O" + str(m.n_leaf) + ";\n  struct Pre {\n";
    for (int i = 0; i < m.n_leaf; ++i) s += "    L" + str(i) + "::Pre l" + str(i) + ";\n";
    s += "  };\n  __device__ __forceinline__ static void pre(const double* th, Pre& s) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        s += "    { const double val[" + str(L.np) + "] = {";
        for (int q = 0; q < L.np; ++q) s += (q ? ", " : "") + (L.src[q] >= 0 ? "th[" + str(L.src[q]) + "]" : hexd(L.cst[q]));
        s += "}; const double aux[2] = {" + hexd(L.aux[0]) + ", " + hexd(L.aux[1]) + "}; L" + str(i) + "::pre(val, aux, s.l" + str(i) + "); }\n";
    }
    // edge values and tangent bases; a table leaf takes sigma(E) of its own table in the ln E slot
    s += "  }\n  __device__ __forceinline__ static void edge(const Pre& s, const GridView& gv, int ec, double E, double lnE, double* v, double* b) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        const std::string second = L.kind == ELISA_COMP_PHOTONABS ? "gv.tab[" + str(2 * i) + "][ec]"
                                                                  : (L.lgs != 0.0 ? "lnE + " + hexd(L.lgs) : std::string("lnE"));
        s += "    L" + str(i) + "::edge(s.l" + str(i) + ", " + (L.gs != 1.0 ? hexd(L.gs) + " * E" : std::string("E")) + ", " + second +
             ", v[" + str(i) + "], b + O" + str(i) + ");\n";
    }
    // scale of a leaf's bin value: out_scale (x grid_scale for the bin width of an additive trapezoid leaf)
    std::vector<double> f(m.n_leaf);
    for (int i = 0; i < m.n_leaf; ++i) f[i] = rule[i] == T_ADD ? m.leaf[i].os * m.leaf[i].gs : (rule[i] == T_DIFF ? m.leaf[i].os : 1.0);
    s += "  }\n  __device__ __forceinline__ static void bins(const double* v, double hd, double* binv) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const std::string id = str(i), nx = "__shfl_down_sync(0xffffffffu, v[" + id + "], 1)";
        const std::string fs = f[i] != 1.0 ? hexd(f[i]) + " * " : std::string();
        if (rule[i] == T_ADD)
            s += "    binv[" + id + "] = " + fs + "(hd * (v[" + id + "] + " + nx + "));\n";
        else if (rule[i] == T_DIFF)
            s += "    binv[" + id + "] = " + fs + "(" + nx + " - v[" + id + "]);\n";
        else
            s += "    binv[" + id + "] = 0.5 * (v[" + id + "] + " + nx + ");\n";
    }
    // the tree forward (values n_k) ...
    s += "  }\n  __device__ __forceinline__ static void tree(const double* binv, double* adj, double& u) {\n";
    struct Node { bool leaf; int idx; int a, b, op; };  // internal: children a, b (node ids)
    std::vector<Node> nodes;
    std::vector<int> stack;
    for (int o = 0; o < m.n_ops; ++o) {
        const int op = m.ops[o];
        if (op >= 0) {
            nodes.push_back({true, op, -1, -1, 0});
            stack.push_back((int)nodes.size() - 1);
        } else {
            if (stack.size() < 2) return std::string();
            const int rb = stack.back(); stack.pop_back();
            const int ra = stack.back(); stack.pop_back();
            nodes.push_back({false, -1, ra, rb, op});
            stack.push_back((int)nodes.size() - 1);
        }
    }
    if (stack.size() != 1) return std::string();
    auto val = [&](int n) { return nodes[n].leaf ? "binv[" + str(nodes[n].idx) + "]" : "n" + str(n); };
    for (int n = 0; n < (int)nodes.size(); ++n)
        if (!nodes[n].leaf)
            s += "    const double n" + str(n) + " = " + val(nodes[n].a) + (nodes[n].op == ELISA_OP_ADD ? " + " : " * ") + val(nodes[n].b) + ";\n";
    const int root = stack.back();
    s += "    u = " + val(root) + ";\n";
    for (int i = 0; i < m.n_leaf; ++i) s += "    adj[" + str(i) + "] = 0.0;\n";
    // ... and in reverse: a node's adjoint goes to its children (times the sibling's value under a product)
    std::vector<std::string> bar(nodes.size());
    bar[root] = "1.0";
    for (int n = (int)nodes.size() - 1; n >= 0; --n) {
        if (bar[n].empty()) return std::string();  // unreachable node: not a tree
        if (nodes[n].leaf) {
            s += "    adj[" + str(nodes[n].idx) + "] += " + bar[n] + ";\n";
            continue;
        }
        const int ca = nodes[n].a, cb = nodes[n].b;
        if (nodes[n].op == ELISA_OP_ADD) {
            bar[ca] = bar[n];
            bar[cb] = bar[n];
        } else {
            s += "    const double a" + str(ca) + " = " + (bar[n] == "1.0" ? std::string() : bar[n] + " * ") + val(cb) + ";\n";
            s += "    const double a" + str(cb) + " = " + (bar[n] == "1.0" ? std::string() : bar[n] + " * ") + val(ca) + ";\n";
            bar[ca] = "a" + str(ca);
            bar[cb] = "a" + str(cb);
        }
    }
    // per leaf: adjoint of the bin -> weights of its two edges -> accumulators of its tangent basis
    s += "  }\n  __device__ __forceinline__ static void accumulate(const double* adj, double gg, double hd, int lane, const double* b, double* acc) {\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const std::string id = str(i);
        const std::string fs = f[i] != 1.0 ? hexd(f[i]) + " * " : std::string();
        std::string t;  // gg = 0 (no bin, or a clipped one) must give weight zero whatever the adjoint overflowed to
        if (rule[i] == T_ADD)
            t = "gg != 0.0 ? " + fs + "(gg * adj[" + id + "] * hd) : 0.0";
        else if (rule[i] == T_DIFF)
            t = "gg != 0.0 ? " + fs + "(gg * adj[" + id + "]) : 0.0";
        else
            t = "gg != 0.0 ? 0.5 * (gg * adj[" + id + "]) : 0.0";
        s += "    { const double w = " + std::string(rule[i] == T_DIFF ? "edge_weight_diff" : "edge_weight_sum") + "(" + t + ", lane);\n"
             "      if (w != 0.0) {\n"
             "#pragma unroll\n"
             "        for (int q = O" + id + "; q < O" + str(i + 1) + "; ++q) acc[q] = fma(w, b[q], acc[q]);\n      } }\n";
    }
    s += "  }\n  __device__ __forceinline__ static void finish(const Pre& s, const double* S, double* grad) {\n"
         "    for (int j = 0; j < NPAR; ++j) grad[j] = 0.0;\n";
    for (int i = 0; i < m.n_leaf; ++i) {
        const LeafDev& L = m.leaf[i];
        s += "    { double d[" + str(L.np) + "]; L" + str(i) + "::finish(s.l" + str(i) + ", S + O" + str(i) + ", d);";
        for (int q = 0; q < L.np; ++q)
            if (L.src[q] >= 0) s += " grad[" + str(L.src[q]) + "] += d[" + str(q) + "];";
        s += " }\n";
    }
    s += "  }\n};\n";
    return s;
}

// tiny_stat >= 0: also instantiate the one-kernel small-batch evaluation (tiny_eval) for that statistic
static std::string generate_unfold_source(const ModelDev& m, int P, int ks, int tiny_stat = -1) {
    std::string s;
    {
        const char* env = getenv("ELISA_B200_UNFOLD_MIN_BLOCKS");  // occupancy experiments
        s += "#define UNFOLD_MIN_BLOCKS " + std::to_string(env ? std::max(1, atoi(env)) : 1) + "\n";
    }
    {
        char buf[32];
        snprintf(buf, sizeof buf, "0x%xu", m.used_mask);
        s += std::string("#define UNFOLD_USED_MASK ") + buf + "\n";
    }
    if (tiny_tune().custom)
        s += "#define TINY_THREADS_CFG " + std::to_string(tiny_tune().threads) + "\n#define TINY_CH_CFG " +
             std::to_string(tiny_tune().ch) + "\n#define TINY_UNROLL_CFG " + std::to_string(tiny_tune().unroll) + "\n";
    s += "#include \"unfold_skel.cuh\"\nusing namespace elisa;\n";
    s += generate_model_struct(m, P, false, "ModelV");
    s += generate_model_struct(m, P, true, "ModelG");
    const std::string k = std::to_string(ks);
    s += "extern \"C\" __global__ void __launch_bounds__(128, UNFOLD_MIN_BLOCKS) unfold_val(const UnfoldCore a) { unfold_generic<ModelV, double, " +
         k + ">(a); }\n";
    s += "extern \"C\" __global__ void __launch_bounds__(128, UNFOLD_MIN_BLOCKS) unfold_grad(const UnfoldCore a) { unfold_generic<ModelG, Dual<" +
         std::to_string(P) + ">, " + k + ">(a); }\n";
    const std::string edge_model = generate_edge_model(m, P);
    const std::string tree_model = edge_model.empty() ? generate_tree_model(m, P) : std::string();
    s += edge_model;
    s += tree_model;
    if (!edge_model.empty())
        s += "extern \"C\" __global__ void __launch_bounds__(128, UNFOLD_MIN_BLOCKS) unfold_contract(const UnfoldCore a) { "
             "unfold_contract_edges<ModelE>(a); }\n";
    else if (!tree_model.empty())
        s += "extern \"C\" __global__ void __launch_bounds__(128, UNFOLD_MIN_BLOCKS) unfold_contract(const UnfoldCore a) { "
             "unfold_contract_tree<ModelT>(a); }\n";
    else
        s += "extern \"C\" __global__ void __launch_bounds__(128, UNFOLD_MIN_BLOCKS) unfold_contract(const UnfoldCore a) { "
             "unfold_contract_generic<ModelG, Dual<" + std::to_string(P) + ">>(a); }\n";
    if (tiny_stat >= 0)
        s += "extern \"C\" __global__ void __launch_bounds__(TINY_THREADS) tiny_eval_kernel(const TinyCall t) { tiny_eval<ModelG, " +
             std::to_string(P) + ", " + std::to_string(tiny_stat) + ">(t); }\n";
    return s;
}

// NVRTC: source -> sm_100a cubin.  Needs no GPU.
static bool compile_unfold_source(const std::string& src, std::vector<char>* cubin, std::string* why) {
    NvrtcApi& api = nvrtc_api();
    if (!api.ok) {
        *why = "libnvrtc.so.12 not found";
        return false;
    }
    const char* headers[] = {k_src_skel, k_src_components, k_src_dual, k_src_elisa_h, k_src_slice, k_src_fastmath, k_src_stats};
    const char* names[] = {"unfold_skel.cuh", "components.cuh", "dual.cuh", "elisa_b200.h", "slice_i8.cuh", "fastmath.cuh", "stats.cuh"};
    nvrtcProgram prog = nullptr;
    if (api.create(&prog, src.c_str(), "elisa_unfold_jit.cu", 7, headers, names) != NVRTC_SUCCESS) {
        *why = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    const nvrtcResult rc = api.compile(prog, 3, opts);
    if (rc != NVRTC_SUCCESS) {
        size_t n = 0;
        api.log_size(prog, &n);
        std::string log(n, '\0');
        if (n) api.log(prog, &log[0]);
        *why = "nvrtc compile failed: " + log.substr(0, 4000);
        api.destroy(&prog);
        return false;
    }
    size_t n = 0;
    api.cubin_size(prog, &n);
    cubin->resize(n);
    api.cubin(prog, cubin->data());
    api.destroy(&prog);
    if (const char* path = getenv("ELISA_B200_DUMP_CUBIN")) {  // for `cuobjdump -sass` of the specialised kernels
        if (FILE* f = fopen(path, "wb")) {
            fwrite(cubin->data(), 1, n, f);
            fclose(f);
        }
    }
    return true;
}

// Compile (or fetch from the in-process cache) the specialised kernels of one model.
// Returns nullptr with `why` set when runtime compilation is not possible.
static const JitKernels* jit_unfold(const ModelDev& m, int P, int ks, std::string* why, int tiny_stat = -1) {
    static std::mutex mu;
    static std::map<std::string, JitKernels*> cache;
    const std::string src = generate_unfold_source(m, P, ks, tiny_stat);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(src);
    if (it != cache.end()) return it->second;
    std::vector<char> cubin;
    if (!compile_unfold_source(src, &cubin, why)) return nullptr;
    JitKernels* k = new JitKernels();
    cudaError_t e = cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->val, k->lib, "unfold_val");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->grad, k->lib, "unfold_grad");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k->contract, k->lib, "unfold_contract");
    if (e == cudaSuccess && tiny_stat >= 0) e = cudaLibraryGetKernel(&k->tiny, k->lib, "tiny_eval_kernel");
    if (e == cudaSuccess && tiny_stat >= 0)
        e = cudaFuncSetAttribute((const void*)k->tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TINY_MAX_SMEM);
    if (e != cudaSuccess) {
        *why = std::string("loading the compiled unfold kernels failed: ") + cudaGetErrorString(e);
        delete k;
        return nullptr;
    }
    cache[src] = k;
    return k;
}

// ---- optional device timing -----------------------------------------------------------
struct ScopedSpan {
    ElisaPlan* plan;
    cudaStream_t st;
    cudaEvent_t b = nullptr;
    ScopedSpan(ElisaPlan* p, int cls, cudaStream_t s) : plan(p), st(s) {
        if (plan == nullptr || !plan->profiling) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (plan->ev_used == plan->ev_pool.size()) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) return;
                plan->ev_pool.push_back(e);
            }
            ev[i] = plan->ev_pool[plan->ev_used++];
        }
        cudaEventRecord(ev[0], st);
        b = ev[1];
        plan->spans.push_back({cls, ev[0], ev[1]});
    }
    ~ScopedSpan() {
        if (b) cudaEventRecord(b, st);
    }
};

static void drain_spans(ElisaPlan* plan) {
    for (auto& sp : plan->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            plan->prof_ms[sp.cls] += ms;
            plan->prof_n[sp.cls] += 1;
        }
    }
    plan->spans.clear();
    plan->ev_used = 0;
}

// ---- launches ------------------------------------------------------------------------
// fused_digits (nullable): set to whether the kernel also wrote the digits of U into plan->A_img /
// a_scale (runtime-specialised kernel, sliced-integer fold, one warp per row); requesting it
// (non-null) is what asks for the fusion.
static int launch_unfold(ElisaPlan* plan, const SpectrumPlan& sp, const double* theta, int n_valid, int n_rows,
                         bool grad, double* U, double* dU, int ld, int e_pad, bool clip, cudaStream_t st,
                         bool* fused_digits = nullptr) {
    UnfoldArgs args;
    UnfoldCore& a = args.core;
    a.gv = sp.grid_view();
    a.theta = theta;
    a.P = plan->P;
    a.n_valid = n_valid;
    a.n_rows = n_rows;
    // enough warps to fill the machine: split the energy axis when there are few rows
    int n_seg = 1;
    const int64_t want_warps = 148 * 32;
    while ((int64_t)n_rows * n_seg < want_warps && (sp.E + n_seg - 1) / n_seg > 62) n_seg *= 2;
    int seg_len = ((sp.E + n_seg - 1) / n_seg + 30) / 31 * 31;
    n_seg = (sp.E + seg_len - 1) / seg_len;
    a.n_seg = n_seg;
    a.seg_len = seg_len;
    a.ld = ld;
    a.e_pad = e_pad;
    a.clip = clip ? 1 : 0;
    a.U = U;
    a.dU = (plan->debug_skip & 32) ? nullptr : dU;
    a.plane = (int64_t)n_rows * ld;
    a.used_mask = sp.model.used_mask;
    a.kb_total = round_up(ld, I8_BK) / I8_BK;
    a.img = nullptr;
    a.scale = nullptr;
    a.G = nullptr;
    a.ldg = 0;
    a.grad_out = nullptr;
    a.grad_stride = 0;
    a.kexp = sp.balanced_fwd ? sp.t_exp : nullptr;
    a.gv.has_point = 0;
    for (int k = 0; k < ELISA_MAX_NORMS; ++k) {
        a.ngv[k] = sp.norm_view(k);
        a.ngv[k].has_point = 0;
        a.nw[k] = sp.norm_grid[k].w;
    }
    if (fused_digits != nullptr) {
        *fused_digits = sp.jit != nullptr && (sp.use_i8() && !plan->force_dmma) && n_seg == 1 && plan->A_img != nullptr;
        if (*fused_digits) {
            a.img = plan->A_img;
            a.scale = plan->a_scale;
        }
    }
    const int64_t warps = (int64_t)n_rows * n_seg;
    const int blocks = (int)((warps + UNFOLD_WARPS - 1) / UNFOLD_WARPS);
    const int threads = UNFOLD_WARPS * 32;
    ScopedSpan span(plan, ELISA_PROFILE_UNFOLD, st);
    if (sp.jit != nullptr) {  // runtime-specialised kernel
        void* kargs[] = {(void*)&a};
        CUDA_TRY(cudaLaunchKernel((const void*)(grad ? sp.jit->grad : sp.jit->val), dim3(blocks), dim3(threads), kargs,
                                  0, st));
        ++plan->launches;
        return ELISA_OK;
    }
    args.model = sp.model;
    if (!grad)
        unfold_kernel<0><<<blocks, threads, 0, st>>>(args);
    else if (plan->P <= 4)
        unfold_kernel<4><<<blocks, threads, 0, st>>>(args);
    else if (plan->P <= 8)
        unfold_kernel<8><<<blocks, threads, 0, st>>>(args);
    else if (plan->P <= 16)
        unfold_kernel<16><<<blocks, threads, 0, st>>>(args);
    else
        unfold_kernel<32><<<blocks, threads, 0, st>>>(args);
    ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

// Gradient by contraction (unfold_contract_generic): grad[row, s, :] = sum_e G[row, e] du_e/dtheta.
static int launch_contract(ElisaPlan* plan, const SpectrumPlan& sp, int s, const double* theta, int n_valid,
                           const double* G, int ldg, double* grad, cudaStream_t st) {
    UnfoldArgs args;
    UnfoldCore& a = args.core;
    memset(&a, 0, sizeof a);
    a.gv = sp.grid_view();
    a.theta = theta;
    a.P = plan->P;
    a.n_valid = n_valid;
    a.n_rows = n_valid;
    a.n_seg = 1;
    a.seg_len = (sp.E + 30) / 31 * 31;
    a.clip = 1;
    a.used_mask = sp.model.used_mask;
    a.gv.has_point = 0;
    for (int k = 0; k < ELISA_MAX_NORMS; ++k) {
        a.ngv[k] = sp.norm_view(k);
        a.ngv[k].has_point = 0;
        a.nw[k] = sp.norm_grid[k].w;
    }
    a.G = G;
    a.ldg = ldg;
    a.grad_out = grad + (size_t)s * plan->P;
    a.grad_stride = (int64_t)plan->S * plan->P;
    const int blocks = (n_valid + UNFOLD_WARPS - 1) / UNFOLD_WARPS;
    const int threads = UNFOLD_WARPS * 32;
    ScopedSpan span(plan, ELISA_PROFILE_CONTRACT, st);
    if (sp.jit != nullptr) {
        void* kargs[] = {(void*)&a};
        CUDA_TRY(cudaLaunchKernel((const void*)sp.jit->contract, dim3(blocks), dim3(threads), kargs, 0, st));
        ++plan->launches;
        return ELISA_OK;
    }
    args.model = sp.model;
    if (plan->P <= 4)
        unfold_contract_kernel<4><<<blocks, threads, 0, st>>>(args);
    else if (plan->P <= 8)
        unfold_contract_kernel<8><<<blocks, threads, 0, st>>>(args);
    else if (plan->P <= 16)
        unfold_contract_kernel<16><<<blocks, threads, 0, st>>>(args);
    else
        unfold_contract_kernel<32><<<blocks, threads, 0, st>>>(args);
    ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

template <class Epi>
static int launch_gemm(ElisaPlan* plan, int cls, const double* A, int lda, const PackedDevice& B, int m_tiles,
                       const Epi& epi, cudaStream_t st) {
    static bool configured = false;  // per instantiation
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_f64_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GEMM_SMEM_BYTES));
        configured = true;
    }
    dim3 grid(B.n_tiles, m_tiles);
    ScopedSpan span(plan, cls, st);
    gemm_f64_kernel<Epi><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(A, lda, B.view(), epi);
    ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

template <int STAT, bool GRAD>
static int launch_stat_gemm(ElisaPlan* plan, const SpectrumPlan& sp, int n_valid, int m_tiles, const double* on,
                            const double* off, cudaStream_t st) {
    EpilogueStat<STAT, GRAD> epi;
    epi.ch = sp.chan_view();
    epi.counts_on = on ? on + sp.counts_off : nullptr;
    epi.counts_off = off ? off + sp.counts_off : nullptr;
    epi.ld_counts = plan->sum_C;
    epi.n_valid = n_valid;
    epi.C = sp.C;
    epi.W = plan->W;
    epi.ldw = sp.C_pad;
    epi.part = plan->part;
    epi.ld_part = plan->ld_part;
    return launch_gemm(plan, ELISA_PROFILE_FOLD, plan->U, sp.E_pad, sp.fwd, m_tiles, epi, st);
}

template <bool GRAD>
static int launch_stat_gemm_any(ElisaPlan* plan, const SpectrumPlan& sp, int n_valid, int m_tiles, const double* on,
                                const double* off, cudaStream_t st) {
    switch (sp.stat) {
        case ELISA_STAT_CHI2: return launch_stat_gemm<ELISA_STAT_CHI2, GRAD>(plan, sp, n_valid, m_tiles, on, off, st);
        case ELISA_STAT_CSTAT: return launch_stat_gemm<ELISA_STAT_CSTAT, GRAD>(plan, sp, n_valid, m_tiles, on, off, st);
        case ELISA_STAT_PSTAT: return launch_stat_gemm<ELISA_STAT_PSTAT, GRAD>(plan, sp, n_valid, m_tiles, on, off, st);
        case ELISA_STAT_PGSTAT: return launch_stat_gemm<ELISA_STAT_PGSTAT, GRAD>(plan, sp, n_valid, m_tiles, on, off, st);
        default: return launch_stat_gemm<ELISA_STAT_WSTAT, GRAD>(plan, sp, n_valid, m_tiles, on, off, st);
    }
}


// ---- sliced-integer fold (fold_i8.cuh) -----------------------------------------------------
#define ELISA_FOR_EACH_KS(X) X(4) X(5) X(6) X(7)

static int launch_slice(ElisaPlan* plan, int ks, const SliceArgs& a, int cls, cudaStream_t st) {
    const int rows = a.row_tiles * a.tile_rows;
    const int blocks = (rows + 7) / 8;
    ScopedSpan span(plan, cls, st);
    switch (ks) {
#define X(K) case K: slice_rows_kernel<K><<<blocks, 256, 0, st>>>(a); break;
        ELISA_FOR_EACH_KS(X)
#undef X
        default: return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
    }
    if (plan) ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

static int launch_lead_zero(ElisaPlan* plan, const I8Operand& op, cudaStream_t st) {
    if (op.lead == nullptr || op.n_images == 0) return ELISA_OK;
    switch (op.ks) {
#define X(K) case K: lead_zero_kernel<K><<<(unsigned)op.n_images, 256, 0, st>>>(op.img, op.n_images, I8_B_SLICE, op.lead); break;
        ELISA_FOR_EACH_KS(X)
#undef X
        default: return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
    }
    if (plan) ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

template <int KS, class Epi, int NI>
static int launch_fold_i8_ni(ElisaPlan* plan, int cls, const FoldI8Args& fa, const Epi& epi, int sm_count,
                             cudaStream_t st);

// env ELISA_B200_I8_ISSUERS=1|2 overrides the epilogue's issuer count (cross-checks)
template <int KS, class Epi>
static int launch_fold_i8_ks(ElisaPlan* plan, int cls, const FoldI8Args& fa, const Epi& epi, int sm_count,
                             cudaStream_t st) {
    const char* env = getenv("ELISA_B200_I8_ISSUERS");
    const int ni = env ? atoi(env) : Epi::ISSUERS;
    if (ni == 1) return launch_fold_i8_ni<KS, Epi, 1>(plan, cls, fa, epi, sm_count, st);
    if (ni == 2) return launch_fold_i8_ni<KS, Epi, 2>(plan, cls, fa, epi, sm_count, st);
    return launch_fold_i8_ni<KS, Epi, Epi::ISSUERS>(plan, cls, fa, epi, sm_count, st);
}

template <int KS, class Epi, int NI>
static int launch_fold_i8_ni(ElisaPlan* plan, int cls, const FoldI8Args& fa, const Epi& epi, int sm_count,
                             cudaStream_t st) {
    static bool configured = false;  // per instantiation
    if (!configured) {
        CUDA_TRY(cudaFuncSetAttribute(fold_i8_kernel<KS, Epi, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      i8_smem_bytes(KS)));
        configured = true;
    }
    FoldI8Args f = fa;
    f.n_groups = (f.n_tiles + I8_GROUP - 1) / I8_GROUP;
    const int grid = std::min(f.m_tiles * f.n_groups, sm_count);  // items: (row tile, column group)
    ScopedSpan span(plan, cls, st);
    fold_i8_kernel<KS, Epi, NI><<<grid, I8_THREADS, i8_smem_bytes(KS), st>>>(f, epi);
    if (plan) ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

template <class Epi>
static int launch_fold_i8(ElisaPlan* plan, int ks, int cls, const FoldI8Args& fa, const Epi& epi, int sm_count,
                          cudaStream_t st) {
    switch (ks) {
#define X(K) case K: return launch_fold_i8_ks<K>(plan, cls, fa, epi, sm_count, st);
        ELISA_FOR_EACH_KS(X)
#undef X
        default: return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
    }
}

static FoldI8Args fold_i8_args(const ElisaPlan* plan, const I8Operand& B, int kb_a, int m_tiles) {
    FoldI8Args fa;
    fa.A_img = plan->A_img;
    fa.a_scale = plan->a_scale;
    fa.kb_a = kb_a;
    fa.B_img = B.img;
    fa.b_scale = B.scale;
    fa.b_tile_img = B.tile_img;
    fa.kb_lo = B.kb_lo;
    fa.kb_hi = B.kb_hi;
    // Skipping of leading zero digit planes: OFF unless ELISA_B200_I8_SKIP=1.  On the bench response it
    // removes 44 % (forward) / 24 % (transposed) of the digit products and the kernel is no faster for
    // it (profiles/r02_digit_skip.md: the folds are bound by the latency of their load / issue
    // handshakes, not by the MMA count), the transposed fold even 6 % slower.
    static const bool skip = getenv("ELISA_B200_I8_SKIP") && atoi(getenv("ELISA_B200_I8_SKIP")) != 0;
    fa.b_lead = skip ? B.lead : nullptr;
    fa.m_tiles = m_tiles;
    fa.n_tiles = B.n_tiles;
    fa.n_groups = 1;  // set at launch
    fa.trap_info = trap_info_dev();
    static const int dbg = getenv("ELISA_B200_I8_DEBUG") ? atoi(getenv("ELISA_B200_I8_DEBUG")) : 0;  // timing experiments
    fa.debug = dbg;
    return fa;
}

// Digit images of a static operand: `dense` is [rows, K] row-major on the HOST (rows = output
// columns of the fold, K = contraction index).  Each 64-row tile stores only the k-blocks
// between its first and last non-zero (a band-sparse response keeps 2-3 of them).
static int build_i8_operand(ElisaPlan* plan, int ks, const std::vector<double>& dense, int rows, int K,
                            I8Operand* out) {
    const int n_tiles = round_up(rows, I8_BN) / I8_BN;
    const int kb_total = round_up(K, I8_BK) / I8_BK;
    std::vector<int64_t> tile_img(n_tiles);
    std::vector<int32_t> kb_lo(n_tiles), kb_hi(n_tiles);
    int64_t n_images = 0;
    for (int t = 0; t < n_tiles; ++t) {
        int lo = K, hi = 0;
        for (int r = t * I8_BN; r < std::min(rows, (t + 1) * I8_BN); ++r) {
            const double* x = dense.data() + (size_t)r * K;
            int a = 0, b = K;
            while (a < K && x[a] == 0.0) ++a;
            while (b > a && x[b - 1] == 0.0) --b;
            if (a < b) {
                lo = std::min(lo, a);
                hi = std::max(hi, b);
            }
        }
        if (lo >= hi) lo = hi = 0;
        kb_lo[t] = lo / I8_BK;
        kb_hi[t] = (hi + I8_BK - 1) / I8_BK;
        tile_img[t] = n_images;
        n_images += kb_hi[t] - kb_lo[t];
    }
    int rc;
    {
        std::vector<double> asum((size_t)n_tiles * I8_BN, 0.0);
        for (int r = 0; r < rows; ++r) {
            const double* x = dense.data() + (size_t)r * K;
            double t = 0.0;
            for (int k = 0; k < K; ++k) t += fabs(x[k]);
            asum[r] = t;
        }
        if ((rc = upload(plan, asum, &out->abs_sum))) return rc;
    }
    if ((rc = upload(plan, tile_img, &out->tile_img))) return rc;
    if ((rc = upload(plan, kb_lo, &out->kb_lo))) return rc;
    if ((rc = upload(plan, kb_hi, &out->kb_hi))) return rc;
    const size_t img_bytes = std::max<size_t>((size_t)n_images * ks * I8_B_SLICE, 16);
    CUDA_TRY(cudaMalloc((void**)&out->img, img_bytes));
    plan->owned.push_back(out->img);
    CUDA_TRY(cudaMalloc((void**)&out->scale, (size_t)n_tiles * I8_BN * 8));
    plan->owned.push_back(out->scale);
    CUDA_TRY(cudaMalloc((void**)&out->lead, std::max<size_t>((size_t)n_images, 16)));
    plan->owned.push_back(out->lead);
    CUDA_TRY(cudaMemset(out->lead, 0, std::max<size_t>((size_t)n_images, 16)));
    out->h_kb_lo = kb_lo;
    out->h_kb_hi = kb_hi;
    double* tmp = nullptr;
    CUDA_TRY(cudaMalloc((void**)&tmp, std::max<size_t>(dense.size() * 8, 16)));
    cudaError_t e = cudaMemcpy(tmp, dense.data(), dense.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        SliceArgs a;
        a.X = tmp;
        a.ld = K;
        a.rows = rows;
        a.cols = K;
        a.row_tiles = n_tiles;
        a.tile_rows = I8_BN;
        a.kb_total = kb_total;
        a.tile_img = out->tile_img;
        a.kb_lo = out->kb_lo;
        a.kb_hi = out->kb_hi;
        a.img = out->img;
        a.scale = out->scale;
        a.kscale = nullptr;
        a.abs_sum = nullptr;
        const bool prof = plan->profiling;
        plan->profiling = false;
        rc = launch_slice(plan, ks, a, ELISA_PROFILE_REDUCE, nullptr);
        plan->profiling = prof;
        if (rc == ELISA_OK) {
            out->ks = ks;
            out->n_images = n_images;
            rc = launch_lead_zero(plan, *out, nullptr);
        }
        if (rc == ELISA_OK) e = cudaDeviceSynchronize();
    }
    plan->owned.push_back(tmp);  // kept: the operand is re-sliced with column scales when the fold is balanced
    out->dense = tmp;
    out->rows = rows;
    out->K = K;
    out->ks = ks;
    if (rc) return rc;
    if (e != cudaSuccess) return fail(ELISA_ERR_CUDA, std::string("slicing the response failed: ") + cudaGetErrorString(e));
    out->n_tiles = n_tiles;
    out->n_images = n_images;
    plan->workspace_bytes += (int64_t)img_bytes + (int64_t)dense.size() * 8;
    return ELISA_OK;
}

template <int KS, int STAT, bool GRAD>
static int launch_stat_slice_ks(ElisaPlan* plan, const StatSliceArgs& a, cudaStream_t st) {
    const int blocks = (a.rows_pad + 7) / 8;
    ScopedSpan span(plan, ELISA_PROFILE_STAT, st);
    stat_slice_kernel<KS, STAT, GRAD><<<blocks, 256, 0, st>>>(a);
    ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

template <int STAT, bool GRAD>
static int launch_stat_slice(ElisaPlan* plan, const StatSliceArgs& a, cudaStream_t st) {
    switch (plan->ks_bwd) {  // the kernel slices W, the A operand of the transposed fold
#define X(K) case K: return launch_stat_slice_ks<K, STAT, GRAD>(plan, a, st);
        ELISA_FOR_EACH_KS(X)
#undef X
        default: return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
    }
}

template <bool GRAD>
static int launch_stat_slice_any(ElisaPlan* plan, const SpectrumPlan& sp, const StatSliceArgs& a, cudaStream_t st) {
    switch (sp.stat) {
        case ELISA_STAT_CHI2: return launch_stat_slice<ELISA_STAT_CHI2, GRAD>(plan, a, st);
        case ELISA_STAT_CSTAT: return launch_stat_slice<ELISA_STAT_CSTAT, GRAD>(plan, a, st);
        case ELISA_STAT_PSTAT: return launch_stat_slice<ELISA_STAT_PSTAT, GRAD>(plan, a, st);
        case ELISA_STAT_PGSTAT: return launch_stat_slice<ELISA_STAT_PGSTAT, GRAD>(plan, a, st);
        default: return launch_stat_slice<ELISA_STAT_WSTAT, GRAD>(plan, a, st);
    }
}

// digits of the per-chunk A operand X [n_rows, ld] (U or W) into plan->A_img / a_scale
static int slice_chunk_operand(ElisaPlan* plan, const double* X, int ld, int cols, int n_valid, int m_tiles, int cls,
                               cudaStream_t st, const double* kscale = nullptr) {
    SliceArgs a;
    a.X = X;
    a.ld = ld;
    a.rows = n_valid;
    a.cols = cols;
    a.row_tiles = m_tiles;
    a.tile_rows = I8_BM;
    a.kb_total = round_up(ld, I8_BK) / I8_BK;
    a.tile_img = nullptr;
    a.kb_lo = nullptr;
    a.kb_hi = nullptr;
    a.img = plan->A_img;
    a.scale = plan->a_scale;
    a.kscale = kscale;
    a.abs_sum = nullptr;
    return launch_slice(plan, plan->ks, a, cls, st);
}

// Re-slice a static operand with column scales (balancing); same images, ranges and buffers.
static int reslice_operand(ElisaPlan* plan, const I8Operand& op, const double* kscale, cudaStream_t st) {
    SliceArgs a;
    a.X = op.dense;
    a.ld = op.K;
    a.rows = op.rows;
    a.cols = op.K;
    a.row_tiles = op.n_tiles;
    a.tile_rows = I8_BN;
    a.kb_total = round_up(op.K, I8_BK) / I8_BK;
    a.tile_img = op.tile_img;
    a.kb_lo = op.kb_lo;
    a.kb_hi = op.kb_hi;
    a.img = op.img;
    a.scale = op.scale;
    a.kscale = kscale;
    a.abs_sum = op.abs_sum;
    const int rc = launch_slice(plan, op.ks, a, ELISA_PROFILE_SLICE, st);
    return rc ? rc : launch_lead_zero(plan, op, st);  // the scaled operand has other leading zeros
}

// forward fold (plain store) -> statistic + digits of W -> transposed fold with the gradient epilogue
static int run_chunk_i8(ElisaPlan* plan, const SpectrumPlan& sp, int s, int n, int m_tiles, bool want_grad,
                        bool have_digits, const double* on, const double* off, double* loglike, cudaStream_t st,
                        bool stop_after_stat = false, const double* theta = nullptr, double* grad = nullptr) {
    int rc;
    const int skip = plan->debug_skip;
    if (!have_digits && !(skip & 1) &&  // the unfold kernel did not slice U itself (interpreter kernel / split rows)
        (rc = slice_chunk_operand(plan, plan->U, sp.E_pad, sp.E, n, m_tiles, ELISA_PROFILE_SLICE, st,
                                  sp.balanced_fwd ? sp.t_inv : nullptr)))
        return rc;
    if (!(skip & 2)) {
        I8EpilogueStore es{plan->W, sp.C_pad, m_tiles * I8_BM, sp.C_pad, plan->a_scale};
        const FoldI8Args fa = fold_i8_args(plan, sp.i8_fwd, round_up(sp.E_pad, I8_BK) / I8_BK, m_tiles);
        if ((rc = launch_fold_i8(plan, plan->ks, ELISA_PROFILE_FOLD, fa, es, plan->sm_count, st))) return rc;
    }
    StatSliceArgs a;
    a.X = plan->W;
    a.ldx = sp.C_pad;
    a.ch = sp.chan_view();
    a.counts_on = on ? on + sp.counts_off : nullptr;
    a.counts_off = off ? off + sp.counts_off : nullptr;
    a.ld_counts = plan->sum_C;
    a.n_valid = n;
    a.rows_pad = m_tiles * I8_BM;
    a.C = sp.C;
    a.loglike = loglike + s;
    a.ll_stride = plan->S;
    a.img = plan->A_img;
    a.scale = plan->a_scale;
    a.kb_total = round_up(sp.C_pad, I8_BK) / I8_BK;
    a.colsum = sp.i8_fwd.abs_sum;
    a.guard = plan->guard;
    a.flag_list = stop_after_stat ? nullptr : plan->flag_list;  // pilot rows are not rows of the call
    a.flag_cap = plan->flag_cap;
    a.row_base = plan->row_base;
    a.guard_eta = ldexp(1.0, -8 * plan->ks);
    a.guard_eta_g = ldexp(1.0, -8 * plan->ks_bwd);
    a.guard_grad = 1.0;
    // trip level: 10x below the 1e-10 tolerance (the estimate is a typical-error model, not a bound:
    // tests/test_gpu_guard_sweep.py holds it against the DMMA fold over random shapes); the
    // BASELINE configurations sit at <= 2e-12
    static const double tol = getenv("ELISA_B200_GUARD_TOL") ? atof(getenv("ELISA_B200_GUARD_TOL")) : 1e-11;
    a.guard_tol = tol;
    a.kinv = sp.balanced_bwd ? sp.r_inv : nullptr;
    if (!(skip & 4)) {
        if (!want_grad) return launch_stat_slice_any<false>(plan, sp, a, st);
        if ((rc = launch_stat_slice_any<true>(plan, sp, a, st))) return rc;
    }
    if (!want_grad || stop_after_stat || (skip & 8)) return ELISA_OK;
    const FoldI8Args fa = fold_i8_args(plan, sp.i8_bwd, round_up(sp.C_pad, I8_BK) / I8_BK, m_tiles);
    if (plan->grad_contract && theta != nullptr && grad != nullptr) {
        // gradient by contraction: the transposed fold stores G = W . R^T like the forward fold stores
        // R u, and the component kernel contracts it with the tangents it recomputes in registers
        I8EpilogueStore es{plan->G, sp.E_pad, m_tiles * I8_BM, sp.E_pad, plan->a_scale};
        if ((rc = launch_fold_i8(plan, plan->ks_bwd, ELISA_PROFILE_TFOLD, fa, es, plan->sm_count, st))) return rc;
        if (skip & 64) return ELISA_OK;
        return launch_contract(plan, sp, s, theta, n, plan->G, sp.E_pad, grad, st);
    }
    const int n_part = I8_CG * ((sp.i8_bwd.n_tiles + I8_GROUP - 1) / I8_GROUP);
    // the gradient epilogue keeps <= 16 row partials in registers: more free parameters take
    // further launches of the transposed fold (the digit products are repeated; models with
    // more than 16 free parameters are rare)
    for (int p0 = 0; p0 < plan->P; p0 += I8_GRAD_PMAX) {
        const uint32_t mask = (sp.model.used_mask >> p0) & 0xffffu;
        if (mask == 0) continue;
        I8EpilogueGrad eg;
        eg.plane = (int64_t)m_tiles * I8_BM * sp.E_pad;
        eg.dU = plan->dU + (int64_t)p0 * eg.plane;
        eg.ldu = sp.E_pad;
        eg.P = std::min(plan->P - p0, I8_GRAD_PMAX);
        eg.used_mask = mask;
        eg.n_part = n_part;
        eg.ld_part = plan->ld_part;
        eg.gpart = plan->gpart + (int64_t)p0 * n_part * plan->ld_part;
        static const int dbg = getenv("ELISA_B200_I8_DEBUG") ? atoi(getenv("ELISA_B200_I8_DEBUG")) : 0;
        eg.dbg = dbg;
        if ((rc = launch_fold_i8(plan, plan->ks_bwd, ELISA_PROFILE_TFOLD, fa, eg, plan->sm_count, st))) return rc;
    }
    return ELISA_OK;
}

// Both static operands of one spectrum: rows of R^T (channels; contraction over energies) for the
// forward fold and rows of R (energies; contraction over channels) for the transposed fold.
static int build_spectrum_i8(ElisaPlan* plan, const ElisaSpectrumDesc& sd, SpectrumPlan* sp, int ks) {
    const int E = sp->E, C = sp->C;
    std::vector<double> R((size_t)E * C, 0.0), Rt((size_t)C * E, 0.0);
    if (sd.response_dense != nullptr) {
        memcpy(R.data(), sd.response_dense, R.size() * 8);
    } else {
        for (int c = 0; c < C; ++c)
            for (int i = sd.csr_indptr[c]; i < sd.csr_indptr[c + 1]; ++i)
                R[(size_t)sd.csr_indices[i] * C + c] += sd.csr_values[i];
    }
    for (int e = 0; e < E; ++e)
        for (int c = 0; c < C; ++c) Rt[(size_t)c * E + e] = R[(size_t)e * C + c];
    int rc;
    if ((rc = build_i8_operand(plan, ks, Rt, C, E, &sp->i8_fwd))) return rc;
    if ((rc = build_i8_operand(plan, plan->ks_bwd, R, E, C, &sp->i8_bwd))) return rc;
    return ELISA_OK;
}

// One (chunk, spectrum) pass.  theta/on/off/loglike/grad already point at the chunk's first row.
static bool tiny_ready(const ElisaPlan* plan, int s) {
    const SpectrumPlan& sp = plan->spec[s];
    return plan->tiny_args != nullptr && sp.tiny_R != nullptr && sp.jit != nullptr && sp.jit->tiny != nullptr;
}

// Small batch of small spectra in one kernel (tiny_eval, unfold_skel.cuh): one CTA per (row, spectrum) for the
// spectra [s0, s0 + count), which share one translation unit (same model expression and statistic).
static int launch_tiny(ElisaPlan* plan, int s0, int count, const double* theta, const double* on, const double* off, int n,
                       double* loglike, double* grad, cudaStream_t st, int64_t i0 = 0, bool fuse_gather = false,
                       bool last_launch = false) {
    TinyCall k;
    memset(&k, 0, sizeof k);
    if (fuse_gather) {  // results go to the peers' full arrays from inside the kernel; the last launch raises the flags
        const ElisaPlan::PeerGather& g = plan->gather;
        k.world = g.world;
        k.rank = g.rank;
        for (int r = 0; r < g.world; ++r) {
            k.peer_ll[r] = g.ll[r] + (g.row_offset + i0) * plan->S;
            k.peer_grad[r] = (grad != nullptr && !g.grad.empty()) ? g.grad[r] + (g.row_offset + i0) * plan->S * plan->P : nullptr;
            k.peer_flags[r] = g.flags[r];
        }
        k.done = g.done;
        k.epoch = last_launch ? g.epoch : 0;
    }
    k.spec = plan->tiny_args;
    k.s0 = s0;
    k.S = plan->S;
    k.P = plan->P;
    k.n_valid = n;
    k.theta = theta;
    k.counts_on = on;
    k.counts_off = off;
    k.ld_counts = plan->sum_C;
    k.loglike = loglike;
    k.grad = grad;
    k.clk = nullptr;
    static const bool clocks = getenv("ELISA_B200_TINY_CLOCKS") != nullptr;  // phase clocks of CTA (0, 0): a diagnostic, synchronises
    static long long* clk_dev = nullptr;
    if (clocks) {
        if (clk_dev == nullptr) CUDA_TRY(cudaMalloc(&clk_dev, 8 * sizeof(long long)));
        k.clk = clk_dev;
    }
    size_t smem = 0;
    for (int s = s0; s < s0 + count; ++s) smem = std::max(smem, tiny_smem_bytes(plan->spec[s].E, plan->P));
    ScopedSpan span(plan, ELISA_PROFILE_TINY, st);
    void* kargs[] = {(void*)&k};
    CUDA_TRY(cudaLaunchKernel((const void*)plan->spec[s0].jit->tiny, dim3(n, count), dim3(tiny_tune().threads), kargs, smem, st));
    if (clocks) {
        long long h[8];
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(h, clk_dev, sizeof h, cudaMemcpyDeviceToHost));
        fprintf(stderr, "tiny_eval clocks (n = %d x %d spectra): pre %lld, model %lld, barrier %lld, fold %lld, statistic %lld, reduce + store %lld\n",
                n, count, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
    }
    ++plan->launches;
    return ELISA_OK;
}

static int run_chunk(ElisaPlan* plan, int s, const double* theta, const double* on, const double* off, int n,
                     double* loglike, double* grad, cudaStream_t st) {
    const SpectrumPlan& sp = plan->spec[s];
    if (plan->tiny_now && tiny_ready(plan, s) && !plan->force_dmma && !plan->debug_skip)
        return launch_tiny(plan, s, 1, theta, on, off, n, loglike, grad, st);
    const int n_rows = round_up(n, GEMM_BM);
    const int m_tiles = n_rows / GEMM_BM;
    const bool want_grad = grad != nullptr;
    const bool i8 = (sp.use_i8() && !plan->force_dmma);
    // gradient by contraction (both fold engines): the forward pass needs the values only (no tangent
    // planes are written); the tangents are recomputed against G after the transposed fold
    const bool contract = want_grad && plan->grad_contract && plan->G != nullptr;
    int rc;
    bool fused_digits = false;
    if (plan->debug_skip & 1) {
        fused_digits = true;  // timing experiment: the workspace keeps what the last full pass left
    } else if ((rc = launch_unfold(plan, sp, theta, n, n_rows, want_grad && !contract, plan->U, plan->dU, sp.E_pad,
                                   sp.E_pad, true, st, i8 ? &fused_digits : nullptr)))
        return rc;
    if (i8) {
        if ((rc = run_chunk_i8(plan, sp, s, n, m_tiles, want_grad, fused_digits, on, off, loglike, st, false,
                               contract ? theta : nullptr, contract ? grad : nullptr)))
            return rc;
        if (!want_grad || contract) return ELISA_OK;  // the statistic kernel wrote loglike itself, the contraction grad
    } else if (contract) {
        if ((rc = launch_stat_gemm_any<true>(plan, sp, n, m_tiles, on, off, st))) return rc;
        EpilogueStore es{plan->G, sp.E_pad};  // G = W . R^T, plain store
        if ((rc = launch_gemm(plan, ELISA_PROFILE_TFOLD, plan->W, sp.C_pad, sp.bwd, m_tiles, es, st))) return rc;
    } else if (want_grad) {
        if ((rc = launch_stat_gemm_any<true>(plan, sp, n, m_tiles, on, off, st))) return rc;
        EpilogueGrad eg;
        eg.dU = plan->dU;
        eg.plane = (int64_t)n_rows * sp.E_pad;
        eg.ldu = sp.E_pad;
        eg.P = plan->P;
        eg.used_mask = sp.model.used_mask;
        eg.n_tiles = sp.bwd.n_tiles;
        eg.gpart = plan->gpart;
        eg.ld_part = plan->ld_part;
        if ((rc = launch_gemm(plan, ELISA_PROFILE_TFOLD, plan->W, sp.C_pad, sp.bwd, m_tiles, eg, st))) return rc;
    } else {
        if ((rc = launch_stat_gemm_any<false>(plan, sp, n, m_tiles, on, off, st))) return rc;
    }
    ReduceArgs ra;
    ra.part = i8 ? nullptr : plan->part;  // sliced-integer fold: loglike is already final
    ra.gpart = plan->gpart;
    ra.tiles_c = i8 ? I8_CG * ((sp.i8_fwd.n_tiles + I8_GROUP - 1) / I8_GROUP) : sp.fwd.n_tiles;  // partial sums per row
    ra.tiles_e = i8 ? I8_CG * ((sp.i8_bwd.n_tiles + I8_GROUP - 1) / I8_GROUP) : sp.bwd.n_tiles;
    ra.ld = plan->ld_part;
    ra.n = n;
    ra.P = plan->P;
    ra.S = plan->S;
    ra.s = s;
    ra.used_mask = sp.model.used_mask;
    ra.loglike = loglike;
    ra.grad = contract ? nullptr : grad;  // DMMA fold with the gradient by contraction: only the log-likelihood partials
    if (!(plan->debug_skip & 16)) {
        ScopedSpan span(plan, ELISA_PROFILE_REDUCE, st);
        reduce_kernel<<<(n + 127) / 128, 128, 0, st>>>(ra);
        ++plan->launches;
        CUDA_TRY(cudaGetLastError());
    }
    if (contract && !(plan->debug_skip & 64)) return launch_contract(plan, sp, s, theta, n, plan->G, sp.E_pad, grad, st);
    return ELISA_OK;
}

// ---- normal equations (batched re-fit) ------------------------------------------------------
template <int STAT>
static int launch_normal_eq(ElisaPlan* plan, const NormalEqArgs& a, cudaStream_t st) {
    const int blocks = (a.n + 7) / 8;
    ScopedSpan span(plan, ELISA_PROFILE_STAT, st);
    if (a.P <= 4)
        normal_eq_kernel<STAT, 4><<<blocks, 256, 0, st>>>(a);
    else if (a.P <= 8)
        normal_eq_kernel<STAT, 8><<<blocks, 256, 0, st>>>(a);
    else  // 136 accumulators per lane: spills, but re-fits with > 8 free parameters are rare
        normal_eq_kernel<STAT, 16><<<blocks, 256, 0, st>>>(a);
    ++plan->launches;
    CUDA_TRY(cudaGetLastError());
    return ELISA_OK;
}

// Unfold with tangents, then forward-fold u and every du/dtheta_j with the plain-store epilogue
// into plan->Xp: plane 0 = R u, plane 1 + j = R du/dtheta_j.
static int fold_tangent_planes(ElisaPlan* plan, int s, const double* theta, int n, cudaStream_t st) {
    const SpectrumPlan& sp = plan->spec[s];
    const int n_rows = round_up(n, GEMM_BM);
    const int m_tiles = n_rows / GEMM_BM;
    const int64_t xplane = (int64_t)plan->chunk * plan->max_C_pad;
    const bool i8 = sp.use_i8() && !plan->force_dmma;
    int rc;
    bool fused_digits = false;
    if ((rc = launch_unfold(plan, sp, theta, n, n_rows, true, plan->U, plan->dU, sp.E_pad, sp.E_pad, true, st,
                            i8 ? &fused_digits : nullptr)))
        return rc;
    for (int j = -1; j < plan->P; ++j) {  // j = -1: the model itself
        if (j >= 0 && !((sp.model.used_mask >> j) & 1u)) continue;
        const double* A = j < 0 ? plan->U : plan->dU + (int64_t)j * n_rows * sp.E_pad;
        double* X = plan->Xp + (int64_t)(1 + j) * xplane;
        if (i8) {
            if (!(j < 0 && fused_digits) &&
                (rc = slice_chunk_operand(plan, A, sp.E_pad, sp.E, n, m_tiles, ELISA_PROFILE_SLICE, st,
                                          sp.balanced_fwd ? sp.t_inv : nullptr)))
                return rc;
            if (j < 0 && plan->a_scale_u != nullptr)  // the guard needs the row scales of U after the tangents took their place
                CUDA_TRY(cudaMemcpyAsync(plan->a_scale_u, plan->a_scale, (size_t)n_rows * 8, cudaMemcpyDeviceToDevice, st));
            I8EpilogueStore es{X, sp.C_pad, m_tiles * I8_BM, sp.C_pad, plan->a_scale};
            const FoldI8Args fa = fold_i8_args(plan, sp.i8_fwd, round_up(sp.E_pad, I8_BK) / I8_BK, m_tiles);
            if ((rc = launch_fold_i8(plan, plan->ks, ELISA_PROFILE_FOLD, fa, es, plan->sm_count, st))) return rc;
        } else {
            EpilogueStore es{X, sp.C_pad};
            if ((rc = launch_gemm(plan, ELISA_PROFILE_FOLD, A, sp.E_pad, sp.fwd, m_tiles, es, st))) return rc;
        }
    }
    return ELISA_OK;
}

static int ensure_tangent_workspace(ElisaPlan* plan) {
    if (plan->Xp != nullptr) return ELISA_OK;
    const size_t bytes = (size_t)(1 + plan->P) * plan->chunk * plan->max_C_pad * 8;
    if (cudaMalloc((void**)&plan->Xp, bytes) != cudaSuccess)
        return fail(ELISA_ERR_NOMEM, "out of device memory (tangent spectra workspace)");
    plan->owned.push_back(plan->Xp);
    plan->workspace_bytes += (int64_t)bytes;
    return ELISA_OK;
}

// One (chunk, spectrum) pass of the normal equations: the folded tangent planes, then the
// per-replica sums.
static int run_normal_eq_chunk(ElisaPlan* plan, int s, const double* theta, const double* on, const double* off,
                               int n, double* loglike, double* grad, double* jtj, cudaStream_t st) {
    const SpectrumPlan& sp = plan->spec[s];
    int rc;
    if ((rc = fold_tangent_planes(plan, s, theta, n, st))) return rc;
    NormalEqArgs a;
    a.X = plan->Xp;
    a.plane = (int64_t)plan->chunk * plan->max_C_pad;
    a.ldx = sp.C_pad;
    a.ch = sp.chan_view();
    a.counts_on = on ? on + sp.counts_off : nullptr;
    a.counts_off = off ? off + sp.counts_off : nullptr;
    a.ld_counts = plan->sum_C;
    a.n = n;
    a.C = sp.C;
    a.P = plan->P;
    a.used_mask = sp.model.used_mask;
    a.accumulate = s > 0 ? 1 : 0;
    a.loglike = loglike;
    a.grad = grad;
    a.jtj = jtj;
    const bool i8 = sp.use_i8() && !plan->force_dmma;
    a.guard = i8 && plan->a_scale_u != nullptr ? plan->guard : nullptr;
    a.u_scale = plan->a_scale_u;
    a.colsum = sp.i8_fwd.abs_sum;
    a.guard_eta = ldexp(1.0, -8 * plan->ks);
    static const double tol = getenv("ELISA_B200_GUARD_TOL") ? atof(getenv("ELISA_B200_GUARD_TOL")) : 1e-11;
    a.guard_tol = tol;
    switch (sp.stat) {
        case ELISA_STAT_CHI2: return launch_normal_eq<ELISA_STAT_CHI2>(plan, a, st);
        case ELISA_STAT_CSTAT: return launch_normal_eq<ELISA_STAT_CSTAT>(plan, a, st);
        case ELISA_STAT_PSTAT: return launch_normal_eq<ELISA_STAT_PSTAT>(plan, a, st);
        case ELISA_STAT_PGSTAT: return launch_normal_eq<ELISA_STAT_PGSTAT>(plan, a, st);
        default: return launch_normal_eq<ELISA_STAT_WSTAT>(plan, a, st);
    }
}

// ---- per-channel forward-mode Jacobian --------------------------------------------------------
struct JacobianArgs {
    const double* X;  // tangent planes as in NormalEqArgs
    int64_t plane;
    int32_t ldx;
    ChanArrays ch;
    const double* counts_on;
    const double* counts_off;
    int64_t ld_counts;
    int32_t n, C, P, stat;
    uint32_t used_mask;
    double* x;      // [n, C] clipped model counts, nullable
    double* ds;     // [n, P, C], nullable
    double* dl_dx;  // [n, C], nullable
};

template <int STAT>
__device__ __forceinline__ void jacobian_one(const JacobianArgs& a, int row, int col) {
    ChannelData d;
    d.on = a.counts_on ? a.counts_on[(size_t)row * a.ld_counts + col] : a.ch.on[col];
    d.off = a.counts_off ? a.counts_off[(size_t)row * a.ld_counts + col] : a.ch.off[col];
    d.sigma = a.ch.sigma[col];
    d.ratio = a.ch.ratio[col];
    d.gof_on = poisson_gof(d.on);
    d.gof_off = poisson_gof(d.off);
    d.den = a.ch.den[col];
    const double sc = a.ch.scale[col];
    const double raw = a.X[(size_t)row * a.ldx + col] * sc;
    double pass;
    // pstat clips model + background, the others the source counts (likelihood.py:208, 301-302)
    const double shift = STAT == ELISA_STAT_PSTAT ? d.ratio * d.off : 0.0;
    const double xc = clip_counts(raw + shift, pass);
    const size_t o = (size_t)row * a.C + col;
    if (a.x) a.x[o] = xc;
    if (a.dl_dx) {
        double dx = 0.0;
        stat_channel<STAT, true>(raw, d, dx);
        a.dl_dx[o] = dx;
    }
    if (a.ds)
        for (int p = 0; p < a.P; ++p)
            a.ds[((size_t)row * a.P + p) * a.C + col] =
                ((a.used_mask >> p) & 1u) ? pass * sc * a.X[(size_t)(1 + p) * a.plane + (size_t)row * a.ldx + col] : 0.0;
}

__global__ void jacobian_kernel(const JacobianArgs a) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = blockIdx.y;
    if (col >= a.C || row >= a.n) return;
    switch (a.stat) {
        case ELISA_STAT_CHI2: jacobian_one<ELISA_STAT_CHI2>(a, row, col); break;
        case ELISA_STAT_CSTAT: jacobian_one<ELISA_STAT_CSTAT>(a, row, col); break;
        case ELISA_STAT_PSTAT: jacobian_one<ELISA_STAT_PSTAT>(a, row, col); break;
        case ELISA_STAT_PGSTAT: jacobian_one<ELISA_STAT_PGSTAT>(a, row, col); break;
        default: jacobian_one<ELISA_STAT_WSTAT>(a, row, col); break;
    }
}

// Balance the integer folds of one spectrum around a pilot batch (<= 128 rows of the call that
// finds the plan unbalanced, gathered by ensure_balanced): power-of-two envelopes of U over the energies and of W over the
// channels, folded into re-sliced response operands.  Everything is enqueued on `st`; nothing is
// read back.  See col_envelope_kernel for why, DESIGN.md section 2 for the accuracy model.
static int calibrate_spectrum(ElisaPlan* plan, int s, const double* theta, const double* on, const double* off,
                              int np, cudaStream_t st) {
    SpectrumPlan& sp = plan->spec[s];
    const bool prof = plan->profiling;
    plan->profiling = false;
    int rc = ELISA_OK;
    do {
        sp.balanced_fwd = sp.balanced_bwd = false;
        if ((rc = launch_unfold(plan, sp, theta, np, I8_BM, false, plan->U, nullptr, sp.E_pad, sp.E_pad, true, st))) break;
        unsigned long long* peak = plan->guard + 3;  // fourth word of the guard block: scratch
        cudaMemsetAsync(peak, 0, 8, st);
        col_peak_kernel<<<(sp.E_pad + 127) / 128, 128, 0, st>>>(plan->U, sp.E_pad, np, sp.E, peak);
        const int ek = round_up(sp.E_pad, I8_BK), ck = round_up(sp.C_pad, I8_BK);
        col_envelope_kernel<<<ek / 128, 128, 0, st>>>(plan->U, sp.E_pad, np, sp.E, ek, 1, peak, sp.t_scale, sp.t_inv,
                                                      sp.t_exp);
        plan->launches += 2;
        if ((rc = reslice_operand(plan, sp.i8_fwd, sp.t_scale, st))) break;
        sp.balanced_fwd = true;
        // forward half of the chunk pass on the pilot rows: W of the pilot lands in plan->W
        if ((rc = run_chunk_i8(plan, sp, 0, np, 1, true, false, on, off, plan->cal_ll, st, true))) break;
        cudaMemsetAsync(peak, 0, 8, st);
        col_peak_kernel<<<(sp.C_pad + 127) / 128, 128, 0, st>>>(plan->W, sp.C_pad, np, sp.C, peak);
        col_envelope_kernel<<<ck / 128, 128, 0, st>>>(plan->W, sp.C_pad, np, sp.C, ck, 4, peak, sp.r_scale, sp.r_inv,
                                                      sp.r_exp);
        plan->launches += 2;
        if ((rc = reslice_operand(plan, sp.i8_bwd, sp.r_scale, st))) break;
        sp.balanced_bwd = true;
        if (cudaMemsetAsync(plan->guard, 0, 3 * sizeof(unsigned long long), st) != cudaSuccess) {
            rc = fail(ELISA_ERR_CUDA, "resetting the guard after calibration failed");
            break;
        }
    } while (false);
    plan->profiling = prof;
    if (rc == ELISA_OK && cudaGetLastError() != cudaSuccess) rc = fail(ELISA_ERR_CUDA, "calibration launch failed");
    return rc;
}

static int ensure_balanced(ElisaPlan* plan, const double* theta, const double* on, const double* off, int64_t n,
                           cudaStream_t st) {
    if (!plan->balance || plan->force_dmma || plan->cal_ll == nullptr) return ELISA_OK;
    bool staged = false;
    for (int s = 0; s < plan->S; ++s) {
        SpectrumPlan& sp = plan->spec[s];
        if (!sp.use_i8() || (sp.balanced_fwd && sp.balanced_bwd && !sp.stale) || sp.t_scale == nullptr) continue;
        if (!staged && n > I8_BM) {
            // the pilot: 128 rows spread evenly over the batch (a scan over a parameter grid changes
            // systematically from its first to its last row), gathered into scratch
            const int64_t stride = n / I8_BM;
            auto gather = [&](double** dst, const double* src, int64_t width) -> int {
                if (*dst == nullptr) {
                    CUDA_TRY(cudaMalloc((void**)dst, (size_t)I8_BM * width * 8));
                    plan->owned.push_back(*dst);
                }
                CUDA_TRY(cudaMemcpy2DAsync(*dst, (size_t)width * 8, src, (size_t)stride * width * 8, (size_t)width * 8,
                                           I8_BM, cudaMemcpyDefault, st));  // counts may live in host memory (host entry point)
                return ELISA_OK;
            };
            int rc = gather(&plan->cal_theta, theta, plan->P);
            if (rc == ELISA_OK && on != nullptr) rc = gather(&plan->cal_on, on, plan->sum_C);
            if (rc == ELISA_OK && off != nullptr) rc = gather(&plan->cal_off, off, plan->sum_C);
            if (rc) return rc;
            theta = plan->cal_theta;
            if (on != nullptr) on = plan->cal_on;
            if (off != nullptr) off = plan->cal_off;
            staged = true;
        }
        // (the flags only ever change inside calibrate_spectrum, together with the operands they
        // describe: kernels captured in a CUDA graph keep pointing at the same, rewritten arrays)
        const int rc = calibrate_spectrum(plan, s, theta, on, off, (int)std::min<int64_t>(n, I8_BM), st);
        if (rc) return rc;
        sp.stale = false;
    }
    return ELISA_OK;
}

// Sets plan->tiny_now for a call of n rows (tiny_call_n rows when the call is evaluated in pieces) and tells
// whether EVERY spectrum then takes the one-kernel path.
static bool decide_tiny(ElisaPlan* plan, int64_t n) {
    plan->tiny_now = plan->tiny_enabled && (plan->tiny_call_n >= 0 ? plan->tiny_call_n : n) <= TINY_MAX_N;
    bool all_tiny = plan->tiny_now && !plan->force_dmma && !plan->debug_skip;
    for (int s = 0; s < plan->S && all_tiny; ++s) all_tiny = tiny_ready(plan, s);
    return all_tiny;
}

// ---- peer gather -------------------------------------------------------------------------
constexpr int GATHER_MAX_WORLD = 16;
struct GatherFlags {
    long long* flags[GATHER_MAX_WORLD];
};
// one thread per peer: this rank's slot of the peer's flag array takes the call's sequence number (after the
// copies of the stream: the peer finds the rows behind the flag)
__global__ void gather_signal_kernel(GatherFlags f, int world, int rank, long long epoch) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    __threadfence_system();
    *reinterpret_cast<volatile long long*>(f.flags[r] + rank) = epoch;
    __threadfence_system();
}
// one thread per peer: wait until its flag in THIS rank's array has reached the sequence number.  Bounded
// (60 s, ELISA_B200_GATHER_TIMEOUT_S): a missing peer leaves a mark instead of hanging the device.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__global__ void gather_wait_kernel(const long long* flags, int world, int rank, long long epoch, int32_t* timeout,
                                   unsigned long long limit_ns) {
    const int r = threadIdx.x;
    if (r >= world || r == rank) return;
    const volatile long long* f = flags + r;
    const unsigned long long t0 = global_ns();
    while (*f < epoch) {
        __nanosleep(200);
        if (global_ns() - t0 > limit_ns) {
            if (timeout != nullptr) *timeout = r + 1;
            break;
        }
    }
    __threadfence_system();
}

static bool gather_on(const ElisaPlan* plan) { return plan->gather.world > 1; }

// rows [lo, hi) of the call's outputs -> every rank's full arrays (copy engines, behind the work on `st` so far)
static int gather_push(ElisaPlan* plan, int64_t lo, int64_t hi, const double* loglike, const double* grad, cudaStream_t st) {
    ElisaPlan::PeerGather& g = plan->gather;
    if (hi <= lo) return ELISA_OK;
    const int64_t S = plan->S, P = plan->P;
    CUDA_TRY(cudaEventRecord(g.ready, st));
    CUDA_TRY(cudaStreamWaitEvent(g.copy, g.ready, 0));
    for (int k = 1; k <= g.world; ++k) {  // the peers first, starting with the next rank: the ranks do not all hit rank 0 at once
        const int r = (g.rank + k) % g.world;
        double* dl = g.ll[r] + (g.row_offset + lo) * S;
        const double* sl = loglike + lo * S;
        if (dl != sl) CUDA_TRY(cudaMemcpyAsync(dl, sl, (size_t)(hi - lo) * S * 8, cudaMemcpyDeviceToDevice, g.copy));
        if (grad != nullptr && !g.grad.empty()) {
            double* dg = g.grad[r] + (g.row_offset + lo) * S * P;
            const double* sg = grad + lo * S * P;
            if (dg != sg) CUDA_TRY(cudaMemcpyAsync(dg, sg, (size_t)(hi - lo) * S * P * 8, cudaMemcpyDeviceToDevice, g.copy));
        }
    }
    return ELISA_OK;
}

// called after every chunk of run_all: pushes every `every` chunks
static int gather_note(ElisaPlan* plan, int64_t lo, int64_t hi, bool last, const double* loglike, const double* grad,
                       cudaStream_t st) {
    ElisaPlan::PeerGather& g = plan->gather;
    if (g.pend_chunks == 0) g.pend_lo = lo;
    g.pend_hi = hi;
    ++g.pend_chunks;
    if (!last && g.pend_chunks < g.every) return ELISA_OK;
    g.pend_chunks = 0;
    return gather_push(plan, g.pend_lo, g.pend_hi, loglike, grad, st);
}

// end of a gathered call: flags to the peers behind the pushes, `st` waits for the pushes and for the peers' flags
static int gather_finish(ElisaPlan* plan, cudaStream_t st) {
    ElisaPlan::PeerGather& g = plan->gather;
    if (g.timeout != nullptr && *g.timeout != 0) {
        const int r = *g.timeout - 1;
        *g.timeout = 0;
        return fail(ELISA_ERR_CUDA, "peer gather: the flag of rank " + std::to_string(r) + " did not arrive in an earlier call (the ranks must make the same sequence of gathered calls)");
    }
    ScopedSpan span(plan, ELISA_PROFILE_GATHER, st);
    if (!g.signalled) {
        GatherFlags f;
        for (int r = 0; r < GATHER_MAX_WORLD; ++r) f.flags[r] = r < g.world ? g.flags[r] : nullptr;
        gather_signal_kernel<<<1, 32, 0, g.copy>>>(f, g.world, g.rank, g.epoch);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(g.pushed, g.copy));
        CUDA_TRY(cudaStreamWaitEvent(st, g.pushed, 0));
        ++plan->launches;
    }
    g.signalled = false;
    gather_wait_kernel<<<1, 32, 0, st>>>(g.flags[g.rank], g.world, g.rank, g.epoch, g.timeout, g.limit_ns);
    CUDA_TRY(cudaGetLastError());
    ++plan->launches;
    return ELISA_OK;
}

// start of a gathered call: its sequence number
static void gather_begin(ElisaPlan* plan) {
    ++plan->gather.epoch;
    plan->gather.signalled = false;
    plan->gather.pend_chunks = 0;
}

static int run_all(ElisaPlan* plan, const double* theta, const double* on, const double* off, int64_t n,
                   double* loglike, double* grad, cudaStream_t st) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr || loglike == nullptr) return fail(ELISA_ERR_INVALID, "theta / loglike is null");
    int dev;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != plan->device) CUDA_TRY(cudaSetDevice(plan->device));
    if (decide_tiny(plan, n)) {
        // every spectrum takes the one-kernel path: spectra sharing a translation unit (a joint fit of like
        // detectors) are ONE launch; nothing else runs (no envelopes, no workspace, no guard)
        int rc = ELISA_OK;
        // peer gather: the kernel stores into the peers' arrays itself and its last CTA raises the flags
        const bool fuse = gather_on(plan) && plan->gather.world <= TINY_MAX_PEERS && plan->gather.done != nullptr;
        for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += plan->chunk) {
            const int nc = (int)std::min<int64_t>(plan->chunk, n - i0);
            for (int s0 = 0; s0 < plan->S && rc == ELISA_OK;) {
                int s1 = s0 + 1;
                while (s1 < plan->S && plan->spec[s1].jit == plan->spec[s0].jit) ++s1;
                rc = launch_tiny(plan, s0, s1 - s0, theta + i0 * plan->P, on ? on + i0 * plan->sum_C : nullptr,
                                 off ? off + i0 * plan->sum_C : nullptr, nc, loglike + i0 * plan->S,
                                 grad ? grad + i0 * plan->S * plan->P : nullptr, st, i0, fuse, i0 + nc >= n && s1 >= plan->S);
                s0 = s1;
            }
            if (rc == ELISA_OK && gather_on(plan) && !fuse)
                rc = gather_note(plan, i0, i0 + nc, i0 + nc >= n, loglike, grad, st);
        }
        if (rc == ELISA_OK && fuse) plan->gather.signalled = true;
        if (dev != plan->device) cudaSetDevice(dev);
        return rc;
    }
    // the integer folds need their envelopes only when some spectrum of this call runs through them
    int rc = ensure_balanced(plan, theta, on, off, n, st);
    const int64_t base = plan->row_base;  // global index of row 0 of this call (host path: of the piece)
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += plan->chunk) {
        const int nc = (int)std::min<int64_t>(plan->chunk, n - i0);
        plan->row_base = base + i0;
        for (int s = 0; s < plan->S && rc == ELISA_OK; ++s)
            rc = run_chunk(plan, s, theta + i0 * plan->P, on ? on + i0 * plan->sum_C : nullptr,
                           off ? off + i0 * plan->sum_C : nullptr, nc, loglike + i0 * plan->S,
                           grad ? grad + i0 * plan->S * plan->P : nullptr, st);
        if (rc == ELISA_OK && gather_on(plan)) rc = gather_note(plan, i0, i0 + nc, i0 + nc >= n, loglike, grad, st);
    }
    plan->row_base = base;
    if (dev != plan->device) cudaSetDevice(dev);
    return rc;
}

constexpr int64_t GRAPH_MAX_N = 1024;

static void patch_memcpy(cudaGraphExec_t exec, cudaGraphNode_t node, void* dst, const void* src, size_t bytes) {
    cudaGraphExecMemcpyNodeSetParams1D(exec, node, dst, src, bytes, cudaMemcpyDeviceToDevice);
}

// last node captured on `st` so far (the memcpy just enqueued)
static cudaGraphNode_t last_captured_node(cudaStream_t st) {
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t* deps = nullptr;
    size_t n_deps = 0;
    if (cudaStreamGetCaptureInfo(st, &status, nullptr, nullptr, &deps, &n_deps) != cudaSuccess || n_deps == 0)
        return nullptr;
    return deps[n_deps - 1];
}

// value (+ gradient) of a small batch through a replayed CUDA graph; falls back to run_all.
static int run_small(ElisaPlan* plan, const double* theta, const double* on, const double* off, int64_t n,
                     double* loglike, double* grad, cudaStream_t st) {
    const bool eligible = plan != nullptr && plan->graphs_enabled && !plan->profiling && !plan->force_dmma && !gather_on(plan) && n > 0 && n <= GRAPH_MAX_N &&
                          n <= plan->chunk && on == nullptr && off == nullptr && theta != nullptr && loglike != nullptr &&
                          plan->g_theta != nullptr;
    if (!eligible) return run_all(plan, theta, on, off, n, loglike, grad, st);
    // every spectrum on the one-kernel path: a launch or two on the caller's stream, nothing to replay
    if (decide_tiny(plan, n)) return run_all(plan, theta, on, off, n, loglike, grad, st);
    {  // envelopes of the balanced integer fold: (re)derived outside any capture, on the caller's stream
        int dev;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev != plan->device) CUDA_TRY(cudaSetDevice(plan->device));
        const int rc = ensure_balanced(plan, theta, on, off, n, st);
        if (dev != plan->device) cudaSetDevice(dev);
        if (rc) return rc;
    }
    ElisaPlan::GraphEntry* e = nullptr;
    for (auto& g : plan->graphs)
        if (g.n == n && g.grad == (grad != nullptr)) e = &g;
    if (e == nullptr) {
        if (plan->graphs.size() >= 16) return run_all(plan, theta, on, off, n, loglike, grad, st);
        plan->graphs.emplace_back();
        e = &plan->graphs.back();
        e->n = n;
        e->grad = grad != nullptr;
    }
    const size_t b_theta = (size_t)n * plan->P * 8, b_ll = (size_t)n * plan->S * 8, b_grad = (size_t)n * plan->S * plan->P * 8;
    if (e->exec == nullptr) {
        if (++e->uses < 2)  // first use of a shape runs directly (one-time kernel configuration, JIT loads)
            return run_all(plan, theta, on, off, n, loglike, grad, st);
        int dev;
        CUDA_TRY(cudaGetDevice(&dev));
        if (dev != plan->device) CUDA_TRY(cudaSetDevice(plan->device));
        cudaStream_t cs = plan->stream;
        int rc = ELISA_OK;
        bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            ok = cudaMemcpyAsync(plan->g_theta, theta, b_theta, cudaMemcpyDeviceToDevice, cs) == cudaSuccess;
            e->in_theta = last_captured_node(cs);
            const int64_t before = plan->launches;
            if (ok && !plan->small_ws.empty() && plan->small_fork != nullptr) {
                // one branch per spectrum: fork from the capture stream, run the spectrum's chain in
                // its own workspace on its own stream, join
                struct Saved { double *U, *dU, *W, *part, *gpart, *a_scale, *G; int8_t* A_img; int ld; } keep = {
                    plan->U, plan->dU, plan->W, plan->part, plan->gpart, plan->a_scale, plan->G, plan->A_img, plan->ld_part};
                ok = cudaEventRecord(plan->small_fork, cs) == cudaSuccess;
                for (int sidx = 0; sidx < plan->S && ok && rc == ELISA_OK; ++sidx) {
                    const ElisaPlan::SmallWs& w = plan->small_ws[sidx];
                    cudaStream_t bs = plan->small_stream[sidx];
                    ok = cudaStreamWaitEvent(bs, plan->small_fork, 0) == cudaSuccess;
                    plan->U = w.U; plan->dU = w.dU; plan->W = w.W; plan->part = w.part; plan->gpart = w.gpart;
                    plan->a_scale = w.a_scale; plan->A_img = w.A_img; plan->G = w.G;
                    plan->ld_part = (int)std::min<int64_t>(GRAPH_MAX_N, plan->chunk);
                    if (ok) rc = run_chunk(plan, sidx, plan->g_theta, nullptr, nullptr, (int)n, plan->g_ll, grad ? plan->g_grad : nullptr, bs);
                    if (ok) ok = cudaEventRecord(plan->small_done[sidx], bs) == cudaSuccess &&
                                 cudaStreamWaitEvent(cs, plan->small_done[sidx], 0) == cudaSuccess;
                }
                plan->U = keep.U; plan->dU = keep.dU; plan->W = keep.W; plan->part = keep.part; plan->gpart = keep.gpart;
                plan->a_scale = keep.a_scale; plan->A_img = keep.A_img; plan->G = keep.G; plan->ld_part = keep.ld;
            } else if (ok) rc = run_all(plan, plan->g_theta, nullptr, nullptr, n, plan->g_ll, grad ? plan->g_grad : nullptr, cs);
            e->kernels = plan->launches - before;
            plan->launches = before;  // counted per replay below
            ok = ok && rc == ELISA_OK;
            if (ok) ok = cudaMemcpyAsync(loglike, plan->g_ll, b_ll, cudaMemcpyDeviceToDevice, cs) == cudaSuccess;
            e->out_ll = last_captured_node(cs);
            if (ok && grad) {
                ok = cudaMemcpyAsync(grad, plan->g_grad, b_grad, cudaMemcpyDeviceToDevice, cs) == cudaSuccess;
                e->out_grad = last_captured_node(cs);
            }
            cudaGraph_t graph = nullptr;
            const bool ended = cudaStreamEndCapture(cs, &graph) == cudaSuccess && graph != nullptr;
            ok = ok && ended && e->in_theta && e->out_ll && (!grad || e->out_grad);
            if (ok) ok = cudaGraphInstantiate(&e->exec, graph, 0) == cudaSuccess;
            if (ok) {
                e->graph = graph;
            } else {
                if (graph) cudaGraphDestroy(graph);
                e->exec = nullptr;
            }
        }
        if (dev != plan->device) cudaSetDevice(dev);
        if (!ok) {  // capture is an optimisation: never fail the call because of it
            cudaGetLastError();
            plan->graphs_enabled = false;
            return run_all(plan, theta, on, off, n, loglike, grad, st);
        }
    }
    patch_memcpy(e->exec, e->in_theta, plan->g_theta, theta, b_theta);
    patch_memcpy(e->exec, e->out_ll, loglike, plan->g_ll, b_ll);
    if (grad) patch_memcpy(e->exec, e->out_grad, grad, plan->g_grad, b_grad);
    CUDA_TRY(cudaGraphLaunch(e->exec, st));
    plan->launches += e->kernels;
    return ELISA_OK;
}

// ---- accuracy guard: per-row fallback ---------------------------------------------------------
__global__ void gather_rows_kernel(const double* src, int64_t width, const int32_t* idx, int m, double* dst) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c < width && i < m) dst[(size_t)i * width + c] = src[(size_t)idx[i] * width + c];
}
__global__ void scatter_rows_kernel(const double* src, int64_t width, const int32_t* idx, int m, double* dst) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c < width && i < m) dst[(size_t)idx[i] * width + c] = src[(size_t)i * width + c];
}

static int read_guard(ElisaPlan* plan, cudaStream_t st, unsigned long long h[4]) {
    CUDA_TRY(cudaMemcpyAsync(plan->guard_host, plan->guard, 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(h, plan->guard_host, 32);
    return ELISA_OK;
}

// Re-evaluate the listed rows (sorted, unique, global indices) with the FP64 DMMA fold and put
// their results in place.  theta / loglike / grad are device arrays of the whole call; the
// per-replica counts are device arrays too, or host arrays when counts_on_host (the host entry
// point streams them piecewise and never holds them on the device as a whole).
static int rerun_rows_dmma(ElisaPlan* plan, const std::vector<int32_t>& rows, const double* theta, const double* on,
                           const double* off, bool counts_on_host, double* loglike, double* grad, cudaStream_t st) {
    const int m = (int)rows.size();
    if (m == 0) return ELISA_OK;
    const int P = plan->P, S = plan->S;
    const int64_t W = plan->sum_C;
    std::vector<void*> tmp;
    auto dalloc = [&](size_t doubles) -> double* {
        void* p = nullptr;
        if (cudaMalloc(&p, std::max<size_t>(doubles * 8, 16)) != cudaSuccess) return nullptr;
        tmp.push_back(p);
        return (double*)p;
    };
    auto release = [&]() {
        for (void* p : tmp) cudaFree(p);
    };
    int32_t* idx = (int32_t*)dalloc((m + 1) / 2);
    double* th_c = dalloc((size_t)m * P);
    double* ll_c = dalloc((size_t)m * S);
    double* g_c = grad ? dalloc((size_t)m * S * P) : nullptr;
    double* on_c = on ? dalloc((size_t)m * W) : nullptr;
    double* off_c = off ? dalloc((size_t)m * W) : nullptr;
    if (!idx || !th_c || !ll_c || (grad && !g_c) || (on && !on_c) || (off && !off_c)) {
        release();
        return fail(ELISA_ERR_NOMEM, "out of device memory (guard fallback rows)");
    }
    int rc = ELISA_OK;
    std::vector<double> pack;
    auto go = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(idx, rows.data(), (size_t)m * 4, cudaMemcpyHostToDevice, st));
        auto gather = [&](const double* src, int64_t width, double* dst) {
            dim3 grid((unsigned)((width + 255) / 256), (unsigned)m);
            gather_rows_kernel<<<grid, 256, 0, st>>>(src, width, idx, m, dst);
            ++plan->launches;
        };
        gather(theta, P, th_c);
        for (int k = 0; k < 2; ++k) {
            const double* src = k == 0 ? on : off;
            double* dst = k == 0 ? on_c : off_c;
            if (src == nullptr) continue;
            if (counts_on_host) {
                pack.resize((size_t)m * W);
                for (int i = 0; i < m; ++i) memcpy(&pack[(size_t)i * W], src + (size_t)rows[i] * W, (size_t)W * 8);
                CUDA_TRY(cudaMemcpyAsync(dst, pack.data(), (size_t)m * W * 8, cudaMemcpyHostToDevice, st));
                CUDA_TRY(cudaStreamSynchronize(st));  // `pack` is reused for the second array
            } else {
                gather(src, W, dst);
            }
        }
        const bool keep = plan->force_dmma;
        const int64_t base = plan->row_base;
        plan->force_dmma = true;
        plan->row_base = 0;
        const int r = run_all(plan, th_c, on_c, off_c, m, ll_c, g_c, st);
        plan->force_dmma = keep;
        plan->row_base = base;
        if (r) return r;
        auto scatter = [&](const double* src, int64_t width, double* dst) {
            dim3 grid((unsigned)((width + 255) / 256), (unsigned)m);
            scatter_rows_kernel<<<grid, 256, 0, st>>>(src, width, idx, m, dst);
            ++plan->launches;
        };
        scatter(ll_c, S, loglike);
        if (grad) scatter(g_c, (int64_t)S * P, grad);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        return ELISA_OK;
    };
    rc = go();
    release();
    return rc;
}

// One guarded evaluation: eval_all() enqueues the whole call on `st` with the plan's current fold
// engine; eval_rows(rows) re-evaluates a set of rows with the DMMA fold.  Synchronises `st`.
template <class EvalAll, class EvalRows>
static int run_guarded(ElisaPlan* plan, int64_t n, cudaStream_t st, EvalAll&& eval_all, EvalRows&& eval_rows) {
    const bool guard_on = plan->guard != nullptr && !plan->force_dmma && !plan->guard_off;
    if (guard_on) CUDA_TRY(cudaMemsetAsync(plan->guard, 0, 32, st));
    int rc = eval_all();
    if (rc != ELISA_OK || !guard_on) return rc;
    unsigned long long h[4] = {0, 0, 0, 0};
    if ((rc = read_guard(plan, st, h))) return rc;
    if (h[0] == 0) return ELISA_OK;
    if (plan->balance && plan->renew_envelopes && (int64_t)h[0] * 4 > n) {
        // more than a quarter of the batch: the balancing envelopes were taken in another region of
        // parameter space -- derive them from this call's rows and repeat it
        CUDA_TRY(cudaMemsetAsync(plan->guard, 0, 32, st));
        elisa_b200_plan_rebalance(plan);
        ++plan->renewals;
        if ((rc = eval_all())) return rc;
        if ((rc = read_guard(plan, st, h))) return rc;
        if ((int64_t)h[0] * 4 > n) plan->renew_envelopes = false;  // fresh envelopes and still flagged: starved data
        if (h[0] == 0) return ELISA_OK;
    }
    ++plan->guard_fallbacks;
    if (h[0] <= (unsigned long long)plan->flag_cap) {
        std::vector<int32_t> rows((size_t)h[0]);
        CUDA_TRY(cudaMemcpy(rows.data(), plan->flag_list, rows.size() * 4, cudaMemcpyDeviceToHost));
        std::sort(rows.begin(), rows.end());
        rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
        plan->fallback_rows += (int64_t)rows.size();
        rc = eval_rows(rows);
    } else {  // the list overflowed: the whole call on the DMMA fold
        plan->force_dmma = true;
        rc = eval_all();
        plan->force_dmma = false;
        plan->fallback_rows += n;
        if (rc == ELISA_OK) CUDA_TRY(cudaStreamSynchronize(st));
    }
    CUDA_TRY(cudaMemsetAsync(plan->guard, 0, 32, st));
    return rc;
}

// The host entry point: theta, loglike and grad are staged on the device for the whole call
// (8 (P + S + S P) bytes per row); per-replica counts, which can be gigabytes, are streamed in
// pieces of `chunk` rows through two buffers, the H2D copy of piece k + 1 under the evaluation
// of piece k, and only for the arrays the caller gave.
static int host_call(ElisaPlan* plan, const double* theta, const double* on, const double* off, int64_t n,
                     double* loglike, double* grad) {
    const int P = plan->P, S = plan->S;
    const int64_t W = plan->sum_C;
    if (n > plan->h_n) {
        for (double** p : {&plan->h_theta, &plan->h_ll, &plan->h_grad}) {
            if (*p) cudaFree(*p);
            *p = nullptr;
        }
        plan->h_n = 0;
        CUDA_TRY(cudaMalloc((void**)&plan->h_theta, (size_t)n * P * 8));
        CUDA_TRY(cudaMalloc((void**)&plan->h_ll, (size_t)n * S * 8));
        CUDA_TRY(cudaMalloc((void**)&plan->h_grad, (size_t)n * S * P * 8));
        plan->h_n = n;
    }
    const int64_t piece = plan->chunk;
    if (on != nullptr || off != nullptr) {
        if (plan->copy_stream == nullptr)
            CUDA_TRY(cudaStreamCreateWithFlags(&plan->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            if (on != nullptr && plan->pipe_on[b] == nullptr)
                CUDA_TRY(cudaMalloc((void**)&plan->pipe_on[b], (size_t)piece * W * 8));
            if (off != nullptr && plan->pipe_off[b] == nullptr)
                CUDA_TRY(cudaMalloc((void**)&plan->pipe_off[b], (size_t)piece * W * 8));
            if (plan->pipe_free[b] == nullptr) {
                CUDA_TRY(cudaEventCreateWithFlags(&plan->pipe_free[b], cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&plan->pipe_ready[b], cudaEventDisableTiming));
            }
        }
    }
    cudaStream_t st = plan->stream;
    double* g_dev = grad ? plan->h_grad : nullptr;
    auto eval_all = [&]() -> int {
        CUDA_TRY(cudaMemcpyAsync(plan->h_theta, theta, (size_t)n * P * 8, cudaMemcpyHostToDevice, st));
        if (on == nullptr && off == nullptr) return run_all(plan, plan->h_theta, nullptr, nullptr, n, plan->h_ll, g_dev, st);
        if (n > piece) {  // the pilot of the balanced fold: rows spread over the whole batch, not over piece 0
            const int rc = ensure_balanced(plan, plan->h_theta, on, off, n, st);
            if (rc) return rc;
        }
        int rc = ELISA_OK;
        int k = 0;
        for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += piece, ++k) {
            const int64_t nc = std::min(piece, n - i0);
            const int b = k & 1;
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(plan->copy_stream, plan->pipe_free[b], 0));
            if (on) CUDA_TRY(cudaMemcpyAsync(plan->pipe_on[b], on + i0 * W, (size_t)nc * W * 8, cudaMemcpyHostToDevice, plan->copy_stream));
            if (off) CUDA_TRY(cudaMemcpyAsync(plan->pipe_off[b], off + i0 * W, (size_t)nc * W * 8, cudaMemcpyHostToDevice, plan->copy_stream));
            CUDA_TRY(cudaEventRecord(plan->pipe_ready[b], plan->copy_stream));
            CUDA_TRY(cudaStreamWaitEvent(st, plan->pipe_ready[b], 0));
            plan->row_base = i0;
            rc = run_all(plan, plan->h_theta + i0 * P, on ? plan->pipe_on[b] : nullptr, off ? plan->pipe_off[b] : nullptr, nc,
                         plan->h_ll + i0 * S, g_dev ? g_dev + i0 * S * P : nullptr, st);
            plan->row_base = 0;
            CUDA_TRY(cudaEventRecord(plan->pipe_free[b], st));
        }
        // the next evaluation of this call (envelope renewal, DMMA repeat) reuses the buffers
        if (rc == ELISA_OK) CUDA_TRY(cudaStreamSynchronize(st));
        return rc;
    };
    plan->tiny_call_n = n;  // pieces of one call take the same kernels
    int rc = run_guarded(plan, n, st, eval_all, [&](const std::vector<int32_t>& rows) {
        return rerun_rows_dmma(plan, rows, plan->h_theta, on, off, true, plan->h_ll, g_dev, st);
    });
    plan->tiny_call_n = -1;
    if (rc == ELISA_OK) {
        CUDA_TRY(cudaMemcpyAsync(loglike, plan->h_ll, (size_t)n * S * 8, cudaMemcpyDeviceToHost, st));
        if (grad) CUDA_TRY(cudaMemcpyAsync(grad, plan->h_grad, (size_t)n * S * P * 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return rc;
}

}  // namespace elisa

// ======================================================================================
// C ABI
// ======================================================================================
extern "C" {

int elisa_b200_abi_version(void) { return ELISA_B200_ABI_VERSION; }

int elisa_b200_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char ipc_handle[64]) {
    if (dev_ptr == nullptr || ipc_handle == nullptr || bytes <= 0) return fail(ELISA_ERR_INVALID, "peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    *dev_ptr = nullptr;
    CUDA_TRY(cudaMalloc(dev_ptr, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaMemset(*dev_ptr, 0, (size_t)bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *dev_ptr);
    if (e != cudaSuccess) {
        cudaFree(*dev_ptr);
        *dev_ptr = nullptr;
        return fail(ELISA_ERR_CUDA, std::string("peer_alloc: ") + cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, 64);
    return ELISA_OK;
}
int elisa_b200_peer_open(const unsigned char ipc_handle[64], void** dev_ptr) {
    if (dev_ptr == nullptr || ipc_handle == nullptr) return fail(ELISA_ERR_INVALID, "peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    *dev_ptr = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return ELISA_OK;
}
int elisa_b200_peer_close(void* dev_ptr) {
    if (dev_ptr != nullptr) CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
    return ELISA_OK;
}
int elisa_b200_peer_free(void* dev_ptr) {
    if (dev_ptr != nullptr) CUDA_TRY(cudaFree(dev_ptr));
    return ELISA_OK;
}
int elisa_b200_plan_set_gather(ElisaPlan* plan, int32_t world, int32_t rank, void* const* ll_all, void* const* grad_all,
                               void* const* flags, int64_t row_offset) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    std::lock_guard<std::mutex> lock(plan->mu);
    ElisaPlan::PeerGather& g = plan->gather;
    if (world <= 1) {
        g.world = 0;
        return ELISA_OK;
    }
    if (world > GATHER_MAX_WORLD || rank < 0 || rank >= world || ll_all == nullptr || flags == nullptr || row_offset < 0)
        return fail(ELISA_ERR_INVALID, "set_gather: bad argument");
    for (int r = 0; r < world; ++r)
        if (ll_all[r] == nullptr || flags[r] == nullptr || (grad_all != nullptr && grad_all[r] == nullptr))
            return fail(ELISA_ERR_INVALID, "set_gather: null array");
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev != plan->device) CUDA_TRY(cudaSetDevice(plan->device));
    if (g.copy == nullptr) {
        CUDA_TRY(cudaStreamCreateWithFlags(&g.copy, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&g.pushed, cudaEventDisableTiming));
        CUDA_TRY(cudaMalloc((void**)&g.done, sizeof(unsigned int)));
        CUDA_TRY(cudaMemset(g.done, 0, sizeof(unsigned int)));
        CUDA_TRY(cudaHostAlloc((void**)&g.timeout, sizeof(int32_t), cudaHostAllocMapped));
        *g.timeout = 0;
        if (const char* env = getenv("ELISA_B200_GATHER_EVERY")) g.every = std::max(1, atoi(env));
        if (const char* env = getenv("ELISA_B200_GATHER_TIMEOUT_S")) g.limit_ns = (unsigned long long)std::max(1, atoi(env)) * 1000000000ull;
    }
    g.ll.assign(world, nullptr);
    g.flags.assign(world, nullptr);
    g.grad.clear();
    if (grad_all != nullptr) g.grad.assign(world, nullptr);
    for (int r = 0; r < world; ++r) {
        g.ll[r] = (double*)ll_all[r];
        g.flags[r] = (long long*)flags[r];
        if (grad_all != nullptr) g.grad[r] = (double*)grad_all[r];
    }
    g.world = world;
    g.rank = rank;
    g.row_offset = row_offset;
    g.pend_chunks = 0;
    if (dev != plan->device) cudaSetDevice(dev);
    return ELISA_OK;
}

int64_t elisa_b200_sizeof(int32_t which) {
    switch (which) {
        case 0: return sizeof(ElisaComponentDesc);
        case 1: return sizeof(ElisaNormDesc);
        case 2: return sizeof(ElisaModelDesc);
        case 3: return sizeof(ElisaSpectrumDesc);
        case 4: return sizeof(ElisaPlanDesc);
        default: return -1;
    }
}

const char* elisa_b200_last_error(void) { return g_last_error.c_str(); }

void elisa_b200_plan_destroy(ElisaPlan* plan) {
    if (plan == nullptr) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(plan->device);
    for (void* p : plan->owned) cudaFree(p);
    for (double* p : {plan->h_theta, plan->h_ll, plan->h_grad})
        if (p) cudaFree(p);
    for (cudaStream_t st : plan->small_stream)
        if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : plan->small_done)
        if (e) cudaEventDestroy(e);
    if (plan->small_fork) cudaEventDestroy(plan->small_fork);
    for (auto& g : plan->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        if (g.graph) cudaGraphDestroy(g.graph);
    }
    for (cudaEvent_t e : plan->ev_pool) cudaEventDestroy(e);
    if (plan->guard_event) cudaEventDestroy(plan->guard_event);
    if (plan->guard_host) cudaFreeHost(plan->guard_host);
    for (int i = 0; i < 2; ++i) {
        if (plan->pipe_on[i]) cudaFree(plan->pipe_on[i]);
        if (plan->pipe_off[i]) cudaFree(plan->pipe_off[i]);
        if (plan->pipe_free[i]) cudaEventDestroy(plan->pipe_free[i]);
        if (plan->pipe_ready[i]) cudaEventDestroy(plan->pipe_ready[i]);
    }
    if (plan->copy_stream) cudaStreamDestroy(plan->copy_stream);
    if (plan->stream) cudaStreamDestroy(plan->stream);
    if (plan->gather.copy) cudaStreamDestroy(plan->gather.copy);
    if (plan->gather.ready) cudaEventDestroy(plan->gather.ready);
    if (plan->gather.pushed) cudaEventDestroy(plan->gather.pushed);
    if (plan->gather.done) cudaFree(plan->gather.done);
    if (plan->gather.timeout) cudaFreeHost(plan->gather.timeout);
    cudaSetDevice(dev);
    delete plan;
}

int elisa_b200_plan_create(const ElisaPlanDesc* desc, ElisaPlan** out) {
    if (out == nullptr) return fail(ELISA_ERR_INVALID, "out is null");
    *out = nullptr;
    if (desc == nullptr) return fail(ELISA_ERR_INVALID, "desc is null");
    if (desc->abi_version != ELISA_B200_ABI_VERSION) return fail(ELISA_ERR_INVALID, "ABI version mismatch");
    if (desc->n_params < 1 || desc->n_params > ELISA_MAX_PARAMS)
        return fail(ELISA_ERR_INVALID, "n_params out of range [1, ELISA_MAX_PARAMS]");
    if (desc->n_spectra < 1 || desc->spectra == nullptr) return fail(ELISA_ERR_INVALID, "no spectra");
    if (desc->chunk < 0) return fail(ELISA_ERR_INVALID, "chunk < 0");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(ELISA_ERR_CUDA, "no CUDA device: elisa_b200 has no CPU fallback");
    if (desc->device < 0 || desc->device >= n_dev) return fail(ELISA_ERR_INVALID, "device ordinal out of range");
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(desc->device));

    ElisaPlan* plan = new (std::nothrow) ElisaPlan();
    if (plan == nullptr) return fail(ELISA_ERR_NOMEM, "out of host memory");
    plan->device = desc->device;
    plan->P = desc->n_params;
    plan->S = desc->n_spectra;
    plan->spec.resize(plan->S);
    int rc = ELISA_OK;
    int64_t coff = 0;
    // fold engine: sliced-integer tcgen05 (default) or FP64 DMMA
    bool force_i8 = (desc->flags & ELISA_FLAG_FOLD_I8) != 0, any_i8 = false;
    {
        bool dmma = (desc->flags & ELISA_FLAG_FOLD_DMMA) != 0;
        int ks = (desc->flags >> ELISA_FLAG_FOLD_SLICES_SHIFT) & 0xf;
        if (const char* env = getenv("ELISA_B200_FOLD")) {
            if (!force_i8 && !dmma) {
                if (strcmp(env, "dmma") == 0) dmma = true;
                if (strcmp(env, "i8") == 0) force_i8 = true;
            }
        }
        if (const char* env = getenv("ELISA_B200_FOLD_SLICES"))
            if (ks == 0) ks = atoi(env);
        if (ks == 0) ks = ELISA_FOLD_SLICES_DEFAULT;
        if (dmma && force_i8) {
            delete plan;
            cudaSetDevice(prev);
            return fail(ELISA_ERR_INVALID, "ELISA_FLAG_FOLD_DMMA and ELISA_FLAG_FOLD_I8 are exclusive");
        }
        int ks_bwd = (desc->flags >> ELISA_FLAG_FOLD_SLICES_BWD_SHIFT) & 0xf;
        if (const char* env = getenv("ELISA_B200_FOLD_SLICES_BWD"))
            if (ks_bwd == 0) ks_bwd = atoi(env);
        if (ks_bwd == 0) ks_bwd = std::min(ks, ELISA_FOLD_SLICES_BWD_DEFAULT);
        if (ks < 4 || ks > 7 || ks_bwd < 4 || ks_bwd > 7) {
            delete plan;
            cudaSetDevice(prev);
            return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
        }
        plan->ks = dmma ? 0 : ks;
        plan->ks_bwd = dmma ? 0 : ks_bwd;
        plan->guard_off = force_i8;
        // one-kernel small-batch path: off when a fold engine was asked for by name, or by flag / environment
        plan->tiny_enabled = !(desc->flags & (ELISA_FLAG_NO_SMALL_BATCH_KERNEL | ELISA_FLAG_FOLD_I8 | ELISA_FLAG_FOLD_DMMA)) && !dmma && !force_i8;
        if (const char* env = getenv("ELISA_B200_TINY")) plan->tiny_enabled = plan->tiny_enabled && strcmp(env, "0") != 0;
        cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, plan->device);
        if (plan->sm_count < 1) plan->sm_count = 148;
    }
    for (int s = 0; s < plan->S && rc == ELISA_OK; ++s) {
        rc = build_spectrum(plan, desc->spectra[s], &plan->spec[s]);
        if (rc) break;
        // specialise the unfold kernel for this model unless told not to
        const char* env = getenv("ELISA_B200_JIT");
        const bool off = (desc->flags & ELISA_FLAG_NO_JIT) || (env && strcmp(env, "0") == 0);
        if (off && model_has_free_shift(plan->spec[s].model)) {
            rc = fail(ELISA_ERR_UNSUPPORTED, "free energy shifts need the runtime-specialised kernel (ELISA_FLAG_NO_JIT / ELISA_B200_JIT=0 given)");
            break;
        }
        if (!off) {
            std::string why;
            plan->spec[s].jit = jit_unfold(plan->spec[s].model, plan->P, plan->ks, &why,
                                           plan->spec[s].tiny_R != nullptr ? plan->spec[s].stat : -1);
            if (plan->spec[s].jit == nullptr && model_has_free_shift(plan->spec[s].model)) {
                rc = fail(ELISA_ERR_UNSUPPORTED, "free energy shifts need the runtime-specialised kernel, which is unavailable: " + why);
                break;
            }
            if (plan->spec[s].jit == nullptr && ((desc->flags & ELISA_FLAG_REQUIRE_JIT) || (env && strcmp(env, "require") == 0))) {
                rc = fail(ELISA_ERR_UNSUPPORTED, "runtime specialisation required but unavailable: " + why);
                break;
            }
        }
        // A narrow band-sparse response leaves the folds little to do, and the sliced-integer
        // engine still slices whole rows of U and W: below ~5 % stored the DMMA fold, which reads
        // them as they are, is the faster one (cfg 4, 2.7 % stored: 5.6e6 vs 5.1e6 evals/s).
        const bool sparse_prefers_dmma = !force_i8 && plan->spec[s].fill < 0.05;
        // A contraction shorter than one 64-wide tile (a heavily grouped spectrum, C < 64, or a
        // coarse photon grid) gives the integer engine nothing to gain -- it pads to 128-deep
        // k-blocks -- and takes away the averaging over the contraction index its normwise error
        // model leans on (a 30 x 6 pgstat case far from the optimum: gradient off by 1.3e-10 with
        // a guard estimate of 6e-12).  Such spectra keep the DMMA fold.
        const bool tiny_prefers_dmma = !force_i8 && std::min(plan->spec[s].E, plan->spec[s].C) < 64;
        if (plan->ks > 0 && !sparse_prefers_dmma && !tiny_prefers_dmma && plan->spec[s].E <= I8_MAX_K &&
            plan->spec[s].C <= I8_MAX_K) {
            if ((rc = build_spectrum_i8(plan, desc->spectra[s], &plan->spec[s], plan->ks))) break;
            any_i8 = true;
        } else if (plan->ks > 0 && force_i8) {
            rc = fail(ELISA_ERR_UNSUPPORTED, "sliced-integer fold requested but E or C exceeds its int32 headroom");
            break;
        }
        plan->spec[s].counts_off = coff;
        coff += plan->spec[s].C;
        plan->max_E_pad = std::max(plan->max_E_pad, plan->spec[s].E_pad);
        plan->max_C_pad = std::max(plan->max_C_pad, plan->spec[s].C_pad);
    }
    plan->sum_C = coff;
    if (rc == ELISA_OK && plan->tiny_enabled) {  // per-spectrum arguments of the one-kernel small-batch path
        std::vector<TinyArgs> ta(plan->S);
        bool any = false;
        for (int s = 0; s < plan->S; ++s) {
            const SpectrumPlan& sp = plan->spec[s];
            TinyArgs& t = ta[s];
            memset(&t, 0, sizeof t);
            if (sp.tiny_R == nullptr || sp.jit == nullptr || sp.jit->tiny == nullptr) continue;
            any = true;
            UnfoldCore& a = t.core;
            a.gv = sp.grid_view();
            a.gv.has_point = 0;
            a.P = plan->P;
            a.n_seg = 1;
            a.clip = 1;
            a.used_mask = sp.model.used_mask;
            for (int k = 0; k < ELISA_MAX_NORMS; ++k) {
                a.ngv[k] = sp.norm_view(k);
                a.ngv[k].has_point = 0;
                a.nw[k] = sp.norm_grid[k].w;
            }
            t.R = sp.tiny_R;
            t.ldr = sp.C;
            t.C = sp.C;
            t.ch = sp.chan_view();
            t.counts_off = sp.counts_off;
        }
        if (any) rc = upload(plan, ta, &plan->tiny_args);
    }
    if (rc == ELISA_OK) {
        // chunk: two 128-row tiles per SM by default (one: 2.7 % slower on the bench -- five launches and their tails per
        // chunk; three: + 0.5 % for 8 GB, tools/chunk_probe.py), bounded by a 6 GiB workspace
        int64_t chunk = desc->chunk > 0 ? desc->chunk : 2 * 148 * GEMM_BM;
        chunk = (chunk + GEMM_BM - 1) / GEMM_BM * GEMM_BM;
        // row partials: one per 64-column tile (DMMA) or per I8_TC-column group (sliced-integer fold)
        const int64_t pmul = any_i8 ? I8_CG : 1;
        const int64_t tiles_c = pmul * plan->max_C_pad / GEMM_BN, tiles_e = pmul * plan->max_E_pad / GEMM_BN;
        const int64_t per_row = 8 * ((int64_t)plan->max_E_pad * (1 + plan->P) + plan->max_C_pad + tiles_c +
                                     tiles_e * plan->P);
        if (desc->chunk == 0) {
            const int64_t budget = (int64_t)6 << 30;
            while (chunk > GEMM_BM && chunk * per_row > budget) chunk = (chunk / 2 + GEMM_BM - 1) / GEMM_BM * GEMM_BM;
        }
        plan->chunk = chunk;
        plan->ld_part = (int)chunk;
        auto alloc = [&](double** p, int64_t n_doubles) -> int {
            CUDA_TRY(cudaMalloc((void**)p, (size_t)n_doubles * 8));
            plan->owned.push_back(*p);
            plan->workspace_bytes += n_doubles * 8;
            return ELISA_OK;
        };
        if (rc == ELISA_OK) rc = alloc(&plan->U, chunk * plan->max_E_pad);
        if (rc == ELISA_OK) rc = alloc(&plan->dU, chunk * plan->max_E_pad * plan->P);
        if (rc == ELISA_OK) rc = alloc(&plan->W, chunk * plan->max_C_pad);
        if (rc == ELISA_OK) {
            const char* env = getenv("ELISA_B200_GRAD");
            plan->grad_contract = !(env && strcmp(env, "planes") == 0);
            if (plan->grad_contract) rc = alloc(&plan->G, chunk * plan->max_E_pad);
        }
        if (rc == ELISA_OK) rc = alloc(&plan->part, chunk * tiles_c);
        if (rc == ELISA_OK) rc = alloc(&plan->gpart, chunk * tiles_e * plan->P);
        if (rc == ELISA_OK && any_i8) {
            const int64_t kmax = round_up(std::max(plan->max_E_pad, plan->max_C_pad), I8_BK);
            const int64_t bytes = chunk * kmax * std::max(plan->ks, plan->ks_bwd);
            if (cudaMalloc((void**)&plan->A_img, (size_t)bytes) != cudaSuccess) {
                rc = fail(ELISA_ERR_NOMEM, "out of device memory (digit images)");
            } else {
                plan->owned.push_back(plan->A_img);
                plan->workspace_bytes += bytes;
                rc = alloc(&plan->a_scale, chunk);
                if (rc == ELISA_OK) {
                    double* g = nullptr;
                    rc = alloc(&g, 4);
                    plan->guard = reinterpret_cast<unsigned long long*>(g);
                    if (rc == ELISA_OK && cudaMemset(g, 0, 32) != cudaSuccess) rc = fail(ELISA_ERR_CUDA, "cudaMemset failed");
                }
                if (rc == ELISA_OK) {  // rows flagged by the guard, for the per-row DMMA fallback
                    double* fl = nullptr;
                    plan->flag_cap = 65536;
                    rc = alloc(&fl, plan->flag_cap / 2);
                    plan->flag_list = reinterpret_cast<int32_t*>(fl);
                }
                if (rc == ELISA_OK) rc = alloc(&plan->a_scale_u, chunk);
                if (rc == ELISA_OK && (cudaHostAlloc((void**)&plan->guard_host, 32, cudaHostAllocDefault) != cudaSuccess ||
                                       cudaEventCreateWithFlags(&plan->guard_event, cudaEventDisableTiming) != cudaSuccess))
                    rc = fail(ELISA_ERR_CUDA, "allocating the guard's host mirror failed");
                if (rc == ELISA_OK) memset(plan->guard_host, 0, 32);
                // balancing of the integer folds: column envelopes per spectrum + pilot scratch
                {
                    const char* env = getenv("ELISA_B200_BALANCE");
                    plan->balance = !(env && atoi(env) == 0) && !(desc->flags & ELISA_FLAG_NO_BALANCE);
                }
                if (rc == ELISA_OK && plan->balance) rc = alloc(&plan->cal_ll, (int64_t)I8_BM * plan->S);
                for (int s = 0; s < plan->S && rc == ELISA_OK && plan->balance; ++s) {
                    SpectrumPlan& sp = plan->spec[s];
                    if (!sp.use_i8()) continue;
                    // padded to whole k-blocks (filled with ones by col_envelope_kernel): the slicers
                    // read them with unguarded vector loads
                    const int64_t ek = round_up(sp.E_pad, I8_BK), ck = round_up(sp.C_pad, I8_BK);
                    if ((rc = alloc(&sp.t_scale, ek)) || (rc = alloc(&sp.t_inv, ek)) || (rc = alloc(&sp.r_scale, ck)) ||
                        (rc = alloc(&sp.r_inv, ck)))
                        break;
                    double* e8 = nullptr;  // byte arrays carved from one FP64 allocation
                    if ((rc = alloc(&e8, (ek + ck + 7) / 8))) break;
                    sp.t_exp = reinterpret_cast<uint8_t*>(e8);
                    sp.r_exp = sp.t_exp + ek;
                }
            }
        }
    }
    if (rc == ELISA_OK) {
        const char* env = getenv("ELISA_B200_GRAPHS");
        plan->graphs_enabled = !(env && atoi(env) == 0);
        auto alloc_small = [&](double** p, int64_t n_doubles) -> int {
            CUDA_TRY(cudaMalloc((void**)p, (size_t)n_doubles * 8));
            plan->owned.push_back(*p);
            return ELISA_OK;
        };
        if (plan->graphs_enabled) {
            rc = alloc_small(&plan->g_theta, GRAPH_MAX_N * plan->P);
            if (rc == ELISA_OK) rc = alloc_small(&plan->g_ll, GRAPH_MAX_N * plan->S);
            if (rc == ELISA_OK) rc = alloc_small(&plan->g_grad, GRAPH_MAX_N * plan->S * plan->P);
            const int64_t rows = std::min<int64_t>(GRAPH_MAX_N, plan->chunk);
            if (rc == ELISA_OK && plan->S > 1 && plan->S <= 32) {
                plan->small_ws.resize(plan->S);
                plan->small_stream.assign(plan->S, nullptr);
                plan->small_done.assign(plan->S, nullptr);
                for (int s = 0; s < plan->S && rc == ELISA_OK; ++s) {
                    const SpectrumPlan& sp = plan->spec[s];
                    ElisaPlan::SmallWs& w = plan->small_ws[s];
                    const int64_t pc = (sp.use_i8() ? I8_CG : 1) * sp.C_pad / GEMM_BN, pe = (sp.use_i8() ? I8_CG : 1) * sp.E_pad / GEMM_BN;
                    rc = alloc_small(&w.U, rows * sp.E_pad);
                    if (rc == ELISA_OK) rc = alloc_small(&w.dU, rows * sp.E_pad * plan->P);
                    if (rc == ELISA_OK) rc = alloc_small(&w.W, rows * sp.C_pad);
                    if (rc == ELISA_OK && plan->grad_contract) rc = alloc_small(&w.G, rows * sp.E_pad);
                    if (rc == ELISA_OK) rc = alloc_small(&w.part, rows * pc);
                    if (rc == ELISA_OK) rc = alloc_small(&w.gpart, rows * pe * plan->P);
                    if (rc == ELISA_OK && sp.use_i8()) {
                        rc = alloc_small(&w.a_scale, rows);
                        const int64_t kmax = round_up(std::max(sp.E_pad, sp.C_pad), I8_BK);
                        double* img = nullptr;
                        if (rc == ELISA_OK) rc = alloc_small(&img, (rows * kmax * std::max(plan->ks, plan->ks_bwd) + 7) / 8);
                        w.A_img = reinterpret_cast<int8_t*>(img);
                    }
                    if (rc == ELISA_OK && (cudaStreamCreateWithFlags(&plan->small_stream[s], cudaStreamNonBlocking) != cudaSuccess ||
                                           cudaEventCreateWithFlags(&plan->small_done[s], cudaEventDisableTiming) != cudaSuccess))
                        rc = fail(ELISA_ERR_CUDA, "creating the per-spectrum streams failed");
                }
                if (rc == ELISA_OK && cudaEventCreateWithFlags(&plan->small_fork, cudaEventDisableTiming) != cudaSuccess)
                    rc = fail(ELISA_ERR_CUDA, "creating the per-spectrum streams failed");
            }
        }
    }
    if (rc == ELISA_OK && cudaStreamCreateWithFlags(&plan->stream, cudaStreamNonBlocking) != cudaSuccess)
        rc = fail(ELISA_ERR_CUDA, "cudaStreamCreate failed");
    cudaSetDevice(prev);
    if (rc != ELISA_OK) {
        const std::string keep = g_last_error;
        elisa_b200_plan_destroy(plan);
        g_last_error = keep;
        return rc;
    }
    *out = plan;
    return ELISA_OK;
}

int elisa_b200_plan_set_channel_width(ElisaPlan* plan, int32_t spectrum, const double* channel_width) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S || channel_width == nullptr)
        return fail(ELISA_ERR_INVALID, "set_channel_width: bad argument");
    std::lock_guard<std::mutex> lock(plan->mu);
    SpectrumPlan& sp = plan->spec[spectrum];
    std::vector<double> inv(sp.C_pad, 1.0);
    for (int c = 0; c < sp.C; ++c) inv[c] = 1.0 / channel_width[c];
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    CUDA_TRY(cudaMemcpy(sp.chan + (size_t)8 * sp.C_pad, inv.data(), (size_t)sp.C_pad * 8, cudaMemcpyHostToDevice));
    cudaSetDevice(prev);
    return ELISA_OK;
}

int elisa_b200_loglike_dev(ElisaPlan* plan, const double* theta, const double* counts_on, const double* counts_off,
                           int64_t n, double* loglike, void* cuda_stream) {
    std::unique_lock<std::mutex> lock;
    if (plan != nullptr) lock = std::unique_lock<std::mutex>(plan->mu);
    if (plan != nullptr && gather_on(plan)) gather_begin(plan);
    int rc = run_small(plan, theta, counts_on, counts_off, n, loglike, nullptr, (cudaStream_t)cuda_stream);
    if (rc == ELISA_OK && plan != nullptr && gather_on(plan)) rc = gather_finish(plan, (cudaStream_t)cuda_stream);
    return rc;
}

int elisa_b200_loglike_grad_dev(ElisaPlan* plan, const double* theta, const double* counts_on,
                                const double* counts_off, int64_t n, double* loglike, double* grad,
                                void* cuda_stream) {
    if (grad == nullptr) return fail(ELISA_ERR_INVALID, "grad is null");
    std::unique_lock<std::mutex> lock;
    if (plan != nullptr) lock = std::unique_lock<std::mutex>(plan->mu);
    if (plan != nullptr && gather_on(plan)) gather_begin(plan);
    int rc = run_small(plan, theta, counts_on, counts_off, n, loglike, grad, (cudaStream_t)cuda_stream);
    if (rc == ELISA_OK && plan != nullptr && gather_on(plan)) rc = gather_finish(plan, (cudaStream_t)cuda_stream);
    return rc;
}

int elisa_b200_unfold_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, int64_t n, double* unfold,
                          void* cuda_stream) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return fail(ELISA_ERR_INVALID, "unfold: bad spectrum");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr || unfold == nullptr) return fail(ELISA_ERR_INVALID, "unfold: null pointer");
    const SpectrumPlan& sp = plan->spec[spectrum];
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    int rc = ELISA_OK;
    const int64_t step = 1 << 20;
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += step) {
        const int nc = (int)std::min<int64_t>(step, n - i0);
        rc = launch_unfold(plan, sp, theta + i0 * plan->P, nc, nc, false, unfold + i0 * sp.E, nullptr, sp.E, sp.E,
                           false, (cudaStream_t)cuda_stream);
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_sites_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, const double* counts_on,
                         const double* counts_off, int64_t n, double* rate, double* model_on, double* ll_on,
                         double* model_off, double* ll_off, void* cuda_stream) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return fail(ELISA_ERR_INVALID, "sites: bad spectrum");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr) return fail(ELISA_ERR_INVALID, "sites: theta is null");
    const SpectrumPlan& sp = plan->spec[spectrum];
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    int rc = ELISA_OK;
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += plan->chunk) {
        const int nc = (int)std::min<int64_t>(plan->chunk, n - i0);
        const int n_rows = round_up(nc, GEMM_BM);
        rc = launch_unfold(plan, sp, theta + i0 * plan->P, nc, n_rows, false, plan->U, nullptr, sp.E_pad, sp.E_pad,
                           true, st);
        if (rc) break;
        EpilogueStore es{plan->W, sp.C_pad};
        rc = launch_gemm(plan, ELISA_PROFILE_FOLD, plan->U, sp.E_pad, sp.fwd, n_rows / GEMM_BM, es, st);
        if (rc) break;
        SitesArgs a;
        a.X = plan->W;
        a.ldx = sp.C_pad;
        a.ch = sp.chan_view();
        a.area = sp.chan + (size_t)7 * sp.C_pad;
        a.inv_width = sp.chan + (size_t)8 * sp.C_pad;
        a.counts_on = counts_on ? counts_on + i0 * plan->sum_C + sp.counts_off : nullptr;
        a.counts_off = counts_off ? counts_off + i0 * plan->sum_C + sp.counts_off : nullptr;
        a.ld_counts = plan->sum_C;
        a.n = nc;
        a.C = sp.C;
        a.stat = sp.stat;
        const size_t o = (size_t)i0 * sp.C;
        a.rate = rate ? rate + o : nullptr;
        a.model_on = model_on ? model_on + o : nullptr;
        a.ll_on = ll_on ? ll_on + o : nullptr;
        a.model_off = model_off ? model_off + o : nullptr;
        a.ll_off = ll_off ? ll_off + o : nullptr;
        dim3 grid((sp.C + 127) / 128, nc);
        ScopedSpan span(plan, ELISA_PROFILE_REDUCE, st);
        sites_kernel<<<grid, 128, 0, st>>>(a);
        ++plan->launches;
        if (cudaGetLastError() != cudaSuccess) rc = fail(ELISA_ERR_CUDA, "sites_kernel launch failed");
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_loglike_grad_host(ElisaPlan* plan, const double* theta, const double* counts_on,
                                 const double* counts_off, int64_t n, double* loglike, double* grad) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr || loglike == nullptr) return fail(ELISA_ERR_INVALID, "theta / loglike is null");
    if (n > 0x7fffffff) return fail(ELISA_ERR_INVALID, "n exceeds 2^31 - 1 rows per call");
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    const int rc = host_call(plan, theta, counts_on, counts_off, n, loglike, grad);
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_loglike_grad_guarded_dev(ElisaPlan* plan, const double* theta, const double* counts_on,
                                        const double* counts_off, int64_t n, double* loglike, double* grad,
                                        void* cuda_stream, int64_t* fallback_rows) {
    if (fallback_rows) *fallback_rows = 0;
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr || loglike == nullptr) return fail(ELISA_ERR_INVALID, "theta / loglike is null");
    if (n > 0x7fffffff) return fail(ELISA_ERR_INVALID, "n exceeds 2^31 - 1 rows per call");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    const int64_t before = plan->fallback_rows;
    if (gather_on(plan)) gather_begin(plan);
    int rc = run_guarded(
        plan, n, st, [&]() { return run_small(plan, theta, counts_on, counts_off, n, loglike, grad, st); },
        [&](const std::vector<int32_t>& rows) {
            return rerun_rows_dmma(plan, rows, theta, counts_on, counts_off, false, loglike, grad, st);
        });
    if (fallback_rows) *fallback_rows = plan->fallback_rows - before;
    if (rc == ELISA_OK && gather_on(plan)) {
        // rows the DMMA fold re-evaluated were pushed with their first values: push the block again
        if (plan->fallback_rows != before) rc = gather_push(plan, 0, n, loglike, grad, st);
        if (rc == ELISA_OK) rc = gather_finish(plan, st);
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_plan_guard_snapshot(ElisaPlan* plan, void* cuda_stream) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    if (plan->guard == nullptr) return ELISA_OK;  // every spectrum folds with DMMA: nothing to guard
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CUDA_TRY(cudaMemcpyAsync(plan->guard_host, plan->guard, 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(plan->guard_event, st));
    plan->guard_snapshot_pending = true;
    return ELISA_OK;
}

int elisa_b200_plan_guard_poll(ElisaPlan* plan, int wait, int64_t* flagged, double* worst) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    if (flagged) *flagged = 0;
    if (worst) *worst = 0.0;
    if (plan->guard == nullptr) return 1;
    if (plan->guard_snapshot_pending) {
        if (wait) {
            CUDA_TRY(cudaEventSynchronize(plan->guard_event));
        } else {
            const cudaError_t e = cudaEventQuery(plan->guard_event);
            if (e == cudaErrorNotReady) return 0;
            if (e != cudaSuccess) return fail(ELISA_ERR_CUDA, std::string("guard event: ") + cudaGetErrorString(e));
        }
        plan->guard_snapshot_pending = false;
    }
    if (flagged) *flagged = (int64_t)plan->guard_host[0];
    if (worst) {
        double w, wg;
        memcpy(&w, &plan->guard_host[1], 8);
        memcpy(&wg, &plan->guard_host[2], 8);
        *worst = std::max(w, wg);
    }
    return 1;
}

int elisa_b200_plan_set_debug(ElisaPlan* plan, int32_t skip_mask) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    plan->debug_skip = skip_mask;
    if (skip_mask) plan->graphs_enabled = false;  // captured graphs hold the full kernel sequence
    return ELISA_OK;
}

int elisa_b200_plan_set_profiling(ElisaPlan* plan, int enable) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    drain_spans(plan);
    plan->profiling = enable != 0;
    for (int i = 0; i < ELISA_PROFILE_CLASSES; ++i) {
        plan->prof_ms[i] = 0.0;
        plan->prof_n[i] = 0;
    }
    return ELISA_OK;
}

int elisa_b200_plan_get_profile(ElisaPlan* plan, double* ms, int64_t* launches) {
    if (plan == nullptr || ms == nullptr || launches == nullptr) return fail(ELISA_ERR_INVALID, "get_profile: null");
    drain_spans(plan);
    for (int i = 0; i < ELISA_PROFILE_CLASSES; ++i) {
        ms[i] = plan->prof_ms[i];
        launches[i] = plan->prof_n[i];
    }
    return ELISA_OK;
}

int64_t elisa_b200_plan_workspace_bytes(const ElisaPlan* plan) { return plan ? plan->workspace_bytes : 0; }
int64_t elisa_b200_plan_launch_count(const ElisaPlan* plan) { return plan ? plan->launches : 0; }
int64_t elisa_b200_plan_chunk(const ElisaPlan* plan) { return plan ? plan->chunk : 0; }
int64_t elisa_b200_specialise_check(const ElisaModelDesc* model, int32_t n_params, char* source_out,
                                    int64_t source_cap) {
    if (model == nullptr || n_params < 1 || n_params > ELISA_MAX_PARAMS) return fail(ELISA_ERR_INVALID, "specialise_check: bad argument");
    ModelDev m;
    int rc = build_model(*model, n_params, &m);
    if (rc) return rc;
    // ELISA_B200_SPECIALISE_TINY=<statistic>: include the one-kernel small-batch evaluation (offline SASS, tests)
    const char* tiny = getenv("ELISA_B200_SPECIALISE_TINY");
    const std::string src = generate_unfold_source(m, n_params, ELISA_FOLD_SLICES_DEFAULT, tiny ? atoi(tiny) : -1);
    if (source_out != nullptr && source_cap > 0) {
        const size_t n = std::min<size_t>(src.size(), (size_t)source_cap - 1);
        memcpy(source_out, src.data(), n);
        source_out[n] = '\0';
    }
    std::vector<char> cubin;
    std::string why;
    if (!compile_unfold_source(src, &cubin, &why)) return fail(ELISA_ERR_UNSUPPORTED, why);
    return (int64_t)cubin.size();
}

int elisa_b200_plan_fold_slices(const ElisaPlan* plan, int32_t spectrum) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return -1;
    return plan->spec[spectrum].use_i8() ? plan->ks : 0;
}
int elisa_b200_plan_fold_products(ElisaPlan* plan, int32_t spectrum, double out[2]) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S || out == nullptr) return fail(ELISA_ERR_INVALID, "fold_products: bad argument");
    std::lock_guard<std::mutex> lock(plan->mu);
    const SpectrumPlan& sp = plan->spec[spectrum];
    out[0] = out[1] = 0.0;
    if (!sp.use_i8()) return ELISA_OK;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    CUDA_TRY(cudaDeviceSynchronize());
    static const bool skip = getenv("ELISA_B200_I8_SKIP") && atoi(getenv("ELISA_B200_I8_SKIP")) != 0;
    for (int k = 0; k < 2; ++k) {
        const I8Operand& op = k == 0 ? sp.i8_fwd : sp.i8_bwd;
        std::vector<uint8_t> lead((size_t)op.n_images);
        if (op.n_images) CUDA_TRY(cudaMemcpy(lead.data(), op.lead, lead.size(), cudaMemcpyDeviceToHost));
        const int kb_total = round_up(op.K, I8_BK) / I8_BK;
        double prod = 0.0;
        int64_t img = 0;
        for (int t = 0; t < op.n_tiles; ++t)
            for (int kb = op.h_kb_lo[t]; kb < op.h_kb_hi[t]; ++kb, ++img) {
                const int z = (skip && kb != op.h_kb_lo[t]) ? std::min<int>(lead[img], op.ks) : 0;
                prod += 0.5 * (op.ks - z) * (op.ks - z + 1);
            }
        out[k] = op.n_tiles && kb_total ? prod / ((double)op.n_tiles * kb_total) : 0.0;
    }
    cudaSetDevice(prev);
    return ELISA_OK;
}

int elisa_b200_plan_small_batch_kernel(const ElisaPlan* plan, int32_t spectrum) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return -1;
    const SpectrumPlan& sp = plan->spec[spectrum];
    return plan->tiny_enabled && sp.tiny_R != nullptr && sp.jit != nullptr && sp.jit->tiny != nullptr ? 1 : 0;
}
int elisa_b200_plan_fold_slices_bwd(const ElisaPlan* plan, int32_t spectrum) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return -1;
    return plan->spec[spectrum].use_i8() ? plan->ks_bwd : 0;
}

int elisa_b200_normal_eq_dev(ElisaPlan* plan, const double* theta, const double* counts_on, const double* counts_off,
                             int64_t n, double* loglike, double* grad, double* jtj, void* cuda_stream) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr || loglike == nullptr || grad == nullptr || jtj == nullptr)
        return fail(ELISA_ERR_INVALID, "normal_eq: null pointer");
    if (plan->P > 16) return fail(ELISA_ERR_UNSUPPORTED, "normal_eq: more than 16 free parameters");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    int rc = ensure_tangent_workspace(plan);  // first use: workspace of the folded model and tangent spectra
    if (rc == ELISA_OK) rc = ensure_balanced(plan, theta, counts_on, counts_off, n, st);
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += plan->chunk) {
        const int nc = (int)std::min<int64_t>(plan->chunk, n - i0);
        for (int s = 0; s < plan->S && rc == ELISA_OK; ++s)
            rc = run_normal_eq_chunk(plan, s, theta + i0 * plan->P, counts_on ? counts_on + i0 * plan->sum_C : nullptr,
                                     counts_off ? counts_off + i0 * plan->sum_C : nullptr, nc, loglike + i0,
                                     grad + i0 * plan->P, jtj + i0 * plan->P * plan->P, st);
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_jacobian_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, const double* counts_on,
                            const double* counts_off, int64_t n, double* x, double* ds, double* dl_dx,
                            void* cuda_stream) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return fail(ELISA_ERR_INVALID, "jacobian: bad spectrum");
    std::lock_guard<std::mutex> lock(plan->mu);
    if (n < 0) return fail(ELISA_ERR_INVALID, "n < 0");
    if (n == 0) return ELISA_OK;
    if (theta == nullptr) return fail(ELISA_ERR_INVALID, "jacobian: theta is null");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const SpectrumPlan& sp = plan->spec[spectrum];
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    int rc = ensure_tangent_workspace(plan);
    if (rc == ELISA_OK) rc = ensure_balanced(plan, theta, counts_on, counts_off, n, st);
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += plan->chunk) {
        const int nc = (int)std::min<int64_t>(plan->chunk, n - i0);
        if ((rc = fold_tangent_planes(plan, spectrum, theta + i0 * plan->P, nc, st))) break;
        JacobianArgs a;
        a.X = plan->Xp;
        a.plane = (int64_t)plan->chunk * plan->max_C_pad;
        a.ldx = sp.C_pad;
        a.ch = sp.chan_view();
        a.counts_on = counts_on ? counts_on + i0 * plan->sum_C + sp.counts_off : nullptr;
        a.counts_off = counts_off ? counts_off + i0 * plan->sum_C + sp.counts_off : nullptr;
        a.ld_counts = plan->sum_C;
        a.n = nc;
        a.C = sp.C;
        a.P = plan->P;
        a.stat = sp.stat;
        a.used_mask = sp.model.used_mask;
        a.x = x ? x + (size_t)i0 * sp.C : nullptr;
        a.ds = ds ? ds + (size_t)i0 * plan->P * sp.C : nullptr;
        a.dl_dx = dl_dx ? dl_dx + (size_t)i0 * sp.C : nullptr;
        dim3 grid((sp.C + 127) / 128, nc);
        ScopedSpan span(plan, ELISA_PROFILE_REDUCE, st);
        jacobian_kernel<<<grid, 128, 0, st>>>(a);
        ++plan->launches;
        if (cudaGetLastError() != cudaSuccess) rc = fail(ELISA_ERR_CUDA, "jacobian_kernel launch failed");
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_csr_fold_dev(ElisaPlan* plan, int32_t spectrum, const double* unfold, int64_t ldu, int64_t n, double* folded,
                            int64_t ldx, void* cuda_stream) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return fail(ELISA_ERR_INVALID, "csr_fold: bad spectrum");
    std::lock_guard<std::mutex> lock(plan->mu);
    const SpectrumPlan& sp = plan->spec[spectrum];
    if (sp.csr_ptr == nullptr) return fail(ELISA_ERR_UNSUPPORTED, "csr_fold: the response of this spectrum is dense (more than 25 % stored)");
    if (n < 0 || unfold == nullptr || folded == nullptr || ldu < sp.E || ldx < sp.C) return fail(ELISA_ERR_INVALID, "csr_fold: bad argument");
    if (n == 0) return ELISA_OK;
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = ELISA_OK;
    for (int64_t i0 = 0; i0 < n && rc == ELISA_OK; i0 += 65535 * CSR_RT) {  // grid.y limit
        const int nc = (int)std::min<int64_t>(65535 * CSR_RT, n - i0);
        dim3 grid((sp.C + 127) / 128, (nc + CSR_RT - 1) / CSR_RT);
        ScopedSpan span(plan, ELISA_PROFILE_FOLD, st);
        csr_fold_kernel<<<grid, 128, 0, st>>>(sp.csr_ptr, sp.csr_idx, sp.csr_val, unfold + i0 * ldu, ldu, nc, sp.C,
                                              folded + i0 * ldx, ldx);
        ++plan->launches;
        if (cudaGetLastError() != cudaSuccess) rc = fail(ELISA_ERR_CUDA, "csr_fold_kernel launch failed");
    }
    cudaSetDevice(prev);
    return rc;
}

int elisa_b200_sample_dev(ElisaPlan* plan, int32_t spectrum, int32_t which, const double* mean, int64_t mean_rows, int64_t n,
                          uint64_t seed, double* out, int64_t ld, void* cuda_stream) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return fail(ELISA_ERR_INVALID, "sample: bad spectrum");
    if (which < 0 || which > 1 || mean == nullptr || out == nullptr || n < 0 || (mean_rows != 1 && mean_rows != n))
        return fail(ELISA_ERR_INVALID, "sample: bad argument");
    if (n == 0) return ELISA_OK;
    const SpectrumPlan& sp = plan->spec[spectrum];
    if (ld < sp.C) return fail(ELISA_ERR_INVALID, "sample: ld < n_channels");
    // sampling distributions of the reference (infer/likelihood.py:209-225, 350-385; helper.py:221-278):
    // 'on' data Normal(model, net_errors) for chi2 and Poisson otherwise; 'off' data Normal(model,
    // back_errors) for pgstat and Poisson for wstat; other statistics have no sampled background
    const double* sigma = nullptr;
    if (which == 0 && sp.stat == ELISA_STAT_CHI2) sigma = sp.chan + (size_t)3 * sp.C_pad;
    if (which == 1) {
        if (sp.stat == ELISA_STAT_PGSTAT) sigma = sp.chan + (size_t)3 * sp.C_pad;
        else if (sp.stat != ELISA_STAT_WSTAT) return fail(ELISA_ERR_INVALID, "sample: this statistic has no sampled background");
    }
    int prev = 0;
    CUDA_TRY(cudaGetDevice(&prev));
    CUDA_TRY(cudaSetDevice(plan->device));
    SampleArgs a;
    a.mean = mean;
    a.mean_rows = mean_rows;
    a.sigma = sigma;
    a.n = n;
    a.C = sp.C;
    a.out = out;
    a.ld = ld;
    a.seed = seed;
    // disjoint streams per (spectrum, on / off array): 2^40 elements each
    a.stream0 = ((uint64_t)(2 * spectrum + which)) << 40;
    const int64_t total = n * sp.C;
    sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a);
    ++plan->launches;
    const cudaError_t e = cudaGetLastError();
    cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(ELISA_ERR_CUDA, std::string("sample_kernel: ") + cudaGetErrorString(e));
    return ELISA_OK;
}

int elisa_b200_plan_rebalance(ElisaPlan* plan) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    for (int s = 0; s < plan->S; ++s) plan->spec[s].stale = true;
    return ELISA_OK;
}

int elisa_b200_plan_set_fold(ElisaPlan* plan, int dmma) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    plan->force_dmma = dmma != 0;
    return ELISA_OK;
}

int64_t elisa_b200_plan_guard_fallbacks(const ElisaPlan* plan) { return plan ? plan->guard_fallbacks : 0; }
int elisa_b200_plan_guard_stats(const ElisaPlan* plan, int64_t out[3]) {
    if (plan == nullptr || out == nullptr) return fail(ELISA_ERR_INVALID, "guard_stats: null");
    out[0] = plan->guard_fallbacks;
    out[1] = plan->fallback_rows;
    out[2] = plan->renewals;
    return ELISA_OK;
}

int elisa_b200_plan_fold_guard(ElisaPlan* plan, int64_t* flagged, double* worst, int reset) {
    if (plan == nullptr) return fail(ELISA_ERR_INVALID, "null plan");
    std::lock_guard<std::mutex> lock(plan->mu);
    unsigned long long h[4] = {0, 0, 0, 0};
    if (plan->guard != nullptr) {
        int prev = 0;
        CUDA_TRY(cudaGetDevice(&prev));
        CUDA_TRY(cudaSetDevice(plan->device));
        CUDA_TRY(cudaMemcpy(h, plan->guard, 32, cudaMemcpyDeviceToHost));
        if (reset) CUDA_TRY(cudaMemset(plan->guard, 0, 32));
        cudaSetDevice(prev);
    }
    if (flagged) *flagged = (int64_t)h[0];
    if (worst) {  // the larger of the value-side and the gradient-side estimate
        double w, wg;
        memcpy(&w, &h[1], 8);
        memcpy(&wg, &h[2], 8);
        *worst = std::max(w, wg);
    }
    return ELISA_OK;
}

int elisa_b200_sliced_gemm_dev(const double* A, int64_t lda, const double* B, int64_t ldb, int32_t m, int32_t n,
                               int32_t k, int32_t slices, double* D, int64_t ldd, double* ms, void* cuda_stream) {
    if (A == nullptr || B == nullptr || D == nullptr || m < 1 || n < 1 || k < 1 || lda < k || ldb < k || ldd < n)
        return fail(ELISA_ERR_INVALID, "sliced_gemm: bad argument");
    if (slices < 4 || slices > 7) return fail(ELISA_ERR_INVALID, "fold slices must be 4..7");
    if (k > I8_MAX_K) return fail(ELISA_ERR_UNSUPPORTED, "sliced_gemm: k exceeds the int32 headroom (16384)");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int m_tiles = round_up(m, I8_BM) / I8_BM, n_tiles = round_up(n, I8_BN) / I8_BN;
    const int kb = round_up(k, I8_BK) / I8_BK;
    int8_t *a_img = nullptr, *b_img = nullptr;
    double *a_sc = nullptr, *b_sc = nullptr;
    int32_t* rng = nullptr;
    int64_t* timg = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = ELISA_OK;
    auto cleanup = [&]() {
        cudaFree(a_img); cudaFree(b_img); cudaFree(a_sc); cudaFree(b_sc); cudaFree(rng); cudaFree(timg);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    };
#define SG_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            cleanup();                                                                                 \
            return fail(ELISA_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
        }                                                                                              \
    } while (0)
    SG_TRY(cudaMalloc((void**)&a_img, (size_t)m_tiles * kb * slices * I8_A_SLICE));
    SG_TRY(cudaMalloc((void**)&b_img, (size_t)n_tiles * kb * slices * I8_B_SLICE));
    SG_TRY(cudaMalloc((void**)&a_sc, (size_t)m_tiles * I8_BM * 8));
    SG_TRY(cudaMalloc((void**)&b_sc, (size_t)n_tiles * I8_BN * 8));
    std::vector<int32_t> h_rng(2 * n_tiles);
    std::vector<int64_t> h_timg(n_tiles);
    for (int t = 0; t < n_tiles; ++t) {
        h_rng[t] = 0;
        h_rng[n_tiles + t] = kb;
        h_timg[t] = (int64_t)t * kb;
    }
    SG_TRY(cudaMalloc((void**)&rng, h_rng.size() * 4));
    SG_TRY(cudaMalloc((void**)&timg, h_timg.size() * 8));
    SG_TRY(cudaMemcpyAsync(rng, h_rng.data(), h_rng.size() * 4, cudaMemcpyHostToDevice, st));
    SG_TRY(cudaMemcpyAsync(timg, h_timg.data(), h_timg.size() * 8, cudaMemcpyHostToDevice, st));
    SG_TRY(cudaStreamSynchronize(st));
    SliceArgs sa;
    sa.X = A; sa.ld = lda; sa.rows = m; sa.cols = k; sa.row_tiles = m_tiles; sa.tile_rows = I8_BM; sa.kb_total = kb;
    sa.tile_img = nullptr; sa.kb_lo = nullptr; sa.kb_hi = nullptr; sa.img = a_img; sa.scale = a_sc;
    sa.kscale = nullptr; sa.abs_sum = nullptr;
    SliceArgs sb = sa;
    sb.X = B; sb.ld = ldb; sb.rows = n; sb.row_tiles = n_tiles; sb.tile_rows = I8_BN; sb.img = b_img; sb.scale = b_sc;
    if ((rc = launch_slice(nullptr, slices, sa, 0, st)) || (rc = launch_slice(nullptr, slices, sb, 0, st))) {
        cleanup();
        return rc;
    }
    FoldI8Args fa;
    fa.A_img = a_img; fa.a_scale = a_sc; fa.kb_a = kb; fa.B_img = b_img; fa.b_scale = b_sc; fa.b_tile_img = timg;
    fa.kb_lo = rng; fa.kb_hi = rng + n_tiles; fa.m_tiles = m_tiles; fa.n_tiles = n_tiles;
    fa.b_lead = nullptr;
    fa.n_groups = 1;  // set at launch
    fa.trap_info = trap_info_dev();
    fa.debug = getenv("ELISA_B200_I8_DEBUG") ? atoi(getenv("ELISA_B200_I8_DEBUG")) : 0;
    I8EpilogueStore es{D, ldd, m, n, a_sc};
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    SG_TRY(cudaEventCreate(&e0));
    SG_TRY(cudaEventCreate(&e1));
    SG_TRY(cudaEventRecord(e0, st));
    rc = launch_fold_i8(nullptr, slices, 0, fa, es, sms, st);
    if (rc) {
        cleanup();
        return rc;
    }
    SG_TRY(cudaEventRecord(e1, st));
    SG_TRY(cudaStreamSynchronize(st));
    float t = 0.f;
    SG_TRY(cudaEventElapsedTime(&t, e0, e1));
    if (ms) *ms = t;
#undef SG_TRY
    cleanup();
    return ELISA_OK;
}

int elisa_b200_plan_is_specialised(const ElisaPlan* plan, int32_t spectrum) {
    if (plan == nullptr || spectrum < 0 || spectrum >= plan->S) return 0;
    return plan->spec[spectrum].jit != nullptr ? 1 : 0;
}

}  // extern "C"
