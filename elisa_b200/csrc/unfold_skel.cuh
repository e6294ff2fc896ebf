// Skeleton of the runtime-specialised unfold kernel.
//
// At plan creation the library writes, for each distinct model expression, a tiny
// translation unit that defines two policy structs, `ModelV` (values) and `ModelG` (values +
// d/dtheta): parameters wired to theta columns or literals, one leaf_bin_pc<KIND> call per
// component -- each in the smallest dual type that holds the leaf's own free parameters --
// and the +/* tree as straight-line scalar code; and instantiates unfold_generic<ModelV, double>
// and unfold_generic<ModelG, Dual<P>>.  It is compiled for sm_100a with NVRTC; every component body is the
// same hand-written code as in the ahead-of-time interpreter (components.cuh), but with
// all dispatch, parameter routing and derivative sparsity resolved at compile time
// (dual seeds are literals, so zero tangents vanish by constant folding).
//
// Thread layout: one warp per (parameter vector, energy segment); lane = energy edge;
// passes of 31 bins (see components.cuh).
#pragma once
#include "components.cuh"
#include "slice_i8.cuh"
#include "stats.cuh"

namespace elisa {

struct UnfoldCore {
    GridView gv;
    const double* theta;  // [N, P]
    int32_t P;
    int32_t n_valid;  // rows < n_valid are evaluated, rows in [n_valid, n_rows) are zero-filled
    int32_t n_rows;
    int32_t n_seg;    // segments the energy axis is split into (one warp per (row, segment))
    int32_t seg_len;  // bins per segment (multiple of 31)
    int32_t ld;       // leading dimension of U / dU rows
    int32_t e_pad;    // columns [E, e_pad) are zero-filled
    int32_t clip;     // apply clip(unfold, 1e-300, 1e300) (infer/likelihood.py:204)
    double* U;        // [n_rows, ld]
    double* dU;       // [P, n_rows, ld] or nullptr
    int64_t plane;    // n_rows * ld
    uint32_t used_mask;
    int32_t kb_total;  // k-blocks per row tile of the digit images
    // digits of U for the sliced-integer fold, written by the kernel itself when the whole row
    // belongs to one warp (n_seg == 1) and the kernel was specialised with KS > 0; else nullptr
    int8_t* img;
    double* scale;
    // balancing of the forward fold: when the kernel slices U itself it stores and slices
    // U[e] * 2^kexp[e] (the FP64 copy is only read back for its digits then); nullptr: no scaling
    const uint8_t* kexp;
    // flux normalisations (PhFlux / EnFlux, models/conv.py:145-256): their own grids and, for
    // EnFlux, the bin weights keV_to_erg * sqrt(e_j e_j+1) (nullptr: 1)
    GridView ngv[ELISA_MAX_NORMS];
    const double* nw[ELISA_MAX_NORMS];
    // contraction mode (unfold_contract): instead of storing U and dU/dtheta the kernel reads
    // G[row, e] = d loglike / d u_e (the transposed fold's plain output) and accumulates
    // grad[row, j] = sum_e G[row, e] * du_e/dtheta_j in registers -- the tangents never exist in
    // memory.  grad_out points at element (row 0, spectrum s, parameter 0); rows are grad_stride apart.
    const double* G;
    int64_t ldg;
    double* grad_out;
    int64_t grad_stride;
};

constexpr int UNFOLD_WARPS = 4;

__device__ __forceinline__ void unfold_zero(const UnfoldCore& a, size_t row_off, int lo, int hi, int lane, bool grad) {
    for (int e = lo + lane; e < hi; e += 32) {
        a.U[row_off + e] = 0.0;
        if (grad && a.dU != nullptr)
            for (int j = 0; j < a.P; ++j)
                if ((a.used_mask >> j) & 1u) a.dU[j * a.plane + row_off + e] = 0.0;
    }
}

// UNFOLD_USED_MASK (defined by the generated translation unit): the theta columns the model
// depends on, known when the kernel is specialised -- the per-plane tests fold away.
template <typename T>
__device__ __forceinline__ double unfold_store(const UnfoldCore& a, size_t off, const T& res, double k = 1.0) {
    const double v = value(res);
    bool pass = true;
    double u = v;
    if (a.clip && !(v >= 1e-300 && v <= 1e300)) {  // clip(unfold, 1e-300, 1e300), likelihood.py:204: rare, one branch
        pass = false;
        u = v < 1e-300 ? 1e-300 : (v > 1e300 ? 1e300 : v);  // jnp.clip propagates NaN
    }
    u *= k;  // a power of two: exact
    a.U[off] = u;
    if constexpr (Scalar<T>::NP > 0) {
#ifdef UNFOLD_USED_MASK
        constexpr uint32_t used = UNFOLD_USED_MASK;
        constexpr int np = Scalar<T>::NP;
#else
        const uint32_t used = a.used_mask;
        const int np = a.P;
#endif
        if (a.dU != nullptr) {  // nullptr: timing experiment (elisa_b200_plan_set_debug bit 32), the tangents are dropped
            double* q = a.dU + off;
#pragma unroll
            for (int j = 0; j < Scalar<T>::NP; ++j) {
                if (j < np && ((used >> j) & 1u)) *q = pass ? res.d[j] : 0.0;
                q += a.plane;
            }
        }
    }
    return u;
}

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
template <int N>
__device__ __forceinline__ Dual<N> warp_sum(Dual<N> x) {
    x.v = warp_sum(x.v);
#pragma unroll
    for (int j = 0; j < N; ++j) x.d[j] = warp_sum(x.d[j]);
    return x;
}

template <class Model, typename T, int KS = 0>
__device__ __forceinline__ void unfold_generic(const UnfoldCore& a) {
    constexpr bool GRAD = Scalar<T>::NP > 0;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * UNFOLD_WARPS + wib;
    const int row = (int)(wid / a.n_seg), seg = (int)(wid % a.n_seg);
    if (row >= a.n_rows) return;
    const int E = a.gv.E;
    const int e_lo = seg * a.seg_len;
    const int e_hi = min(e_lo + a.seg_len, E);
    const bool last_seg = seg == a.n_seg - 1;
    const size_t row_off = (size_t)row * a.ld;
    const bool slice = KS > 0 && a.img != nullptr;  // host guarantees n_seg == 1 then
    if (row >= a.n_valid) {  // padding rows: zeros, so that the fold sees finite operands
        unfold_zero(a, row_off, e_lo, last_seg ? a.e_pad : e_hi, lane, GRAD);
        if constexpr (KS > 0) {
            if (slice) {
                if (lane == 0) a.scale[row] = 0.0;
                slice_row_digits<KS>(a.U + row_off, false, false, E, 0.0, 0.0, row % I8_BM, I8_BM, 0, a.kb_total,
                                     a.img, (int64_t)(row / I8_BM) * a.kb_total, lane);
            }
        }
        return;
    }
    typename Model::Pre pre;
    Model::pre(a.theta + (size_t)row * a.P, pre);
    // flux normalisations: every warp of the row sums the normalised sub-expression over the
    // normalisation's own grid (nothing to do for models without PhFlux / EnFlux)
    Model::norms(pre, a, a.theta + (size_t)row * a.P, lane);
    // the grid values of a pass are loaded during the previous one
    GridView gv = a.gv;
    gv.has_point = 1;
    auto load_point = [&](int e, double& pE, double& pl, double& pEm, double& plm) {
        const int ec = e <= E ? e : E, em = e < E ? e : E - 1;
        pE = a.gv.egrid[ec];
        pl = a.gv.lngrid[ec];
        pEm = a.gv.emid[em];
        plm = a.gv.lnmid[em];
    };
    double nE, nl, nEm, nlm;
    load_point(e_lo + lane, nE, nl, nEm, nlm);
    // balancing factor of the lane's bin, loaded one pass ahead like the grid values
    const bool scaled = slice && a.kexp != nullptr;
    int nk = scaled ? a.kexp[min(e_lo + lane, E - 1)] : 0;
    // largest |u| of the row for the digit scale, tracked on the HIGH WORD of the doubles (an integer
    // max instead of a NaN-aware FP64 max per bin; a non-finite value shows as a high word >=
    // 0x7ff00000).  The scale is then taken from the upper end of that high word's interval: at
    // most 2^-20 above the true maximum, far inside the headroom of the top digit (120 of 127).
    int amax_hi = 0;
    for (int e0 = e_lo; e0 < e_hi; e0 += 31) {
        const int e = e0 + lane;
        gv.pE = nE;
        gv.pl = nl;
        gv.pEm = nEm;
        gv.plm = nlm;
        const double ck = __hiloint2double((1023 + nk) << 20, 0);  // 2^nk
        if (e0 + 31 < e_hi) {
            load_point(e + 31, nE, nl, nEm, nlm);
            if (scaled) nk = a.kexp[min(e + 31, E - 1)];
        }
        const T res = Model::bin(pre, gv, e);
        if (lane < 31 && e < e_hi) {
            const double u = unfold_store<T>(a, row_off + e, res, ck);
            if constexpr (KS > 0) amax_hi = max(amax_hi, __double2hiint(u) & 0x7fffffff);
        }
    }
    if (last_seg) unfold_zero(a, row_off, E, a.e_pad, lane, GRAD);
    if constexpr (KS > 0) {
        if (slice) {  // digits of the finished row, read back while it is still in L2
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax_hi = max(amax_hi, __shfl_xor_sync(0xffffffffu, amax_hi, o));
            const bool bad = amax_hi >= 0x7ff00000;
            const double amax = amax_hi == 0 ? 0.0 : __hiloint2double(amax_hi, (int)0xffffffff);
            double f1, f2;
            const double sc = slice_scale<KS>(amax, bad, f1, f2);
            if (lane == 0) a.scale[row] = sc;
            __syncwarp();
            slice_row_digits<KS>(a.U + row_off, true, (a.ld & 1) == 0, E, f1, f2, row % I8_BM, I8_BM, 0, a.kb_total,
                                 a.img, (int64_t)(row / I8_BM) * a.kb_total, lane);
        }
    }
}

// Gradient by contraction: grad[row, j] = sum_e G[row, e] * d u_e / d theta_j with the tangents
// recomputed on the fly (forward duals, as in unfold_generic) and consumed in registers.
// Replaces the round trip of the P tangent planes through HBM (written by the unfold kernel,
// read once by the transposed fold's epilogue: 1.5 GB per chunk of the bench, which made the
// unfold kernel WRITE-bound -- 0.61 ms with the stores, 0.41 ms without, tools/power_probe.py).
// One warp per row (the host keeps n_seg == 1); lane-serial partial sums, then an xor tree: the
// summation order is fixed and independent of the batch.
template <class Model, typename T>
__device__ __forceinline__ void unfold_contract_generic(const UnfoldCore& a) {
    constexpr int NP = Scalar<T>::NP;
    static_assert(NP > 0, "the contraction needs the dual type");
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * UNFOLD_WARPS + wib;
    const int row = (int)wid;
    if (row >= a.n_valid) return;
    const int E = a.gv.E;
    typename Model::Pre pre;
    Model::pre(a.theta + (size_t)row * a.P, pre);
    Model::norms(pre, a, a.theta + (size_t)row * a.P, lane);
    GridView gv = a.gv;
    gv.has_point = 1;
    auto load_point = [&](int e, double& pE, double& pl, double& pEm, double& plm) {
        const int ec = e <= E ? e : E, em = e < E ? e : E - 1;
        pE = a.gv.egrid[ec];
        pl = a.gv.lngrid[ec];
        pEm = a.gv.emid[em];
        plm = a.gv.lnmid[em];
    };
#ifdef UNFOLD_USED_MASK
    constexpr uint32_t used = UNFOLD_USED_MASK;  // compile-time: unused columns cost nothing
#else
    const uint32_t used = a.used_mask;
#endif
    const double* g_row = a.G + (size_t)row * a.ldg;
    double nE, nl, nEm, nlm;
    load_point(lane, nE, nl, nEm, nlm);
    double ng = (lane < 31 && lane < E) ? g_row[lane] : 0.0;
    double acc[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j] = 0.0;
    for (int e0 = 0; e0 < E; e0 += 31) {
        const int e = e0 + lane;
        gv.pE = nE;
        gv.pl = nl;
        gv.pEm = nEm;
        gv.plm = nlm;
        double g = ng;
        if (e0 + 31 < E) {
            load_point(e + 31, nE, nl, nEm, nlm);
            ng = (lane < 31 && e + 31 < E) ? g_row[e + 31] : 0.0;
        }
        const T res = Model::bin(pre, gv, e);
        // lane 31 and lanes past the grid hold no bin (their duals may be anything); clip(unfold,
        // 1e-300, 1e300) (likelihood.py:204) passes no derivative where it clips
        // (as a ZERO tangent, so that a NaN in G -- a NaN parameter vector -- still reaches the sum)
        if (lane < 31 && e < E) {
            const bool pass = !(a.clip && !(res.v >= 1e-300 && res.v <= 1e300));
#pragma unroll
            for (int j = 0; j < NP; ++j)
                if ((used >> j) & 1u) acc[j] = fma(g, pass ? res.d[j] : 0.0, acc[j]);
        }
    }
    double* out = a.grad_out + (size_t)row * a.grad_stride;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        if (j < a.P) {
            const double t = ((used >> j) & 1u) ? warp_sum(acc[j]) : 0.0;
            if (lane == 0) out[j] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Gradient by contraction on EDGES.  For a model that is a plain sum of leaves whose bin value is
// LINEAR in two neighbouring edge values --
//   trapezoid continuum leaves:  u_e = 0.5 (E_{e+1} - E_e) (g(E_e) + g(E_{e+1})),
//   antiderivative leaves:       u_e = h(E_{e+1}) - h(E_e)           (Comp::EDGE_DIFF, the power law)
// -- the contraction  grad_j = sum_e G_e du_e/dtheta_j  is re-ordered into a sum over edges,
//   grad_j = sum_k w_k dg(E_k)/dtheta_j,   w_k = 0.5 (G_{k-1} dE_{k-1} + G_k dE_k)   (trapezoid)
//                                          w_k = G_{k-1} - G_k                       (difference),
// so the tangents are never shuffled between lanes (one shuffle of the weight instead of one per
// tangent), and components with closed-form tangents (Comp::basis / Comp::finish: dg/dp_q =
// sum_i M_qi b_i(E) with M constant along the energy axis) accumulate S_i = sum_k w_k b_i(E_k) and
// apply M once per row -- no dual arithmetic, no structurally zero multiply-adds.  Components
// without a closed form are evaluated in Dual<NP> on the edge and contribute b_q = dg/dp_q.
// The clip of the unfolded model (likelihood.py:204) passes no derivative: the bin value is formed
// from the edge values for that test alone.
template <bool B, class X, class Y>
struct Cond {
    using type = X;
};
template <class X, class Y>
struct Cond<false, X, Y> {
    using type = Y;
};
template <class C, class = void>
struct HasBasis {
    static constexpr bool value = false;
    static constexpr int NB = C::NP;
};
template <class C>
struct HasBasis<C, decltype((void)C::NB)> {
    static constexpr bool value = true;
    static constexpr int NB = C::NB;
};
template <class C, class = void>
struct IsEdgeDiff {
    static constexpr bool value = false;
};
template <class C>
struct IsEdgeDiff<C, decltype((void)C::EDGE_DIFF)> {
    static constexpr bool value = true;
};

template <int KIND>
struct EdgeLeaf {
    using C = Comp<KIND>;
    static constexpr bool CLOSED = HasBasis<C>::value;
    static constexpr bool DIFF = IsEdgeDiff<C>::value;
    static constexpr int NP = C::NP, NC = C::NC, NB = HasBasis<C>::NB;
    static_assert(CLOSED || ((C::RULE == RULE_CONTINUUM_ADD || C::RULE == RULE_CONTINUUM_MUL) && C::NG == 1),
                  "edge contraction: unsupported component");
    using T = typename Cond<CLOSED, double, Dual<NP>>::type;
    struct Pre {
        T p[NP];
        T c[NC];
    };
    __device__ __forceinline__ static void pre(const double* val, const double* aux, Pre& s) {
#pragma unroll
        for (int k = 0; k < NP; ++k) s.p[k] = Flat<T>::seed(val[k], k);
#pragma unroll
        for (int k = 0; k < NC; ++k) s.c[k] = Flat<T>::seed(0.0, -1);
        C::pre(s.p, aux, s.c);
    }
    __device__ __forceinline__ static void edge(const Pre& s, double E, double lnE, double& v, double* b) {
        if constexpr (CLOSED) {
            C::basis(s.p, s.c, E, lnE, v, b);
        } else {
            T g[1];
            C::edge(s.p, s.c, E, lnE, g);
            v = g[0].v;
#pragma unroll
            for (int q = 0; q < NP; ++q) b[q] = g[0].d[q];
        }
    }
    __device__ __forceinline__ static void finish(const Pre& s, const double* S, double* d) {
        if constexpr (CLOSED) {
            C::finish(s.p, s.c, S, d);
        } else {
#pragma unroll
            for (int q = 0; q < NP; ++q) d[q] = S[q];
        }
    }
};

// Constant (models/mul.py:54-56): its bin value is the parameter itself -- as an edge value with the trapezoid
// rule of a multiplicative leaf, (f + f) / 2 = f exactly, tangent 1
template <>
struct EdgeLeaf<ELISA_COMP_CONSTANT> {
    static constexpr bool CLOSED = true, DIFF = false;
    static constexpr int NP = 1, NC = 0, NB = 1;
    struct Pre {
        double f;
    };
    __device__ __forceinline__ static void pre(const double* val, const double*, Pre& s) { s.f = val[0]; }
    __device__ __forceinline__ static void edge(const Pre& s, double, double, double& v, double* b) {
        v = s.f;
        b[0] = 1.0;
    }
    __device__ __forceinline__ static void finish(const Pre&, const double* S, double* d) { d[0] = S[0]; }
};

// Model (generated, ModelE): Pre, pre(theta, Pre&), NACC accumulators of which bit i of DMASK marks the
// difference-weighted ones, HAS_T / HAS_D, edge(pre, gv, E, lnE, vT, vD, b), finish(pre, S, grad[NPAR]).
template <class Model>
__device__ __forceinline__ void unfold_contract_edges(const UnfoldCore& a) {
    constexpr int NA = Model::NACC;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = (int)((int64_t)blockIdx.x * UNFOLD_WARPS + wib);
    if (row >= a.n_valid) return;
    const int E = a.gv.E;
    typename Model::Pre pre;
    Model::pre(a.theta + (size_t)row * a.P, pre);
    const double* g_row = a.G + (size_t)row * a.ldg;
    double nE = a.gv.egrid[min(lane, E)], nl = a.gv.lngrid[min(lane, E)];
    double ng = (lane < 31 && lane < E) ? g_row[lane] : 0.0;
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    for (int e0 = 0; e0 < E; e0 += 31) {
        const int e = e0 + lane;
        const double pE = nE, pl = nl, g = ng;
        if (e0 + 31 < E) {  // the next pass's edge and weight, loaded under this pass's arithmetic
            const int ec = min(e + 31, E);
            nE = a.gv.egrid[ec];
            nl = a.gv.lngrid[ec];
            ng = (lane < 31 && e + 31 < E) ? g_row[e + 31] : 0.0;
        }
        double vT = 0.0, vD = 0.0, b[NA];
        Model::edge(pre, a.gv, pE, pl, vT, vD, b);
        const double hd = 0.5 * (__shfl_down_sync(0xffffffffu, pE, 1) - pE);
        double u = 0.0;
        if constexpr (Model::HAS_T) u = hd * (vT + __shfl_down_sync(0xffffffffu, vT, 1));
        if constexpr (Model::HAS_D) u += __shfl_down_sync(0xffffffffu, vD, 1) - vD;
        // lane 31 and lanes past the grid hold no bin (g = 0); a clipped bin passes no derivative: its
        // weight is g * 0 -- zero, and a zero weight skips the accumulation, so that 0 * inf of an
        // overflowing tangent does not reach the sum, while a NaN in G (a NaN parameter vector makes
        // the bin NaN, which fails the clip test too) stays NaN and does
        const bool pass = !(a.clip && !(u >= 1e-300 && u <= 1e300));
        const double gg = g * (pass ? 1.0 : 0.0);
        double wT = 0.0, wD = 0.0;
        if constexpr (Model::HAS_T) {
            const double gt = gg * hd;
            const double up = __shfl_up_sync(0xffffffffu, gt, 1);
            wT = lane ? gt + up : gt;
        }
        if constexpr (Model::HAS_D) {
            const double up = __shfl_up_sync(0xffffffffu, gg, 1);
            wD = lane ? up - gg : -gg;
        }
        if (Model::HAS_T && wT != 0.0) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
                if (!((Model::DMASK >> i) & 1u)) acc[i] = fma(wT, b[i], acc[i]);
        }
        if (Model::HAS_D && wD != 0.0) {
#pragma unroll
            for (int i = 0; i < NA; ++i)
                if ((Model::DMASK >> i) & 1u) acc[i] = fma(wD, b[i], acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = warp_sum(acc[i]);
    double grad[Model::NPAR];
    Model::finish(pre, acc, grad);
    double* out = a.grad_out + (size_t)row * a.grad_stride;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < Model::NPAR; ++j) out[j] = grad[j];
    }
}

// The same for an expression TREE of sums and products (an absorbed continuum, PhAbs * (PowerLaw + Blackbody): what
// people fit).  The bin value of every leaf is linear in two neighbouring edge values -- additive trapezoid leaves,
// the power law's antiderivative, and multiplicative trapezoid leaves, (m(E_e) + m(E_e+1)) / 2 -- and the `+` / `*`
// tree acts on bin values, so  grad_j = sum_e G_e du_e/dtheta_j  goes through the tree in REVERSE per bin (the adjoint
// of a leaf's bin value is G_e times the product of the factors it is multiplied with) and then, leaf by leaf, onto
// the edges exactly as above, with the leaf's own weights.  Model (generated, ModelT): Pre, pre(theta, Pre&),
// NLEAF, NACC, edge(pre, gv, ec, E, lnE, v[NLEAF], b[NACC]) -- edge value and tangent basis of every leaf;
// bins(v, hd, binv) -- the leaves' bin values (shuffles: every lane calls it); tree(binv, adj, u) -- the model's bin
// value u and the adjoints of the leaves' bin values for a unit adjoint of u; accumulate(adj, gg, hd, lane, b, acc);
// finish(pre, S, grad[NPAR]).
template <class Model>
__device__ __forceinline__ void unfold_contract_tree(const UnfoldCore& a) {
    constexpr int NA = Model::NACC, NL = Model::NLEAF;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int row = (int)((int64_t)blockIdx.x * UNFOLD_WARPS + wib);
    if (row >= a.n_valid) return;
    const int E = a.gv.E;
    typename Model::Pre pre;
    Model::pre(a.theta + (size_t)row * a.P, pre);
    const double* g_row = a.G + (size_t)row * a.ldg;
    double nE = a.gv.egrid[min(lane, E)], nl = a.gv.lngrid[min(lane, E)];
    double ng = (lane < 31 && lane < E) ? g_row[lane] : 0.0;
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = 0.0;
    for (int e0 = 0; e0 < E; e0 += 31) {
        const int e = e0 + lane;
        const double pE = nE, pl = nl, g = ng;
        if (e0 + 31 < E) {  // the next pass's edge and weight, loaded under this pass's arithmetic
            const int ec = min(e + 31, E);
            nE = a.gv.egrid[ec];
            nl = a.gv.lngrid[ec];
            ng = (lane < 31 && e + 31 < E) ? g_row[e + 31] : 0.0;
        }
        double v[NL], b[NA], binv[NL], adj[NL], u;
        Model::edge(pre, a.gv, min(e, E), pE, pl, v, b);
        const double hd = 0.5 * (__shfl_down_sync(0xffffffffu, pE, 1) - pE);
        Model::bins(v, hd, binv);
        Model::tree(binv, adj, u);
        // as in unfold_contract_edges: no bin in lane 31 and past the grid (g = 0), a clipped bin passes no
        // derivative, zero weights skip the accumulation, a NaN in G stays NaN
        const bool pass = !(a.clip && !(u >= 1e-300 && u <= 1e300));
        const double gg = g * (pass ? 1.0 : 0.0);
        Model::accumulate(adj, gg, hd, lane, b, acc);
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i] = warp_sum(acc[i]);
    double grad[Model::NPAR];
    Model::finish(pre, acc, grad);
    double* out = a.grad_out + (size_t)row * a.grad_stride;
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < Model::NPAR; ++j) out[j] = grad[j];
    }
}
// weight of an edge from the per-bin adjoint t of the lane's bin (bin e lies between the edges e and e + 1)
__device__ __forceinline__ double edge_weight_sum(double t, int lane) {   // trapezoid: both neighbouring bins
    const double up = __shfl_up_sync(0xffffffffu, t, 1);
    return lane ? t + up : t;
}
__device__ __forceinline__ double edge_weight_diff(double t, int lane) {  // difference of edge values: F(E_e+1) - F(E_e)
    const double up = __shfl_up_sync(0xffffffffu, t, 1);
    return lane ? up - t : -t;
}

// ---------------------------------------------------------------------------------------
// Small batches in ONE kernel (NUTS chains, a handful of walkers: BASELINE configs[2]).
// The chunk pipeline is five dependent launches per spectrum -- unfold, fold, statistic + digits,
// transposed fold, contraction -- each ~10 us at 64 rows, bound by launch and tail latency, not by
// work.  For a small spectrum the whole evaluation of one parameter vector fits one CTA:
//   1. the warps evaluate the model in dual numbers, u_e and du_e/dtheta_j into shared memory;
//   2. thread c folds its channel in FP64, x_c = sum_e R[e, c] u_e and t_jc = sum_e R[e, c] du_e/dtheta_j
//      (forward mode: (1 + P) FMAs per response element, read coalesced from L2);
//   3. the statistic of the channel gives l_c and w_c = dl/dx_c; loglike = sum_c l_c and
//      grad_j = sum_c w_c t_jc are reduced in a fixed order (lane-serial, xor tree, warps in order).
// Plain FP64 FMAs: no digit slicing, no accuracy guard needed.  One CTA per (parameter vector, spectrum):
// the spectra of a joint fit that share a model and a statistic (one translation unit) are ONE launch.
struct TinyArgs {      // per spectrum, fixed at plan creation (an array of these lives in device memory)
    UnfoldCore core;   // gv, norms, used_mask, clip (the per-call fields of it are not used)
    const double* R;   // [E, ldr] dense response, row = energy bin
    int32_t ldr, C;
    ChanArrays ch;
    int64_t counts_off;  // column of this spectrum in the per-row counts arrays
};
constexpr int TINY_MAX_PEERS = 8;
struct TinyCall {      // per launch: spectra [s0, s0 + gridDim.y) of the plan, one CTA per (row, spectrum)
    const TinyArgs* spec;  // the plan's array, indexed by spectrum
    int32_t s0, S, P, n_valid;
    const double* theta;       // [n, P]
    const double* counts_on;   // [n, ld_counts] per-row counts (nullptr: the plan's)
    const double* counts_off;
    int64_t ld_counts;
    double* loglike;   // [n, S]
    double* grad;      // [n, S, P] or nullptr (value only)
    long long* clk;    // nullptr, or 8 slots: thread 0 of CTA (0, 0) leaves clock64() at the phase boundaries (ELISA_B200_TINY_CLOCKS)
    // peer gather fused into the kernel (elisa_b200_plan_set_gather): every result is also stored into the other
    // ranks' full arrays over NVLink (peer-mapped CUDA IPC memory), and the LAST CTA of the call's last launch
    // raises this rank's flag at every peer -- evaluation and all-gather are one kernel
    int32_t world, rank;                  // world <= 1: no peers
    double* peer_ll[TINY_MAX_PEERS];      // rank r's full loglike array at this launch's first row
    double* peer_grad[TINY_MAX_PEERS];    // ... gradient array (nullptr: not gathered)
    long long* peer_flags[TINY_MAX_PEERS];
    unsigned int* done;                   // CTAs of this launch that have stored their results (device counter, self-resetting)
    long long epoch;                      // sequence number to raise at the peers; 0: not the last launch of the call
};

// Launch shape: the generated translation unit may override the three numbers (the host reads the same ones from
// ELISA_B200_TINY_TUNE = "threads,channels per thread,unroll" -- a tuning hook, tools/tiny_tune.py).
#ifndef TINY_THREADS_CFG
#define TINY_THREADS_CFG 256
#endif
#ifndef TINY_CH_CFG
#define TINY_CH_CFG 0
#endif
#ifndef TINY_UNROLL_CFG
#define TINY_UNROLL_CFG 4
#endif
constexpr int TINY_THREADS = TINY_THREADS_CFG;
constexpr int TINY_UNROLL = TINY_UNROLL_CFG;  // energies per step of the fold loop: UNROLL x CH independent loads in flight per thread
// channels per thread of the fold (register blocking: one broadcast read of u_e, du_e/dtheta serves CH channels);
// 1 + NP accumulators per channel live in registers
__host__ __device__ constexpr int tiny_ch(int NP, int cfg = TINY_CH_CFG) { return cfg > 0 ? cfg : (NP <= 4 ? 4 : 2); }

// shared memory (doubles): u [E] | du/dtheta [NP][E] | partial folds [warps][1 + NP][CB] | warp sums [warps][1 + NP]
// with CB = 32 * channels per thread = channels per pass
__host__ __device__ inline size_t tiny_smem_doubles(int E, int NP, int threads = TINY_THREADS, int ch = 0) {
    const int cb = 32 * (ch > 0 ? ch : tiny_ch(NP));
    return (size_t)(1 + NP) * E + (size_t)(threads / 32) * (1 + NP) * cb + (size_t)(threads / 32) * (1 + NP);
}

template <class Model, int NP, int STAT>
__device__ __forceinline__ void tiny_eval(const TinyCall& k) {
    using T = Dual<NP>;
    extern __shared__ double tiny_smem[];
    const int spec = k.s0 + blockIdx.y;
    const TinyArgs& t = k.spec[spec];  // global memory: uniform loads
    const UnfoldCore& a = t.core;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int row = blockIdx.x;
    if (row >= k.n_valid) return;
    const int E = a.gv.E;
    const bool want_grad = k.grad != nullptr;
    const double* theta = k.theta + (size_t)row * k.P;
    const bool clk = k.clk != nullptr && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0;
#define TINY_CLK(i) do { if (clk) k.clk[i] = clock64(); } while (0)
    TINY_CLK(0);
#ifdef UNFOLD_USED_MASK
    constexpr uint32_t used = UNFOLD_USED_MASK;
#else
    const uint32_t used = a.used_mask;
#endif
    constexpr int NW = TINY_THREADS / 32;         // warps = groups the energies are dealt to
    constexpr int CH = tiny_ch(NP);               // channels per thread
    constexpr int CB = 32 * CH;                   // channels per pass
    static_assert(CB <= TINY_THREADS, "one thread per channel of a pass does the statistic");
    double* su = tiny_smem;                       // [E]
    double* sd = su + E;                          // [NP][E]
    double* part = sd + (size_t)NP * E;           // [NW][1 + NP][CB]
    double* red = part + (size_t)NW * (1 + NP) * CB;  // [warps][1 + NP]
    // 1. the model in dual numbers; warp w takes the passes w, w + 8, ...
    {
        typename Model::Pre pre;
        Model::pre(theta, pre);
        Model::norms(pre, a, theta, lane);
        TINY_CLK(1);
        GridView gv = a.gv;
        gv.has_point = 0;
        for (int e0 = 31 * wib; e0 < E; e0 += 31 * (TINY_THREADS / 32)) {
            const int e = e0 + lane;
            const T res = Model::bin(pre, gv, e);
            if (lane < 31 && e < E) {
                const double v = res.v;
                const bool pass = !(a.clip && !(v >= 1e-300 && v <= 1e300));
                su[e] = pass ? v : (v < 1e-300 ? 1e-300 : (v > 1e300 ? 1e300 : v));  // jnp.clip propagates NaN
#pragma unroll
                for (int j = 0; j < NP; ++j)
                    if ((used >> j) & 1u) sd[(size_t)j * E + e] = pass ? res.d[j] : 0.0;
            }
        }
    }
    TINY_CLK(2);
    __syncthreads();
    TINY_CLK(3);
    // 2 + 3. fold and statistic, CB channels per pass.  Warp w folds the energies [w E / NW, (w + 1) E / NW) of the
    // channels c0 + lane + 32 q, q < CH: u_e and du_e/dtheta are broadcast reads of shared memory, and with one
    // channel per thread those 1 + NP reads per response element were what bound the loop (r02c clocks: 8.2 k of
    // 14 k cycles); blocked over CH channels a read serves CH (1 + NP) multiply-adds, and UNROLL x CH loads of the
    // response are in flight per thread.  Then thread c of the pass adds the NW partial folds in a fixed order
    // and does the statistic.
    const int e_lo = (int)((int64_t)E * wib / NW), e_hi = (int)((int64_t)E * (wib + 1) / NW);
    double ll = 0.0, g[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) g[j] = 0.0;
    const double* con = k.counts_on ? k.counts_on + (size_t)row * k.ld_counts + t.counts_off : nullptr;
    const double* coff = k.counts_off ? k.counts_off + (size_t)row * k.ld_counts + t.counts_off : nullptr;
    const int C = t.C, ldr = t.ldr;
    const double* R = t.R;
    const ChanArrays ch = t.ch;
    for (int c0 = 0; c0 < C; c0 += CB) {
        double xq[CH], tq[CH][NP];
        bool ok[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            xq[q] = 0.0;
            ok[q] = c0 + lane + 32 * q < C;
#pragma unroll
            for (int j = 0; j < NP; ++j) tq[q][j] = 0.0;
        }
        {
            const double* r = R + c0 + lane;
            auto step = [&](const double (&rv)[CH], int e) {
                const double u = su[e];
#pragma unroll
                for (int q = 0; q < CH; ++q) xq[q] = fma(rv[q], u, xq[q]);
                if (want_grad) {
#pragma unroll
                    for (int j = 0; j < NP; ++j)
                        if ((used >> j) & 1u) {
                            const double dj = sd[(size_t)j * E + e];
#pragma unroll
                            for (int q = 0; q < CH; ++q) tq[q][j] = fma(rv[q], dj, tq[q][j]);
                        }
                }
            };
            int e = e_lo;
            for (; e + TINY_UNROLL <= e_hi; e += TINY_UNROLL) {
                double rv[TINY_UNROLL][CH];
#pragma unroll
                for (int i = 0; i < TINY_UNROLL; ++i)
#pragma unroll
                    for (int q = 0; q < CH; ++q) rv[i][q] = ok[q] ? __ldg(r + (size_t)(e + i) * ldr + 32 * q) : 0.0;
#pragma unroll
                for (int i = 0; i < TINY_UNROLL; ++i) step(rv[i], e + i);
            }
            for (; e < e_hi; ++e) {
                double rv[CH];
#pragma unroll
                for (int q = 0; q < CH; ++q) rv[q] = ok[q] ? __ldg(r + (size_t)e * ldr + 32 * q) : 0.0;
                step(rv, e);
            }
        }
        if (c0 == 0) TINY_CLK(4);
        if (c0 > 0) __syncthreads();  // the previous pass's partial folds have been consumed
#pragma unroll
        for (int q = 0; q < CH; ++q) {
            part[((size_t)wib * (1 + NP)) * CB + lane + 32 * q] = xq[q];
#pragma unroll
            for (int j = 0; j < NP; ++j)
                if ((used >> j) & 1u) part[((size_t)wib * (1 + NP) + 1 + j) * CB + lane + 32 * q] = tq[q][j];
        }
        __syncthreads();
        const int c = c0 + tid;
        if (tid < CB && c < C) {
            double x = 0.0, tj[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) tj[j] = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                x += part[((size_t)w * (1 + NP)) * CB + tid];
#pragma unroll
                for (int j = 0; j < NP; ++j)
                    if ((used >> j) & 1u) tj[j] += part[((size_t)w * (1 + NP) + 1 + j) * CB + tid];
            }
            ChannelData d;
            const double sc = ch.scale[c];
            d.on = ch.on[c];
            d.gof_on = ch.gof_on[c];
            d.off = d.sigma = d.ratio = d.gof_off = d.den = 0.0;
            if (STAT == ELISA_STAT_CHI2) d.sigma = ch.sigma[c];
            if (STAT == ELISA_STAT_PSTAT || STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) {
                d.off = ch.off[c];
                d.ratio = ch.ratio[c];
            }
            if (STAT == ELISA_STAT_PGSTAT) d.sigma = ch.sigma[c];
            if (STAT == ELISA_STAT_WSTAT) d.gof_off = ch.gof_off[c];
            if (STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) d.den = ch.den[c];
            if (con != nullptr) {
                d.on = con[c];
                if (STAT != ELISA_STAT_CHI2) d.gof_on = poisson_gof(d.on);
            }
            if ((STAT == ELISA_STAT_PGSTAT || STAT == ELISA_STAT_WSTAT) && coff != nullptr) {
                d.off = coff[c];
                if (STAT == ELISA_STAT_WSTAT) d.gof_off = poisson_gof(d.off);
            }
            double dx = 0.0;
            ll += stat_channel<STAT, true>(x * sc, d, dx);
            const double w = dx * sc;
#pragma unroll
            for (int j = 0; j < NP; ++j)
                if ((used >> j) & 1u) g[j] = fma(w, tj[j], g[j]);
        }
    }
    // fixed-order reduction: lane-serial sums above, xor tree, then the warps of group 0 in order
    TINY_CLK(5);
    ll = warp_sum(ll);
    if (lane == 0) red[wib * (1 + NP)] = ll;
    if (want_grad) {
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const double s = ((used >> j) & 1u) ? warp_sum(g[j]) : 0.0;
            if (lane == 0) red[wib * (1 + NP) + 1 + j] = s;
        }
    }
    __syncthreads();
    if (tid <= NP && (tid == 0 || want_grad)) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < (CB < TINY_THREADS ? CB : TINY_THREADS) / 32; ++w) s += red[w * (1 + NP) + tid];
        const size_t i_ll = (size_t)row * k.S + spec, i_g = ((size_t)row * k.S + spec) * k.P + tid - 1;
        if (tid == 0)
            k.loglike[i_ll] = s;
        else if (tid - 1 < k.P)
            k.grad[i_g] = s;
        if (k.world > 1) {
#pragma unroll 1
            for (int r = 0; r < k.world; ++r) {
                if (tid == 0) {
                    if (k.peer_ll[r] + i_ll != k.loglike + i_ll) k.peer_ll[r][i_ll] = s;
                } else if (tid - 1 < k.P && k.peer_grad[r] != nullptr) {
                    if (k.peer_grad[r] + i_g != k.grad + i_g) k.peer_grad[r][i_g] = s;
                }
            }
        }
    }
    if (k.world > 1 && k.epoch != 0) {
        // results of this CTA are on their way to the peers; the last CTA to get here raises the flags behind them
        __shared__ bool last;
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const unsigned int total = gridDim.x * gridDim.y;
            last = atomicAdd(k.done, 1u) == total - 1;
            if (last) *k.done = 0u;
        }
        __syncthreads();
        if (last && tid < k.world && tid != k.rank) {
            __threadfence_system();
            *reinterpret_cast<volatile long long*>(k.peer_flags[tid] + k.rank) = k.epoch;
        }
    }
    TINY_CLK(6);
#undef TINY_CLK
}

}  // namespace elisa
