/* elisa_b200 -- C ABI of the B200-native forward-folding likelihood path.
 *
 * Drop-in boundary for ONE hot path of wcxve/elisa: for a batch of N parameter
 * vectors, evaluate the built-in spectral components over each spectrum's photon
 * energy bins, fold through the instrument response, add the scaled background, and
 * reduce the per-channel fit statistic to one log-likelihood per (sample, spectrum),
 * with its gradient with respect to the parameters.
 *
 * The reference (pure Python on JAX) has no FFI of its own for this path; the seam
 * this library sits behind is the likelihood-factory seam
 *   elisa/infer/helper.py:313-323,332-367   likelihood_wrapper[stat](data, model.eval)
 *   elisa/infer/likelihood.py:184-450       chi2 / cstat / pstat / pgstat / wstat closures
 *   elisa/models/model.py:202-219,363-436   CompiledModel.eval (+ vmap batching)
 *   elisa/infer/helper.py:494-512,571-594   get_loglike / deviance reductions
 *   elisa/infer/fit.py:487                  jax.grad(deviance)  (value + gradient)
 * INTEGRATION.md shows the jax.ffi / ctypes binding a maintainer adds on the reference
 * side.  All pointers are plain C; no torch / XLA types appear in any signature.
 *
 * Conventions
 *  - all floating point data is IEEE double, C-contiguous (the reference runs with
 *    jax_enable_x64, src/elisa/__init__.py:35);
 *  - theta is [N, P]: row n is one parameter vector in CONSTRAINED space (transforms,
 *    priors and log-Jacobians stay in numpyro, infer/helper.py:409-426);
 *  - loglike is [N, S] (get_loglike()['group'], infer/helper.py:494-512); deviance is
 *    -2 * loglike and the joint total is the row sum;
 *  - grad is [N, S, P]: d loglike[n, s] / d theta[n, p];
 *  - "_dev" entry points take DEVICE pointers and only enqueue work on the given
 *    stream (no allocation, no host synchronisation); batches of <= 1024 parameter vectors
 *    without per-replica counts are launch-bound, so their kernel sequence is captured into a
 *    CUDA graph on the second call of a shape and replayed afterwards, the spectra of a joint
 *    fit as parallel branches (env
 *    ELISA_B200_GRAPHS=0 disables it; results are bit-identical either way); "_host" entry points take HOST
 *    pointers, stage through the plan's pinned buffers and return when results are
 *    in the caller's memory;
 *  - every function returns 0 on success or a negative ELISA_ERR_* code and never
 *    throws or aborts; NaN/Inf in results are data, not errors
 *    (callers filter them: infer/helper.py:659-664).
 *  - plans are immutable after creation except for their private workspace.  Entry points may
 *    be called on one plan from several host threads (numpyro chain_method='parallel'): calls
 *    that touch the workspace are serialised by a mutex inside the plan, and work enqueued on
 *    DIFFERENT streams by successive calls is the caller's to order (the workspace is shared).
 *    For calls that should overlap on the device create one plan per thread / stream.
 */
#ifndef ELISA_B200_H_
#define ELISA_B200_H_

#ifdef __CUDACC_RTC__ /* runtime-compiled device code (csrc/unfold_skel.cuh): no host headers */
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ELISA_B200_ABI_VERSION 6

/* ---- error codes ---------------------------------------------------------------- */
#define ELISA_OK 0
#define ELISA_ERR_INVALID -1     /* bad descriptor / argument                         */
#define ELISA_ERR_CUDA -2        /* a CUDA runtime call failed (see last_error)       */
#define ELISA_ERR_UNSUPPORTED -3 /* valid in the reference, not handled by this path  */
#define ELISA_ERR_NOMEM -4

/* ---- fit statistics: elisa/infer/likelihood.py:35 (Statistic literal) ----------- */
#define ELISA_STAT_CHI2 0   /* likelihood.py:184-227 */
#define ELISA_STAT_CSTAT 1  /* likelihood.py:230-272 */
#define ELISA_STAT_PSTAT 2  /* likelihood.py:275-321 */
#define ELISA_STAT_PGSTAT 3 /* likelihood.py:324-387 */
#define ELISA_STAT_WSTAT 4  /* likelihood.py:390-450 */

/* ---- built-in components: elisa/models/add.py:20-37, elisa/models/mul.py:24-35 -- */
/* parameter order inside a component is the reference's _config order.             */
#define ELISA_COMP_BAND 0             /* add.py:100-121  (alpha, beta, Ec, K)             */
#define ELISA_COMP_BANDEP 1           /* add.py:177-200  (alpha, beta, Ep, K)             */
#define ELISA_COMP_BENTPL 2           /* add.py:231-236  (alpha, Eb, K)                   */
#define ELISA_COMP_BLACKBODY 3        /* add.py:266-286  (kT, K)                          */
#define ELISA_COMP_BLACKBODYRAD 4     /* add.py:317-338  (kT, K)                          */
#define ELISA_COMP_BROKENPL 5         /* add.py:378-393  (alpha1, Eb, alpha2, K)          */
#define ELISA_COMP_COMPT 6            /* add.py:473-479  (alpha, Ep, K)                   */
#define ELISA_COMP_CUTOFFPL 7         /* add.py:434-439  (alpha, Ec, K)                   */
#define ELISA_COMP_GAUSS 8            /* add.py:511-516  (El, sigma, K)                   */
#define ELISA_COMP_LORENTZ 9          /* add.py:553-561  (El, sigma, K)                   */
#define ELISA_COMP_OTTB 10            /* add.py:590-594  (kT, K)                          */
#define ELISA_COMP_OTTS 11            /* add.py:625-629  (Ec, K)                          */
#define ELISA_COMP_PLENFLUX 12        /* add.py:682-708  (alpha, F)  aux = emin, emax     */
#define ELISA_COMP_PLPHFLUX 13        /* add.py:682-708  (alpha, F)  aux = emin, emax     */
#define ELISA_COMP_POWERLAW 14        /* add.py:40-50,655-657  (alpha, K)                 */
#define ELISA_COMP_SMOOTHLYBROKENPL 15 /* add.py:846-861 (alpha1, Eb, alpha2, rho, K)     */
#define ELISA_COMP_CONSTANT 16        /* mul.py:54-56    (f)                              */
#define ELISA_COMP_EDGE 17            /* mul.py:88-94    (Ec, D)                          */
#define ELISA_COMP_EXPABS 18          /* mul.py:116-118  (Ec)                             */
#define ELISA_COMP_EXPFAC 19          /* mul.py:154-159  (A, f, Ec)                       */
#define ELISA_COMP_GABS 20            /* mul.py:195-204  (El, sigma, tau)                 */
#define ELISA_COMP_HIGHECUT 21        /* mul.py:236-240  (Ec, Ef)                         */
#define ELISA_COMP_PLABS 22           /* mul.py:266-271  (alpha, K)                       */
#define ELISA_COMP_PHOTONABS 23       /* mul.py:274-299,360-388 (nH); PhAbs / TBAbs / WAbs:
                                        * exp(-nH sigma(E)), sigma from the cross-section table
                                        * of the descriptor, log-log interpolated, extrapolated  */
#define ELISA_COMP_COUNT 24

/* numerical bin integration: elisa/models/model.py:1731-1761 */
#define ELISA_METHOD_TRAPZ 0
#define ELISA_METHOD_SIMPSON 1

#define ELISA_MAX_COMP_PARAMS 5
#define ELISA_MAX_COMPONENTS 16 /* leaves of one model expression             */
#define ELISA_MAX_PARAMS 32     /* columns of theta (normal_eq_dev: <= 16)    */

/* postfix opcodes of the model expression (models/model.py:1285-1295): an entry >= 0
 * pushes the bin values of component #entry, ELISA_OP_ADD / ELISA_OP_MUL pop two
 * operands and push lhs + rhs / lhs * rhs (op applied to BIN values).            */
#define ELISA_OP_ADD -1
#define ELISA_OP_MUL -2
/* Flux normalisation PhFlux / EnFlux (models/conv.py:145-256): the unary opcode
 * ELISA_OP_NORM0 - k replaces the operand on top of the stack, an additive sub-expression
 * N(E), by F_k / flux_k * N(E), flux_k = sum over the bins of ElisaNormDesc #k's own grid of
 * that same sub-expression (times keV_to_erg * sqrt(e_j e_j+1) for EnFlux).              */
#define ELISA_OP_NORM0 -16
#define ELISA_MAX_NORMS 2
#define ELISA_MAX_SHIFTS 2 /* FREE energy shifts (z / v a fit parameter) around one component */
#define ELISA_SHIFT_Z 1
#define ELISA_SHIFT_ZA 2
#define ELISA_SHIFT_V 3
#define ELISA_NORM_PHFLUX 0 /* conv.py:185-195 */
#define ELISA_NORM_ENFLUX 1 /* conv.py:244-256 */

#ifndef __CUDACC_RTC__ /* host-side descriptors and entry points */
typedef struct ElisaComponentDesc {
    int32_t kind;   /* ELISA_COMP_*                                                  */
    int32_t method; /* ELISA_METHOD_* (ignored by analytic-integral components)      */
    /* where each parameter comes from: >= 0 -> column of theta (two slots naming one
     * column = linked parameters); -1 -> the constant in param_const (fixed values,
     * models/model.py:2248)                                                         */
    int32_t param_src[ELISA_MAX_COMP_PARAMS];
    double param_const[ELISA_MAX_COMP_PARAMS];
    double aux[2]; /* PLPhFlux / PLEnFlux: emin, emax (add.py:662-680)               */
    /* ELISA_COMP_PHOTONABS: one cross-section table of the reference's tables/xsect.hdf5
     * (its 'energy' dataset and '{phabs|tbabs|wabs}/{xsect}/{abund}'), table_len >= 2
     * non-decreasing positive energies; NULL / 0 for every other kind.  The plan evaluates
     * sigma on the photon grid at creation exactly as mul.py:274-283 does
     * (exp(interp(log e, log energy, log xsect)) with linear extrapolation).          */
    const double* table_energy;
    const double* table_xsect;
    int32_t table_len;
    /* 1: the two arrays already hold ln(energy), ln(xsect) -- what `_make_interp` keeps
     * (mul.py:275-276).  The reference takes np.log of the float32 datasets of xsect.hdf5, i.e.
     * FLOAT32 logarithms promoted to float64 afterwards; a binding that wants its numbers
     * reproduces that rounding on its side and passes the result here.  Energies may repeat
     * (the float32 table has 87 duplicates): jnp.interp never selects a zero-width segment
     * except at the clipped ends, where it returns the left value.                       */
    int32_t table_log;
    /* Energy shifts ZAShift / ZMShift / VAShift / VMShift with a FIXED z or v
     * (models/conv.py:294-300, 335-341, 378-386, 424-432): the component is evaluated on
     * photon_egrid * grid_scale (the product of the factors 1 + z, 1 - v/c of the shifts
     * around it) and its bin values are multiplied by out_scale (the product of ZAShift's
     * 1 / (1 + z) around an additive component).  0 reads as 1.  A component that occurs
     * inside and outside a shift is two entries with the same param_src.  norm_* are the
     * same two factors on the flux grid of the enclosing ELISA_OP_NORM (shifts INSIDE the
     * normalised sub-expression only).                                               */
    double grid_scale, out_scale;
    double norm_grid_scale, norm_out_scale;
    /* The same shifts with a FREE z or v (an ordinary fit parameter, conv.py:294-300): up to
     * ELISA_MAX_SHIFTS entries around the component.  shift_kind[k]: 0 = none, ELISA_SHIFT_Z = the grid
     * is multiplied by 1 + z, ELISA_SHIFT_ZA = the same and the bin values are divided by 1 + z
     * (ZAShift around an additive component), ELISA_SHIFT_V = the grid is multiplied by 1 - v / c
     * (c = 299792.458 km/s); shift_src[k] = the theta column of z / v.  The factor then is a dual
     * number: the component is differentiated with respect to its energy grid as well.  Needs the
     * runtime-specialised kernel; not available for table components or together with flux
     * normalisations (ELISA_ERR_UNSUPPORTED).                                                  */
    int32_t shift_kind[ELISA_MAX_SHIFTS];
    int32_t shift_src[ELISA_MAX_SHIFTS];
} ElisaComponentDesc;

/* One PhFlux / EnFlux normalisation (models/conv.py:21-143).                        */
typedef struct ElisaNormDesc {
    int32_t kind;      /* ELISA_NORM_*                                                */
    int32_t f_src;     /* the flux parameter F: column of theta, or -1 -> f_const     */
    double f_const;
    const double* grid; /* [n_grid] flux_egrid: geomspace / linspace(emin, emax, ngrid),
                         * conv.py:81-84; positive, increasing                        */
    int32_t n_grid;     /* >= 2                                                       */
} ElisaNormDesc;

typedef struct ElisaModelDesc {
    int32_t n_components;
    const ElisaComponentDesc* components;
    int32_t n_ops;
    const int32_t* ops; /* postfix program, see ELISA_OP_*                           */
    int32_t n_norms;    /* <= ELISA_MAX_NORMS; normalisations may not nest           */
    const ElisaNormDesc* norms;
} ElisaModelDesc;

/* One dataset: the fields of FixedData the path reads (data/base.py:1893-1975).     */
typedef struct ElisaSpectrumDesc {
    int32_t n_energies; /* E = len(photon_egrid) - 1                                 */
    int32_t n_channels; /* C                                                         */
    const double* photon_egrid; /* [E + 1]                                           */
    /* dense response, FixedData.response_matrix layout: [E, C] row-major; or NULL   */
    const double* response_dense;
    /* sparse response (FixedData.sparse_matrix, used when response_sparse): CSR by
     * CHANNEL row of R.T = what BCSR.from_scipy_sparse(sparse.T) holds
     * (infer/likelihood.py:177-181); or NULL                                        */
    const int32_t* csr_indptr;  /* [C + 1]                                           */
    const int32_t* csr_indices; /* [nnz] energy-bin index                            */
    const double* csr_values;   /* [nnz]                                             */
    const double* area_scale;   /* [C]                                               */
    double exposure;            /* spec_exposure                                     */
    int32_t stat;               /* ELISA_STAT_*                                      */
    /* observed data; which arrays a statistic reads (likelihood.py):
     *   chi2  : on = net_counts,  on_error = net_errors
     *   cstat : on = spec_counts
     *   pstat : on = spec_counts, off = back_counts, back_ratio
     *   pgstat: on = spec_counts, off = back_counts, off_error = back_errors, back_ratio
     *   wstat : on = spec_counts, off = back_counts, back_ratio                     */
    const double* on_counts;  /* [C]                                                 */
    const double* on_error;   /* [C] or NULL                                         */
    const double* off_counts; /* [C] or NULL                                         */
    const double* off_error;  /* [C] or NULL                                         */
    const double* back_ratio; /* [C] or NULL                                         */
    ElisaModelDesc model;
} ElisaSpectrumDesc;

typedef struct ElisaPlanDesc {
    int32_t abi_version; /* ELISA_B200_ABI_VERSION                                   */
    int32_t device;      /* CUDA device ordinal                                      */
    int32_t n_params;    /* P, columns of theta                                      */
    int32_t n_spectra;   /* S                                                        */
    const ElisaSpectrumDesc* spectra;
    int64_t chunk; /* samples processed per internal pass (0 = default); sets the
                    * private workspace: chunk * (2 E + C) doubles for the widest
                    * spectrum                                                       */
    int32_t flags; /* ELISA_FLAG_* bits, 0 = defaults                                */
} ElisaPlanDesc;

/* The component-evaluation kernel is specialised per model expression at plan creation
 * with NVRTC (csrc/unfold_skel.cuh); without libnvrtc, or with ELISA_FLAG_NO_JIT /
 * env ELISA_B200_JIT=0, the ahead-of-time interpreter kernel runs the same program.
 * ELISA_FLAG_REQUIRE_JIT / env ELISA_B200_JIT=require make a failed specialisation an error. */
#define ELISA_FLAG_NO_JIT 1
#define ELISA_FLAG_REQUIRE_JIT 2
/* Fold engine.  Default: the sliced-integer fold (csrc/fold_i8.cuh) -- operands split into
 * `slices` balanced base-256 digits, digit products on the int8 tcgen05 tensor cores with
 * exact int32 accumulation in TMEM, recombined in FP64: FP64-class accuracy (normwise
 * ~2^(-8 slices + 3); 6 slices hold the 1e-10 tolerance of the statistic and gradient) at
 * several times the FP64 DMMA rate.  By default a spectrum whose packed response is under
 * 5 % stored (narrow band-sparse RSPs), or with fewer than 64 energies or channels, keeps
 * the DMMA fold, which is as fast or faster there;
 * elisa_b200_plan_fold_slices() tells which engine each spectrum got.
 * ELISA_FLAG_FOLD_DMMA / env ELISA_B200_FOLD=dmma
 * selects the native FP64 DMMA fold (csrc/gemm_f64.cuh); ELISA_FLAG_FOLD_I8 / env
 * ELISA_B200_FOLD=i8 makes a spectrum the integer fold cannot take (E or C > 16384) an
 * error instead of a silent DMMA fold, overrides the 5 % rule, and switches off the guard fallback of the "_host"
 * entry point (see elisa_b200_plan_fold_guard).  Slices: bits 8-11 of flags (4..7; 0 = default 6;
 * 4 is an FP32-class variant, ~1e-6) or env ELISA_B200_FOLD_SLICES.                       */
#define ELISA_FLAG_FOLD_DMMA 4
#define ELISA_FLAG_FOLD_I8 8
#define ELISA_FLAG_NO_BALANCE 16 /* integer fold without column balancing (see elisa_b200_plan_rebalance) */
/* Calls of <= 1024 parameter vectors on a small spectrum (E * C * (1 + free parameters) <= 2^18, <= 8 parameters,
 * runtime-specialised model) are evaluated by ONE kernel per spectrum in plain FP64 (csrc/unfold_skel.cuh,
 * tiny_eval) instead of the chunk pipeline; this flag, a fold engine requested by flag or environment, or env
 * ELISA_B200_TINY=0 switch that off.  elisa_b200_plan_small_batch_kernel() tells what a spectrum got.        */
#define ELISA_FLAG_NO_SMALL_BATCH_KERNEL 32
#define ELISA_FLAG_FOLD_SLICES_SHIFT 8
/* Slices of the TRANSPOSED fold (W . R^T, the gradient side): bits 12-15 of flags (4..7; 0 =
 * default) or env ELISA_B200_FOLD_SLICES_BWD.  The gradient contraction tolerates one digit
 * less than the statistic (profiles/r02_slices.md), so the default is 5 there: 15 digit
 * products instead of 21.                                                               */
#define ELISA_FLAG_FOLD_SLICES_BWD_SHIFT 12
#define ELISA_FOLD_SLICES_DEFAULT 6
#define ELISA_FOLD_SLICES_BWD_DEFAULT 6

typedef struct ElisaPlan ElisaPlan;

/* Uploads and owns every replicated constant (grids, responses, observed data, the
 * component program) and the workspace.  Replaces the closure construction of
 * likelihood.py:184-196 etc. (jnp.array(...) of each FixedData field).              */
int elisa_b200_plan_create(const ElisaPlanDesc* desc, ElisaPlan** out);
void elisa_b200_plan_destroy(ElisaPlan* plan);

/* counts_on / counts_off: NULL, or per-replica observed data [N, sum_s C_s]
 * (spectra concatenated along the last axis in plan order) replacing on_counts /
 * off_counts -- what substituting '{name}_Non_data' / '{name}_Noff_data' does
 * upstream (infer/helper.py:608-613).                                               */
int elisa_b200_loglike_dev(ElisaPlan* plan, const double* theta, const double* counts_on,
                           const double* counts_off, int64_t n, double* loglike, void* cuda_stream);
int elisa_b200_loglike_grad_dev(ElisaPlan* plan, const double* theta, const double* counts_on,
                                const double* counts_off, int64_t n, double* loglike, double* grad,
                                void* cuda_stream);
/* per-channel sites of spectrum `spectrum` (likelihood.py:206-225): any output may be
 * NULL.  model_on/ll_on/model_off/ll_off/rate are [N, C_s].                         */
int elisa_b200_sites_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, const double* counts_on,
                         const double* counts_off, int64_t n, double* rate, double* model_on, double* ll_on,
                         double* model_off, double* ll_off, void* cuda_stream);
/* the unfolded model (N, E_s) of spectrum `spectrum`: CompiledModel.eval,
 * models/model.py:415-436, after the clip of likelihood.py:204                      */
int elisa_b200_unfold_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, int64_t n, double* unfold,
                          void* cuda_stream);

/* Gauss-Newton normal equations of the deviance residuals r_c = sqrt(-2 l_c), the least-squares
 * problem the reference's batched re-fit hands to Levenberg-Marquardt (infer/helper.py:588-594,
 * 601-625: fit_once / batch_fit, used by boot() and ppc()).  Per parameter vector, summed over
 * all spectra of the plan:
 *   loglike[n]    = sum_c l_c                                   (= -0.5 * |r|^2)
 *   grad[n, p]    = d loglike / d theta_p                       (= -(J^T r)_p)
 *   jtj[n, p, q]  = sum_c (dr_c/dtheta_p)(dr_c/dtheta_q)        (P x P, symmetric, both halves)
 * from forward-folded tangent spectra (one forward fold per free parameter, no transposed
 * fold).  counts_on / counts_off as in loglike_dev (per-replica simulated data).
 * The first call allocates the plan's tangent workspace ((1 + P) * chunk * C doubles).     */
int elisa_b200_normal_eq_dev(ElisaPlan* plan, const double* theta, const double* counts_on, const double* counts_off,
                             int64_t n, double* loglike, double* grad, double* jtj, void* cuda_stream);

/* Host-buffer variants: H2D copy of theta (and counts), compute, D2H copy of the
 * results, synchronous.  grad may be NULL (value only).                             */
int elisa_b200_loglike_grad_host(ElisaPlan* plan, const double* theta, const double* counts_on,
                                 const double* counts_off, int64_t n, double* loglike, double* grad);

/* 1 / channel_width for the '{name}' site of sites_dev (likelihood.py:206): optional;
 * without it `rate` is source_rate itself.                                          */
int elisa_b200_plan_set_channel_width(ElisaPlan* plan, int32_t spectrum, const double* channel_width);

/* Optional device timing per kernel class: while enabled every launch is bracketed by
 * CUDA events on its stream; get_profile synchronises those events and returns the
 * summed milliseconds and launch counts since set_profiling(plan, 1).               */
#define ELISA_PROFILE_UNFOLD 0 /* component evaluation                              */
#define ELISA_PROFILE_FOLD 1   /* forward fold + statistic epilogue                 */
#define ELISA_PROFILE_TFOLD 2  /* transposed fold + gradient epilogue               */
#define ELISA_PROFILE_REDUCE 3 /* partial reductions / per-channel sites            */
#define ELISA_PROFILE_SLICE 4  /* digit slicing of U and W (sliced-integer fold only)    */
#define ELISA_PROFILE_STAT 5   /* statistic + digits of W (sliced-integer fold only)       */
#define ELISA_PROFILE_CONTRACT 6 /* gradient by contraction: tangents recomputed against G (sliced-integer fold only) */
#define ELISA_PROFILE_TINY 7   /* one-kernel evaluation of a small batch of a small spectrum (unfold + FP64 fold + statistic + gradient) */
#define ELISA_PROFILE_GATHER 8 /* peer gather: what is left of it after the last chunk (tail pushes + waiting for the peers) */
#define ELISA_PROFILE_CLASSES 9
int elisa_b200_plan_set_profiling(ElisaPlan* plan, int enable);
int elisa_b200_plan_get_profile(ElisaPlan* plan, double* ms, int64_t* launches);

/* introspection */
/* parameter vectors processed per internal pass                                    */
int64_t elisa_b200_plan_chunk(const ElisaPlan* plan);
int64_t elisa_b200_plan_workspace_bytes(const ElisaPlan* plan);
/* kernels launched by this plan since creation (all entry points)                  */
int64_t elisa_b200_plan_launch_count(const ElisaPlan* plan);
/* Generate and compile (NVRTC, sm_100a; needs no GPU) the specialised component kernel
 * of `model`; returns the cubin size in bytes or a negative error.  source_out, if not
 * NULL, receives the generated translation unit (NUL-terminated, truncated to the cap). */
int64_t elisa_b200_specialise_check(const ElisaModelDesc* model, int32_t n_params, char* source_out,
                                    int64_t source_cap);
/* slices of the sliced-integer fold used for `spectrum` (ELISA_FLAG_FOLD_*), 0 if it folds
 * with FP64 DMMA, -1 on a bad argument; _bwd: of its transposed fold                     */
int elisa_b200_plan_fold_slices(const ElisaPlan* plan, int32_t spectrum);
int elisa_b200_plan_fold_slices_bwd(const ElisaPlan* plan, int32_t spectrum);
/* 1 when calls of up to 1024 parameter vectors evaluate this spectrum in ONE kernel (model in dual numbers,
 * FP64 fold in forward mode, statistic, gradient; small spectra with a runtime-specialised model: the NUTS
 * shape), 0 when they run the chunk pipeline.  Environment ELISA_B200_TINY=0 turns the path off.            */
int elisa_b200_plan_small_batch_kernel(const ElisaPlan* plan, int32_t spectrum);
/* int8 digit products the integer fold EXECUTES per FP64 multiply-add of the dense contraction,
 * averaged over the (column tile, k-block) pairs of the forward (out[0]) and the transposed
 * (out[1]) operand: slices (slices + 1) / 2 for a response without structure (21 for 6 digits),
 * less when k-blocks hold no non-zeros at all (band-sparse responses) or their most significant
 * digit planes are zero (everything outside the redistribution core of an ordinary response),
 * which the kernel skips.  Depends on the balancing envelopes: call after a first evaluation.
 * Synchronises the device.  0 for a spectrum that folds with DMMA.                          */
int elisa_b200_plan_fold_products(ElisaPlan* plan, int32_t spectrum, double out[2]);
/* Accuracy guard of the sliced-integer fold.  Its error is bounded relative to (largest
 * element of the balanced row of U) x (sum over the balanced column of the response), see
 * elisa_b200_plan_rebalance below; rows far from the spectrum the envelopes were taken on, or
 * data with genuinely starved channels (counts where the model predicts ~nothing), lose
 * accuracy.  While evaluating, the statistic kernel estimates to first order the log-likelihood
 * error this may have caused, sum_c |dl/dx_c| * 2^(-8 slices) * max_e|u'| * sum_e|t_e R[e,c]|,
 * and counts the rows where it exceeds 3e-11 * max(|loglike|, 1) (env ELISA_B200_GUARD_TOL); on
 * the gradient side it flags rows in which a few channels tower over the rest of the balanced
 * W = dl/d(Ru) (2^(-8 slices) * C * max|W'| / sum|W'| above the same level).  This call
 * returns that count and the largest relative estimate since the last reset (synchronise the
 * stream of the evaluation first).  Rows are never flagged on the BASELINE configurations
 * (estimates <= 2e-12).  A caller that sees flagged > 0 first renews the envelopes
 * (elisa_b200_plan_rebalance) and repeats the call; if it is flagged again, it re-evaluates
 * after elisa_b200_plan_set_fold(plan, 1) -- the Python host side (elisa_b200.Likelihood,
 * fold=None) and loglike_grad_host do both themselves.                                     */
int elisa_b200_plan_fold_guard(ElisaPlan* plan, int64_t* flagged, double* worst, int reset);
/* Balancing of the integer fold.  The engine's error is relative to (largest element of a row
 * of U or W) x (sum over a row of the response operand); a steep spectrum or a few starved
 * channels would cost the small columns digits.  The first evaluation call of a plan therefore
 * takes 128 of its parameter vectors, spread evenly over the batch, as a pilot, forms
 * power-of-two envelopes t_e of U over the energies and r_c of W over the channels (each
 * relative to its peak, floored 28 bits below it), and re-slices the two response operands as
 * t_e R[e, c] and r_c R[e, c] -- on the call's stream, nothing read back; from then on U / t and
 * W / r are what gets sliced (products unchanged, exactly).  The envelopes stay until
 * elisa_b200_plan_rebalance() marks them stale (the next call re-derives them from its own
 * rows): call it when the region of parameter space changes by orders of magnitude in flux.
 * ELISA_FLAG_NO_BALANCE / env ELISA_B200_BALANCE=0 switch balancing off.                                       */
int elisa_b200_plan_rebalance(ElisaPlan* plan);

/* dmma != 0: evaluate with the FP64 DMMA fold of the same plan (both sets of operands live in a
 * plan) until switched back -- what a caller does with a flagged "_dev" call.  The "_host" entry
 * point, which synchronises anyway, consults the guard and does this itself (unless the plan was
 * created with ELISA_FLAG_FOLD_I8); plan_guard_fallbacks counts how often it did.          */
int elisa_b200_plan_set_fold(ElisaPlan* plan, int dmma);
int64_t elisa_b200_plan_guard_fallbacks(const ElisaPlan* plan);
/* counters of the guarded entry points since plan creation: out[0] calls in which the DMMA fold
 * answered for some rows, out[1] rows it answered in total, out[2] envelope renewals        */
int elisa_b200_plan_guard_stats(const ElisaPlan* plan, int64_t out[3]);
/* The fold engine on its own: D[m, n] = A[m, k] . B[n, k]^T in FP64-class accuracy on the
 * int8 tensor cores (device pointers, row-major, k <= 16384, slices 4..7).  Allocates its
 * digit images, runs on `cuda_stream`, synchronises; `ms` (nullable) receives the device
 * time of the contraction kernel alone.  A validation / benchmarking utility: the plan
 * entry points above run the same kernel with the statistic and gradient epilogues.     */
int elisa_b200_sliced_gemm_dev(const double* A, int64_t lda, const double* B, int64_t ldb, int32_t m, int32_t n,
                               int32_t k, int32_t slices, double* D, int64_t ldd, double* ms, void* cuda_stream);
/* Guarded evaluation with DEVICE pointers: like loglike_grad_dev (grad may be NULL), then
 * consults the accuracy guard -- this entry point SYNCHRONISES `cuda_stream`.  The statistic
 * kernel keeps a list of the flagged rows; only those rows are re-evaluated with the FP64 DMMA
 * fold (gathered into a compact batch, scattered back), so that a batch with a few starved rows
 * keeps the integer fold's rate.  When more than a quarter of the rows are flagged the balancing
 * envelopes are renewed from the call's own rows first (elisa_b200_plan_rebalance) and the
 * call repeated; a list overflow (> 65536 rows) falls back to the DMMA fold for the whole call.
 * fallback_rows (nullable) receives the number of rows answered by the DMMA fold.
 * loglike_grad_host does the same.                                                       */
int elisa_b200_loglike_grad_guarded_dev(ElisaPlan* plan, const double* theta, const double* counts_on,
                                        const double* counts_off, int64_t n, double* loglike, double* grad,
                                        void* cuda_stream, int64_t* fallback_rows);
/* Asynchronous view of the guard for callers that must not synchronise (a leapfrog loop
 * replaying the small-batch graph): snapshot enqueues a copy of the guard words into pinned
 * host memory on `cuda_stream` followed by an event; poll returns 1 and fills flagged / worst
 * once that copy has landed (wait != 0: block until it has), 0 if it is still in flight, and
 * resets nothing -- the counters are cumulative since the last elisa_b200_plan_fold_guard
 * reset.                                                                                  */
int elisa_b200_plan_guard_snapshot(ElisaPlan* plan, void* cuda_stream);
int elisa_b200_plan_guard_poll(ElisaPlan* plan, int wait, int64_t* flagged, double* worst);
/* Per-channel forward-mode Jacobian of spectrum `spectrum`: the folded model counts
 * x[n, c] = clip(source_counts) (likelihood.py:207-208) and ds[n, p, c] = d x[n, c] / d theta_p
 * from forward-folded tangent spectra -- what jax.jacfwd of the '{name}_Non_model' site, and
 * through dl/dx the LM residual Jacobian (infer/helper.py:588-594, 616-622), are made of.
 * dl_dx (nullable) [n, c] = d loglike_c / d x_c of the spectrum's statistic (on + off terms),
 * so that d residual_c / d theta_p = -dl_dx * ds / residual_c.  n_params <= 16.            */
int elisa_b200_jacobian_dev(ElisaPlan* plan, int32_t spectrum, const double* theta, const double* counts_on,
                            const double* counts_off, int64_t n, double* x, double* ds, double* dl_dx,
                            void* cuda_stream);
/* The row-wise sparse fold on its own: folded[n, c] = sum_k values[c, k] * unfold[n, indices[c, k]]
 * over the CSR-by-channel-row response of `spectrum` (a dense response_dense with at most 25 %
 * non-zeros is converted at plan creation) -- the BCSR product of infer/likelihood.py:177-181,205
 * without area / exposure factors.  One thread per channel row, FP64 FMA pipe (the north star's
 * band-sparse kernel).  The evaluation entry points fold band-sparse responses with the
 * block-banded DMMA kernel instead, which is faster (profiles/r02_cfg4.md); this entry point is the
 * reference implementation it was raced against and a building block for callers that hold the
 * unfolded model already.  ELISA_ERR_UNSUPPORTED for a dense response.                      */
int elisa_b200_csr_fold_dev(ElisaPlan* plan, int32_t spectrum, const double* unfold, int64_t ldu, int64_t n, double* folded,
                            int64_t ldx, void* cuda_stream);
/* Device-side simulator (infer/helper.py:221-278, simulate; :808-876, simulate_and_fit): draws n data
 * sets of spectrum `spectrum` from model counts `mean` ([mean_rows, C]; mean_rows = 1: one model for all
 * replicas, = n: one per replica -- the '{name}_Non_model' / '{name}_Noff_model' sites of
 * elisa_b200_sites_dev) into out[row * ld + col].  which = 0: the 'on' data, Poisson(mean), or
 * Normal(mean, net_errors) for chi2; which = 1: the background, Normal(mean, back_errors) for pgstat,
 * Poisson(mean) for wstat.  Counter-based generator (Philox4x32-10), one stream per (seed, spectrum,
 * array, replica, channel): reproducible and independent of the launch shape.  NumPy's stream cannot
 * be reproduced, parity is distributional.                                                  */
int elisa_b200_sample_dev(ElisaPlan* plan, int32_t spectrum, int32_t which, const double* mean, int64_t mean_rows, int64_t n,
                          uint64_t seed, double* out, int64_t ld, void* cuda_stream);
/* ---- peer gather (the N axis is sharded over the GPUs of one box, only the per-sample results are
 * gathered).  The reference's counterpart is the concatenation jax.pmap / shard_map do after the per-device work
 * (infer/helper.py:706-716).  Instead of a collective AFTER the evaluation, a plan with a gather set pushes the
 * rows of every few finished chunks into EVERY rank's copy of the full result arrays while the next chunks are
 * being computed: peer-to-peer copies on the copy engines over NVLink / NVSwitch -- no SM, no NCCL kernel that
 * would compete with the persistent fold CTAs for the SMs -- and the call ends with one flag per peer and a wait
 * for theirs.  One process per GPU; the arrays are CUDA IPC allocations:
 *   elisa_b200_peer_alloc   cudaMalloc on the current device + its IPC handle (64 bytes, to be sent to the peers
 *                           by whatever the host side uses: torch.distributed objects, MPI, a file)
 *   elisa_b200_peer_open    maps a peer's allocation into this process (peer access enabled lazily)
 *   elisa_b200_peer_close / elisa_b200_peer_free
 *   elisa_b200_plan_set_gather(plan, world, rank, ll_all, grad_all, flags, row_offset)
 *       ll_all[r] / grad_all[r]: base of rank r's FULL arrays [n_total, S] / [n_total, S, P] (grad_all may be NULL)
 *       as mapped in THIS process (entry `rank` is this rank's own allocation); flags[r]: rank r's int64[world],
 *       zero-initialised; row_offset: global index of row 0 of this rank's block.  world <= 1 switches it off.
 *   Then elisa_b200_loglike_dev / _loglike_grad_dev / _loglike_grad_guarded_dev additionally (a) copy rows
 *   [i0, i1) of their loglike / grad OUTPUT arguments to ll_all[r] + (row_offset + i0) S (and grad likewise) of
 *   every r (skipping a destination that IS the source: hand in views of the own full arrays and nothing is
 *   copied locally), (b) after the last chunk -- and after the guard's DMMA re-evaluation of flagged rows, which
 *   re-pushes the block -- write the call's sequence number into flags[r][rank] of every peer and make the
 *   stream wait until flags[rank][r] of every peer has reached it: when the call's work on the stream is done,
 *   this rank's full arrays hold every rank's rows.  Every rank must make the same sequence of gathered calls.
 *   A rank may run ahead by one call: keep two sets of arrays and alternate them if a consumer reads the
 *   results while the next call is already running (elisa_b200/parallel.py does).  The gathered calls of one plan
 *   belong on ONE stream (their pushes share a copy stream, an event pair and, on the small-batch path, a device
 *   counter).  A peer whose flag has not arrived after 60 s (ELISA_B200_GATHER_TIMEOUT_S) is reported by the NEXT
 *   gathered call (ELISA_ERR_CUDA, naming the rank) -- the wait is bounded so that a lost rank cannot hang the device. */
int elisa_b200_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char ipc_handle[64]);
int elisa_b200_peer_open(const unsigned char ipc_handle[64], void** dev_ptr);
int elisa_b200_peer_close(void* dev_ptr);
int elisa_b200_peer_free(void* dev_ptr);
int elisa_b200_plan_set_gather(ElisaPlan* plan, int32_t world, int32_t rank, void* const* ll_all, void* const* grad_all,
                               void* const* flags, int64_t row_offset);
/* Timing experiments only: bit mask of kernel classes to SKIP in the evaluation entry points
 * (1 unfold, 2 forward fold, 4 statistic, 8 transposed fold, 16 reduce, 32 the dU/dtheta stores of
 * the specialised unfold kernel); results are garbage
 * while a bit is set.  tools/power_probe.py uses it to time one class in isolation.        */
int elisa_b200_plan_set_debug(ElisaPlan* plan, int32_t skip_mask);
/* 1 if spectrum's component kernel is the runtime-specialised one, 0 if interpreted */
int elisa_b200_plan_is_specialised(const ElisaPlan* plan, int32_t spectrum);
int elisa_b200_abi_version(void);
/* sizeof of the descriptor structs as this library was compiled, for bindings that declare
 * them by hand: 0 ElisaComponentDesc, 1 ElisaNormDesc, 2 ElisaModelDesc, 3 ElisaSpectrumDesc,
 * 4 ElisaPlanDesc; -1 otherwise */
int64_t elisa_b200_sizeof(int32_t which);
/* thread-local, never NULL */
const char* elisa_b200_last_error(void);

#endif /* !__CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* ELISA_B200_H_ */
