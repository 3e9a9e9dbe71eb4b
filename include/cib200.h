/* cib200.h - C-ABI of the B200 (sm_100a) uniform-grid evaluation library.
 *
 * This is the drop-in boundary for the hot path of ConvolutionInterpolations.jl v0.17.0.
 * The reference is pure Julia and has no FFI today; every entry point below names the Julia
 * method whose body a maintainer replaces by a `ccall` (see INTEGRATION.md for the glue).
 * Paths are relative to the reference's src/.
 *
 * Conventions
 *   - plain C, no exceptions; every function returns a cib200_status (0 = OK);
 *     cib200_last_error() returns the thread-local message of the last failure.
 *   - all arrays are column-major (axis 0 contiguous) = Julia memory order.
 *   - dtype CIB200_F64 is the reference precision; CIB200_F32 is the opt-in mode (coefficients,
 *     queries and results in float; cell / table-row arithmetic stays in double).
 *   - host pointers are borrowed for the duration of the call; device memory behind a handle is
 *     owned by the library.  `stream` is a cudaStream_t (NULL = legacy default stream).
 *   - the library binds to the CUDA device that is current when a handle is created / a sweep
 *     is called (one process per GPU); there is NO CPU fallback: without a usable device every
 *     compute entry point fails with CIB200_ECUDA.
 */
#ifndef CIB200_H
#define CIB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CIB200_MAX_ITP_DIM 3
#define CIB200_MAX_SWEEP_DIM 8
#define CIB200_MAX_ND 8 /* N-D point evaluation (cib200_itpnd_*): 4..8 dimensions */
#define CIB200_N_PRE 101

typedef enum {
  CIB200_OK = 0,
  CIB200_EOOB = 1,         /* query outside the domain with extrap = Throw (extrapolate_point.jl:156-159) */
  CIB200_ENAN = 2,         /* NaN/Inf coordinate: Julia raises InexactError from floor(Int, x)          */
  CIB200_EINVAL = 3,       /* malformed arguments                                                     */
  CIB200_EUNSUPPORTED = 4, /* mode outside the accelerated path (glue must keep the Julia method)     */
  CIB200_ECUDA = 5,        /* CUDA runtime error / no device                                          */
  CIB200_ENOMEM = 6
} cib200_status;

typedef enum { CIB200_F64 = 0, CIB200_F32 = 1 } cib200_dtype;
typedef enum { CIB200_HOST = 0, CIB200_DEVICE = 1 } cib200_memkind;
/* b-series: derivatives up to the shipped tables; a-series (Keys-type, a4 is the reference's default kernel):
 * derivative 0 only (their K'' tables are not shipped: derivatives take the reference's :linear-subgrid path) */
typedef enum {
  CIB200_B5 = 0, CIB200_B7 = 1, CIB200_B9 = 2, CIB200_B11 = 3, CIB200_B13 = 4,
  CIB200_A3 = 5, CIB200_A4 = 6, CIB200_A5 = 7, CIB200_A7 = 8
} cib200_kernel;
typedef enum { CIB200_BC_DETECT = 0, CIB200_BC_POLY = 1, CIB200_BC_LINEAR = 2, CIB200_BC_QUADRATIC = 3 } cib200_bc;
typedef enum { CIB200_THROW = 0, CIB200_FLAT = 1, CIB200_LINE = 2 } cib200_extrap;
/* subgrid = how the 101-row tables are interpolated between rows: cubic Hermite (K, K'), quintic Hermite (K, K', K'')
 * - cubic_hermite_interpolation.jl:15-28, quintic_hermite_interpolation.jl:16-37 - or linearly (K only: the path of
 * top derivatives and of every non-all-cubic 3-D interpolant).  The reference's constructors generate >= 10 000 table rows
 * for :linear axes (fast_convolution_interpolation_constructor.jl:117-118): pass them through the *_create_tables /
 * cib200_resample_tables entry points; without them a :linear axis reads the 101-row shipped table (1e-4 relative away) */
typedef enum { CIB200_SG_CUBIC = 0, CIB200_SG_QUINTIC = 1, CIB200_SG_LINEAR = 2 } cib200_subgrid;

/* evaluation strategies of cib200_itp_eval (3-D only; 1-D/2-D always use the direct kernel) */
typedef enum {
  CIB200_EVAL_AUTO = 0,
  CIB200_EVAL_GATHER = 1, /* one warp resolves one query's stencil straight from global memory    */
  CIB200_EVAL_BINNED = 2  /* queries binned by coefficient brick, bricks staged in smem by TMA     */
} cib200_eval_mode;

/* One uniform fast interpolant = the numeric fields of FastConvolutionInterpolation
 * (data_structures/fast_convolution_interpolation.jl:56-102) that the eager cubic-subgrid
 * evaluators read.  Built by fast_convolution_interpolation_constructor.jl:99-174. */
typedef struct cib200_itp_desc {
  int32_t ndim;                          /* 1..3                                                    */
  int32_t dtype;                         /* cib200_dtype of coefs / queries / results               */
  int64_t coef_shape[CIB200_MAX_ITP_DIM]; /* size(itp.coefs) = n_d + 2 (eqs_d - 1)                    */
  const void *coefs;                     /* itp.coefs, ghost-padded, host (or device) memory        */
  const double *knots[CIB200_MAX_ITP_DIM]; /* itp.knots[d] (extended, length coef_shape[d]), host    */
  double h[CIB200_MAX_ITP_DIM];           /* itp.h                                                   */
  int32_t kernel[CIB200_MAX_ITP_DIM];     /* cib200_kernel per axis                                  */
  int32_t deriv[CIB200_MAX_ITP_DIM];      /* derivative order per axis, 0..max shipped               */
  int32_t extrap;                        /* cib200_extrap (etp.et)                                  */
  int32_t perdim;                        /* 0: HigherOrderKernel evaluators, 1: HigherOrderMixedKernel
                                            (only changes how the final scale is associated)        */
  int32_t subgrid[CIB200_MAX_ITP_DIM];    /* cib200_subgrid per axis (itp.subgrid, resolved by _build_subgrids);
                                            quintic: ndim <= 2 and K^(deriv+2) shipped; linear: K^(deriv)
                                            shipped; ndim == 3: all cubic or all linear                */
} cib200_itp_desc;

typedef struct cib200_itp cib200_itp; /* opaque */

/* ---- library ------------------------------------------------------------------------------ */
int32_t cib200_version(void);
const char *cib200_last_error(void);
/* devices visible to the process / select the device the next cib200_itp_create binds to (callers without a CUDA
 * binding of their own - the Julia glue - use these to build one handle per GPU for cib200_itp_eval_multi) */
int32_t cib200_device_count(void);
int32_t cib200_set_device(int32_t device);
/* number of kernel launches issued by this library since process start (bench bookkeeping) */
int64_t cib200_launch_count(void);
/* 101 x 2*eqs row-major table K^(deriv) of `kernel` (host pointer into the library image); the same
 * numbers as precomputed_kernels/kernel_<k>.jl (KERNEL_PRE_<k>_d<deriv>; deriv = -1: the antiderivative table
 * KERNEL_PRE_<k>_i1). NULL if not shipped. */
const double *cib200_kernel_table(int32_t kernel, int32_t deriv, int32_t *eqs_out);

/* ---- subsystem 1: ghost-padded coefficient grid, built on the HOST -------------------------- */
/* replaces create_convolutional_coefs (convolution_coefs/create_convolutional_coefs.jl:141-160).
 * values: n[0] x ... x n[ndim-1]; coefs_out: (n[d] + 2 (eqs_d - 1)) per axis, caller-allocated. */
int32_t cib200_build_coefs(int32_t ndim, int32_t dtype, const int64_t *n, const void *values,
                           const double *h, const int32_t *kernel, const int32_t *bc_left,
                           const int32_t *bc_right, void *coefs_out);
/* The same on the DEVICE: one streaming pass per axis (copy the lines into the padded array, append
 * the eqs - 1 ghosts per side from the line mean and its first / last 8 values).  values / coefs_out are both of
 * `memkind` (HOST: staged through device memory; DEVICE: asynchronous on `stream`).  A device-resident coefs_out can
 * be passed as cib200_itp_desc.coefs (cib200_itp_create copies from host or device memory).  Ghost values equal the
 * host build's up to the association of the line mean (<= 1e-15 relative). */
int32_t cib200_build_coefs_device(int32_t ndim, int32_t dtype, const int64_t *n, const void *values,
                                  const double *h, const int32_t *kernel, const int32_t *bc_left,
                                  const int32_t *bc_right, void *coefs_out, int32_t memkind, void *stream);
/* replaces extend_vector (convolution/extend_vector.jl:10-18); out has n + 2 n_extra entries */
int32_t cib200_extend_knots(const double *knots, int64_t n, int32_t n_extra, double *out);

/* ---- subsystem 2+3: batched point evaluation ------------------------------------------------ */
/* uploads coefficient grid, knots and tables once (replaces nothing: called by the glue right after
 * _build_fast_uniform_convolution, fast_convolution_interpolation_constructor.jl:99-174). */
int32_t cib200_itp_create(const cib200_itp_desc *desc, cib200_itp **handle);
/* The same for an interpolant with :linear-subgrid axes as the reference builds them: for such an axis the constructor
 * raises the table length to max(precompute, 10_000) rows (fast_convolution_interpolation_constructor.jl:117-118) and
 * generates K^(deriv) from exact rationals (precompute_kernel_and_range.jl:40-157).  tables[d]: itp.kernel_pre[d] of a
 * :linear axis (n_pre[d] x 2 eqs doubles in Julia memory order, i.e. column-major; host memory, copied), NULL for the
 * other axes and for :linear axes that should read the 101-row shipped table.  Large tables stay in global memory and
 * the tile-binned strategies are not used for such interpolants. */
int32_t cib200_itp_create_tables(const cib200_itp_desc *desc, const double *const *tables, const int32_t *n_pre,
                                 cib200_itp **handle);
int32_t cib200_itp_destroy(cib200_itp *handle);
/* copies the handle's ghost-padded grid to host memory (fills itp.coefs when the grid was built on the device) */
int32_t cib200_itp_get_coefs(cib200_itp *handle, void *coefs_out);
int32_t cib200_itp_set_eval_mode(cib200_itp *handle, int32_t mode);

/* itp.(xs...) : replaces the per-element loop over
 * (etp::ConvolutionExtrapolation)(x::Vararg{Number,N}) (convolution_extrapolate/extrapolation_operator.jl:31-63)
 * -> extrapolate_point (extrapolate_point.jl:41-160) -> eager evaluators
 * (convolution_fast_interpolation_eager_{1d,2d,3d}.jl, ..._perdim_{2d,3d}.jl).
 * x: ndim SoA coordinate arrays of nq elements, out: nq results, all of `memkind`.
 * first_bad (may be NULL): lowest query index that raised CIB200_EOOB / CIB200_ENAN, else -1.
 * With memkind = DEVICE the call is asynchronous on `stream` unless first_bad != NULL or
 * extrap = Throw (the status needs the device's answer). */
int32_t cib200_itp_eval(cib200_itp *handle, int64_t nq, const void *const *x, void *out,
                        int32_t memkind, void *stream, int64_t *first_bad);

/* The same over several GPUs of ONE process (host buffers): handles[i] was created with device i current and holds a
 * replica of the grid; the batch is split into n_handles contiguous slices evaluated concurrently, one host thread and
 * one H2D / kernel / D2H pipeline per device, no collective.  first_bad / status: the lowest offending
 * query of the whole batch, as in a sequential pass. */
int32_t cib200_itp_eval_multi(cib200_itp *const *handles, int32_t n_handles, int64_t nq,
                              const void *const *x, void *out, int64_t *first_bad);

/* itp(xs, ys, zs) with vectors: Cartesian product
 * ((etp)(x::Vararg{AbstractArray,N}), extrapolation_operator.jl:66-99). out: n[0] x ... column-major */
int32_t cib200_itp_eval_grid(cib200_itp *handle, const int64_t *n, const void *const *x, void *out,
                             int32_t memkind, void *stream, int64_t *first_bad);

/* ---- N >= 4 dimensions: the reference's HigherDimension evaluators ------------------------------
 * itp.(xs...) for 4..CIB200_MAX_ND axes: replaces the per-element loop over the N-D functors
 *   convolution_fast_interpolation_eager_nd.jl:97-140   (perdim = 0: one kernel, one derivative order)
 *   convolution_fast_interpolation_perdim_nd.jl:4-36    (perdim = 1: per-axis kernels / derivative orders)
 * behind extrapolation_operator.jl:31-99 / extrapolate_point.jl:98-160.  For N >= 4 _build_subgrids
 * (fast_convolution_interpolation_constructor.jl:456-494) resolves every axis to the :linear subgrid; with
 * precompute = 101 the tables are the shipped K^(deriv).  coefs: ghost-padded grid (cib200_build_coefs takes up to
 * CIB200_MAX_SWEEP_DIM axes), knots[d]: extended knot vector of length coef_shape[d], all host memory, copied. */
typedef struct cib200_itpnd cib200_itpnd; /* opaque */
int32_t cib200_itpnd_create(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                            const double *const *knots, const double *h, const int32_t *kernel,
                            const int32_t *deriv, int32_t extrap, int32_t perdim, cib200_itpnd **handle);
/* with the generated tables of the (all :linear) axes, as cib200_itp_create_tables; perdim = 0 reads tables[0] on every axis */
int32_t cib200_itpnd_create_tables(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                                   const double *const *knots, const double *h, const int32_t *kernel,
                                   const int32_t *deriv, int32_t extrap, int32_t perdim, const double *const *tables,
                                   const int32_t *n_pre, cib200_itpnd **handle);
int32_t cib200_itpnd_destroy(cib200_itpnd *handle);
/* same contracts as cib200_itp_eval / cib200_itp_eval_grid */
int32_t cib200_itpnd_eval(cib200_itpnd *handle, int64_t nq, const void *const *x, void *out,
                          int32_t memkind, void *stream, int64_t *first_bad);
int32_t cib200_itpnd_eval_grid(cib200_itpnd *handle, const int64_t *n, const void *const *x, void *out,
                               int32_t memkind, void *stream, int64_t *first_bad);

/* ---- antiderivatives: derivative = -1 on every axis (or on some: cib200_itg_create_tables), 1..3 dimensions
 *       ---------------------------------------------------------------------------
 * itp.(xs...) / itp(xs, ys, zs) for the FastIntegralOrder functors
 *   convolution_fast_interpolation/convolution_fast_integration_1d.jl:17-38, _2d.jl:18-80, _3d.jl:4-100
 *   (+ eval_sg_kt, eval_subgrid_helpers.jl:22-32) behind extrapolation_operator.jl:31-99 / extrapolate_point.jl:41-160.
 * cib200_itg_create replaces _build_tails (convolution/fast_convolution_interpolation_constructor.jl:176-364) and
 * _compute_left_values (:414-448): desc is the same cib200_itp_desc with deriv[d] = -1 on every axis, dtype = F64,
 * subgrid cubic or quintic (tables K~ = KERNEL_PRE_<k>_i1, K, K'); coefs may be host or device memory.  The prefix-sum
 * arrays are built on the device; the reference's suffix-sum arrays are identically zero wherever its evaluators read
 * them and are not materialised (cib200_integral.cu).  Results are anchored to zero at the first interior knot of each
 * axis (itp.anchor).  N >= 4 (every axis :linear on generated K~ tables): CIB200_EUNSUPPORTED. */
typedef struct cib200_itg cib200_itg; /* opaque */
int32_t cib200_itg_create(const cib200_itp_desc *desc, cib200_itg **handle);
/* MIXED interpolants (FastMixedIntegralOrder; convolution_fast_mixed_1_integral_dim.jl:11-131,
 * convolution_fast_mixed_2_integral_dims.jl): deriv[d] = -1 on one or two axes, >= 0 on the others, N <= 3.  The
 * non-integral axes are on the :linear subgrid (_build_subgrids, constructor.jl:474-491: desc.subgrid[d] = LINEAR) and
 * read generated tables: tables[d] = itp.kernel_pre[d] (n_pre[d] x 2 eqs doubles, Julia memory order, host), NULL on
 * the integral axes (NULL on a :linear axis: the 101-row shipped K^(deriv), not what the reference builds).
 * tables == NULL with every deriv[d] = -1 is cib200_itg_create. */
int32_t cib200_itg_create_tables(const cib200_itp_desc *desc, const double *const *tables, const int32_t *n_pre,
                                 cib200_itg **handle);
int32_t cib200_itg_destroy(cib200_itg *handle);
/* same contracts as cib200_itp_eval / cib200_itp_eval_grid */
int32_t cib200_itg_eval(cib200_itg *handle, int64_t nq, const void *const *x, void *out, int32_t memkind,
                        void *stream, int64_t *first_bad);
int32_t cib200_itg_eval_grid(cib200_itg *handle, const int64_t *n, const void *const *x, void *out,
                             int32_t memkind, void *stream, int64_t *first_bad);
/* copies to host memory: what = 0: itp.left_values[index + 1] (coef_shape[index] doubles); what = 1: the prefix array
 * saturated on the axes of the bit mask `index` (padded grid size): 1-D mask 1 = tail1_left[1]; 2-D 1, 2 =
 * tail1_left[1], tail1_left[2], 3 = tail2_ll; 3-D 1, 2, 4 = tail1_left[1..3] (= tail3_face_l), 6, 5, 3 =
 * tail3_edge_ll[1..3], 7 = tail3_corner_lll */
int32_t cib200_itg_get(cib200_itg *handle, int32_t what, int32_t index, void *out);

/* ---- antiderivative / mixed interpolants in N >= 4 dimensions (the remainder) --------------
 * itp.(xs...) / itp(xs...) for
 *   convolution_fast_interpolation/convolution_fast_integration_nd.jl:1-28          derivative = -1 on every axis
 *   convolution_fast_interpolation/convolution_fast_mixed_3_integral_dims.jl:8-225  exactly three integral axes
 *   convolution_fast_interpolation/convolution_fast_mixed_N_integral_dims.jl:15-82  four or more integral axes, mixed
 * behind extrapolation_operator.jl:31-99.  Float64; every axis is on the :linear subgrid (_build_subgrids,
 * fast_convolution_interpolation_constructor.jl:456-494): tables[d] = itp.kernel_pre[d] (n_pre[d] x 2 eqs doubles, Julia
 * memory order, host; NULL: the 101-row shipped K~ / K^(deriv)), deriv[d] = -1 marks the integral axes.
 * sum_extent[d]: how many leading (padded) entries of an integral axis the functor visits - itp.domain_size[d] for the
 * every-axis-integral functor, whose loops run over 1:domain_size[d] (convolution_fast_integration_nd.jl:20), size(coefs, d)
 * for the other two (NULL: size(coefs, d) everywhere).  left_values are recomputed by the library
 * (_compute_left_values, constructor.jl:414-448); the tail arrays of the three-axis functor are not needed: the kernel
 * evaluates the sum they accelerate directly (one CTA per query, O(grid) work - what the other two functors do anyway). */
typedef struct cib200_itgnd cib200_itgnd; /* opaque */
int32_t cib200_itgnd_create(int32_t ndim, const int64_t *coef_shape, const double *coefs, const double *const *knots,
                            const double *h, const int32_t *kernel, const int32_t *deriv, int32_t extrap,
                            const double *const *tables, const int32_t *n_pre, const int64_t *sum_extent,
                            cib200_itgnd **handle);
int32_t cib200_itgnd_destroy(cib200_itgnd *handle);
/* same contracts as cib200_itp_eval / cib200_itp_eval_grid */
int32_t cib200_itgnd_eval(cib200_itgnd *handle, int64_t nq, const void *const *x, void *out, int32_t memkind,
                          void *stream, int64_t *first_bad);
int32_t cib200_itgnd_eval_grid(cib200_itgnd *handle, const int64_t *n, const void *const *x, void *out,
                               int32_t memkind, void *stream, int64_t *first_bad);

/* Measurement hook (bench.py's roofline): with profiling enabled the library brackets the passes of every evaluation
 * with CUDA events on the launching stream; cib200_itp_last_pass_ms then returns, for the most recent call, the
 * device time in ms of {binning passes, cell sort, evaluation kernel, spill kernel} (gather strategy: {0, 0, kernel, 0}).
 * No counterpart in the reference. */
int32_t cib200_itp_set_profiling(cib200_itp *handle, int32_t enable);
int32_t cib200_itp_last_pass_ms(cib200_itp *handle, float *ms4);
/* which evaluation kernel the most recent 3-D point-evaluation call on this handle ran (cib200_strategy below;
 * host-buffer calls: the strategy of the last chunk).  Measurement / test hook, no counterpart in the reference. */
typedef enum {
  CIB200_STRATEGY_NONE = 0,
  CIB200_STRATEGY_GATHER = 1,    /* eval_thread_kernel / eval_warp3d_kernel straight from global memory             */
  CIB200_STRATEGY_BINNED = 2,    /* tile binning + eval_binned_kernel (TMA-staged bricks)                            */
  CIB200_STRATEGY_CELLRUN = 3,   /* + cell sort, eval_cellrun_kernel (register-resident stencils)                    */
  CIB200_STRATEGY_LANEQUERY = 4, /* + cell sort, eval_lanequery_kernel (experiments builds only)                     */
  CIB200_STRATEGY_CELLMMA = 5,   /* + cell sort with group list, eval_cellmma_kernel (FP64 tensor cores, DMMA)       */
  CIB200_STRATEGY_CELLMMA_BLOCK = 6 /* the same with rows from 2 x 2 x 2 cell blocks, eval_celloct_kernel (sparser batches) */
} cib200_strategy;
int32_t cib200_itp_last_strategy(cib200_itp *handle);

/* bytes of device scratch the BINNED strategy needs for nq queries (0 for the other modes) */
int64_t cib200_itp_scratch_bytes(cib200_itp *handle, int64_t nq);

/* ---- subsystem 4: separable axis sweeps ------------------------------------------------------ */
/* replaces convolution_resample_internal (convolution/convolution_resample.jl:240-386), cubic subgrid.
 * values: n_in[0] x ...; out: n_out[0] x ...; knots_in[d] / knots_out[d]: HOST arrays (always). */
int32_t cib200_resample(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out,
                        const void *values, void *out, const double *const *knots_in,
                        const double *const *knots_out, const int32_t *kernel, const int32_t *deriv,
                        const int32_t *bc_left, const int32_t *bc_right, int32_t memkind, void *stream);

/* the same with a per-axis subgrid (cib200_subgrid; quintic branch convolution_resample.jl:333-349); subgrid == NULL
 * means cubic on every axis */
int32_t cib200_resample_sg(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out,
                           const void *values, void *out, const double *const *knots_in,
                           const double *const *knots_out, const int32_t *kernel, const int32_t *deriv,
                           const int32_t *bc_left, const int32_t *bc_right, const int32_t *subgrid,
                           int32_t memkind, void *stream);

/* the same with the generated tables of :linear axes: the kw method raises their length to max(precompute, 10_000)
 * rows (convolution_resample.jl:228-229).  tables[d]: kernel_pre of a :linear axis (n_pre[d] x 2 eqs doubles, Julia
 * memory order, host), NULL for the other axes; tables == NULL: every axis on the 101-row shipped tables */
int32_t cib200_resample_tables(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out,
                               const void *values, void *out, const double *const *knots_in,
                               const double *const *knots_out, const int32_t *kernel, const int32_t *deriv,
                               const int32_t *bc_left, const int32_t *bc_right, const int32_t *subgrid,
                               const double *const *tables, const int32_t *n_pre, int32_t memkind, void *stream);

/* replaces convolution_smooth (convolution/convolution_gaussian_smoothing.jl:45-126).
 * h[d] = knots[d][2] - knots[d][1]. Fails with CIB200_EINVAL when the half-width E >= n[d] (:50-56). */
int32_t cib200_smooth(int32_t ndim, int32_t dtype, const int64_t *n, const void *values, void *out,
                      const double *h, double B, int32_t memkind, void *stream);
int32_t cib200_smooth_halfwidth(double B);

/* ---- pinned host memory helpers for the glue (GC-rooted Julia arrays can be registered) ------- */
int32_t cib200_host_register(void *ptr, size_t bytes);
int32_t cib200_host_unregister(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* CIB200_H */
