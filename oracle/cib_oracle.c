/* cib_oracle.c - CPU ORACLE (test infrastructure, NOT product code). See cib_oracle.h.
 *
 * Literal restatement of the reference's uniform-grid hot path.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/src).  Compile with
 * -ffp-contract=off: the reference (Julia without @fastmath) never contracts a*b+c.
 *
 * Known places where bit-equality with Julia is not knowable from the source and the oracle
 * makes a stated choice (all at the 1-ulp-of-a-term level, far below the 1e-12 budget):
 *   - sum(slice)/n            Julia mapreduce is @simd-reassociated; here: sequential sum
 *   - mul!(ghost, M, y)       OpenBLAS gemv order unspecified; here: sequential over columns
 *   - x^n for n>=4, exp       Julia's own pow_body / exp; here: pow_body restated, libm exp
 */
#include "cib_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "cib200_tables.inc" /* generated reference constants, one copy in the tree (see Makefile) */

#define N_PRE 101

static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int orc_get_threads(void) { return g_threads; }

/* data_structures/degree_to_equations.jl:13-25 */
int orc_eqs(int kernel) {
  switch (kernel) {
    case ORC_K_B5: return 5;
    case ORC_K_B7: return 6;
    case ORC_K_B9: return 7;
    case ORC_K_B11: return 8;
    case ORC_K_B13: return 9;
    case ORC_K_A3: return 2;
    case ORC_K_A4: return 3;
    case ORC_K_A5: return 3;
    case ORC_K_A7: return 4;
    default: return -1;
  }
}

int orc_max_table_derivative(int kernel) {
  switch (kernel) {
    case ORC_K_B5: return 3;
    case ORC_K_B7: return 4;
    case ORC_K_B9: return 5;
    case ORC_K_B11: return 6;
    case ORC_K_B13: return 6;
    case ORC_K_A3: case ORC_K_A4: case ORC_K_A5: case ORC_K_A7: return 1;
    default: return -1;
  }
}

/* precomputed_kernels/precomputed_kernel_tables.jl:18-249: kernel_pre = K^(d), kernel_d1_pre = K^(d+1) */
const double *orc_table(int kernel, int d) {
  static const double *const b5[] = {CIB_TAB_b5_d0, CIB_TAB_b5_d1, CIB_TAB_b5_d2, CIB_TAB_b5_d3};
  static const double *const b7[] = {CIB_TAB_b7_d0, CIB_TAB_b7_d1, CIB_TAB_b7_d2, CIB_TAB_b7_d3,
                                     CIB_TAB_b7_d4};
  static const double *const b9[] = {CIB_TAB_b9_d0, CIB_TAB_b9_d1, CIB_TAB_b9_d2,
                                     CIB_TAB_b9_d3, CIB_TAB_b9_d4, CIB_TAB_b9_d5};
  static const double *const b11[] = {CIB_TAB_b11_d0, CIB_TAB_b11_d1, CIB_TAB_b11_d2, CIB_TAB_b11_d3,
                                      CIB_TAB_b11_d4, CIB_TAB_b11_d5, CIB_TAB_b11_d6};
  static const double *const b13[] = {CIB_TAB_b13_d0, CIB_TAB_b13_d1, CIB_TAB_b13_d2, CIB_TAB_b13_d3,
                                      CIB_TAB_b13_d4, CIB_TAB_b13_d5, CIB_TAB_b13_d6};
  static const double *const a3[] = {CIB_TAB_a3_d0, CIB_TAB_a3_d1};
  static const double *const a4[] = {CIB_TAB_a4_d0, CIB_TAB_a4_d1};
  static const double *const a5[] = {CIB_TAB_a5_d0, CIB_TAB_a5_d1};
  static const double *const a7[] = {CIB_TAB_a7_d0, CIB_TAB_a7_d1};
  if (d < 0 || d > orc_max_table_derivative(kernel)) return NULL;
  switch (kernel) {
    case ORC_K_A3: return a3[d];
    case ORC_K_A4: return a4[d];
    case ORC_K_A5: return a5[d];
    case ORC_K_A7: return a7[d];
    case ORC_K_B5: return b5[d];
    case ORC_K_B7: return b7[d];
    case ORC_K_B9: return b9[d];
    case ORC_K_B11: return b11[d];
    case ORC_K_B13: return b13[d];
    default: return NULL;
  }
}

/* ---------------------------------------------------------------------------------------------
 * Julia Base: ^(x::Float64, n::Integer) -> pow_body (base/math.jl, Julia >= 1.8), restated.
 * n == 0 -> 1 ; n == 3 -> x*x*x ; otherwise compensated power by squaring.
 * ------------------------------------------------------------------------------------------- */
static double jl_pow_int(double x, int64_t n) {
  if (n == 0) return 1.0;
  double y = 1.0, xnlo = -0.0, ynlo = 0.0;
  if (n == 3) return x * x * x;
  if (n < 0) {
    double rx = 1.0 / x;
    if (n == -2) return rx * rx;
    if (isfinite(x)) xnlo = -fma(x, rx, -1.0) * rx;
    x = rx;
    n = -n;
  }
  while (n > 1) {
    if (n & 1) {
      double err = fma(y, xnlo, x * ynlo);
      double p = x * y;
      ynlo = fma(x, y, -p);
      y = p;
      ynlo += err;
    }
    double err = x * 2 * xnlo;
    double p = x * x;
    xnlo = fma(x, x, -p);
    x = p;
    xnlo += err;
    n >>= 1;
  }
  double err = fma(y, xnlo, x * ynlo);
  if (isfinite(x) && isfinite(err)) return fma(x, y, err);
  return x * y;
}

/* ---------------------------------------------------------------------------------------------
 * Ghost matrices: convolution_coefs/polynomial_boundary.jl:29-131 (POLYNOMIAL_GHOST_COEFFS),
 * :157-179 (LINEAR/QUADRATIC).  Every row j is the Lagrange extrapolation to x = -j from the
 * nodes 0..m-1 (m = 2 linear, 3 quadratic, 4 for the a-series, 6 for b5, 8 for b7..b13); regenerated here in exact
 * integer arithmetic and compared with the reference's literals in tests/test_oracle_kat.py.
 * ------------------------------------------------------------------------------------------- */
static void lagrange_row(int m, int j, double *row) {
  for (int k = 0; k < m; ++k) {
    int64_t num = 1, den = 1;
    for (int i = 0; i < m; ++i) {
      if (i == k) continue;
      num *= (int64_t)(-j - i);
      den *= (int64_t)(k - i);
    }
    row[k] = (double)(num / den); /* exact: the quotient is an integer */
  }
}

/* minimum_polynomial_bc_points, convolution_interpolation_function.jl:183-195 */
static int poly_ncol(int kernel) { return (kernel >= ORC_K_A3 && kernel <= ORC_K_A7) ? 4 : (kernel == ORC_K_B5 ? 6 : 8); }

/* exported for the KAT test: fills rows x ncol (row-major), returns ncol */
int orc_ghost_matrix(int which /*0 poly,1 linear,2 quadratic*/, int kernel, int rows, double *out) {
  int m = which == 1 ? 2 : which == 2 ? 3 : poly_ncol(kernel);
  for (int j = 1; j <= rows; ++j) lagrange_row(m, j, out + (size_t)(j - 1) * m);
  return m;
}

/* One side of one boundary line.
 * convolution_coefs/create_convolutional_coefs.jl:232-299 (left), :302-369 (right),
 * fill_ghost_points_polynomial! :408-441, fill_ghost_points_gaussian! :443-494.
 * s = mean-centred slice in natural (left-to-right) order, n points.  gv[j-1] = ghost j. */
static void ghost_side(const double *s, int64_t n, double y_mean, double h, int eqs, int kernel,
                       int bc, int gaussian, int right, double *gv) {
  const int ng = eqs - 1;
  if (gaussian) {
    /* :443-494 ; n_fit = min(max(6, eqs ÷ 3), n) */
    int64_t n_fit = eqs / 3 > 6 ? eqs / 3 : 6;
    if (n_fit > n) n_fit = n;
    double sx = 0, sy = 0, sxx = 0, sxy = 0;
    for (int64_t k = 1; k <= n_fit; ++k) {
      double xk = (double)k;
      double yk = right ? s[n - k] : s[k - 1];
      sx += xk;
      sy += yk;
      sxx += xk * xk;
      sxy += xk * yk;
    }
    double nf = (double)n_fit;
    double denom = nf * sxx - sx * sx;
    double b = (nf * sxy - sx * sy) / denom;
    double a = (sy - b * sx) / nf;
    for (int k = 1; k <= ng; ++k) {
      double y_temp = a + b * (double)(1 - k) - y_mean; /* :463 */
      gv[k - 1] = y_mean + y_temp;                      /* :469 */
    }
    return;
  }
  const int n_interior = poly_ncol(kernel); /* size(ghost_matrix, 2), :225 */
  const int few_points = n < n_interior;     /* :226 */
  int use_poly;
  if (few_points || !(bc == ORC_BC_DETECT || bc == ORC_BC_POLY)) {
    use_poly = 0; /* :247-248 */
  } else {
    const int ns = n_interior;
    const double *p = right ? s + (n - ns) : s; /* last / first ns points, natural order */
    double d1[8], d2[8];
    double mn1 = INFINITY, mx1 = -INFINITY;
    for (int d = 1; d <= ns - 1; ++d) {
      double v = p[d] - p[d - 1];
      d1[d - 1] = v;
      mn1 = fmin(mn1, v);
      mx1 = fmax(mx1, v);
    }
    int periodic = (mn1 * mx1) / (h * h) < -1.0 / 10; /* :259 */
    double mn2 = INFINITY, mx2 = -INFINITY;
    for (int d = 1; d <= ns - 2; ++d) {
      double v = d1[d] - d1[d - 1];
      d2[d - 1] = v;
      mn2 = fmin(mn2, v);
      mx2 = fmax(mx2, v);
    }
    int high_curv = fabs(mx2 - mn2) / h > 1.0 / 2; /* :268 */
    int high_d3 = 0;
    if (ns > 4) {
      double mn3 = INFINITY, mx3 = -INFINITY;
      for (int d = 1; d <= ns - 3; ++d) {
        double v = d2[d] - d2[d - 1];
        mn3 = fmin(mn3, v);
        mx3 = fmax(mx3, v);
      }
      high_d3 = fabs(mx3 - mn3) / h > 1.0 / 2; /* :277 */
    }
    use_poly = (bc == ORC_BC_POLY) || (bc == ORC_BC_DETECT && !periodic && !high_curv && !high_d3);
  }
  int m; /* matrix columns */
  if (use_poly) m = n_interior;                        /* :286-287 */
  else if (bc == ORC_BC_LINEAR) m = 2;                 /* :290-292 */
  else if (bc == ORC_BC_QUADRATIC) m = 3;
  else m = 2;                                          /* :293-295 fallback -> linear */
  double y[8], row[8];
  for (int k = 1; k <= m; ++k) y[k - 1] = right ? s[n - k] : s[k - 1]; /* :428-430 reversal */
  for (int j = 1; j <= ng; ++j) {
    lagrange_row(m, j, row);
    double acc = 0.0;
    for (int k = 0; k < m; ++k) acc += row[k] * y[k];
    gv[j - 1] = y_mean + acc; /* :424 / :438 */
  }
}

/* one full line: both sides.  line0 -> first interior element, stride in elements.
 * ghost j is written at first_interior - (j + shift) / last_interior + (j + shift):
 * shift = 0 for the interpolant and resample; shift = 1 reproduces convolution_smooth's
 * buffer (padding E instead of E-1 => get_boundary_indices lands one slot outside the
 * interior, create_convolutional_coefs.jl:77-81 with convolution_gaussian_smoothing.jl:83-105). */
static void ghost_fill_line(double *line0, int64_t stride, int64_t n, double h, int eqs, int kernel,
                            int bc_l, int bc_r, int gaussian, int shift, double *slice) {
  double gv[32 + 64];
  double *gvp = gv;
  double *heap = NULL;
  if (eqs - 1 > 96) { heap = (double *)malloc(sizeof(double) * (size_t)(eqs - 1)); gvp = heap; }
  for (int side = 0; side < 2; ++side) {
    for (int64_t j = 0; j < n; ++j) slice[j] = line0[j * stride]; /* :233-239 / :303-309 */
    double sum = 0.0;
    for (int64_t j = 0; j < n; ++j) sum += slice[j];
    double y_mean = sum / (double)n; /* :242 */
    for (int64_t j = 0; j < n; ++j) slice[j] -= y_mean;
    ghost_side(slice, n, y_mean, h, eqs, kernel, side ? bc_r : bc_l, gaussian, side, gvp);
    for (int j = 1; j <= eqs - 1; ++j) {
      if (side == 0) line0[-(int64_t)(j + shift) * stride] = gvp[j - 1];
      else line0[(n - 1 + j + shift) * stride] = gvp[j - 1];
    }
  }
  free(heap);
}

/* create_convolutional_coefs.jl:141-160 */
int orc_create_coefs(int ndim, const int64_t *n, const double *vs, const double *h,
                     const int32_t *kernel, const int32_t *bc_left, const int32_t *bc_right,
                     double *c) {
  if (ndim < 1 || ndim > ORC_MAXDIM) return ORC_EARG;
  int eqs[ORC_MAXDIM];
  int64_t cs[ORC_MAXDIM], cstride[ORC_MAXDIM];
  int64_t ctot = 1, vtot = 1, nmax = 0;
  for (int d = 0; d < ndim; ++d) {
    eqs[d] = orc_eqs(kernel[d]);
    if (eqs[d] < 0) return ORC_EARG;
    cs[d] = n[d] + 2 * (eqs[d] - 1);
    cstride[d] = ctot;
    ctot *= cs[d];
    vtot *= n[d];
    if (n[d] > nmax) nmax = n[d];
  }
  memset(c, 0, sizeof(double) * (size_t)ctot);
  /* interior copy :143-145 */
  for (int64_t lin = 0; lin < vtot; ++lin) {
    int64_t rem = lin, off = 0;
    for (int d = 0; d < ndim; ++d) {
      int64_t id = rem % n[d];
      rem /= n[d];
      off += (id + eqs[d] - 1) * cstride[d];
    }
    c[off] = vs[lin];
  }
  /* axis by axis, over the FULL padded extent of the other axes (:77-81, :151-158) */
  for (int fd = 0; fd < ndim; ++fd) {
    int64_t nlines = ctot / cs[fd];
#pragma omp parallel num_threads(g_threads)
    {
      double *slice = (double *)malloc(sizeof(double) * (size_t)nmax);
#pragma omp for schedule(static)
      for (int64_t ln = 0; ln < nlines; ++ln) {
        int64_t rem = ln, off = 0;
        for (int d = 0; d < ndim; ++d) {
          if (d == fd) continue;
          int64_t id = rem % cs[d];
          rem /= cs[d];
          off += id * cstride[d];
        }
        off += (int64_t)(eqs[fd] - 1) * cstride[fd];
        ghost_fill_line(c + off, cstride[fd], n[fd], h[fd], eqs[fd], kernel[fd], bc_left[fd],
                        bc_right[fd], 0, 0, slice);
      }
      free(slice);
    }
  }
  return ORC_OK;
}

/* convolution_fast_interpolation/cubic_hermite_interpolation.jl:15-28 */
static inline double cubic_hermite(double t, double f0, double f1, double d0, double d1, double h) {
  double t2 = t * t;
  double h00 = 1.0 + t2 * (-3.0 + 2.0 * t);
  double h01 = t2 * (3.0 - 2.0 * t);
  double h10 = t + t2 * (-2.0 + t);
  double h11 = t2 * (t - 1.0);
  return h00 * f0 + h01 * f1 + h * (h10 * d0 + h11 * d1);
}

/* convolution_fast_interpolation/quintic_hermite_interpolation.jl:16-37 (Julia's 10t3 etc. are plain products;
 * a - b + c chains associate left to right) */
static inline double quintic_hermite(double t, double f0, double f1, double d0, double d1, double dd0, double dd1,
                                     double h) {
  double t2 = t * t;
  double t3 = t2 * t;
  double t4 = t3 * t;
  double t5 = t4 * t;
  double h00 = 1.0 - 10 * t3 + 15 * t4 - 6 * t5;
  double h01 = 10 * t3 - 15 * t4 + 6 * t5;
  double h10 = t - 6 * t3 + 8 * t4 - 3 * t5;
  double h11 = -4 * t3 + 7 * t4 - 3 * t5;
  double h20 = t2 - 3 * t3 + 3 * t4 - t5;
  double h21 = t3 - 2 * t4 + t5;
  double h2 = h * h;
  return h00 * f0 + h01 * f1 + h * (h10 * d0 + h11 * d1) + (h2 / 2) * (h20 * dd0 + h21 * dd1);
}

typedef double v4d __attribute__((vector_size(32), aligned(32)));

typedef struct {
  int64_t i;   /* 1-based cell index into the extended knots / padded coefs */
  int64_t idx; /* 1-based lower table row */
  double t;
} axis_pos;

/* eager_3d.jl:115-137 (same lines in eager_1d.jl:60-71, eager_2d.jl:138-156, perdim_2d.jl:17-49).
 * returns 0 ok, 1 if floor(Int, .) would throw InexactError (NaN / Inf) */
static inline int locate(double x, const double *knots, int64_t nk, double h, int eqs, int n_pre, axis_pos *p) {
  double i_float = (x - knots[0]) / h + 1.0;
  if (!(fabs(i_float) < 9.0e18)) return 1;
  int64_t i = (int64_t)floor(i_float);
  if (i < eqs) i = eqs;
  if (i > nk - eqs) i = nk - eqs;
  double x_diff_right = 1.0 - (x - knots[i - 1]) / h;
  double cidx = x_diff_right * (double)(n_pre - 1) + 1.0;
  if (!(fabs(cidx) < 9.0e18)) return 1;
  int64_t idx = (int64_t)floor(cidx);
  if (idx < 1) idx = 1;
  if (idx > n_pre - 1) idx = n_pre - 1;
  p->i = i;
  p->idx = idx;
  p->t = cidx - (double)idx;
  return 0;
}

/* the position variant of eager_2d.jl:93-107 ((linear, linear) on the single-kernel path; eager_nd.jl:97-115 uses the
 * same): table row from floor(diff_right * (n_pre - 1)) + 1 and t from the stored pre_range values,
 * pre_range[r] = Float64((r-1)//(n_pre-1)) (pre_range.jl) */
static inline int locate_prerange(double x, const double *knots, int64_t nk, double h, int eqs, int n_pre, axis_pos *p) {
  double i_float = (x - knots[0]) / h + 1.0;
  if (!(fabs(i_float) < 9.0e18)) return 1;
  int64_t i = (int64_t)floor(i_float);
  if (i < eqs) i = eqs;
  if (i > nk - eqs) i = nk - eqs;
  double diff_left = (x - knots[i - 1]) / h;
  double diff_right = 1.0 - diff_left;
  double f = diff_right * (double)(n_pre - 1);
  if (!(fabs(f) < 9.0e18)) return 1;
  int64_t idx = (int64_t)floor(f) + 1;
  if (idx < 1) idx = 1;
  if (idx > n_pre - 1) idx = n_pre - 1;
  double r0 = (double)(idx - 1) / (double)(n_pre - 1), r1 = (double)idx / (double)(n_pre - 1);
  p->i = i;
  p->idx = idx;
  p->t = (diff_right - r0) / (r1 - r0);
  return 0;
}

typedef struct {
  const double *kp[3];
  const double *kd[3];
  const double *ke[3]; /* K^(d+2): quintic subgrid only */
  int eqs[3];
  int n_pre[3];        /* rows of kp[d]: 101, or the generated length of a :linear axis */
} itp_tabs;

#define TAB(T, row1, col1, eqs) ((T)[((row1)-1) * 2 * (eqs) + ((col1)-1)])

/* the un-wrapped functor itp(x...) : eager_1d.jl:92-115, eager_2d.jl:136-212 / perdim_2d.jl:182-215,
 * eager_3d.jl:181-305 / perdim_3d.jl:12-159 */
static int itp_core(const orc_itp *it, const itp_tabs *tb, const double *x, double *res) {
  const double h_pre = 1.0 / (double)(N_PRE - 1);
  axis_pos p[3];
  const int prerange = it->ndim == 2 && !it->perdim && it->subgrid[0] == ORC_SG_LINEAR && it->subgrid[1] == ORC_SG_LINEAR;
  for (int d = 0; d < it->ndim; ++d)
    if (prerange ? locate_prerange(x[d], it->knots[d], it->cshape[d], it->h[d], tb->eqs[d], tb->n_pre[d], &p[d])
                 : locate(x[d], it->knots[d], it->cshape[d], it->h[d], tb->eqs[d], tb->n_pre[d], &p[d]))
      return ORC_ENAN;

  if (it->ndim == 1 && it->subgrid[0] == ORC_SG_LINEAR) { /* eager_1d.jl:73-90 */
    const int e = tb->eqs[0];
    double result_lower = 0, result_upper = 0;
    for (int j = -(e - 1); j <= e; ++j) {
      double coef = it->coefs[p[0].i + j - 1];
      result_lower += coef * TAB(tb->kp[0], p[0].idx, j + e, e);
      result_upper += coef * TAB(tb->kp[0], p[0].idx + 1, j + e, e);
    }
    double result = (1.0 - p[0].t) * result_lower + p[0].t * result_upper;
    double scale = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    *res = result * scale;
    return ORC_OK;
  }

  if (it->ndim == 3 && it->subgrid[0] == ORC_SG_LINEAR) {
    /* eager_3d.jl:139-176 (linear, linear, linear) / perdim_3d.jl:161-194 ("fallback to all linear"): 8 accumulators
     * r[bx][by][bz] += ((coef * kx) * ky) * kz in the n -> m -> l loop order, trilinear blend (@fastmath there) */
    const int ex = tb->eqs[0], ey = tb->eqs[1], ez = tb->eqs[2];
    const int64_t s1 = it->cshape[0], s2 = it->cshape[0] * it->cshape[1];
    double r[2][2][2];
    memset(r, 0, sizeof r);
    for (int n = -(ez - 1); n <= ez; ++n) {
      double kz[2] = {TAB(tb->kp[2], p[2].idx, n + ez, ez), TAB(tb->kp[2], p[2].idx + 1, n + ez, ez)};
      for (int m = -(ey - 1); m <= ey; ++m) {
        double ky[2] = {TAB(tb->kp[1], p[1].idx, m + ey, ey), TAB(tb->kp[1], p[1].idx + 1, m + ey, ey)};
        for (int l = -(ex - 1); l <= ex; ++l) {
          double coef = it->coefs[(p[0].i + l - 1) + (p[1].i + m - 1) * s1 + (p[2].i + n - 1) * s2];
          double kx[2] = {TAB(tb->kp[0], p[0].idx, l + ex, ex), TAB(tb->kp[0], p[0].idx + 1, l + ex, ex)};
          for (int bz = 0; bz < 2; ++bz)
            for (int by = 0; by < 2; ++by)
              for (int bx = 0; bx < 2; ++bx) r[bx][by][bz] += coef * kx[bx] * ky[by] * kz[bz];
        }
      }
    }
    double tx = p[0].t, ty = p[1].t, tz = p[2].t, ox = 1.0 - tx, oy = 1.0 - ty, oz = 1.0 - tz;
    double result = ox * oy * oz * r[0][0][0] + tx * oy * oz * r[1][0][0] + ox * ty * oz * r[0][1][0] +
                    tx * ty * oz * r[1][1][0] + ox * oy * tz * r[0][0][1] + tx * oy * tz * r[1][0][1] +
                    ox * ty * tz * r[0][1][1] + tx * ty * tz * r[1][1][1];
    double sc[3];
    for (int d = 0; d < 3; ++d) sc[d] = jl_pow_int(-1.0 / it->h[d], it->deriv[d]);
    if (it->perdim) {
      double s = 1.0;
      for (int d = 0; d < 3; ++d) s *= sc[d];
      *res = result * s;
    } else {
      *res = result * (sc[0] * sc[1] * sc[2]); /* result * prod(...) */
    }
    return ORC_OK;
  }

  if (it->ndim == 1 && it->subgrid[0] == ORC_SG_QUINTIC) { /* eager_1d.jl:117-145 */
    const int e = tb->eqs[0];
    double sum_f0 = 0, sum_f1 = 0, sum_d0 = 0, sum_d1 = 0, sum_dd0 = 0, sum_dd1 = 0;
    for (int j = -(e - 1); j <= e; ++j) {
      double coef = it->coefs[p[0].i + j - 1];
      int col = j + e;
      sum_f0 += coef * TAB(tb->kp[0], p[0].idx, col, e);
      sum_f1 += coef * TAB(tb->kp[0], p[0].idx + 1, col, e);
      sum_d0 += coef * TAB(tb->kd[0], p[0].idx, col, e);
      sum_d1 += coef * TAB(tb->kd[0], p[0].idx + 1, col, e);
      sum_dd0 += coef * TAB(tb->ke[0], p[0].idx, col, e);
      sum_dd1 += coef * TAB(tb->ke[0], p[0].idx + 1, col, e);
    }
    double result = quintic_hermite(p[0].t, sum_f0, sum_f1, sum_d0, sum_d1, sum_dd0, sum_dd1, h_pre);
    double scale = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    *res = result * scale;
    return ORC_OK;
  }

  if (it->ndim == 2 && (it->subgrid[0] != ORC_SG_CUBIC || it->subgrid[1] != ORC_SG_CUBIC)) {
    /* eager_2d.jl:214-311 (quintic, quintic); perdim_2d.jl (quintic, cubic) :218-258, (cubic, quintic), (quintic,
     * quintic): accumulators s[x-type][y-type][x-bracket][y-bracket], type 0 = f, 1 = d, 2 = e, each receiving
     * (coef * kx) * ky in the reference's m -> l loop order; Hermite along x first, then along y */
    const int ex = tb->eqs[0], ey = tb->eqs[1];
    const int64_t s1 = it->cshape[0];
    /* linear axes (perdim_2d.jl:57-76, :78-113, :115-147 ..., eager_2d.jl:91-134): one type (f), linear blend */
    const int nx = it->subgrid[0] == ORC_SG_QUINTIC ? 3 : (it->subgrid[0] == ORC_SG_LINEAR ? 1 : 2);
    const int ny = it->subgrid[1] == ORC_SG_QUINTIC ? 3 : (it->subgrid[1] == ORC_SG_LINEAR ? 1 : 2);
    const double *tx_tab[3] = {tb->kp[0], tb->kd[0], tb->ke[0]};
    const double *ty_tab[3] = {tb->kp[1], tb->kd[1], tb->ke[1]};
    double s[3][3][2][2];
    memset(s, 0, sizeof s);
    for (int m = -(ey - 1); m <= ey; ++m) {
      int col_y = m + ey;
      double ky[3][2];
      for (int ty = 0; ty < ny; ++ty) {
        ky[ty][0] = TAB(ty_tab[ty], p[1].idx, col_y, ey);
        ky[ty][1] = TAB(ty_tab[ty], p[1].idx + 1, col_y, ey);
      }
      for (int l = -(ex - 1); l <= ex; ++l) {
        double coef = it->coefs[(p[0].i + l - 1) + (p[1].i + m - 1) * s1];
        int col_x = l + ex;
        for (int tx = 0; tx < nx; ++tx) {
          double kx0 = TAB(tx_tab[tx], p[0].idx, col_x, ex), kx1 = TAB(tx_tab[tx], p[0].idx + 1, col_x, ex);
          for (int ty = 0; ty < ny; ++ty) {
            s[tx][ty][0][0] += coef * kx0 * ky[ty][0];
            s[tx][ty][1][0] += coef * kx1 * ky[ty][0];
            s[tx][ty][0][1] += coef * kx0 * ky[ty][1];
            s[tx][ty][1][1] += coef * kx1 * ky[ty][1];
          }
        }
      }
    }
    double vy[3][2]; /* [y-type][y-bracket] after the Hermite along x */
    for (int ty = 0; ty < ny; ++ty)
      for (int by = 0; by < 2; ++by)
        vy[ty][by] = nx == 3 ? quintic_hermite(p[0].t, s[0][ty][0][by], s[0][ty][1][by], s[1][ty][0][by], s[1][ty][1][by],
                                               s[2][ty][0][by], s[2][ty][1][by], h_pre)
                   : nx == 2 ? cubic_hermite(p[0].t, s[0][ty][0][by], s[0][ty][1][by], s[1][ty][0][by], s[1][ty][1][by], h_pre)
                             : (1.0 - p[0].t) * s[0][ty][0][by] + p[0].t * s[0][ty][1][by];
    double result = ny == 3 ? quintic_hermite(p[1].t, vy[0][0], vy[0][1], vy[1][0], vy[1][1], vy[2][0], vy[2][1], h_pre)
                  : ny == 2 ? cubic_hermite(p[1].t, vy[0][0], vy[0][1], vy[1][0], vy[1][1], h_pre)
                            : (1.0 - p[1].t) * vy[0][0] + p[1].t * vy[0][1];
    double sc0 = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    double sc1 = jl_pow_int(-1.0 / it->h[1], it->deriv[1]);
    if (it->perdim) {
      double sc = 1.0;
      sc *= sc0;
      sc *= sc1;
      *res = result * sc;
    } else {
      *res = result * sc0 * sc1;
    }
    return ORC_OK;
  }

  if (it->ndim == 1) {
    const int e = tb->eqs[0];
    double sum_f0 = 0, sum_f1 = 0, sum_di0 = 0, sum_di1 = 0;
    for (int j = -(e - 1); j <= e; ++j) {
      double coef = it->coefs[p[0].i + j - 1];
      int col = j + e;
      sum_f0 += coef * TAB(tb->kp[0], p[0].idx, col, e);
      sum_f1 += coef * TAB(tb->kp[0], p[0].idx + 1, col, e);
      sum_di0 += coef * TAB(tb->kd[0], p[0].idx, col, e);
      sum_di1 += coef * TAB(tb->kd[0], p[0].idx + 1, col, e);
    }
    double result = cubic_hermite(p[0].t, sum_f0, sum_f1, sum_di0, sum_di1, h_pre);
    double scale = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    *res = result * scale;
    return ORC_OK;
  }

  if (it->ndim == 2) {
    const int ex = tb->eqs[0], ey = tb->eqs[1];
    const int64_t s1 = it->cshape[0];
    double s[2][2][2][2]; /* [x-type][y-type][x-bracket][y-bracket], type 0=f 1=d */
    memset(s, 0, sizeof s);
    for (int m = -(ey - 1); m <= ey; ++m) {
      int col_y = m + ey;
      double ky[2][2] = {{TAB(tb->kp[1], p[1].idx, col_y, ey), TAB(tb->kp[1], p[1].idx + 1, col_y, ey)},
                         {TAB(tb->kd[1], p[1].idx, col_y, ey), TAB(tb->kd[1], p[1].idx + 1, col_y, ey)}};
      for (int l = -(ex - 1); l <= ex; ++l) {
        double coef = it->coefs[(p[0].i + l - 1) + (p[1].i + m - 1) * s1];
        int col_x = l + ex;
        double kx[2][2] = {{TAB(tb->kp[0], p[0].idx, col_x, ex), TAB(tb->kp[0], p[0].idx + 1, col_x, ex)},
                           {TAB(tb->kd[0], p[0].idx, col_x, ex), TAB(tb->kd[0], p[0].idx + 1, col_x, ex)}};
        for (int tx = 0; tx < 2; ++tx)
          for (int ty = 0; ty < 2; ++ty)
            for (int bx = 0; bx < 2; ++bx)
              for (int by = 0; by < 2; ++by) s[tx][ty][bx][by] += coef * kx[tx][bx] * ky[ty][by];
      }
    }
    double val_y0 = cubic_hermite(p[0].t, s[0][0][0][0], s[0][0][1][0], s[1][0][0][0], s[1][0][1][0], h_pre);
    double val_y1 = cubic_hermite(p[0].t, s[0][0][0][1], s[0][0][1][1], s[1][0][0][1], s[1][0][1][1], h_pre);
    double dval_y0 = cubic_hermite(p[0].t, s[0][1][0][0], s[0][1][1][0], s[1][1][0][0], s[1][1][1][0], h_pre);
    double dval_y1 = cubic_hermite(p[0].t, s[0][1][0][1], s[0][1][1][1], s[1][1][0][1], s[1][1][1][1], h_pre);
    double result = cubic_hermite(p[1].t, val_y0, val_y1, dval_y0, dval_y1, h_pre);
    double sc0 = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    double sc1 = jl_pow_int(-1.0 / it->h[1], it->deriv[1]);
    if (it->perdim) { /* perdim_2d.jl:24-30,214: result * (1*sc0*sc1) */
      double sc = 1.0;
      sc *= sc0;
      sc *= sc1;
      *res = result * sc;
    } else { /* eager_2d.jl:212 */
      *res = result * sc0 * sc1;
    }
    return ORC_OK;
  }

  /* 3-D: eager_3d.jl:181-305 / perdim_3d.jl:12-159.  The 64 accumulators s_{x}{y}{z}_{bx}{by}{bz} are kept in a
   * flat array s[iz*16 + iy*4 + ix] with i? = type*2 + bracket (0: f0, 1: f1, 2: d0, 3: d1); every accumulator
   * receives ((coef*kx)*ky)*kz in the reference's n -> m -> l loop order, so each sum is bit-identical to the
   * literal code while the compiler is free to vectorise ACROSS accumulators. */
  const int ex = tb->eqs[0], ey = tb->eqs[1], ez = tb->eqs[2];
  const int64_t s1 = it->cshape[0], s2 = it->cshape[0] * it->cshape[1];
  double s[64];
  v4d sv[16];
  memset(sv, 0, sizeof sv);
  double kxs[18][4] __attribute__((aligned(32))); /* the four x-table entries per tap (loop-invariant loads of eager_3d.jl:224-229) */
  for (int l = -(ex - 1); l <= ex; ++l) {
    int col_x = l + ex;
    kxs[col_x - 1][0] = TAB(tb->kp[0], p[0].idx, col_x, ex);
    kxs[col_x - 1][1] = TAB(tb->kp[0], p[0].idx + 1, col_x, ex);
    kxs[col_x - 1][2] = TAB(tb->kd[0], p[0].idx, col_x, ex);
    kxs[col_x - 1][3] = TAB(tb->kd[0], p[0].idx + 1, col_x, ex);
  }
  for (int n = -(ez - 1); n <= ez; ++n) {
    int col_z = n + ez;
    const double kz4[4] = {TAB(tb->kp[2], p[2].idx, col_z, ez), TAB(tb->kp[2], p[2].idx + 1, col_z, ez),
                           TAB(tb->kd[2], p[2].idx, col_z, ez), TAB(tb->kd[2], p[2].idx + 1, col_z, ez)};
    for (int m = -(ey - 1); m <= ey; ++m) {
      int col_y = m + ey;
      const double ky4[4] = {TAB(tb->kp[1], p[1].idx, col_y, ey), TAB(tb->kp[1], p[1].idx + 1, col_y, ey),
                             TAB(tb->kd[1], p[1].idx, col_y, ey), TAB(tb->kd[1], p[1].idx + 1, col_y, ey)};
      const double *crow = it->coefs + (p[0].i - ex) + (p[1].i + m - 1) * s1 + (p[2].i + n - 1) * s2;
      for (int c = 0; c < 2 * ex; ++c) {
        const double coef = crow[c];
        const double *kx4 = kxs[c];
        /* cx = coef*kx (4 lanes), then per (iy, iz): s[iz][iy][0..3] += (cx*ky[iy])*kz[iz] - same products,
         * same order of the two multiplications and the addition as the reference's scalar statements */
        v4d cx = (v4d){coef, coef, coef, coef} * *(const v4d *)kx4;
        for (int j = 0; j < 4; ++j) {
          v4d cxy = cx * (v4d){ky4[j], ky4[j], ky4[j], ky4[j]};
          for (int k = 0; k < 4; ++k) sv[k * 4 + j] += cxy * (v4d){kz4[k], kz4[k], kz4[k], kz4[k]};
        }
      }
    }
  }
  memcpy(s, sv, sizeof s);
#define S3(tx, bx, ty, by, tz, bz) s[((tz)*2 + (bz)) * 16 + ((ty)*2 + (by)) * 4 + ((tx)*2 + (bx))]
  double v[2][2][2][2]; /* [yt][zt][by][bz] */
  for (int ty = 0; ty < 2; ++ty)
    for (int tz = 0; tz < 2; ++tz)
      for (int by = 0; by < 2; ++by)
        for (int bz = 0; bz < 2; ++bz)
          v[ty][tz][by][bz] = cubic_hermite(p[0].t, S3(0, 0, ty, by, tz, bz), S3(0, 1, ty, by, tz, bz),
                                            S3(1, 0, ty, by, tz, bz), S3(1, 1, ty, by, tz, bz), h_pre);
#undef S3
  double w[2][2]; /* [zt][bz] */
  for (int tz = 0; tz < 2; ++tz)
    for (int bz = 0; bz < 2; ++bz)
      w[tz][bz] = cubic_hermite(p[1].t, v[0][tz][0][bz], v[0][tz][1][bz], v[1][tz][0][bz], v[1][tz][1][bz], h_pre);
  double result = cubic_hermite(p[2].t, w[0][0], w[0][1], w[1][0], w[1][1], h_pre);
  *res = result * jl_pow_int(-1.0 / it->h[0], it->deriv[0]) * jl_pow_int(-1.0 / it->h[1], it->deriv[1]) *
         jl_pow_int(-1.0 / it->h[2], it->deriv[2]);
  return ORC_OK;
}

/* extrapolation_operator.jl:31-63 + extrapolate_point.jl:41-94 (1-D), :98-160 (N-D) */
static int etp_point(const orc_itp *it, const itp_tabs *tb, const double *x, double *res) {
  const int N = it->ndim;
  double lo[3], hi[3];
  int within = 1;
  for (int d = 0; d < N; ++d) { /* _domain_bounds, extrapolate_point.jl:172-176 */
    lo[d] = it->knots[d][tb->eqs[d] - 1];
    hi[d] = it->knots[d][it->cshape[d] - 1 - (tb->eqs[d] - 1)];
  }
  for (int d = 0; d < N; ++d)
    if (x[d] > hi[d] || x[d] < lo[d]) { within = 0; break; }
  if (within) return itp_core(it, tb, x, res);
  double cl[3];
  for (int d = 0; d < N; ++d) cl[d] = x[d] < lo[d] ? lo[d] : (x[d] > hi[d] ? hi[d] : x[d]);
  if (it->extrap == ORC_EXTRAP_LINE) {
    double dx[3], dist_sq = 0.0;
    for (int d = 0; d < N; ++d) {
      dx[d] = x[d] - cl[d];
      dist_sq += dx[d] * dx[d];
    }
    if (dist_sq < 2.220446049250313e-16) return itp_core(it, tb, cl, res);
    double h_min = it->h[0];
    for (int d = 1; d < N; ++d) h_min = it->h[d] < h_min ? it->h[d] : h_min;
    double eps = h_min / 100.0;
    double dist = sqrt(dist_sq);
    double xin[3];
    for (int d = 0; d < N; ++d) xin[d] = cl[d] - eps * dx[d] / dist;
    double f_b, f_in;
    int rc = itp_core(it, tb, cl, &f_b);
    if (rc) return rc;
    rc = itp_core(it, tb, xin, &f_in);
    if (rc) return rc;
    double slope = (f_b - f_in) / eps;
    *res = f_b + slope * dist;
    return ORC_OK;
  }
  if (it->extrap == ORC_EXTRAP_FLAT) return itp_core(it, tb, cl, res);
  return ORC_EOOB; /* Throw */
}

int orc_itp_eval(const orc_itp *it, int64_t nq, const double *const *x, double *out, int64_t *first_bad) {
  if (it->ndim < 1 || it->ndim > 3) return ORC_EARG;
  itp_tabs tb;
  for (int d = 0; d < it->ndim; ++d) {
    tb.eqs[d] = orc_eqs(it->kernel[d]);
    tb.kp[d] = orc_table(it->kernel[d], it->deriv[d]);
    tb.n_pre[d] = N_PRE;
    if (it->subgrid[d] == ORC_SG_LINEAR && it->tab[d] && it->n_pre[d] >= 2) { /* generated table of a :linear axis */
      tb.kp[d] = it->tab[d];
      tb.n_pre[d] = it->n_pre[d];
    }
    tb.kd[d] = orc_table(it->kernel[d], it->deriv[d] + 1);
    tb.ke[d] = it->subgrid[d] == ORC_SG_QUINTIC ? orc_table(it->kernel[d], it->deriv[d] + 2) : NULL;
    if (tb.eqs[d] < 0 || !tb.kp[d] || (!tb.kd[d] && it->subgrid[d] != ORC_SG_LINEAR)) return ORC_EARG;
    if (it->subgrid[d] == ORC_SG_QUINTIC && (!tb.ke[d] || it->ndim > 2)) return ORC_EARG;
    if (it->ndim == 3 && (it->subgrid[d] == ORC_SG_LINEAR) != (it->subgrid[0] == ORC_SG_LINEAR)) return ORC_EARG;
  }
  int64_t bad = -1;
  int bad_rc = ORC_OK;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int64_t q = 0; q < nq; ++q) {
    double xq[3] = {0, 0, 0};
    for (int d = 0; d < it->ndim; ++d) xq[d] = x[d][q];
    double r = NAN;
    int rc = etp_point(it, &tb, xq, &r);
    out[q] = r;
    if (rc) {
#pragma omp critical
      {
        if (bad < 0 || q < bad) { bad = q; bad_rc = rc; }
      }
    }
  }
  if (first_bad) *first_bad = bad;
  return bad_rc;
}

/* ---- N >= 4 dimensions ------------------------------------------------------------------------ */
typedef struct {
  const double *kp[ORC_MAXND];
  int eqs[ORC_MAXND];
  int n_pre[ORC_MAXND];
} nd_tabs;

/* the un-wrapped N-D functor: eager_nd.jl:97-140 (HigherOrderKernel: pre_range positions, itp.eqs[1] and
 * itp.kernel_pre[1] on every axis) / perdim_nd.jl:4-36 (cidx positions, per-axis tables) */
static int itp_core_nd(const orc_itp_nd *it, const nd_tabs *tb, const double *x, double *res) {
  const int N = it->ndim;
  axis_pos p[ORC_MAXND];
  for (int d = 0; d < N; ++d) {
    const int eqs = it->perdim ? tb->eqs[d] : tb->eqs[0];
    /* eager_nd.jl reads itp.pre_range[1] / itp.kernel_pre[1] on every axis */
    const int npre = it->perdim ? tb->n_pre[d] : tb->n_pre[0];
    int bad = it->perdim ? locate(x[d], it->knots[d], it->cshape[d], it->h[d], eqs, npre, &p[d])
                         : locate_prerange(x[d], it->knots[d], it->cshape[d], it->h[d], eqs, npre, &p[d]);
    if (bad) return ORC_ENAN;
  }
  int64_t stride[ORC_MAXND];
  stride[0] = 1;
  for (int d = 1; d < N; ++d) stride[d] = stride[d - 1] * it->cshape[d - 1];
  int off[ORC_MAXND]; /* offsets[d] + eqs - 1 = 0 .. 2 eqs - 1, first axis fastest (Iterators.product) */
  for (int d = 0; d < N; ++d) off[d] = 0;
  double result = 0.0;
  for (;;) {
    int64_t lin = 0;
    for (int d = 0; d < N; ++d) {
      const int eqs = it->perdim ? tb->eqs[d] : tb->eqs[0];
      lin += (p[d].i - 1 + (off[d] - (eqs - 1))) * stride[d];
    }
    const double coef = it->coefs[lin];
    double kernel_val = 1.0;
    for (int d = 0; d < N; ++d) {
      const int eqs = it->perdim ? tb->eqs[d] : tb->eqs[0];
      const double *kp = it->perdim ? tb->kp[d] : tb->kp[0];
      const int col = off[d] + 1; /* offsets[d] + eqs, 1-based */
      const double k_lower = TAB(kp, p[d].idx, col, eqs);
      const double k_upper = TAB(kp, p[d].idx + 1, col, eqs);
      kernel_val *= (1.0 - p[d].t) * k_lower + p[d].t * k_upper;
    }
    result += coef * kernel_val;
    int d = 0;
    for (; d < N; ++d) {
      const int eqs = it->perdim ? tb->eqs[d] : tb->eqs[0];
      if (++off[d] < 2 * eqs) break;
      off[d] = 0;
    }
    if (d == N) break;
  }
  double s;
  if (it->perdim) { /* _perdim_scale, perdim_2d.jl:24-30 */
    s = 1.0;
    for (int d = 0; d < N; ++d) s *= jl_pow_int(-1.0 / it->h[d], it->deriv[d]);
  } else { /* prod(generator), eager_nd.jl:139 */
    s = jl_pow_int(-1.0 / it->h[0], it->deriv[0]);
    for (int d = 1; d < N; ++d) s *= jl_pow_int(-1.0 / it->h[d], it->deriv[d]);
  }
  *res = result * s;
  return ORC_OK;
}

/* extrapolation_operator.jl:31-63 + extrapolate_point.jl:98-160 */
static int etp_point_nd(const orc_itp_nd *it, const nd_tabs *tb, const double *x, double *res) {
  const int N = it->ndim;
  double lo[ORC_MAXND], hi[ORC_MAXND];
  int within = 1;
  for (int d = 0; d < N; ++d) {
    lo[d] = it->knots[d][tb->eqs[d] - 1];
    hi[d] = it->knots[d][it->cshape[d] - 1 - (tb->eqs[d] - 1)];
  }
  for (int d = 0; d < N; ++d)
    if (x[d] > hi[d] || x[d] < lo[d]) { within = 0; break; }
  if (within) return itp_core_nd(it, tb, x, res);
  double cl[ORC_MAXND];
  for (int d = 0; d < N; ++d) cl[d] = x[d] < lo[d] ? lo[d] : (x[d] > hi[d] ? hi[d] : x[d]);
  if (it->extrap == ORC_EXTRAP_LINE) {
    double dx[ORC_MAXND], dist_sq = 0.0;
    for (int d = 0; d < N; ++d) {
      dx[d] = x[d] - cl[d];
      dist_sq += dx[d] * dx[d];
    }
    if (dist_sq < 2.220446049250313e-16) return itp_core_nd(it, tb, cl, res);
    double h_min = it->h[0];
    for (int d = 1; d < N; ++d) h_min = it->h[d] < h_min ? it->h[d] : h_min;
    double eps = h_min / 100.0;
    double dist = sqrt(dist_sq);
    double xin[ORC_MAXND];
    for (int d = 0; d < N; ++d) xin[d] = cl[d] - eps * dx[d] / dist;
    double f_b, f_in;
    int rc = itp_core_nd(it, tb, cl, &f_b);
    if (rc) return rc;
    rc = itp_core_nd(it, tb, xin, &f_in);
    if (rc) return rc;
    double slope = (f_b - f_in) / eps;
    *res = f_b + slope * dist;
    return ORC_OK;
  }
  if (it->extrap == ORC_EXTRAP_FLAT) return itp_core_nd(it, tb, cl, res);
  return ORC_EOOB;
}

int orc_itp_eval_nd(const orc_itp_nd *it, int64_t nq, const double *const *x, double *out, int64_t *first_bad) {
  if (it->ndim < 4 || it->ndim > ORC_MAXND) return ORC_EARG;
  nd_tabs tb;
  for (int d = 0; d < it->ndim; ++d) {
    tb.eqs[d] = orc_eqs(it->kernel[d]);
    tb.kp[d] = orc_table(it->kernel[d], it->deriv[d]);
    tb.n_pre[d] = N_PRE;
    if (it->tab[d] && it->n_pre[d] >= 2) {
      tb.kp[d] = it->tab[d];
      tb.n_pre[d] = it->n_pre[d];
    }
    if (tb.eqs[d] < 0 || !tb.kp[d]) return ORC_EARG;
    if (!it->perdim && (it->kernel[d] != it->kernel[0] || it->deriv[d] != it->deriv[0])) return ORC_EARG;
  }
  int64_t bad = -1;
  int bad_rc = ORC_OK;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int64_t q = 0; q < nq; ++q) {
    double xq[ORC_MAXND];
    for (int d = 0; d < it->ndim; ++d) xq[d] = x[d][q];
    double r = NAN;
    int rc = etp_point_nd(it, &tb, xq, &r);
    out[q] = r;
    if (rc) {
#pragma omp critical
      {
        if (bad < 0 || q < bad) { bad = q; bad_rc = rc; }
      }
    }
  }
  if (first_bad) *first_bad = bad;
  return bad_rc;
}

int orc_resample_tables(int ndim, const int64_t *n_in, const int64_t *n_out, const double *values,
                        double *out, const double *const *knots_in, const double *const *knots_out,
                        const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                        const int32_t *bc_right, const int32_t *subgrid, const double *const *tab,
                        const int32_t *n_pre_axis);
/* convolution/convolution_resample.jl:240-386, cubic subgrid branch :350-361 */
int orc_resample(int ndim, const int64_t *n_in, const int64_t *n_out, const double *values,
                 double *out, const double *const *knots_in, const double *const *knots_out,
                 const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                 const int32_t *bc_right) {
  int32_t sg[ORC_MAXDIM] = {0, 0, 0, 0};
  return orc_resample_sg(ndim, n_in, n_out, values, out, knots_in, knots_out, kernel, deriv, bc_left, bc_right, sg);
}

int orc_resample_sg(int ndim, const int64_t *n_in, const int64_t *n_out, const double *values,
                    double *out, const double *const *knots_in, const double *const *knots_out,
                    const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                    const int32_t *bc_right, const int32_t *subgrid) {
  return orc_resample_tables(ndim, n_in, n_out, values, out, knots_in, knots_out, kernel, deriv, bc_left, bc_right,
                             subgrid, NULL, NULL);
}

/* tab[d] / n_pre[d]: the generated table of a :linear axis (precompute_actual = max(precompute, 10_000),
 * convolution_resample.jl:228-229), row-major; NULL = the 101-row shipped table */
int orc_resample_tables(int ndim, const int64_t *n_in, const int64_t *n_out, const double *values,
                        double *out, const double *const *knots_in, const double *const *knots_out,
                        const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                        const int32_t *bc_right, const int32_t *subgrid, const double *const *tab,
                        const int32_t *n_pre_axis) {
  if (ndim < 1 || ndim > ORC_MAXDIM) return ORC_EARG;
  int64_t cur_shape[ORC_MAXDIM];
  int64_t tot = 1;
  for (int d = 0; d < ndim; ++d) { cur_shape[d] = n_in[d]; tot *= n_in[d]; }
  double *current = (double *)malloc(sizeof(double) * (size_t)tot);
  memcpy(current, values, sizeof(double) * (size_t)tot);
  const double h_pre = 1.0 / (double)(N_PRE - 1);
  for (int d = 0; d < ndim; ++d) {
    const int eqs_d = orc_eqs(kernel[d]);
    const double *kernel_pre = orc_table(kernel[d], deriv[d]);
    const double *kd1_pre = orc_table(kernel[d], deriv[d] + 1);
    const int quintic = subgrid[d] == ORC_SG_QUINTIC;
    const double *kd2_pre = quintic ? orc_table(kernel[d], deriv[d] + 2) : NULL;
    const int linear = subgrid[d] == ORC_SG_LINEAR;
    int n_pre = N_PRE;
    if (linear && tab && n_pre_axis && tab[d] && n_pre_axis[d] >= 2) {
      kernel_pre = tab[d];
      n_pre = n_pre_axis[d];
    }
    if (eqs_d < 0 || !kernel_pre || (!kd1_pre && !linear) || (quintic && !kd2_pre)) { free(current); return ORC_EARG; }
    const int64_t nin = cur_shape[d], nout = n_out[d];
    const double h_d = knots_in[d][1] - knots_in[d][0];
    const int64_t n_coef = nin + 2 * (eqs_d - 1);
    const double x_expanded_1 = knots_in[d][0] - (double)(eqs_d - 1) * h_d;
    int64_t next_shape[ORC_MAXDIM], ntot = 1, stride_in = 1, stride_out = 1;
    for (int j = 0; j < ndim; ++j) { next_shape[j] = j == d ? nout : cur_shape[j]; ntot *= next_shape[j]; }
    for (int j = 0; j < d; ++j) { stride_in *= cur_shape[j]; stride_out *= next_shape[j]; }
    double *next = (double *)calloc((size_t)ntot, sizeof(double));
    const int64_t inner = stride_in; /* product of dims < d (equal in/out) */
    int64_t outer = 1;
    for (int j = d + 1; j < ndim; ++j) outer *= cur_shape[j];
    const double scale = jl_pow_int(-1.0 / h_d, deriv[d]);
#pragma omp parallel num_threads(g_threads)
    {
      double *coef_buf = (double *)malloc(sizeof(double) * (size_t)n_coef);
      double *slice = (double *)malloc(sizeof(double) * (size_t)nin);
#pragma omp for schedule(static) collapse(2)
      for (int64_t o = 0; o < outer; ++o)
        for (int64_t in = 0; in < inner; ++in) {
          const int64_t offset_in = in + o * inner * nin;
          const int64_t offset_out = in + o * inner * nout;
          memset(coef_buf, 0, sizeof(double) * (size_t)n_coef);
          for (int64_t k = 1; k <= nin; ++k) coef_buf[k + eqs_d - 1 - 1] = current[offset_in + (k - 1) * stride_in];
          ghost_fill_line(coef_buf + (eqs_d - 1), 1, nin, h_d, eqs_d, kernel[d], bc_left[d], bc_right[d], 0, 0, slice);
          for (int64_t k = 1; k <= nout; ++k) {
            double xq = knots_out[d][k - 1];
            double i_float = (xq - x_expanded_1) / h_d + 1.0;
            int64_t i = (int64_t)floor(i_float);
            if (i < eqs_d) i = eqs_d;
            if (i > n_coef - eqs_d) i = n_coef - eqs_d;
            double knot_i = x_expanded_1 + (double)(i - 1) * h_d;
            double x_diff_left = (xq - knot_i) / h_d;
            double x_diff_right = 1.0 - x_diff_left;
            double continuous_idx = x_diff_right * (double)(n_pre - 1) + 1.0;
            int64_t idx = (int64_t)floor(continuous_idx);
            if (idx < 1) idx = 1;
            if (idx > n_pre - 1) idx = n_pre - 1;
            double t = continuous_idx - (double)idx;
            double sum_f0 = 0, sum_f1 = 0, sum_d0 = 0, sum_d1 = 0, sum_dd0 = 0, sum_dd1 = 0;
            for (int l = -(eqs_d - 1); l <= eqs_d; ++l) {
              double coef = coef_buf[i + l - 1];
              int col = l + eqs_d;
              sum_f0 += coef * TAB(kernel_pre, idx, col, eqs_d);
              sum_f1 += coef * TAB(kernel_pre, idx + 1, col, eqs_d);
              if (linear) continue; /* :363-373 */
              sum_d0 += coef * TAB(kd1_pre, idx, col, eqs_d);
              sum_d1 += coef * TAB(kd1_pre, idx + 1, col, eqs_d);
              if (quintic) { /* :333-349 */
                sum_dd0 += coef * TAB(kd2_pre, idx, col, eqs_d);
                sum_dd1 += coef * TAB(kd2_pre, idx + 1, col, eqs_d);
              }
            }
            double v = quintic ? quintic_hermite(t, sum_f0, sum_f1, sum_d0, sum_d1, sum_dd0, sum_dd1, h_pre)
                     : linear  ? (1.0 - t) * sum_f0 + t * sum_f1
                               : cubic_hermite(t, sum_f0, sum_f1, sum_d0, sum_d1, h_pre);
            next[offset_out + (k - 1) * stride_out] = v * scale; /* :378-380 */
          }
        }
      free(coef_buf);
      free(slice);
    }
    free(current);
    current = next;
    cur_shape[d] = nout;
  }
  int64_t otot = 1;
  for (int d = 0; d < ndim; ++d) otot *= n_out[d];
  memcpy(out, current, sizeof(double) * (size_t)otot);
  free(current);
  return ORC_OK;
}

/* convolution_gaussian_smoothing.jl:46 */
int orc_smooth_halfwidth(double B) { return (int)ceil(sqrt(-log(1e-12 / 2) / B)); }

/* convolution_kernels/convolution_kernel_gaussian.jl:23-31 */
static double theta(double B, int terms) {
  double q = exp(-B);
  double sum = 1.0;
  for (int n = 1; n <= terms; ++n) {
    double term = 2 * jl_pow_int(q, (int64_t)n * n);
    sum += term;
  }
  return sum;
}

/* convolution/convolution_gaussian_smoothing.jl:45-126 */
int orc_smooth(int ndim, const int64_t *n, const double *values, double *out, const double *h, double B) {
  if (ndim < 1 || ndim > ORC_MAXDIM) return ORC_EARG;
  const int E = orc_smooth_halfwidth(B);
  int64_t tot = 1;
  for (int d = 0; d < ndim; ++d) {
    if (E >= n[d]) return ORC_EARG; /* :50-56 */
    tot *= n[d];
  }
  const double norm = theta(B, E);
  double *kw = (double *)malloc(sizeof(double) * (size_t)(2 * E + 1));
  for (int l = -E; l <= E; ++l) kw[l + E] = exp(-B * ((double)l * (double)l)) / norm; /* :59-63 */
  double *current = (double *)malloc(sizeof(double) * (size_t)tot);
  memcpy(current, values, sizeof(double) * (size_t)tot);
  for (int d = 0; d < ndim; ++d) {
    const int64_t nd = n[d];
    const int64_t n_coef = nd + 2 * E;
    int64_t stride = 1, outer = 1;
    for (int j = 0; j < d; ++j) stride *= n[j];
    for (int j = d + 1; j < ndim; ++j) outer *= n[j];
    double *next = (double *)calloc((size_t)tot, sizeof(double));
#pragma omp parallel num_threads(g_threads)
    {
      double *coef_buf = (double *)malloc(sizeof(double) * (size_t)n_coef);
      double *slice = (double *)malloc(sizeof(double) * (size_t)nd);
#pragma omp for schedule(static) collapse(2)
      for (int64_t o = 0; o < outer; ++o)
        for (int64_t in = 0; in < stride; ++in) {
          const int64_t off = in + o * stride * nd;
          memset(coef_buf, 0, sizeof(double) * (size_t)n_coef);
          for (int64_t k = 1; k <= nd; ++k) coef_buf[k + E - 1] = current[off + (k - 1) * stride]; /* :97-100 */
          /* :103-105 -> Gaussian branch, ghosts shifted one slot outward (shift = 1), the two
           * slots next to the interior stay 0 */
          ghost_fill_line(coef_buf + E, 1, nd, h[d], E, ORC_K_GAUSS, ORC_BC_DETECT, ORC_BC_DETECT, 1, 1, slice);
          for (int64_t k = 1; k <= nd; ++k) {
            double result = 0.0;
            int64_t i_base = k + E;
            for (int l = -E; l <= E; ++l) result += coef_buf[i_base + l - 1] * kw[l + E];
            next[off + (k - 1) * stride] = result;
          }
        }
      free(coef_buf);
      free(slice);
    }
    free(current);
    current = next;
  }
  memcpy(out, current, sizeof(double) * (size_t)tot);
  free(current);
  free(kw);
  return ORC_OK;
}

/* =============================================================================================
 * Antiderivative fast path (derivative = -1 on EVERY axis, 1-D / 2-D / 3-D): FastIntegralOrder functors
 *   convolution_fast_interpolation/convolution_fast_integration_{1d,2d,3d}.jl, eval_subgrid_helpers.jl:22-32,
 *   tails built by convolution/fast_convolution_interpolation_constructor.jl:176-364 (_build_tails_dispatch),
 *   _compute_left_values :414-448, _suffix_sum :452-454.
 * Tables: kernel_pre = K~ (KERNEL_PRE_<k>_i1), kernel_d1_pre = K, kernel_d2_pre = K'
 * (precomputed_kernel_tables.jl:93-96).  Everything below is a literal restatement, including the right-hand tails
 * (which are identically zero: the clamped cell index keeps every right-tail index >= 2 eqs + 1, where
 * -1/2 - left_values[j] == 0) and the scan order of tail3_corner_rlr (:351, suffix along int_dims[2] and cumsum along
 * int_dims[1] although the weights are (r, l, r) - harmless for the same reason).
 * ============================================================================================= */
const double *orc_table_i1(int kernel) {
  switch (kernel) {
    case ORC_K_A3: return CIB_TAB_a3_i1;
    case ORC_K_A4: return CIB_TAB_a4_i1;
    case ORC_K_A5: return CIB_TAB_a5_i1;
    case ORC_K_A7: return CIB_TAB_a7_i1;
    case ORC_K_B5: return CIB_TAB_b5_i1;
    case ORC_K_B7: return CIB_TAB_b7_i1;
    case ORC_K_B9: return CIB_TAB_b9_i1;
    case ORC_K_B11: return CIB_TAB_b11_i1;
    case ORC_K_B13: return CIB_TAB_b13_i1;
    default: return NULL;
  }
}

int orc_itg_create_mixed(int ndim, const int64_t *cshape, const double *coefs, const double *const *knots,
                         const double *h, const int32_t *kernel, const int32_t *subgrid, const int32_t *deriv,
                         const double *const *tab, const int32_t *n_pre, int extrap, orc_itg **out);
struct orc_itg {
  int ndim, extrap;
  int64_t cshape[3], cstride[3], ntot;
  double *coefs;
  double *knots[3];
  double h[3];
  int eqs[3], subgrid[3];
  const double *kp[3], *kd[3], *ke[3];
  double *lv[3];
  /* mixed interpolants (FastMixedIntegralOrder): deriv[d] = -1 on the integral axes, >= 0 elsewhere; the other axes read
   * the generated :linear table lin[d] (n_lin[d] rows, row-major) of K^(deriv[d]) */
  int deriv[3], n_int;
  double *lin[3];
  int n_lin[3];
  /* tails, indexed by a base-3 code over the axes: digit 0 = free (local stencil), 1 = left (prefix), 2 = right
   * (suffix); code = s0 + 3 s1 + 9 s2.  1-D: tail1_left = [1], tail1_right = [2].  2-D: tail1_left[1] = [1],
   * tail1_right[1] = [2], tail1_left[2] = [3], tail1_right[2] = [6], tail2_ll = [4], tail2_rl = [5], tail2_lr = [7],
   * tail2_rr = [8].  3-D likewise (faces one non-zero digit, edges two, corners three). */
  double *tail[27];
};

static double sg_blend(const orc_itg *g, int d, double t, int64_t idx, int64_t col) {
  const int e = g->eqs[d];
  const double h_pre = 1.0 / (double)(N_PRE - 1);
  if (g->subgrid[d] == ORC_SG_QUINTIC)
    return quintic_hermite(t, TAB(g->kp[d], idx, col, e), TAB(g->kp[d], idx + 1, col, e), TAB(g->kd[d], idx, col, e),
                           TAB(g->kd[d], idx + 1, col, e), TAB(g->ke[d], idx, col, e), TAB(g->ke[d], idx + 1, col, e),
                           h_pre);
  if (g->subgrid[d] == ORC_SG_CUBIC)
    return cubic_hermite(t, TAB(g->kp[d], idx, col, e), TAB(g->kp[d], idx + 1, col, e), TAB(g->kd[d], idx, col, e),
                         TAB(g->kd[d], idx + 1, col, e), h_pre);
  return (1.0 - t) * TAB(g->kp[d], idx, col, e) + t * TAB(g->kp[d], idx + 1, col, e);
}

/* eval_sg_kt, eval_subgrid_helpers.jl:22-32 (jd 1-based) */
static double eval_sg_kt(const orc_itg *g, int d, int64_t jd, double xd) {
  const int e = g->eqs[d];
  double xjd = g->knots[d][e - 1] + (double)(jd - e) * g->h[d];
  double sj = (xd - xjd) / g->h[d];
  if (fabs(sj) >= (double)e) return 0.5 * (double)((sj > 0) - (sj < 0));
  double col_float = (double)e + sj;
  int64_t col = (int64_t)floor(col_float) + 1;
  if (col < 1) col = 1;
  if (col > 2 * e) col = 2 * e;
  double t = (col_float - (double)(col - 1)) * (double)(N_PRE - 1) + 1.0;
  int64_t idx = (int64_t)floor(t);
  if (idx < 1) idx = 1;
  if (idx > N_PRE - 1) idx = N_PRE - 1;
  t -= (double)idx;
  return sg_blend(g, d, t, idx, col);
}

/* _compute_left_values, fast_convolution_interpolation_constructor.jl:414-448 */
static void compute_left_values(const orc_itg *g, int d, double *lv) {
  const int e = g->eqs[d];
  for (int64_t j = 1; j <= g->cshape[d]; ++j) {
    double sj = (double)(e - j);
    if (fabs(sj) >= (double)e) {
      lv[j - 1] = 0.5 * (double)((sj > 0) - (sj < 0));
      continue;
    }
    double col_float = (double)e + sj;
    int64_t col = (int64_t)floor(col_float) + 1;
    if (col < 1) col = 1;
    if (col > 2 * e) col = 2 * e;
    double x_diff_right = col_float - (double)(col - 1);
    double cidx = x_diff_right * (double)(N_PRE - 1) + 1.0;
    int64_t idx = (int64_t)floor(cidx);
    if (idx < 1) idx = 1;
    if (idx > N_PRE - 1) idx = N_PRE - 1;
    lv[j - 1] = sg_blend(g, d, cidx - (double)idx, idx, col);
  }
}

/* Julia Base accumulate_pairwise! (base/accumulate.jl): what cumsum(v::Vector; dims=1) runs for floats */
static double jl_accumulate_pairwise(double *c, const double *v, double s, int64_t i1, int64_t n, int64_t stride) {
  if (n < 128) {
    double s_ = v[i1 * stride];
    c[i1 * stride] = s + s_;
    for (int64_t i = i1 + 1; i < i1 + n; ++i) {
      s_ = s_ + v[i * stride];
      c[i * stride] = s + s_;
    }
    return s_;
  }
  int64_t n2 = n >> 1;
  double s_ = jl_accumulate_pairwise(c, v, s, i1, n2, stride);
  s_ = s_ + jl_accumulate_pairwise(c, v, s + s_, i1 + n2, n - n2, stride);
  return s_;
}

/* cumsum(A, dims = ax) (reverse = 0) or _suffix_sum(A, ax) = reverse(cumsum(reverse(A))) (reverse = 1), in place.
 * N-D arrays: Base._accumulate! walks the axis sequentially; a Vector takes the pairwise route. */
static void scan_axis(const orc_itg *g, double *a, int ax, int reverse) {
  const int64_t n = g->cshape[ax], st = g->cstride[ax];
  if (g->ndim == 1) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n), *res = (double *)malloc(sizeof(double) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) tmp[i] = reverse ? a[n - 1 - i] : a[i];
    res[0] = tmp[0];
    if (n > 1) jl_accumulate_pairwise(res, tmp, tmp[0], 1, n - 1, 1);
    for (int64_t i = 0; i < n; ++i) a[i] = reverse ? res[n - 1 - i] : res[i];
    free(tmp);
    free(res);
    return;
  }
  for (int64_t o = 0; o < g->ntot; ++o) {
    if ((o / st) % n != 0) continue; /* first element of a line along ax */
    double *line = a + o;
    if (!reverse)
      for (int64_t i = 1; i < n; ++i) line[i * st] = line[(i - 1) * st] + line[i * st];
    else
      for (int64_t i = n - 2; i >= 0; --i) line[i * st] = line[(i + 1) * st] + line[i * st];
  }
}

/* coefs .* w_{d1} .* w_{d2} ... (ascending axes), side[d]: 0 none, 1 -> (1/2 - lv), 2 -> (-1/2 - lv) */
static double *weighted(const orc_itg *g, const int *side) {
  double *a = (double *)malloc(sizeof(double) * (size_t)g->ntot);
  for (int64_t o = 0; o < g->ntot; ++o) {
    double v = g->coefs[o];
    for (int d = 0; d < g->ndim; ++d) {
      if (!side[d]) continue;
      int64_t j = (o / g->cstride[d]) % g->cshape[d];
      v = v * ((side[d] == 1 ? 0.5 : -0.5) - g->lv[d][j]);
    }
    a[o] = v;
  }
  return a;
}

void orc_itg_destroy(orc_itg *g) {
  if (!g) return;
  free(g->coefs);
  for (int d = 0; d < 3; ++d) { free(g->knots[d]); free(g->lv[d]); free(g->lin[d]); }
  for (int c = 0; c < 27; ++c) free(g->tail[c]);
  free(g);
}

int orc_itg_create(int ndim, const int64_t *cshape, const double *coefs, const double *const *knots, const double *h,
                   const int32_t *kernel, const int32_t *subgrid, int extrap, orc_itg **out) {
  const int32_t dv[3] = {-1, -1, -1};
  return orc_itg_create_mixed(ndim, cshape, coefs, knots, h, kernel, subgrid, dv, NULL, NULL, extrap, out);
}

/* deriv[d] = -1: integral axis (subgrid cubic / quintic on the shipped K~ tables); deriv[d] >= 0: interpolation /
 * derivative axis on the :linear subgrid (_build_subgrids, constructor.jl:474-491) reading tab[d] (n_pre[d] rows,
 * row-major; NULL = the 101-row shipped K^(deriv)) */
int orc_itg_create_mixed(int ndim, const int64_t *cshape, const double *coefs, const double *const *knots,
                         const double *h, const int32_t *kernel, const int32_t *subgrid, const int32_t *deriv,
                         const double *const *tab, const int32_t *n_pre, int extrap, orc_itg **out) {
  if (ndim < 1 || ndim > 3) return ORC_EARG;
  orc_itg *g = (orc_itg *)calloc(1, sizeof(orc_itg));
  for (int d = 0; d < ndim; ++d) {
    g->deriv[d] = deriv[d];
    g->n_int += deriv[d] == -1;
  }
  if (g->n_int == 0) { free(g); return ORC_EARG; }
  g->ndim = ndim;
  g->extrap = extrap;
  g->ntot = 1;
  for (int d = 0; d < ndim; ++d) {
    g->cshape[d] = cshape[d];
    g->cstride[d] = g->ntot;
    g->ntot *= cshape[d];
    g->h[d] = h[d];
    g->eqs[d] = orc_eqs(kernel[d]);
    g->subgrid[d] = subgrid[d];
    g->kp[d] = orc_table_i1(kernel[d]);
    g->kd[d] = orc_table(kernel[d], 0);
    g->ke[d] = orc_table(kernel[d], 1);
    if (deriv[d] >= 0) { /* :linear axis of a mixed interpolant */
      const double *src = (tab && n_pre && tab[d] && n_pre[d] >= 2) ? tab[d] : orc_table(kernel[d], deriv[d]);
      g->n_lin[d] = (tab && n_pre && tab[d] && n_pre[d] >= 2) ? n_pre[d] : N_PRE;
      if (g->eqs[d] < 0 || !src) { orc_itg_destroy(g); return ORC_EARG; }
      const size_t nel = (size_t)g->n_lin[d] * 2 * (size_t)g->eqs[d];
      g->lin[d] = (double *)malloc(sizeof(double) * nel);
      memcpy(g->lin[d], src, sizeof(double) * nel);
    } else if (g->eqs[d] < 0 || !g->kp[d] || !g->kd[d] || (subgrid[d] == ORC_SG_QUINTIC && !g->ke[d]) ||
               subgrid[d] == ORC_SG_LINEAR /* integral axes: shipped K~ tables only */) {
      orc_itg_destroy(g);
      return ORC_EARG;
    }
    g->knots[d] = (double *)malloc(sizeof(double) * (size_t)cshape[d]);
    memcpy(g->knots[d], knots[d], sizeof(double) * (size_t)cshape[d]);
    g->lv[d] = (double *)malloc(sizeof(double) * (size_t)cshape[d]);
  }
  g->coefs = (double *)malloc(sizeof(double) * (size_t)g->ntot);
  memcpy(g->coefs, coefs, sizeof(double) * (size_t)g->ntot);
  for (int d = 0; d < ndim; ++d)
    if (g->deriv[d] == -1) compute_left_values(g, d, g->lv[d]);

  /* scan recipes: {axis (0-based), reverse} in the order the reference applies them */
  typedef struct { int n; int ax[3]; int rev[3]; } recipe;
  int ncode = ndim == 1 ? 3 : (ndim == 2 ? 9 : 27);
  for (int code = 1; code < ncode; ++code) {
    int side[3] = {code % 3, (code / 3) % 3, (code / 9) % 3};
    int nz = (side[0] != 0) + (side[1] != 0) + (side[2] != 0);
    int on_int_axes = 1; /* tails exist along the integral axes only (int_dims, constructor.jl:198, 235) */
    for (int d = 0; d < ndim; ++d)
      if (side[d] && g->deriv[d] != -1) on_int_axes = 0;
    if (!on_int_axes) continue;
    recipe r = {0, {0, 0, 0}, {0, 0, 0}};
    if (nz == 1) { /* tail1_left / tail1_right (:207-222), tail3_face_* (:330-331) */
      for (int d = 0; d < 3; ++d)
        if (side[d]) { r.n = 1; r.ax[0] = d; r.rev[0] = side[d] == 2; }
    } else if (nz == 2) { /* tail2_* (:265-268), tail3_edge_* (:332-335) */
      int a = -1, b = -1; /* a < b: the two saturated axes */
      for (int d = 0; d < 3; ++d)
        if (side[d]) { if (a < 0) a = d; else b = d; }
      r.n = 2;
      if (side[a] == 1 && side[b] == 1) { /* ll: cumsum(cumsum(w, dims=a), dims=b) */
        r.ax[0] = a; r.rev[0] = 0; r.ax[1] = b; r.rev[1] = 0;
      } else { /* rl: suffix(cumsum(w, b), a); lr: cumsum(suffix(w, b), a); rr: suffix(suffix(w, b), a) */
        r.ax[0] = b; r.rev[0] = side[b] == 2; r.ax[1] = a; r.rev[1] = side[a] == 2;
      }
    } else { /* corners (:336-343), digits (s0, s1, s2) */
      r.n = 3;
      const int key = side[0] * 100 + side[1] * 10 + side[2];
      switch (key) {
        case 111: r.ax[0] = 0; r.rev[0] = 0; r.ax[1] = 1; r.rev[1] = 0; r.ax[2] = 2; r.rev[2] = 0; break; /* lll */
        case 211: r.ax[0] = 1; r.rev[0] = 0; r.ax[1] = 2; r.rev[1] = 0; r.ax[2] = 0; r.rev[2] = 1; break; /* rll */
        case 121: r.ax[0] = 2; r.rev[0] = 0; r.ax[1] = 1; r.rev[1] = 1; r.ax[2] = 0; r.rev[2] = 0; break; /* lrl */
        case 112: r.ax[0] = 2; r.rev[0] = 1; r.ax[1] = 0; r.rev[1] = 0; r.ax[2] = 1; r.rev[2] = 0; break; /* llr */
        case 221: r.ax[0] = 2; r.rev[0] = 0; r.ax[1] = 1; r.rev[1] = 1; r.ax[2] = 0; r.rev[2] = 1; break; /* rrl */
        case 212: r.ax[0] = 1; r.rev[0] = 1; r.ax[1] = 0; r.rev[1] = 0; r.ax[2] = 2; r.rev[2] = 1; break; /* rlr (sic) */
        case 122: r.ax[0] = 2; r.rev[0] = 1; r.ax[1] = 1; r.rev[1] = 1; r.ax[2] = 0; r.rev[2] = 0; break; /* lrr */
        default:  r.ax[0] = 2; r.rev[0] = 1; r.ax[1] = 1; r.rev[1] = 1; r.ax[2] = 0; r.rev[2] = 1; break; /* rrr */
      }
    }
    double *a = weighted(g, side);
    for (int s = 0; s < r.n; ++s) scan_axis(g, a, r.ax[s], r.rev[s]);
    g->tail[code] = a;
  }
  *out = g;
  return ORC_OK;
}

/* the un-wrapped functors: convolution_fast_integration_1d.jl:17-38, _2d.jl:18-80, _3d.jl:4-100 */
static int itg_core_mixed(const orc_itg *g, const double *x, double *res);
static int itg_core(const orc_itg *g, const double *x, double *res) {
  const int N = g->ndim;
  if (g->n_int < N) return itg_core_mixed(g, x, res);
  int64_t i[3] = {1, 1, 1}, l[3], r[3];
  int l_ok[3], r_ok[3];
  for (int d = 0; d < N; ++d) {
    double i_float = (x[d] - g->knots[d][0]) / g->h[d] + 1.0;
    if (!(fabs(i_float) < 9.0e18)) return ORC_ENAN;
    int64_t id = (int64_t)floor(i_float);
    if (id < g->eqs[d]) id = g->eqs[d];
    if (id > g->cshape[d] - g->eqs[d]) id = g->cshape[d] - g->eqs[d];
    i[d] = id;
    l[d] = id - g->eqs[d];
    l_ok[d] = l[d] >= 1;
    r[d] = id + g->eqs[d] + 1;
    r_ok[d] = r[d] <= g->cshape[d];
  }
#define LVW(d, j) (eval_sg_kt(g, d, j, x[d]) - g->lv[d][(j)-1])
#define A1(code, a) (g->tail[code][(a)-1])
#define A2(code, a, b) (g->tail[code][((a)-1) + ((b)-1) * g->cstride[1]])
#define A3(code, a, b, c) (g->tail[code][((a)-1) + ((b)-1) * g->cstride[1] + ((c)-1) * g->cstride[2]])
  double result = 0.0;
  if (N == 1) {
    for (int64_t j = i[0] - g->eqs[0] + 1; j <= i[0] + g->eqs[0]; ++j)
      result += g->coefs[j - 1] * LVW(0, j);
    double left_tail = l_ok[0] ? A1(1, l[0]) : 0.0;
    double right_tail = r_ok[0] ? A1(2, r[0]) : 0.0;
    *res = (result + left_tail + right_tail) * g->h[0];
    return ORC_OK;
  }
  if (N == 2) {
    const int e1 = g->eqs[0], e2 = g->eqs[1];
    for (int64_t j2 = i[1] - e2 + 1; j2 <= i[1] + e2; ++j2) {
      double lv2 = LVW(1, j2);
      for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1)
        result += g->coefs[(j1 - 1) + (j2 - 1) * g->cstride[1]] * LVW(0, j1) * lv2;
      double tl1 = l_ok[0] ? A2(1, l[0], j2) : 0.0;
      double tr1 = r_ok[0] ? A2(2, r[0], j2) : 0.0;
      result += (tl1 + tr1) * lv2;
    }
    if (l_ok[1]) {
      for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1) result += A2(3, j1, l[1]) * LVW(0, j1);
      result += l_ok[0] ? A2(4, l[0], l[1]) : 0.0;
      result += r_ok[0] ? A2(5, r[0], l[1]) : 0.0;
    }
    if (r_ok[1]) {
      for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1) result += A2(6, j1, r[1]) * LVW(0, j1);
      result += l_ok[0] ? A2(7, l[0], r[1]) : 0.0;
      result += r_ok[0] ? A2(8, r[0], r[1]) : 0.0;
    }
    *res = result * g->h[0] * g->h[1];
    return ORC_OK;
  }
  const int e1 = g->eqs[0], e2 = g->eqs[1], e3 = g->eqs[2];
  const int64_t l1 = l[0], l2 = l[1], l3 = l[2], r1 = r[0], r2 = r[1], r3 = r[2];
  const int l1_ok = l_ok[0], l2_ok = l_ok[1], l3_ok = l_ok[2], r1_ok = r_ok[0], r2_ok = r_ok[1], r3_ok = r_ok[2];
  for (int64_t j3 = i[2] - e3 + 1; j3 <= i[2] + e3; ++j3) {
    double lv3 = LVW(2, j3);
    for (int64_t j2 = i[1] - e2 + 1; j2 <= i[1] + e2; ++j2) {
      double lv2 = LVW(1, j2);
      double lv23 = lv2 * lv3;
      for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1)
        result += g->coefs[(j1 - 1) + (j2 - 1) * g->cstride[1] + (j3 - 1) * g->cstride[2]] * LVW(0, j1) * lv23;
      result += (l1_ok ? A3(1, l1, j2, j3) : 0.0) * lv23;
      result += (r1_ok ? A3(2, r1, j2, j3) : 0.0) * lv23;
    }
    for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1) {
      double lv1 = LVW(0, j1);
      result += (l2_ok ? A3(3, j1, l2, j3) : 0.0) * lv1 * lv3;
      result += (r2_ok ? A3(6, j1, r2, j3) : 0.0) * lv1 * lv3;
    }
    result += (l1_ok && l2_ok ? A3(4, l1, l2, j3) : 0.0) * lv3;
    result += (r1_ok && l2_ok ? A3(5, r1, l2, j3) : 0.0) * lv3;
    result += (l1_ok && r2_ok ? A3(7, l1, r2, j3) : 0.0) * lv3;
    result += (r1_ok && r2_ok ? A3(8, r1, r2, j3) : 0.0) * lv3;
  }
  for (int64_t j2 = i[1] - e2 + 1; j2 <= i[1] + e2; ++j2) {
    double lv2 = LVW(1, j2);
    for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1) {
      double lv1 = LVW(0, j1);
      result += (l3_ok ? A3(9, j1, j2, l3) : 0.0) * lv1 * lv2;
      result += (r3_ok ? A3(18, j1, j2, r3) : 0.0) * lv1 * lv2;
    }
  }
  for (int64_t j2 = i[1] - e2 + 1; j2 <= i[1] + e2; ++j2) {
    double lv2 = LVW(1, j2);
    result += (l1_ok && l3_ok ? A3(1 + 9, l1, j2, l3) : 0.0) * lv2;
    result += (r1_ok && l3_ok ? A3(2 + 9, r1, j2, l3) : 0.0) * lv2;
    result += (l1_ok && r3_ok ? A3(1 + 18, l1, j2, r3) : 0.0) * lv2;
    result += (r1_ok && r3_ok ? A3(2 + 18, r1, j2, r3) : 0.0) * lv2;
  }
  for (int64_t j1 = i[0] - e1 + 1; j1 <= i[0] + e1; ++j1) {
    double lv1 = LVW(0, j1);
    result += (l2_ok && l3_ok ? A3(3 + 9, j1, l2, l3) : 0.0) * lv1;
    result += (r2_ok && l3_ok ? A3(6 + 9, j1, r2, l3) : 0.0) * lv1;
    result += (l2_ok && r3_ok ? A3(3 + 18, j1, l2, r3) : 0.0) * lv1;
    result += (r2_ok && r3_ok ? A3(6 + 18, j1, r2, r3) : 0.0) * lv1;
  }
  result += (l1_ok && l2_ok && l3_ok ? A3(1 + 3 + 9, l1, l2, l3) : 0.0);
  result += (r1_ok && l2_ok && l3_ok ? A3(2 + 3 + 9, r1, l2, l3) : 0.0);
  result += (l1_ok && r2_ok && l3_ok ? A3(1 + 6 + 9, l1, r2, l3) : 0.0);
  result += (l1_ok && l2_ok && r3_ok ? A3(1 + 3 + 18, l1, l2, r3) : 0.0);
  result += (r1_ok && r2_ok && l3_ok ? A3(2 + 6 + 9, r1, r2, l3) : 0.0);
  result += (r1_ok && l2_ok && r3_ok ? A3(2 + 3 + 18, r1, l2, r3) : 0.0);
  result += (l1_ok && r2_ok && r3_ok ? A3(1 + 6 + 18, l1, r2, r3) : 0.0);
  result += (r1_ok && r2_ok && r3_ok ? A3(2 + 6 + 18, r1, r2, r3) : 0.0);
  *res = result * g->h[0] * g->h[1] * g->h[2];
  return ORC_OK;
#undef LVW
#undef A1
#undef A2
#undef A3
}

/* FastMixedIntegralOrder functors, N <= 3: convolution_fast_mixed_1_integral_dim.jl:11-131 (one integral axis),
 * convolution_fast_mixed_2_integral_dims.jl (two): Iterators.product over the stencils of the non-integral axes (first
 * axis fastest) with the :linear kernel product, inside it the K~ centre loop(s) and the tail lookups of the integral
 * axes; scale = prod(h_d over integral axes, (-1/h_d)^DO_d elsewhere), accumulated from one(T) in axis order */
static int itg_core_mixed(const orc_itg *g, const double *x, double *res) {
  const int N = g->ndim;
  int64_t i[3] = {1, 1, 1}, l[3] = {0, 0, 0}, r[3] = {0, 0, 0}, idx_sg[3] = {0, 0, 0};
  int l_ok[3] = {0, 0, 0}, r_ok[3] = {0, 0, 0}, id[2] = {-1, -1}, nid = 0;
  double t_sg[3] = {0, 0, 0};
  for (int d = 0; d < N; ++d) {
    double i_float = (x[d] - g->knots[d][0]) / g->h[d] + 1.0;
    if (!(fabs(i_float) < 9.0e18)) return ORC_ENAN;
    int64_t v = (int64_t)floor(i_float);
    if (v < g->eqs[d]) v = g->eqs[d];
    if (v > g->cshape[d] - g->eqs[d]) v = g->cshape[d] - g->eqs[d];
    i[d] = v;
    l[d] = v - g->eqs[d];
    r[d] = v + g->eqs[d] + 1;
    l_ok[d] = l[d] >= 1;
    r_ok[d] = r[d] <= g->cshape[d];
    if (g->deriv[d] == -1) {
      id[nid++] = d;
    } else {
      double diff_right = 1.0 - (x[d] - g->knots[d][v - 1]) / g->h[d];
      double cidx = diff_right * (double)(g->n_lin[d] - 1) + 1.0;
      if (!(fabs(cidx) < 9.0e18)) return ORC_ENAN;
      int64_t k = (int64_t)floor(cidx);
      if (k < 1) k = 1;
      if (k > g->n_lin[d] - 1) k = g->n_lin[d] - 1;
      idx_sg[d] = k;
      t_sg[d] = cidx - (double)k;
    }
  }
#define LVW(d, j) (eval_sg_kt(g, d, j, x[d]) - g->lv[d][(j)-1])
#define AT(code, a) (g->tail[code][((a)[0] - 1) + ((a)[1] - 1) * g->cstride[1] + ((a)[2] - 1) * g->cstride[2]])
  static const int P3[3] = {1, 3, 9};
  double result = 0.0;
  int64_t lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1}, c[3];
  for (int d = 0; d < N; ++d)
    if (g->deriv[d] != -1) { lo[d] = i[d] - g->eqs[d] + 1; hi[d] = i[d] + g->eqs[d]; }
  for (c[2] = lo[2]; c[2] <= hi[2]; ++c[2])
    for (c[1] = lo[1]; c[1] <= hi[1]; ++c[1])
      for (c[0] = lo[0]; c[0] <= hi[0]; ++c[0]) {
        double kt_prod = 1.0;
        for (int d = 0; d < N; ++d) {
          if (g->deriv[d] == -1) continue;
          int64_t col = c[d] - i[d] + g->eqs[d];
          double k0 = TAB(g->lin[d], idx_sg[d], col, g->eqs[d]), k1 = TAB(g->lin[d], idx_sg[d] + 1, col, g->eqs[d]);
          kt_prod *= (1.0 - t_sg[d]) * k0 + t_sg[d] * k1;
        }
        int64_t a[3] = {c[0], c[1], c[2]};
        if (nid == 1) {
          const int D = id[0];
          for (int64_t j = i[D] - g->eqs[D] + 1; j <= i[D] + g->eqs[D]; ++j) {
            double lv = LVW(D, j);
            a[D] = j;
            result += g->coefs[(a[0] - 1) + (a[1] - 1) * g->cstride[1] + (a[2] - 1) * g->cstride[2]] * lv * kt_prod;
          }
          double tl = 0.0, tr = 0.0;
          if (l_ok[D]) { a[D] = l[D]; tl = AT(1 * P3[D], a); }
          if (r_ok[D]) { a[D] = r[D]; tr = AT(2 * P3[D], a); }
          result += (tl + tr) * kt_prod;
        } else {
          const int D1 = id[0], D2 = id[1];
          for (int64_t j2 = i[D2] - g->eqs[D2] + 1; j2 <= i[D2] + g->eqs[D2]; ++j2) {
            double lv2 = LVW(D2, j2);
            a[D2] = j2;
            for (int64_t j1 = i[D1] - g->eqs[D1] + 1; j1 <= i[D1] + g->eqs[D1]; ++j1) {
              double lv1 = LVW(D1, j1);
              a[D1] = j1;
              result += g->coefs[(a[0] - 1) + (a[1] - 1) * g->cstride[1] + (a[2] - 1) * g->cstride[2]] * lv1 * lv2 * kt_prod;
            }
            double tl1 = 0.0, tr1 = 0.0;
            if (l_ok[D1]) { a[D1] = l[D1]; tl1 = AT(1 * P3[D1], a); }
            if (r_ok[D1]) { a[D1] = r[D1]; tr1 = AT(2 * P3[D1], a); }
            result += (tl1 + tr1) * lv2 * kt_prod;
          }
          for (int64_t j1 = i[D1] - g->eqs[D1] + 1; j1 <= i[D1] + g->eqs[D1]; ++j1) {
            double lv1 = LVW(D1, j1);
            a[D1] = j1;
            double tl2 = 0.0, tr2 = 0.0;
            if (l_ok[D2]) { a[D2] = l[D2]; tl2 = AT(1 * P3[D2], a); }
            if (r_ok[D2]) { a[D2] = r[D2]; tr2 = AT(2 * P3[D2], a); }
            result += (tl2 + tr2) * lv1 * kt_prod;
          }
          double c_ll = 0.0, c_rl = 0.0, c_lr = 0.0, c_rr = 0.0;
          if (l_ok[D1] && l_ok[D2]) { a[D1] = l[D1]; a[D2] = l[D2]; c_ll = AT(1 * P3[D1] + 1 * P3[D2], a); }
          if (r_ok[D1] && l_ok[D2]) { a[D1] = r[D1]; a[D2] = l[D2]; c_rl = AT(2 * P3[D1] + 1 * P3[D2], a); }
          if (l_ok[D1] && r_ok[D2]) { a[D1] = l[D1]; a[D2] = r[D2]; c_lr = AT(1 * P3[D1] + 2 * P3[D2], a); }
          if (r_ok[D1] && r_ok[D2]) { a[D1] = r[D1]; a[D2] = r[D2]; c_rr = AT(2 * P3[D1] + 2 * P3[D2], a); }
          result += (c_ll + c_rl + c_lr + c_rr) * kt_prod;
        }
      }
  double scale = 1.0;
  for (int d = 0; d < N; ++d)
    scale *= g->deriv[d] == -1 ? g->h[d] : jl_pow_int(-1.0 / g->h[d], g->deriv[d]);
  *res = result * scale;
  return ORC_OK;
#undef LVW
#undef AT
}

/* extrapolation_operator.jl:31-63 + extrapolate_point.jl:41-160 around the integral functor */
static int itg_etp_point(const orc_itg *g, const double *x, double *res) {
  const int N = g->ndim;
  double lo[3], hi[3], cl[3];
  int within = 1;
  for (int d = 0; d < N; ++d) {
    lo[d] = g->knots[d][g->eqs[d] - 1];
    hi[d] = g->knots[d][g->cshape[d] - 1 - (g->eqs[d] - 1)];
  }
  for (int d = 0; d < N; ++d)
    if (x[d] > hi[d] || x[d] < lo[d]) { within = 0; break; }
  if (within) return itg_core(g, x, res);
  for (int d = 0; d < N; ++d) cl[d] = x[d] < lo[d] ? lo[d] : (x[d] > hi[d] ? hi[d] : x[d]);
  if (g->extrap == ORC_EXTRAP_LINE) {
    double dx[3], dist_sq = 0.0;
    for (int d = 0; d < N; ++d) {
      dx[d] = x[d] - cl[d];
      dist_sq += dx[d] * dx[d];
    }
    if (dist_sq < 2.220446049250313e-16) return itg_core(g, cl, res);
    double h_min = g->h[0];
    for (int d = 1; d < N; ++d) h_min = g->h[d] < h_min ? g->h[d] : h_min;
    double eps = h_min / 100.0, dist = sqrt(dist_sq), xin[3], f_b, f_in;
    for (int d = 0; d < N; ++d) xin[d] = cl[d] - eps * dx[d] / dist;
    int rc = itg_core(g, cl, &f_b);
    if (rc) return rc;
    rc = itg_core(g, xin, &f_in);
    if (rc) return rc;
    *res = f_b + (f_b - f_in) / eps * dist;
    return ORC_OK;
  }
  if (g->extrap == ORC_EXTRAP_FLAT) return itg_core(g, cl, res);
  return ORC_EOOB;
}

int orc_itg_eval(const orc_itg *g, int64_t nq, const double *const *x, double *out, int64_t *first_bad) {
  int64_t bad = -1;
  int bad_rc = ORC_OK;
#pragma omp parallel for schedule(static) num_threads(g_threads)
  for (int64_t q = 0; q < nq; ++q) {
    double xq[3] = {0, 0, 0}, r = NAN;
    for (int d = 0; d < g->ndim; ++d) xq[d] = x[d][q];
    int rc = itg_etp_point(g, xq, &r);
    out[q] = r;
    if (rc) {
#pragma omp critical
      {
        if (bad < 0 || q < bad) { bad = q; bad_rc = rc; }
      }
    }
  }
  if (first_bad) *first_bad = bad;
  return bad_rc;
}

/* largest |value| the evaluators can ever read from a right-hand tail (indices r >= 2 eqs + 1 on the saturated axes):
 * identically zero - the property the CUDA path relies on when it keeps the left-hand tails only */
double orc_itg_right_tail_maxabs(const orc_itg *g) {
  double m = 0.0;
  int ncode = g->ndim == 1 ? 3 : (g->ndim == 2 ? 9 : 27);
  for (int code = 1; code < ncode; ++code) {
    int side[3] = {code % 3, (code / 3) % 3, (code / 9) % 3};
    if (side[0] != 2 && side[1] != 2 && side[2] != 2) continue;
    if (!g->tail[code]) continue;
    for (int64_t o = 0; o < g->ntot; ++o) {
      int ok = 1;
      for (int d = 0; d < g->ndim; ++d) {
        int64_t j = (o / g->cstride[d]) % g->cshape[d] + 1;
        if (side[d] == 2 && j < 2 * g->eqs[d] + 1) ok = 0;
      }
      if (ok && fabs(g->tail[code][o]) > m) m = fabs(g->tail[code][o]);
    }
  }
  return m;
}
/* copies left_values[d] / one tail array (code as in struct orc_itg) for tests */
int orc_itg_get(const orc_itg *g, int what, int idx, double *out) {
  if (what == 0) { memcpy(out, g->lv[idx], sizeof(double) * (size_t)g->cshape[idx]); return ORC_OK; }
  if (what == 1 && idx > 0 && idx < 27 && g->tail[idx]) { memcpy(out, g->tail[idx], sizeof(double) * (size_t)g->ntot); return ORC_OK; }
  return ORC_EARG;
}
