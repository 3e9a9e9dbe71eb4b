// cib200_eval.cu - K1: batched point evaluation of the fast uniform interpolant (direct strategy).
//
// Replaces the per-element Julia loop
//   (etp::ConvolutionExtrapolation)(x...)            convolution_extrapolate/extrapolation_operator.jl:31-63
//   -> extrapolate_point                             convolution_extrapolate/extrapolate_point.jl:41-160
//   -> (itp::FastConvolutionInterpolation)(x...)     convolution_fast_interpolation_eager_{1d,2d,3d}.jl,
//                                                    ..._perdim_{2d,3d}.jl   (cubic subgrid branches)
//
// Two kernels:
//   eval_thread_kernel<T,NDIM,E>  one thread per query (1-D, 2-D, and 3-D with mixed kernels)
//   eval_warp3d_kernel<T,E>       3-D: a warp owns a tile of 32 queries.  Phase 1: every lane loads
//       one query (coalesced SoA loads), classifies it (inside / Flat / Line / Throw) and computes the
//       cell + table-row positions of its 1 or 2 evaluation points.  Phase 2: the warp walks the 32
//       queries; for each, lanes are laid out as (tap l along the contiguous axis) x (G rows), so one
//       load instruction fetches G complete 2E-wide stencil rows; per-axis Hermite-blended weights
//       live in per-warp shared memory; partial sums are combined with warp shuffles.
//
// Arithmetic: positions exactly as the reference (see cib200_device.cuh); the stencil sum uses the
// factorised form (per-axis blended weights, FMA) - same polynomial, different association; measured
// deviation from the literal 64-accumulator form is a few ulp of sum|c w| (tests/test_parity_*.py).
#include <math_constants.h>

#include "cib200_device.cuh"

namespace cib {

struct Plan {
  int ntask;  // 0: failed (err set) or out of range of the batch, 1: one evaluation, 2: Line (boundary, inward)
  int err;    // 0 none, 1 OOB (Throw), 2 NaN/Inf
  double dist;
};

// classification + positions of one query: extrapolation_operator.jl:39-61, extrapolate_point.jl:98-160
template <int NDIM>
__device__ __forceinline__ Plan make_plan(const ItpParams &p, const double *x, AxisPos *pa, AxisPos *pb) {
  Plan pl;
  pl.ntask = 0;
  pl.err = 0;
  pl.dist = 0.0;
  bool inside = true;
#pragma unroll
  for (int d = 0; d < NDIM; ++d)
    if (x[d] > p.hi[d] || x[d] < p.lo[d]) inside = false;
  bool ok = true;
  if (inside) {
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
      ok &= locate_from_knot(x[d], __ldg(p.knots[d]), p.h[d], p.eqs[d], p.nk[d], p.knots[d], pa[d]);
    pl.ntask = 1;
  } else if (p.extrap == CIB200_THROW) {
    pl.err = 1;
    return pl;
  } else {
    double cl[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) cl[d] = x[d] < p.lo[d] ? p.lo[d] : (x[d] > p.hi[d] ? p.hi[d] : x[d]);
    double dist_sq = 0.0;
    double dx[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      dx[d] = __dsub_rn(x[d], cl[d]);
      dist_sq = __dadd_rn(dist_sq, __dmul_rn(dx[d], dx[d]));
    }
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
      ok &= locate_from_knot(cl[d], __ldg(p.knots[d]), p.h[d], p.eqs[d], p.nk[d], p.knots[d], pa[d]);
    pl.ntask = 1;
    if (p.extrap == CIB200_LINE && !(dist_sq < 2.220446049250313e-16)) {
      double dist = __dsqrt_rn(dist_sq);
#pragma unroll
      for (int d = 0; d < NDIM; ++d) {
        double xin = __dsub_rn(cl[d], __ddiv_rn(__dmul_rn(p.eps_line, dx[d]), dist));
        ok &= locate_from_knot(xin, __ldg(p.knots[d]), p.h[d], p.eqs[d], p.nk[d], p.knots[d], pb[d]);
      }
      pl.ntask = 2;
      pl.dist = dist;
    }
  }
  if (!ok) {
    pl.ntask = 0;
    pl.err = 2;
  }
  return pl;
}

// final scale, associated like the reference (eager_1d.jl:113-114, eager_2d.jl:212, eager_3d.jl:304,
// perdim_2d.jl:24-30/214, perdim_3d.jl:159)
struct Scale {
  double s[3];
  int perdim2d;
};
template <int NDIM>
__device__ __forceinline__ double apply_scale(double r, const Scale &sc) {
  if (NDIM == 1) return __dmul_rn(r, sc.s[0]);
  if (NDIM == 2) {
    if (sc.perdim2d) return __dmul_rn(r, __dmul_rn(__dmul_rn(1.0, sc.s[0]), sc.s[1]));
    return __dmul_rn(__dmul_rn(r, sc.s[0]), sc.s[1]);
  }
  return __dmul_rn(__dmul_rn(__dmul_rn(r, sc.s[0]), sc.s[1]), sc.s[2]);
}

__device__ __forceinline__ double finish_line(double fb, double fin, double eps, double dist) {
  double slope = __ddiv_rn(__dsub_rn(fb, fin), eps);  // extrapolate_point.jl:146-148
  return __dadd_rn(fb, __dmul_rn(slope, dist));
}

__device__ __forceinline__ void report(EvalStatus *st, int err, long long q) {
  if (err == 1) atomic_min_u64(&st->first_oob, (unsigned long long)q);
  if (err == 2) atomic_min_u64(&st->first_nan, (unsigned long long)q);
}

template <int NDIM>
__device__ __forceinline__ void decode_coords(const Coords &c, long long q, long long *qd) {
  if (!c.grid) return;
  long long r = q;
#pragma unroll
  for (int d = 0; d < NDIM; ++d) {
    qd[d] = r % c.n[d];
    r /= c.n[d];
  }
}

// ---------------------------------------------------------------------------------------------
// thread-per-query kernel
// ---------------------------------------------------------------------------------------------
template <typename T, int NDIM, int E>
__device__ __forceinline__ double core_thread(const ItpParams &p, const double *const *tk, const double *const *td,
                                              const AxisPos *pos) {
  const T *__restrict__ coefs = (const T *)p.coefs;
  if (E > 0) {
    constexpr int W = 2 * (E > 0 ? E : 1);
    double w[NDIM][W];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      Hermite hb = hermite_basis(pos[d].t);
#pragma unroll
      for (int l = 0; l < W; ++l) w[d][l] = blended_weight(hb, tk[d], td[d], pos[d].row, l, W);
    }
    if (NDIM == 1) {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < W; ++l) acc = fma((double)coefs[pos[0].base + l], w[0][l], acc);
      return acc;
    }
    if (NDIM == 2) {
      double acc = 0.0;
      const T *c0 = coefs + pos[0].base + (long long)pos[1].base * p.s1;
#pragma unroll 2
      for (int m = 0; m < W; ++m) {
        double a = 0.0;
#pragma unroll
        for (int l = 0; l < W; ++l) a = fma((double)c0[l + (long long)m * p.s1], w[0][l], a);
        acc = fma(a, w[NDIM > 1 ? 1 : 0][m], acc);
      }
      return acc;
    }
    double acc = 0.0;
    const T *c0 = coefs + pos[0].base + (long long)pos[NDIM > 1 ? 1 : 0].base * p.s1 +
                  (long long)pos[NDIM > 2 ? 2 : 0].base * p.s2;
    for (int n = 0; n < W; ++n) {
      double an = 0.0;
      for (int m = 0; m < W; ++m) {
        double a = 0.0;
        const T *cr = c0 + (long long)m * p.s1 + (long long)n * p.s2;
#pragma unroll
        for (int l = 0; l < W; ++l) a = fma((double)cr[l], w[0][l], a);
        an = fma(a, w[NDIM > 1 ? 1 : 0][m], an);
      }
      acc = fma(an, w[NDIM > 2 ? 2 : 0][n], acc);
    }
    return acc;
  }
  // generic: per-axis eqs at run time, weights recomputed from the shared-memory tables on the fly
  Hermite hb[NDIM];
#pragma unroll
  for (int d = 0; d < NDIM; ++d) hb[d] = hermite_basis(pos[d].t);
  const int w0 = 2 * p.eqs[0];
  if (NDIM == 1) {
    double acc = 0.0;
    for (int l = 0; l < w0; ++l)
      acc = fma((double)coefs[pos[0].base + l], blended_weight(hb[0], tk[0], td[0], pos[0].row, l, w0), acc);
    return acc;
  }
  const int w1 = 2 * p.eqs[NDIM > 1 ? 1 : 0];
  if (NDIM == 2) {
    double acc = 0.0;
    const T *c0 = coefs + pos[0].base + (long long)pos[NDIM > 1 ? 1 : 0].base * p.s1;
    for (int m = 0; m < w1; ++m) {
      double a = 0.0;
      for (int l = 0; l < w0; ++l)
        a = fma((double)c0[l + (long long)m * p.s1], blended_weight(hb[0], tk[0], td[0], pos[0].row, l, w0), a);
      acc = fma(a, blended_weight(hb[NDIM > 1 ? 1 : 0], tk[1], td[1], pos[NDIM > 1 ? 1 : 0].row, m, w1), acc);
    }
    return acc;
  }
  const int w2 = 2 * p.eqs[NDIM > 2 ? 2 : 0];
  double acc = 0.0;
  const T *c0 = coefs + pos[0].base + (long long)pos[NDIM > 1 ? 1 : 0].base * p.s1 +
                (long long)pos[NDIM > 2 ? 2 : 0].base * p.s2;
  for (int n = 0; n < w2; ++n) {
    double an = 0.0;
    for (int m = 0; m < w1; ++m) {
      double a = 0.0;
      const T *cr = c0 + (long long)m * p.s1 + (long long)n * p.s2;
      for (int l = 0; l < w0; ++l)
        a = fma((double)cr[l], blended_weight(hb[0], tk[0], td[0], pos[0].row, l, w0), a);
      an = fma(a, blended_weight(hb[NDIM > 1 ? 1 : 0], tk[1], td[1], pos[NDIM > 1 ? 1 : 0].row, m, w1), an);
    }
    acc = fma(an, blended_weight(hb[NDIM > 2 ? 2 : 0], tk[2], td[2], pos[NDIM > 2 ? 2 : 0].row, n, w2), acc);
  }
  return acc;
}

template <typename T, int NDIM, int E>
__global__ void __launch_bounds__(256) eval_thread_kernel(ItpParams p, Coords c, Scale sc, T *__restrict__ out,
                                                          long long nq, EvalStatus *status) {
  extern __shared__ double smem[];
  const double *tk[3];
  const double *td[3];
  stage_tables(p, smem, tk, td);
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    long long qd[3] = {0, 0, 0};
    decode_coords<NDIM>(c, q, qd);
    double x[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) x[d] = load_coord<T>(c, d, q, qd[d]);
    AxisPos pa[NDIM], pb[NDIM];
    Plan pl = make_plan<NDIM>(p, x, pa, pb);
    double r = CUDART_NAN;
    if (pl.ntask >= 1) {
      double fb = apply_scale<NDIM>(core_thread<T, NDIM, E>(p, tk, td, pa), sc);
      r = fb;
      if (pl.ntask == 2) {
        double fin = apply_scale<NDIM>(core_thread<T, NDIM, E>(p, tk, td, pb), sc);
        r = finish_line(fb, fin, p.eps_line, pl.dist);
      }
    } else {
      report(status, pl.err, q + c.qoff);
    }
    out[q] = (T)r;
  }
}

// ---------------------------------------------------------------------------------------------
// 3-D warp-cooperative kernel
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ double ldg_as_double(const T *p) {
  return (double)__ldg(p);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int E>
__global__ void __launch_bounds__(256, 2) eval_warp3d_kernel(ItpParams p, Coords c, Scale sc, T *__restrict__ out,
                                                             long long nq, EvalStatus *status) {
  constexpr int W = 2 * E;           // taps per axis
  constexpr int G = 32 / W;          // stencil rows fetched per load instruction
  constexpr int ACTIVE = W * G;      // lanes that carry a tap
  constexpr int ROWS = W * W;        // (m, n) rows of the brick
  constexpr int ITERS = (ROWS + G - 1) / G;
  constexpr int CHUNK = ITERS < 36 ? ITERS : 32;

  extern __shared__ double smem[];
  const double *tk[3];
  const double *td[3];
  stage_tables(p, smem, tk, td);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *wyz = smem + tables_smem_doubles(p) + warp * (ROWS + 3 * W);
  double *wax = wyz + ROWS;  // [3][W] blended weights of the current evaluation point
  __syncthreads();

  const int l = lane % W, g = lane / W;
  const bool active = lane < ACTIVE;
  const T *__restrict__ coefs = (const T *)p.coefs;
  const long long s1 = p.s1, s2 = p.s2;
  const long long wrap_adj = s2 - (long long)W * s1;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long nbatch = (nq + 31) >> 5;

  for (long long b = (long long)blockIdx.x * (blockDim.x >> 5) + warp; b < nbatch; b += nwarps) {
    const long long q = (b << 5) + lane;
    // ---- phase 1: one query per lane ----------------------------------------------------------
    AxisPos pa[3], pb[3];
    Plan pl;
    pl.ntask = 0;
    pl.err = 0;
    pl.dist = 0.0;
    if (q < nq) {
      long long qd[3] = {0, 0, 0};
      decode_coords<3>(c, q, qd);
      double x[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) x[d] = load_coord<T>(c, d, q, qd[d]);
      pl = make_plan<3>(p, x, pa, pb);
    }
    double res_a = 0.0, res_b = 0.0;
    // ---- phase 2: the warp resolves the stencils one evaluation point at a time ---------------
    unsigned todo = __ballot_sync(0xffffffffu, pl.ntask > 0);
    while (todo) {
      const int j = __ffs(todo) - 1;
      todo &= todo - 1;
      const int nt = __shfl_sync(0xffffffffu, pl.ntask, j);
      for (int t = 0; t < nt; ++t) {
        int base[3], row[3];
        double tt[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          base[d] = __shfl_sync(0xffffffffu, t == 0 ? pa[d].base : pb[d].base, j);
          row[d] = __shfl_sync(0xffffffffu, t == 0 ? pa[d].row : pb[d].row, j);
          tt[d] = __shfl_sync(0xffffffffu, t == 0 ? pa[d].t : pb[d].t, j);
        }
        for (int idx = lane; idx < 3 * W; idx += 32) {
          const int a = idx / W, col = idx - a * W;
          const double ta = a == 0 ? tt[0] : (a == 1 ? tt[1] : tt[2]);
          const int ra = a == 0 ? row[0] : (a == 1 ? row[1] : row[2]);
          Hermite hb = hermite_basis(ta);
          wax[idx] = blended_weight(hb, a == 0 ? tk[0] : (a == 1 ? tk[1] : tk[2]),
                                    a == 0 ? td[0] : (a == 1 ? td[1] : td[2]), ra, col, W);
        }
        __syncwarp();
        for (int r = lane; r < ROWS; r += 32) wyz[r] = wax[W + r % W] * wax[2 * W + r / W];
        const double wx = active ? wax[l] : 0.0;
        __syncwarp();
        const T *ptr = coefs + base[0] + l + ((long long)base[1] + g) * s1 + (long long)base[2] * s2;
        double acc = 0.0;
#pragma unroll 1
        for (int it0 = 0; it0 < ITERS; it0 += CHUNK) {
          double cv[CHUNK];
#pragma unroll
          for (int u = 0; u < CHUNK; ++u) {
            const int r0 = (it0 + u) * G;  // row of lane group 0 in this step
            const int m0 = r0 % W, n0 = r0 / W;
            const bool wrap = m0 + g >= W;
            const bool valid = active && (it0 + u) < ITERS && (r0 + g) < ROWS;
            const long long off = (long long)m0 * s1 + (long long)n0 * s2 + (wrap ? wrap_adj : 0);
            cv[u] = valid ? ldg_as_double(ptr + off) : 0.0;
          }
#pragma unroll
          for (int u = 0; u < CHUNK; ++u) {
            const int r = (it0 + u) * G + g;
            const double wv = (active && (it0 + u) < ITERS && r < ROWS) ? wyz[r] : 0.0;
            acc = fma(cv[u], wv, acc);
          }
        }
        const double total = warp_sum(acc * wx);
        if (lane == j) {
          if (t == 0) res_a = total;
          else res_b = total;
        }
        __syncwarp();
      }
    }
    // ---- finish: one result per lane, coalesced store -----------------------------------------
    if (q < nq) {
      double r = CUDART_NAN;
      if (pl.ntask >= 1) {
        double fb = apply_scale<3>(res_a, sc);
        r = fb;
        if (pl.ntask == 2) r = finish_line(fb, apply_scale<3>(res_b, sc), p.eps_line, pl.dist);
      } else {
        report(status, pl.err, q + c.qoff);
      }
      out[q] = (T)r;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------------
static double jl_pow_int(double x, long long n) {  // Julia Base pow_body (math.jl), restated
  if (n == 0) return 1.0;
  if (n == 1) return x;
  if (n == 2) return x * x;
  if (n == 3) return x * x * x;
  double y = 1.0, xnlo = -0.0, ynlo = 0.0;
  while (n > 1) {
    if (n & 1) {
      double err = fma(y, xnlo, x * ynlo);
      double pr = x * y;
      ynlo = fma(x, y, -pr);
      y = pr;
      ynlo += err;
    }
    double err = x * 2 * xnlo;
    double pr = x * x;
    xnlo = fma(x, x, -pr);
    x = pr;
    xnlo += err;
    n >>= 1;
  }
  double err = fma(y, xnlo, x * ynlo);
  return fma(x, y, err);
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CIB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return CIB200_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int NDIM, int E>
static int launch_thread(const ItpParams &p, const Coords &c, const Scale &sc, void *out, long long nq,
                         EvalStatus *st, cudaStream_t stream) {
  size_t smem = sizeof(double) * (size_t)tables_smem_doubles(p);
  auto kern = eval_thread_kernel<T, NDIM, E>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  long long blocks = (nq + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 256, smem, stream>>>(p, c, sc, (T *)out, nq, st);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

template <typename T, int E>
static int launch_warp3d(const ItpParams &p, const Coords &c, const Scale &sc, void *out, long long nq,
                         EvalStatus *st, cudaStream_t stream) {
  constexpr int W = 2 * E;
  size_t smem = sizeof(double) * ((size_t)tables_smem_doubles(p) + 8 * (W * W + 3 * W));
  auto kern = eval_warp3d_kernel<T, E>;
  int rc = set_smem(kern, smem);
  if (rc) return rc;
  long long nbatch = (nq + 31) / 32;
  long long blocks = (nbatch + 7) / 8;
  long long cap = (long long)sm_count() * 2 * 4;  // 2 resident CTAs / SM, a few waves of persistence
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 256, smem, stream>>>(p, c, sc, (T *)out, nq, st);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

template <typename T>
static int launch_eval_t(const cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st) {
  const ItpParams &p = h->p;
  Scale sc;
  for (int d = 0; d < 3; ++d) sc.s[d] = d < p.ndim ? jl_pow_int(-1.0 / p.h[d], h->deriv[d]) : 1.0;
  sc.perdim2d = (p.ndim == 2 && h->perdim) ? 1 : 0;
  if (nq <= 0) return CIB200_OK;
  const bool uniform = (p.ndim == 1) || (p.ndim == 2 && p.eqs[0] == p.eqs[1]) ||
                       (p.ndim == 3 && p.eqs[0] == p.eqs[1] && p.eqs[1] == p.eqs[2]);
  const int e = uniform ? p.eqs[0] : 0;
  if (p.ndim == 1) {
    if (e == 5) return launch_thread<T, 1, 5>(p, c, sc, out, nq, h->d_status, st);
    if (e == 9) return launch_thread<T, 1, 9>(p, c, sc, out, nq, h->d_status, st);
    return launch_thread<T, 1, 0>(p, c, sc, out, nq, h->d_status, st);
  }
  if (p.ndim == 2) {
    if (e == 5) return launch_thread<T, 2, 5>(p, c, sc, out, nq, h->d_status, st);
    if (e == 9) return launch_thread<T, 2, 9>(p, c, sc, out, nq, h->d_status, st);
    return launch_thread<T, 2, 0>(p, c, sc, out, nq, h->d_status, st);
  }
  switch (e) {
    case 5: return launch_warp3d<T, 5>(p, c, sc, out, nq, h->d_status, st);
    case 6: return launch_warp3d<T, 6>(p, c, sc, out, nq, h->d_status, st);
    case 7: return launch_warp3d<T, 7>(p, c, sc, out, nq, h->d_status, st);
    case 8: return launch_warp3d<T, 8>(p, c, sc, out, nq, h->d_status, st);
    case 9: return launch_warp3d<T, 9>(p, c, sc, out, nq, h->d_status, st);
    default: return launch_thread<T, 3, 0>(p, c, sc, out, nq, h->d_status, st);
  }
}

int launch_eval(const cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st) {
  if (h->dtype == CIB200_F32) return launch_eval_t<float>(h, c, out, nq, st);
  return launch_eval_t<double>(h, c, out, nq, st);
}

}  // namespace cib
