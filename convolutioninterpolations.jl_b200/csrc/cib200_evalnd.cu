// cib200_evalnd.cu - batched point evaluation of the fast uniform interpolant in N >= 4 dimensions.
//
// Replaces the per-element loop over the N-D functors
//   HigherOrderKernel (one kernel, one derivative order on all axes)
//       convolution_fast_interpolation/convolution_fast_interpolation_eager_nd.jl:97-140
//   HigherOrderMixedKernel (per-axis kernels / derivative orders)
//       convolution_fast_interpolation/convolution_fast_interpolation_perdim_nd.jl:4-36
// behind the extrapolation wrapper (extrapolation_operator.jl:31-99, extrapolate_point.jl:98-160).
// For N >= 4 the reference resolves every axis to the :linear subgrid (_build_subgrids,
// fast_convolution_interpolation_constructor.jl:456-494); with precompute = 101 the generated tables are the shipped
// K^(d) tables, so a stencil weight is (1 - t) K[row, col] + t K[row + 1, col].
//
// One warp owns one query: lane d < N resolves axis d (bounds test, clamp, cell, table row, t); the (2 eqs)^N
// stencil is walked as  outer axes (3..N-1, warp-uniform odometer)  x  a shared-memory table of the (axis 1, axis 2)
// offsets and weight products  x  axis 0 across the lanes (G = 32 / (2 eqs_0) complete rows per load instruction:
// the rows of the coefficient grid are read as contiguous segments).  The tensor-product weights are applied in the
// factorised order (axis 0 last) - the same polynomial as the reference's sum of coef * prod_d w_d, associated
// differently (a few ulp of sum |c w|).
#include <math_constants.h>
#include <string.h>

#include <limits>

#include <algorithm>
#include <vector>

#include "cib200_device.cuh"

namespace cib {

constexpr int MAXND = CIB200_MAX_ND;
constexpr int ND_WMAX = 18;                  // taps per axis of the widest kernel (b13)
constexpr int ND_RMAX = ND_WMAX * ND_WMAX;   // entries of the (axis 1, axis 2) table

struct NdParams {
  const void *coefs;
  int ndim, perdim, extrap;
  int nk[MAXND], eqs[MAXND];
  long long stride[MAXND];       // element strides of the padded grid (column-major)
  const double *knots[MAXND];    // device, extended knots
  const double *tab[MAXND];      // device, K^(deriv_d) of axis d, npre[d] x 2 eqs_d
  int npre[MAXND];               // 101 (shipped) or the generated length of the :linear axes (>= 10 000 in the reference)
  double h[MAXND], lo[MAXND], hi[MAXND], knot0[MAXND];
  double scale, eps_line;
};
struct NdCoords {
  const void *x[MAXND];
  long long n[MAXND];
  int grid;
  long long qoff;
};

template <typename T>
__global__ void __launch_bounds__(128) eval_nd_kernel(const NdParams p, const NdCoords c, T *__restrict__ out,
                                                      long long nq, EvalStatus *st) {
  constexpr int WPB = 4;
  __shared__ double s_w[WPB][MAXND][ND_WMAX];
  __shared__ double s_wm[WPB][ND_RMAX];
  __shared__ long long s_om[WPB][ND_RMAX];
  __shared__ double s_t[WPB][MAXND];
  __shared__ int s_row[WPB][MAXND];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nd = p.ndim;
  const T *__restrict__ coefs = (const T *)p.coefs;
  double(*w)[ND_WMAX] = s_w[wib];
  double *wm = s_wm[wib];
  long long *om = s_om[wib];
  const int W0 = 2 * p.eqs[0], W1 = 2 * p.eqs[1], W2 = 2 * p.eqs[2];
  const int G = 32 / W0, R = W1 * W2;
  const int l0 = lane % W0, g0 = lane / W0;
  const bool active = g0 < G;
  int n_outer = 1;
  for (int d = 3; d < nd; ++d) n_outer *= 2 * p.eqs[d];
  // this lane's axis (lanes >= nd mirror axis 0 and are ignored)
  const int ax = lane < nd ? lane : 0;
  const double lo = p.lo[ax], hi = p.hi[ax], hh = p.h[ax], k0 = p.knot0[ax];
  const int eq = p.eqs[ax], nk = p.nk[ax];
  const int npre = p.perdim ? p.npre[ax] : p.npre[0];  // eager_nd.jl reads itp.pre_range[1] on every axis
  const double *__restrict__ kn = p.knots[ax];
  const long long sax = p.stride[ax];
  const unsigned axmask = nd >= 32 ? 0xffffffffu : ((1u << nd) - 1u);

  const long long nwarps = (long long)gridDim.x * WPB;
  for (long long q = (long long)blockIdx.x * WPB + wib; q < nq; q += nwarps) {
    // ---- coordinate of this lane's axis ----
    long long qi = q;
    if (c.grid) {
      long long r = q;
      for (int d = 0; d < ax; ++d) r /= c.n[d];
      qi = r % c.n[ax];
    }
    const double x = (double)((const T *)c.x[ax])[qi];
    // ---- classification: extrapolation_operator.jl:39-61, extrapolate_point.jl:98-160 ----
    const bool outside = (x > hi) || (x < lo);
    const bool inside = (__ballot_sync(0xffffffffu, outside) & axmask) == 0u;
    int err = 0, ntask = 1;
    double dist = 0.0, cl = x, dx = 0.0;
    if (!inside) {
      if (p.extrap == CIB200_THROW) {
        err = 1;
      } else {
        cl = x < lo ? lo : (x > hi ? hi : x);
        dx = __dsub_rn(x, cl);
        if (p.extrap == CIB200_LINE) {
          const double dx2 = __dmul_rn(dx, dx);
          double dist_sq = 0.0;
          for (int d = 0; d < nd; ++d) dist_sq = __dadd_rn(dist_sq, __shfl_sync(0xffffffffu, dx2, d));
          if (!(dist_sq < 2.220446049250313e-16)) {
            ntask = 2;
            dist = __dsqrt_rn(dist_sq);
          }
        }
      }
    }
    double res[2] = {0.0, 0.0};
    if (!err) {
      for (int task = 0; task < ntask; ++task) {
        const double xe = task == 0 ? cl : __dsub_rn(cl, __ddiv_rn(__dmul_rn(p.eps_line, dx), dist));
        AxisPos pos;
        const bool ok = p.perdim ? locate_from_knot(xe, k0, hh, eq, nk, kn, pos, npre) : locate_prerange(xe, k0, hh, eq, nk, kn, pos, npre);
        if ((__ballot_sync(0xffffffffu, !ok) & axmask) != 0u) {
          err = 2;
          break;
        }
        // base offset of the stencil corner: exact integer sum over the axes
        long long boff = lane < nd ? (long long)pos.base * sax : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) boff += __shfl_xor_sync(0xffffffffu, boff, o);
        __syncwarp();
        if (lane < nd) {
          s_t[wib][lane] = pos.t;
          s_row[wib][lane] = pos.row;
        }
        __syncwarp();
        // per-axis stencil weights: (1 - t) K[row, col] + t K[row + 1, col]   (eager_nd.jl:133-135, perdim_nd.jl:30-32)
        for (int idx = lane; idx < nd * ND_WMAX; idx += 32) {
          const int d = idx / ND_WMAX, col = idx - d * ND_WMAX;
          const int wd = 2 * p.eqs[d];
          if (col < wd) {
            const double t = s_t[wib][d];
            const double *kr = p.tab[d] + s_row[wib][d] * wd + col;
            w[d][col] = __dadd_rn(__dmul_rn(__dsub_rn(1.0, t), __ldg(kr)), __dmul_rn(t, __ldg(kr + wd)));
          }
        }
        __syncwarp();
        for (int r = lane; r < R; r += 32) {
          const int m = r % W1, n = r / W1;
          om[r] = (long long)m * p.stride[1] + (long long)n * p.stride[2];
          wm[r] = w[1][m] * w[2][n];
        }
        __syncwarp();
        const T *__restrict__ cb = coefs + boff + l0;
        double acc = 0.0;
        for (int o = 0; o < n_outer; ++o) {
          int rem = o;
          long long oo = 0;
          double wo = 1.0;
          for (int d = 3; d < nd; ++d) {
            const int wd = 2 * p.eqs[d];
            const int i = rem % wd;
            rem /= wd;
            oo += (long long)i * p.stride[d];
            wo *= w[d][i];
          }
          double a = 0.0;
          if (active) {
            const T *__restrict__ co = cb + oo;
            for (int r = g0; r < R; r += G) a = fma((double)__ldg(co + om[r]), wm[r], a);
          }
          acc = fma(a, wo, acc);
        }
        acc = active ? acc * w[0][l0] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        res[task] = __dmul_rn(acc, p.scale);
      }
    }
    if (lane == 0) {
      if (err) {
        if (err == 1) atomic_min_u64(&st->first_oob, (unsigned long long)(q + c.qoff));
        if (err == 2) atomic_min_u64(&st->first_nan, (unsigned long long)(q + c.qoff));
        out[q] = (T)CUDART_NAN;
      } else if (ntask == 2) {
        const double slope = __ddiv_rn(__dsub_rn(res[0], res[1]), p.eps_line);  // extrapolate_point.jl:146-148
        out[q] = (T)__dadd_rn(res[0], __dmul_rn(slope, dist));
      } else {
        out[q] = (T)res[0];
      }
    }
  }
}

}  // namespace cib

// opaque handle behind cib200_itpnd*
struct cib200_itpnd {
  int device, dtype;
  cib::NdParams p;
  long long cshape[cib::MAXND];
  void *d_coefs;
  double *d_knots[cib::MAXND];
  double *d_tabs[cib::MAXND];
  cib::EvalStatus *d_status;
  cudaStream_t stream;  // host-memory calls
};

using namespace cib;

static size_t nd_esize(int dtype) { return dtype == CIB200_F32 ? 4 : 8; }

static int nd_launch(cib200_itpnd *h, const NdCoords &c, void *out, long long nq, cudaStream_t st) {
  if (nq <= 0) return CIB200_OK;
  const long long want = (nq + 3) / 4;
  const int grid = (int)std::min<long long>(want, (long long)sm_count() * 16);
  if (h->dtype == CIB200_F32)
    eval_nd_kernel<float><<<grid, 128, 0, st>>>(h->p, c, (float *)out, nq, h->d_status);
  else
    eval_nd_kernel<double><<<grid, 128, 0, st>>>(h->p, c, (double *)out, nq, h->d_status);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

static int nd_status(cib200_itpnd *h, cudaStream_t st, int64_t *first_bad, const char *who) {
  EvalStatus hs;
  CIB_CUDA(cudaMemcpyAsync(&hs, h->d_status, sizeof hs, cudaMemcpyDeviceToHost, st));
  CIB_CUDA(cudaStreamSynchronize(st));
  const unsigned long long none = std::numeric_limits<unsigned long long>::max();
  int code = CIB200_OK;
  long long idx = -1;
  if (hs.first_oob != none || hs.first_nan != none) {
    if (hs.first_nan < hs.first_oob) {
      code = CIB200_ENAN;
      idx = (long long)hs.first_nan;
    } else {
      code = CIB200_EOOB;
      idx = (long long)hs.first_oob;
    }
    set_error("%s: query %lld %s", who, idx,
              code == CIB200_EOOB ? "is outside the interpolation domain (extrap=Throw)" : "has a NaN/Inf coordinate");
  }
  if (first_bad) *first_bad = idx;
  return code;
}

static int nd_eval_common(cib200_itpnd *h, NdCoords c, long long nq, void *out, int memkind, void *stream,
                          int64_t *first_bad, const char *who) {
  int dev;
  CIB_CUDA(cudaGetDevice(&dev));
  if (dev != h->device) CIB_CUDA(cudaSetDevice(h->device));
  const size_t es = nd_esize(h->dtype);
  const int nd = h->p.ndim;
  if (first_bad) *first_bad = -1;
  if (nq == 0) return CIB200_OK;
  if (memkind == CIB200_DEVICE) {
    cudaStream_t st = (cudaStream_t)stream;
    CIB_CUDA(cudaMemsetAsync(h->d_status, 0xff, sizeof(EvalStatus), st));
    int rc = nd_launch(h, c, out, nq, st);
    if (rc) return rc;
    if (first_bad || h->p.extrap == CIB200_THROW) return nd_status(h, st, first_bad, who);
    return CIB200_OK;
  }
  if (memkind != CIB200_HOST) {
    set_error("%s: unknown memkind %d", who, memkind);
    return CIB200_EINVAL;
  }
  cudaStream_t st = h->stream;
  CIB_CUDA(cudaMemsetAsync(h->d_status, 0xff, sizeof(EvalStatus), st));
  // host buffers: bounded chunks through one staging allocation (the N-D stencil makes the kernel, not PCIe, the cost)
  const long long chunk = c.grid ? nq : std::min<long long>(nq, 1LL << 22);
  size_t xbytes = 0;
  for (int d = 0; d < nd; ++d) xbytes += es * (size_t)(c.grid ? c.n[d] : chunk);
  char *stage = nullptr;
  CIB_CUDA(cudaMalloc((void **)&stage, xbytes + es * (size_t)chunk));
  int rc = CIB200_OK;
  for (long long q0 = 0; q0 < nq && !rc; q0 += chunk) {
    const long long m = std::min(chunk, nq - q0);
    NdCoords cc = c;
    char *dst = stage;
    for (int d = 0; d < nd; ++d) {
      const size_t nb = es * (size_t)(c.grid ? c.n[d] : m);
      cudaError_t e = cudaMemcpyAsync(dst, (const char *)c.x[d] + (c.grid ? 0 : es * (size_t)q0), nb, cudaMemcpyHostToDevice, st);
      if (e != cudaSuccess) {
        set_error("%s: H2D failed: %s", who, cudaGetErrorString(e));
        rc = CIB200_ECUDA;
        break;
      }
      cc.x[d] = dst;
      dst += es * (size_t)(c.grid ? c.n[d] : chunk);
    }
    if (rc) break;
    cc.qoff = q0;
    void *dout = stage + xbytes;
    rc = nd_launch(h, cc, dout, m, st);
    if (rc) break;
    cudaError_t e = cudaMemcpyAsync((char *)out + es * (size_t)q0, dout, es * (size_t)m, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("%s: D2H failed: %s", who, cudaGetErrorString(e));
      rc = CIB200_ECUDA;
    }
  }
  int code = rc ? rc : nd_status(h, st, first_bad, who);
  cudaFree(stage);
  return code;
}

extern "C" {

static int32_t itpnd_create_impl(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                                 const double *const *knots, const double *h, const int32_t *kernel, const int32_t *deriv,
                                 int32_t extrap, int32_t perdim, const double *const *tables, const int32_t *n_pre,
                                 cib200_itpnd **handle);
int32_t cib200_itpnd_create(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                            const double *const *knots, const double *h, const int32_t *kernel, const int32_t *deriv,
                            int32_t extrap, int32_t perdim, cib200_itpnd **handle) {
  return itpnd_create_impl(ndim, dtype, coef_shape, coefs, knots, h, kernel, deriv, extrap, perdim, nullptr, nullptr, handle);
}
int32_t cib200_itpnd_create_tables(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                                   const double *const *knots, const double *h, const int32_t *kernel,
                                   const int32_t *deriv, int32_t extrap, int32_t perdim, const double *const *tables,
                                   const int32_t *n_pre, cib200_itpnd **handle) {
  return itpnd_create_impl(ndim, dtype, coef_shape, coefs, knots, h, kernel, deriv, extrap, perdim, tables, n_pre, handle);
}
static int32_t itpnd_create_impl(int32_t ndim, int32_t dtype, const int64_t *coef_shape, const void *coefs,
                                 const double *const *knots, const double *h, const int32_t *kernel, const int32_t *deriv,
                                 int32_t extrap, int32_t perdim, const double *const *tables, const int32_t *n_pre,
                                 cib200_itpnd **handle) {
  if (!coef_shape || !coefs || !knots || !h || !kernel || !deriv || !handle) {
    set_error("cib200_itpnd_create: null argument");
    return CIB200_EINVAL;
  }
  if (ndim < 4 || ndim > CIB200_MAX_ND) {
    set_error("cib200_itpnd_create: ndim=%d outside 4..%d (1..3 dimensions use cib200_itp_create)", ndim, CIB200_MAX_ND);
    return CIB200_EUNSUPPORTED;
  }
  if (dtype != CIB200_F64 && dtype != CIB200_F32) {
    set_error("cib200_itpnd_create: unknown dtype %d", dtype);
    return CIB200_EINVAL;
  }
  if (extrap < CIB200_THROW || extrap > CIB200_LINE) {
    set_error("cib200_itpnd_create: unknown extrap %d", extrap);
    return CIB200_EINVAL;
  }
  cib200_itpnd *hd = new cib200_itpnd();
  memset(hd, 0, sizeof *hd);
  NdParams &p = hd->p;
  p.ndim = ndim;
  p.perdim = perdim ? 1 : 0;
  p.extrap = extrap;
  hd->dtype = dtype;
  auto fail = [&](int rc) {
    cib200_itpnd_destroy(hd);
    return rc;
  };
  if (cudaGetDevice(&hd->device) != cudaSuccess) {
    set_error("cib200_itpnd_create: no usable CUDA device (there is no CPU fallback)");
    delete hd;
    return CIB200_ECUDA;
  }
  long long tot = 1;
  double hmin = 0.0, scale = 1.0;
  for (int d = 0; d < ndim; ++d) {
    const int eqs = kernel_eqs(kernel[d]);
    if (eqs < 0) {
      set_error("cib200_itpnd_create: unknown kernel id %d on axis %d", kernel[d], d);
      return fail(CIB200_EUNSUPPORTED);
    }
    if (!perdim && (kernel[d] != kernel[0] || deriv[d] != deriv[0])) {
      set_error("cib200_itpnd_create: perdim = 0 needs one kernel and one derivative order on all axes (eager_nd.jl)");
      return fail(CIB200_EINVAL);
    }
    if (deriv[d] < 0 || deriv[d] > kernel_max_deriv_table(kernel[d]) || !host_table(kernel[d], deriv[d])) {
      set_error("cib200_itpnd_create: derivative %d on axis %d needs K^(%d), which is not a shipped table", deriv[d], d, deriv[d]);
      return fail(CIB200_EUNSUPPORTED);
    }
    if (coef_shape[d] < 2 * eqs || !knots[d] || !(h[d] > 0)) {
      set_error("cib200_itpnd_create: bad axis %d", d);
      return fail(CIB200_EINVAL);
    }
    p.eqs[d] = eqs;
    p.nk[d] = (int)coef_shape[d];
    hd->cshape[d] = coef_shape[d];
    p.stride[d] = tot;
    tot *= coef_shape[d];
    p.h[d] = h[d];
    p.knot0[d] = knots[d][0];
    p.lo[d] = knots[d][eqs - 1];                              // _domain_bounds, extrapolate_point.jl:172-176
    p.hi[d] = knots[d][coef_shape[d] - 1 - (eqs - 1)];
    hmin = d == 0 ? h[d] : std::min(hmin, h[d]);
    // eager_nd.jl:139 prod((-1/h_d)^DO_d) (left to right) ; perdim_nd.jl:35 _perdim_scale: s = 1; s *= (-1/h_d)^DO_d
    const double sd = jl_pow_int(-1.0 / h[d], deriv[d]);
    scale = (d == 0 && !perdim) ? sd : scale * sd;
  }
  p.scale = scale;
  p.eps_line = hmin / 100.0;  // extrapolate_point.jl:138-139
  const size_t es = nd_esize(dtype);
  cudaError_t e = cudaMalloc(&hd->d_coefs, es * (size_t)tot);
  if (e == cudaSuccess) e = cudaMemcpy(hd->d_coefs, coefs, es * (size_t)tot, cudaMemcpyHostToDevice);
  for (int d = 0; d < ndim && e == cudaSuccess; ++d) {
    e = cudaMalloc((void **)&hd->d_knots[d], sizeof(double) * (size_t)coef_shape[d]);
    if (e == cudaSuccess) e = cudaMemcpy(hd->d_knots[d], knots[d], sizeof(double) * (size_t)coef_shape[d], cudaMemcpyHostToDevice);
    // generated table of the axis (Julia memory order n_pre x 2 eqs, column-major = itp.kernel_pre[d]) or the shipped one
    const bool gen = tables && n_pre && tables[d] && n_pre[d] >= 2;
    p.npre[d] = gen ? n_pre[d] : CIB_TAB_ROWS;
    const size_t nel = (size_t)p.npre[d] * 2 * (size_t)p.eqs[d], tb = sizeof(double) * nel;
    if (e == cudaSuccess) e = cudaMalloc((void **)&hd->d_tabs[d], tb);
    if (e == cudaSuccess && gen) {
      std::vector<double> rm(nel);
      const int w = 2 * p.eqs[d];
      for (int r = 0; r < p.npre[d]; ++r)
        for (int cc = 0; cc < w; ++cc) rm[(size_t)r * w + cc] = tables[d][(size_t)cc * p.npre[d] + r];
      e = cudaMemcpy(hd->d_tabs[d], rm.data(), tb, cudaMemcpyHostToDevice);
    } else if (e == cudaSuccess) {
      e = cudaMemcpy(hd->d_tabs[d], host_table(kernel[d], deriv[d]), tb, cudaMemcpyHostToDevice);
    }
    p.knots[d] = hd->d_knots[d];
    p.tab[d] = hd->d_tabs[d];
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&hd->d_status, sizeof(EvalStatus));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hd->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cib200_itpnd_create: %s", cudaGetErrorString(e));
    return fail(e == cudaErrorMemoryAllocation ? CIB200_ENOMEM : CIB200_ECUDA);
  }
  p.coefs = hd->d_coefs;
  *handle = hd;
  return CIB200_OK;
}

int32_t cib200_itpnd_destroy(cib200_itpnd *h) {
  if (!h) return CIB200_OK;
  if (h->d_coefs) cudaFree(h->d_coefs);
  for (int d = 0; d < MAXND; ++d) {
    if (h->d_knots[d]) cudaFree(h->d_knots[d]);
    if (h->d_tabs[d]) cudaFree(h->d_tabs[d]);
  }
  if (h->d_status) cudaFree(h->d_status);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CIB200_OK;
}

int32_t cib200_itpnd_eval(cib200_itpnd *h, int64_t nq, const void *const *x, void *out, int32_t memkind, void *stream,
                          int64_t *first_bad) {
  if (!h || nq < 0 || (nq > 0 && (!x || !out))) {
    set_error("cib200_itpnd_eval: bad arguments");
    return CIB200_EINVAL;
  }
  NdCoords c;
  memset(&c, 0, sizeof c);
  for (int d = 0; d < h->p.ndim; ++d) {
    if (nq > 0 && !x[d]) {
      set_error("cib200_itpnd_eval: x[%d] is null", d);
      return CIB200_EINVAL;
    }
    c.x[d] = nq > 0 ? x[d] : nullptr;
    c.n[d] = 1;
  }
  return nd_eval_common(h, c, nq, out, memkind, stream, first_bad, "cib200_itpnd_eval");
}

int32_t cib200_itpnd_eval_grid(cib200_itpnd *h, const int64_t *n, const void *const *x, void *out, int32_t memkind,
                               void *stream, int64_t *first_bad) {
  if (!h || !n || !x || !out) {
    set_error("cib200_itpnd_eval_grid: bad arguments");
    return CIB200_EINVAL;
  }
  NdCoords c;
  memset(&c, 0, sizeof c);
  long long nq = 1;
  for (int d = 0; d < h->p.ndim; ++d) {
    if (n[d] < 0 || (n[d] > 0 && !x[d])) {
      set_error("cib200_itpnd_eval_grid: bad axis %d", d);
      return CIB200_EINVAL;
    }
    c.x[d] = x[d];
    c.n[d] = n[d];
    nq *= n[d];
  }
  c.grid = 1;
  return nd_eval_common(h, c, nq, out, memkind, stream, first_bad, "cib200_itpnd_eval_grid");
}

}  // extern "C"
