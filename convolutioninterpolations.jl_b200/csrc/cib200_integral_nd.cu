// cib200_integral_nd.cu - antiderivative / mixed interpolants in N >= 4 dimensions (the remainder).
//
// Replaces the per-element loop over the reference's three functors for such interpolants
//   convolution_fast_interpolation/convolution_fast_integration_nd.jl:1-28          every axis integral (no tails)
//   convolution_fast_interpolation/convolution_fast_mixed_3_integral_dims.jl:8-225  exactly three integral axes (tails)
//   convolution_fast_interpolation/convolution_fast_mixed_N_integral_dims.jl:15-82  four or more integral axes, mixed
// behind the extrapolation wrapper (extrapolation_operator.jl:31-63, extrapolate_point.jl:98-160).
//
// All three evaluate the same object,  sum_idx coefs[idx] prod_d w_d(idx_d; x_d),  with per-axis weights
//   integral axis:    w_d(j) = K~(s_j) - left_values_d[j]   (K~ through the :linear subgrid, +-1/2 outside the support)
//   other axes:       w_d(j) = K^(DO_d) at the 2 eqs stencil points, 0 elsewhere (:linear subgrid)
// over the ghost-padded grid.  The N-axis functor and the brute-force mixed functor do it literally (O(grid) per
// evaluation); the three-axis functor gets the saturated regions from prefix / suffix sums, which is the same sum in
// another association (oracle.IntegralInterpolantND restates it with all its tail arrays and agrees with the brute
// force to 1e-15).  Quirk kept: the N-axis functor loops over 1:domain_size[d] - the INTERIOR extents - while indexing
// the padded arrays, i.e. it never visits the last 2 (eqs - 1) padded entries of an axis; `sum_extent` carries that.
//
// One CTA per query: the weight vectors of all axes go to shared memory, then the CTA contracts the grid against them
// (axis 0 innermost, the other axes as a mixed-radix counter restricted to the non-zero ranges) and reduces.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <limits>
#include <vector>

#include "cib200_evalcore.cuh"

namespace cib {

static constexpr int IND_MAXND = CIB200_MAX_ND;

struct IntNdParams {
  int ndim, extrap;
  int nk[IND_MAXND];      // padded extents
  int eqs[IND_MAXND];
  int deriv[IND_MAXND];   // -1: integral axis
  int npre[IND_MAXND];
  int sum_extent[IND_MAXND];  // integral axes: leading padded entries the functor visits
  int woff[IND_MAXND];    // offset of the axis' weight vector in shared memory
  long long stride[IND_MAXND];
  const double *coefs;
  const double *knots[IND_MAXND];
  const double *tab[IND_MAXND];  // row-major n_pre x 2 eqs
  const double *lv[IND_MAXND];   // left_values of integral axes
  double h[IND_MAXND], lo[IND_MAXND], hi[IND_MAXND], knot0[IND_MAXND], anchor_knot[IND_MAXND];
  double scale, eps_line;
};

// (1 - t) T[i, col] + t T[i+1, col]: the :linear subgrid lookup shared by both kinds of axes
__device__ __forceinline__ double lin_lookup(const double *__restrict__ tab, int w, double cidx, int npre, int col) {
  long long i = fabs(cidx) < 9.0e18 ? (long long)floor(cidx) : 1;
  i = i < 1 ? 1 : (i > npre - 1 ? npre - 1 : i);
  const double t = __dsub_rn(cidx, (double)i);
  const double a = tab[(i - 1) * w + (col - 1)], b = tab[i * w + (col - 1)];
  return __dadd_rn(__dmul_rn(__dsub_rn(1.0, t), a), __dmul_rn(t, b));
}

// weight of padded index j (1-based) on axis d for coordinate x
__device__ __forceinline__ double axis_weight_nd(const IntNdParams &p, int d, int j, double x) {
  const int eqs = p.eqs[d], w = 2 * eqs;
  if (p.deriv[d] == -1) {
    if (j > p.sum_extent[d]) return 0.0;
    // eval_sg_kt, eval_subgrid_helpers.jl:22-32
    const double xjd = __dadd_rn(p.anchor_knot[d], __dmul_rn((double)(j - eqs), p.h[d]));
    const double sj = __ddiv_rn(__dsub_rn(x, xjd), p.h[d]);
    double kt;
    if (fabs(sj) >= (double)eqs) {
      kt = sj > 0 ? 0.5 : (sj < 0 ? -0.5 : 0.0);
    } else {
      const double col_float = __dadd_rn((double)eqs, sj);
      int col = (int)floor(col_float) + 1;
      col = col < 1 ? 1 : (col > w ? w : col);
      const double cidx = __dadd_rn(__dmul_rn(__dsub_rn(col_float, (double)(col - 1)), (double)(p.npre[d] - 1)), 1.0);
      kt = lin_lookup(p.tab[d], w, cidx, p.npre[d], col);
    }
    return __dsub_rn(kt, p.lv[d][j - 1]);
  }
  // derivative / interpolation axis: convolution_fast_mixed_N_integral_dims.jl:52-68
  const double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(x, p.knot0[d]), p.h[d]), 1.0);
  long long pos = fabs(i_float) < 9.0e18 ? (long long)floor(i_float) : eqs;
  pos = pos < eqs ? eqs : (pos > p.nk[d] - eqs ? p.nk[d] - eqs : pos);
  const int col = j - (int)pos + eqs;
  if (col < 1 || col > w) return 0.0;
  const double diff_right = __dsub_rn(1.0, __ddiv_rn(__dsub_rn(x, p.knots[d][pos - 1]), p.h[d]));
  const double cidx = __dadd_rn(__dmul_rn(diff_right, (double)(p.npre[d] - 1)), 1.0);
  return lin_lookup(p.tab[d], w, cidx, p.npre[d], col);
}

// block-wide: the functor at coordinates xs[0..ndim) (shared memory); every thread returns the value
__device__ double intnd_core(const IntNdParams &p, const double *xs, double *wsm, double *red) {
  const int nd = p.ndim, tid = threadIdx.x, nth = blockDim.x;
  __syncthreads();  // wsm / red free
  int lo[IND_MAXND], len[IND_MAXND];
  for (int d = 0; d < nd; ++d) {
    for (int j = tid; j < p.nk[d]; j += nth) wsm[p.woff[d] + j] = axis_weight_nd(p, d, j + 1, xs[d]);
    if (p.deriv[d] == -1) {
      lo[d] = 0;
      len[d] = p.sum_extent[d];
    } else {
      const double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(xs[d], p.knot0[d]), p.h[d]), 1.0);
      long long pos = fabs(i_float) < 9.0e18 ? (long long)floor(i_float) : p.eqs[d];
      pos = pos < p.eqs[d] ? p.eqs[d] : (pos > p.nk[d] - p.eqs[d] ? p.nk[d] - p.eqs[d] : pos);
      lo[d] = (int)pos - p.eqs[d];  // 0-based first stencil index: padded index pos - eqs + 1
      len[d] = 2 * p.eqs[d];
    }
  }
  __syncthreads();
  long long rows = 1;
  for (int d = 1; d < nd; ++d) rows *= len[d];
  double acc = 0.0;
  for (long long r = tid; r < rows; r += nth) {
    long long q = r, off = 0;
    double wprod = 1.0;
    for (int d = 1; d < nd; ++d) {
      const int i = lo[d] + (int)(q % len[d]);
      q /= len[d];
      off += (long long)i * p.stride[d];
      wprod *= wsm[p.woff[d] + i];
    }
    if (wprod == 0.0) continue;
    const double *c = p.coefs + off + lo[0];
    const double *w0 = wsm + p.woff[0] + lo[0];
    double a = 0.0;
    for (int i = 0; i < len[0]; ++i) a = fma(c[i], w0[i], a);
    acc = fma(a, wprod, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (nth >> 5); ++w) tot += red[w];
  return __dmul_rn(tot, p.scale);
}

struct IntNdCoords {
  const double *x[IND_MAXND];
  long long n[IND_MAXND];
  int grid;
  long long qoff;
};

__global__ void __launch_bounds__(128) intnd_eval_kernel(IntNdParams p, IntNdCoords c, long long nq, double *__restrict__ out,
                                                         EvalStatus *status) {
  extern __shared__ double sm[];
  double *xs = sm;               // [2][IND_MAXND]: boundary point, inward point
  double *red = sm + 2 * IND_MAXND;  // [4]
  double *wsm = red + 4;
  __shared__ int s_task;
  __shared__ double s_dist;
  for (long long q = blockIdx.x; q < nq; q += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      // classification: extrapolation_operator.jl:39-61, extrapolate_point.jl:98-160
      double x[IND_MAXND];
      long long r = q;
      bool inside = true, finite = true;
      for (int d = 0; d < p.ndim; ++d) {
        long long qd = q;
        if (c.grid) {
          qd = r % c.n[d];
          r /= c.n[d];
        }
        x[d] = c.x[d][qd];
        if (!(fabs(x[d]) < 1.0e300)) finite = false;
        if (x[d] > p.hi[d] || x[d] < p.lo[d]) inside = false;
      }
      int task = 1;
      double dist = 0.0;
      if (!finite) {
        task = 0;
        report(status, 2, q + c.qoff);
      } else if (!inside) {
        if (p.extrap == CIB200_THROW) {
          task = 0;
          report(status, 1, q + c.qoff);
        } else {
          double cl[IND_MAXND], dist_sq = 0.0;
          for (int d = 0; d < p.ndim; ++d) {
            cl[d] = x[d] < p.lo[d] ? p.lo[d] : (x[d] > p.hi[d] ? p.hi[d] : x[d]);
            const double dx = __dsub_rn(x[d], cl[d]);
            dist_sq = __dadd_rn(dist_sq, __dmul_rn(dx, dx));
          }
          if (p.extrap == CIB200_LINE && !(dist_sq < 2.220446049250313e-16)) {
            dist = __dsqrt_rn(dist_sq);
            for (int d = 0; d < p.ndim; ++d)
              xs[IND_MAXND + d] = __dsub_rn(cl[d], __ddiv_rn(__dmul_rn(p.eps_line, __dsub_rn(x[d], cl[d])), dist));
            task = 2;
          }
          for (int d = 0; d < p.ndim; ++d) x[d] = cl[d];
        }
      }
      for (int d = 0; d < p.ndim; ++d) xs[d] = x[d];
      s_task = task;
      s_dist = dist;
    }
    __syncthreads();
    const int task = s_task;
    double v = CUDART_NAN;
    if (task >= 1) {
      v = intnd_core(p, xs, wsm, red);
      if (task == 2) v = finish_line(v, intnd_core(p, xs + IND_MAXND, wsm, red), p.eps_line, s_dist);
    }
    if (threadIdx.x == 0) out[q] = v;
  }
}

}  // namespace cib

using namespace cib;

struct cib200_itgnd {
  int device;
  IntNdParams p;
  double *d_coefs;
  double *d_knots[IND_MAXND], *d_tabs[IND_MAXND], *d_lv[IND_MAXND];
  EvalStatus *d_status;
  cudaStream_t stream;
  size_t smem;
};

// _compute_left_values (constructor.jl:414-448), :linear subgrid, on the host from the row-major table
static void host_left_values(int eqs, int n, const double *tab, int npre, std::vector<double> &lv) {
  lv.assign((size_t)n, 0.0);
  const int w = 2 * eqs;
  for (int j = 1; j <= n; ++j) {
    const double sj = (double)(eqs - j);
    if (fabs(sj) >= (double)eqs) {
      lv[(size_t)j - 1] = sj > 0 ? 0.5 : (sj < 0 ? -0.5 : 0.0);
      continue;
    }
    const double col_float = (double)eqs + sj;
    int col = (int)floor(col_float) + 1;
    col = std::min(std::max(col, 1), w);
    const double cidx = (col_float - (double)(col - 1)) * (double)(npre - 1) + 1.0;
    long long i = (long long)floor(cidx);
    i = std::min<long long>(std::max<long long>(i, 1), npre - 1);
    const double t = cidx - (double)i;
    lv[(size_t)j - 1] = (1.0 - t) * tab[(i - 1) * w + (col - 1)] + t * tab[i * w + (col - 1)];
  }
}

extern "C" {

int32_t cib200_itgnd_destroy(cib200_itgnd *h) {
  if (!h) return CIB200_OK;
  if (h->d_coefs) cudaFree(h->d_coefs);
  for (int d = 0; d < IND_MAXND; ++d) {
    if (h->d_knots[d]) cudaFree(h->d_knots[d]);
    if (h->d_tabs[d]) cudaFree(h->d_tabs[d]);
    if (h->d_lv[d]) cudaFree(h->d_lv[d]);
  }
  if (h->d_status) cudaFree(h->d_status);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return CIB200_OK;
}

int32_t cib200_itgnd_create(int32_t ndim, const int64_t *coef_shape, const double *coefs, const double *const *knots,
                            const double *h, const int32_t *kernel, const int32_t *deriv, int32_t extrap,
                            const double *const *tables, const int32_t *n_pre, const int64_t *sum_extent,
                            cib200_itgnd **handle) {
  if (!coef_shape || !coefs || !knots || !h || !kernel || !deriv || !handle) {
    set_error("cib200_itgnd_create: null argument");
    return CIB200_EINVAL;
  }
  if (ndim < 1 || ndim > CIB200_MAX_ND || extrap < CIB200_THROW || extrap > CIB200_LINE) {
    set_error("cib200_itgnd_create: bad ndim / extrap");
    return CIB200_EINVAL;
  }
  cib200_itgnd *hd = new cib200_itgnd();
  memset(hd, 0, sizeof *hd);
  IntNdParams &p = hd->p;
  p.ndim = ndim;
  p.extrap = extrap;
  auto fail = [&](int rc) {
    cib200_itgnd_destroy(hd);
    return rc;
  };
  if (cudaGetDevice(&hd->device) != cudaSuccess) {
    set_error("cib200_itgnd_create: no usable CUDA device (there is no CPU fallback)");
    delete hd;
    return CIB200_ECUDA;
  }
  long long tot = 1;
  int woff = 0, nint = 0;
  double hmin = 0.0, scale = 1.0;
  for (int d = 0; d < ndim; ++d) {
    const int eqs = kernel_eqs(kernel[d]);
    if (eqs < 0 || coef_shape[d] < 2 * eqs || coef_shape[d] > (1 << 20) || !knots[d] || !(h[d] > 0)) {
      set_error("cib200_itgnd_create: bad axis %d", d);
      return fail(CIB200_EINVAL);
    }
    if (deriv[d] < -1 || deriv[d] > kernel_max_deriv_table(kernel[d])) {
      set_error("cib200_itgnd_create: derivative %d on axis %d has no table", deriv[d], d);
      return fail(CIB200_EUNSUPPORTED);
    }
    nint += deriv[d] == -1;
    p.eqs[d] = eqs;
    p.deriv[d] = deriv[d];
    p.nk[d] = (int)coef_shape[d];
    p.sum_extent[d] = (int)(sum_extent ? std::min<int64_t>(std::max<int64_t>(sum_extent[d], 0), coef_shape[d]) : coef_shape[d]);
    p.stride[d] = tot;
    tot *= coef_shape[d];
    p.woff[d] = woff;
    woff += p.nk[d];
    p.h[d] = h[d];
    p.knot0[d] = knots[d][0];
    p.anchor_knot[d] = knots[d][eqs - 1];  // itp.knots[d][eqs_d]
    p.lo[d] = knots[d][eqs - 1];
    p.hi[d] = knots[d][coef_shape[d] - 1 - (eqs - 1)];
    hmin = d == 0 ? h[d] : std::min(hmin, h[d]);
    scale *= deriv[d] == -1 ? h[d] : jl_pow_int(-1.0 / h[d], deriv[d]);  // mixed_N :71-79 ; nd: prod(itp.h)
  }
  if (nint == 0) {
    set_error("cib200_itgnd_create: no integral axis (use cib200_itpnd_create)");
    return fail(CIB200_EINVAL);
  }
  p.scale = scale;
  p.eps_line = hmin / 100.0;
  cudaError_t e = cudaMalloc((void **)&hd->d_coefs, sizeof(double) * (size_t)tot);
  if (e == cudaSuccess) e = cudaMemcpy(hd->d_coefs, coefs, sizeof(double) * (size_t)tot, cudaMemcpyHostToDevice);
  for (int d = 0; d < ndim && e == cudaSuccess; ++d) {
    e = cudaMalloc((void **)&hd->d_knots[d], sizeof(double) * (size_t)coef_shape[d]);
    if (e == cudaSuccess) e = cudaMemcpy(hd->d_knots[d], knots[d], sizeof(double) * (size_t)coef_shape[d], cudaMemcpyHostToDevice);
    // the axis' table: the caller's generated one (Julia memory order n_pre x 2 eqs, column-major) or the shipped 101 rows
    const bool gen = tables && n_pre && tables[d] && n_pre[d] >= 2;
    p.npre[d] = gen ? n_pre[d] : CIB_TAB_ROWS;
    const int w = 2 * p.eqs[d];
    std::vector<double> rm((size_t)p.npre[d] * w);
    if (gen) {
      for (int r = 0; r < p.npre[d]; ++r)
        for (int cc = 0; cc < w; ++cc) rm[(size_t)r * w + cc] = tables[d][(size_t)cc * p.npre[d] + r];
    } else {
      const double *src = deriv[d] == -1 ? host_table_i1(kernel[d]) : host_table(kernel[d], deriv[d]);
      if (!src) {
        set_error("cib200_itgnd_create: no shipped table for axis %d", d);
        return fail(CIB200_EUNSUPPORTED);
      }
      memcpy(rm.data(), src, sizeof(double) * rm.size());
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&hd->d_tabs[d], sizeof(double) * rm.size());
    if (e == cudaSuccess) e = cudaMemcpy(hd->d_tabs[d], rm.data(), sizeof(double) * rm.size(), cudaMemcpyHostToDevice);
    if (deriv[d] == -1 && e == cudaSuccess) {
      std::vector<double> lv;
      host_left_values(p.eqs[d], p.nk[d], rm.data(), p.npre[d], lv);
      e = cudaMalloc((void **)&hd->d_lv[d], sizeof(double) * lv.size());
      if (e == cudaSuccess) e = cudaMemcpy(hd->d_lv[d], lv.data(), sizeof(double) * lv.size(), cudaMemcpyHostToDevice);
    }
    p.knots[d] = hd->d_knots[d];
    p.tab[d] = hd->d_tabs[d];
    p.lv[d] = hd->d_lv[d];
  }
  if (e == cudaSuccess) e = cudaMalloc((void **)&hd->d_status, sizeof(EvalStatus));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hd->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    set_error("cib200_itgnd_create: %s", cudaGetErrorString(e));
    return fail(e == cudaErrorMemoryAllocation ? CIB200_ENOMEM : CIB200_ECUDA);
  }
  p.coefs = hd->d_coefs;
  hd->smem = sizeof(double) * (size_t)(2 * IND_MAXND + 4 + woff);
  if (hd->smem > 200 * 1024) {
    set_error("cib200_itgnd_create: the axes' weight vectors do not fit in shared memory");
    return fail(CIB200_EUNSUPPORTED);
  }
  *handle = hd;
  return CIB200_OK;
}

static int32_t itgnd_eval_common(cib200_itgnd *h, const IntNdCoords &c_in, long long nq, double *out, int32_t memkind,
                                 void *stream, int64_t *first_bad, const char *who) {
  if (first_bad) *first_bad = -1;
  if (nq == 0) return CIB200_OK;
  int dev = -1;
  CIB_CUDA(cudaGetDevice(&dev));
  if (dev != h->device) CIB_CUDA(cudaSetDevice(h->device));
  const int nd = h->p.ndim;
  cudaStream_t st = memkind == CIB200_DEVICE ? (cudaStream_t)stream : h->stream;
  IntNdCoords c = c_in;
  std::vector<void *> tmp;
  auto cleanup = [&]() {
    for (void *p : tmp) cudaFree(p);
    if (dev != h->device) cudaSetDevice(dev);
  };
  double *d_out = out;
  if (memkind == CIB200_HOST) {
    for (int d = 0; d < nd; ++d) {
      const size_t n = (size_t)(c_in.grid ? c_in.n[d] : nq);
      void *p = nullptr;
      if (cudaMalloc(&p, sizeof(double) * std::max<size_t>(n, 1)) != cudaSuccess) {
        cleanup();
        set_error("%s: out of device memory", who);
        return CIB200_ENOMEM;
      }
      tmp.push_back(p);
      cudaMemcpyAsync(p, c_in.x[d], sizeof(double) * n, cudaMemcpyHostToDevice, st);
      c.x[d] = (const double *)p;
    }
    void *p = nullptr;
    if (cudaMalloc(&p, sizeof(double) * (size_t)nq) != cudaSuccess) {
      cleanup();
      set_error("%s: out of device memory", who);
      return CIB200_ENOMEM;
    }
    tmp.push_back(p);
    d_out = (double *)p;
  } else if (memkind != CIB200_DEVICE) {
    set_error("%s: unknown memkind %d", who, memkind);
    return CIB200_EINVAL;
  }
  cudaMemsetAsync(h->d_status, 0xff, sizeof(EvalStatus), st);
  if (h->smem > 48 * 1024) cudaFuncSetAttribute(intnd_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem);
  const unsigned grid = (unsigned)std::min<long long>(nq, 8LL * sm_count());
  intnd_eval_kernel<<<grid, 128, h->smem, st>>>(h->p, c, nq, d_out, h->d_status);
  CIB_LAUNCHED();
  cudaError_t e = cudaGetLastError();
  int32_t rc = CIB200_OK;
  if (e == cudaSuccess && memkind == CIB200_HOST)
    e = cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)nq, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && (memkind == CIB200_HOST || first_bad || h->p.extrap == CIB200_THROW)) {
    EvalStatus hs;
    e = cudaMemcpyAsync(&hs, h->d_status, sizeof hs, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
      const unsigned long long none = std::numeric_limits<unsigned long long>::max();
      if (hs.first_oob != none || hs.first_nan != none) {
        const bool nan_first = hs.first_nan < hs.first_oob;
        rc = nan_first ? CIB200_ENAN : CIB200_EOOB;
        const long long idx = (long long)(nan_first ? hs.first_nan : hs.first_oob);
        if (first_bad) *first_bad = idx;
        set_error("%s: query %lld %s", who, idx, nan_first ? "has a NaN/Inf coordinate" : "is outside the interpolation domain (extrap=Throw)");
      }
    }
  }
  if (e != cudaSuccess) {
    set_error("%s: %s", who, cudaGetErrorString(e));
    rc = CIB200_ECUDA;
  }
  cleanup();
  return rc;
}

int32_t cib200_itgnd_eval(cib200_itgnd *h, int64_t nq, const void *const *x, void *out, int32_t memkind, void *stream,
                          int64_t *first_bad) {
  if (!h || nq < 0 || (nq > 0 && (!x || !out))) {
    set_error("cib200_itgnd_eval: bad arguments");
    return CIB200_EINVAL;
  }
  IntNdCoords c;
  memset(&c, 0, sizeof c);
  for (int d = 0; d < h->p.ndim; ++d) {
    if (nq > 0 && !x[d]) {
      set_error("cib200_itgnd_eval: x[%d] is null", d);
      return CIB200_EINVAL;
    }
    c.x[d] = (const double *)x[d];
    c.n[d] = 1;
  }
  return itgnd_eval_common(h, c, nq, (double *)out, memkind, stream, first_bad, "cib200_itgnd_eval");
}

int32_t cib200_itgnd_eval_grid(cib200_itgnd *h, const int64_t *n, const void *const *x, void *out, int32_t memkind,
                               void *stream, int64_t *first_bad) {
  if (!h || !n || !x || !out) {
    set_error("cib200_itgnd_eval_grid: bad arguments");
    return CIB200_EINVAL;
  }
  IntNdCoords c;
  memset(&c, 0, sizeof c);
  long long nq = 1;
  for (int d = 0; d < h->p.ndim; ++d) {
    if (n[d] < 0 || (n[d] > 0 && !x[d])) {
      set_error("cib200_itgnd_eval_grid: bad axis %d", d);
      return CIB200_EINVAL;
    }
    c.x[d] = (const double *)x[d];
    c.n[d] = n[d];
    nq *= n[d];
  }
  c.grid = 1;
  return itgnd_eval_common(h, c, nq, (double *)out, memkind, stream, first_bad, "cib200_itgnd_eval_grid");
}

}  // extern "C"
