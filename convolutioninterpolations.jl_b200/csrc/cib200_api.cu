// cib200_api.cu - the C-ABI (include/cib200.h): handles, uploads, host-side coefficient builder,
// host<->device pipelining of batched evaluation.  No CPU evaluation path exists in this library.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "cib200_internal.cuh"
#include "cib200_tables.inc"

namespace cib {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

int kernel_eqs(int kernel) {
  switch (kernel) {
    case CIB200_B5: return 5;
    case CIB200_B7: return 6;
    case CIB200_B9: return 7;
    case CIB200_B11: return 8;
    case CIB200_B13: return 9;
    case CIB200_A3: return 2;
    case CIB200_A4: return 3;
    case CIB200_A5: return 3;
    case CIB200_A7: return 4;
    default: return -1;
  }
}
int kernel_max_deriv_table(int kernel) {
  switch (kernel) {
    case CIB200_B5: return 3;
    case CIB200_B7: return 4;
    case CIB200_B9: return 5;
    case CIB200_B11: return 6;
    case CIB200_B13: return 6;
    case CIB200_A3: case CIB200_A4: case CIB200_A5: case CIB200_A7: return 1;
    default: return -1;
  }
}
int kernel_max_eval_deriv(int kernel) {
  switch (kernel) {
    case CIB200_B5: return 2;
    case CIB200_B7: return 3;
    case CIB200_B9: return 4;
    case CIB200_B11: return 5;
    case CIB200_B13: return 5;
    case CIB200_A3: case CIB200_A4: case CIB200_A5: case CIB200_A7: return 0;
    default: return -1;
  }
}
// antiderivative tables K~ (KERNEL_PRE_<k>_i1, precomputed_kernel_tables.jl:93-96); companions are K^(0), K^(1)
const double *host_table_i1(int kernel) {
  switch (kernel) {
    case CIB200_A3: return CIB_TAB_a3_i1;
    case CIB200_A4: return CIB_TAB_a4_i1;
    case CIB200_A5: return CIB_TAB_a5_i1;
    case CIB200_A7: return CIB_TAB_a7_i1;
    case CIB200_B5: return CIB_TAB_b5_i1;
    case CIB200_B7: return CIB_TAB_b7_i1;
    case CIB200_B9: return CIB_TAB_b9_i1;
    case CIB200_B11: return CIB_TAB_b11_i1;
    case CIB200_B13: return CIB_TAB_b13_i1;
    default: return nullptr;
  }
}
const double *host_table(int kernel, int d) {
  static const double *const b5[] = {CIB_TAB_b5_d0, CIB_TAB_b5_d1, CIB_TAB_b5_d2, CIB_TAB_b5_d3};
  static const double *const b7[] = {CIB_TAB_b7_d0, CIB_TAB_b7_d1, CIB_TAB_b7_d2, CIB_TAB_b7_d3, CIB_TAB_b7_d4};
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
  if (d < 0 || d > kernel_max_deriv_table(kernel)) return nullptr;
  switch (kernel) {
    case CIB200_A3: return a3[d];
    case CIB200_A4: return a4[d];
    case CIB200_A5: return a5[d];
    case CIB200_A7: return a7[d];
    case CIB200_B5: return b5[d];
    case CIB200_B7: return b7[d];
    case CIB200_B9: return b9[d];
    case CIB200_B11: return b11[d];
    case CIB200_B13: return b13[d];
    default: return nullptr;
  }
}

// ---------------------------------------------------------------------------------------------
// Host-side ghost fill (north_star subsystem 1).  Restates, independently of the oracle,
// apply_boundary_conditions_for_dim! (convolution_coefs/create_convolutional_coefs.jl:200-370) with
// fill_ghost_points_polynomial! (:408-441); matrices of polynomial_boundary.jl:29-179 are the
// Lagrange extrapolation weights to x = -j from nodes 0..m-1 and are generated, not tabulated.
// Compiled with -ffp-contract=off so the mean / matrix-vector sums are plain mul+add.
// ---------------------------------------------------------------------------------------------
static void extrap_row(int m, int j, double *row) {
  for (int k = 0; k < m; ++k) {
    long long num = 1, den = 1;
    for (int i = 0; i < m; ++i)
      if (i != k) {
        num *= (long long)(-j - i);
        den *= (long long)(k - i);
      }
    row[k] = (double)(num / den);
  }
}

// minimum_polynomial_bc_points, convolution_interpolation_function.jl:183-195 = columns of the kernel's ghost matrix
int kernel_poly_nodes(int kernel) { return kernel >= CIB200_A3 ? 4 : (kernel == CIB200_B5 ? 6 : 8); }
static inline int poly_nodes(int kernel) { return kernel_poly_nodes(kernel); }

// decides the matrix width for one side of one line; p = first (left) / last (right) `ns` centred values
static int choose_nodes(const double *s, long long n, double h, int kernel, int bc, bool right) {
  const int ns = poly_nodes(kernel);
  bool use_poly = false;
  if (n >= ns && (bc == CIB200_BC_DETECT || bc == CIB200_BC_POLY)) {
    if (bc == CIB200_BC_POLY) {
      use_poly = true;
    } else {
      const double *p = right ? s + (n - ns) : s;
      double d1[8], d2[8];
      double mn = INFINITY, mx = -INFINITY;
      for (int d = 0; d < ns - 1; ++d) {
        d1[d] = p[d + 1] - p[d];
        mn = std::fmin(mn, d1[d]);
        mx = std::fmax(mx, d1[d]);
      }
      bool periodic = (mn * mx) / (h * h) < -1.0 / 10;
      mn = INFINITY, mx = -INFINITY;
      for (int d = 0; d < ns - 2; ++d) {
        d2[d] = d1[d + 1] - d1[d];
        mn = std::fmin(mn, d2[d]);
        mx = std::fmax(mx, d2[d]);
      }
      bool high_curv = std::fabs(mx - mn) / h > 1.0 / 2;
      bool high_d3 = false;
      if (ns > 4) {
        mn = INFINITY, mx = -INFINITY;
        for (int d = 0; d < ns - 3; ++d) {
          double v = d2[d + 1] - d2[d];
          mn = std::fmin(mn, v);
          mx = std::fmax(mx, v);
        }
        high_d3 = std::fabs(mx - mn) / h > 1.0 / 2;
      }
      use_poly = !periodic && !high_curv && !high_d3;
    }
  }
  if (use_poly) return ns;
  if (bc == CIB200_BC_QUADRATIC) return 3;
  return 2;  // :linear, failed :detect, too few points
}

template <typename T>
static void fill_line(T *line0, long long stride, long long n, double h, int eqs, int kernel, int bcl,
                      int bcr, double *slice) {
  double sum = 0.0;
  for (long long j = 0; j < n; ++j) {
    slice[j] = (double)line0[j * stride];
    sum += slice[j];
  }
  const double mean = sum / (double)n;
  for (long long j = 0; j < n; ++j) slice[j] -= mean;
  for (int side = 0; side < 2; ++side) {
    const int m = choose_nodes(slice, n, h, kernel, side ? bcr : bcl, side != 0);
    double y[8], row[8];
    for (int k = 0; k < m; ++k) y[k] = side ? slice[n - 1 - k] : slice[k];
    for (int j = 1; j <= eqs - 1; ++j) {
      extrap_row(m, j, row);
      double acc = 0.0;
      for (int k = 0; k < m; ++k) acc += row[k] * y[k];
      const double gval = mean + acc;
      if (side == 0) line0[-(long long)j * stride] = (T)gval;
      else line0[(n - 1 + j) * stride] = (T)gval;
    }
  }
}

template <typename T>
static int build_coefs_t(int ndim, const int64_t *n, const T *vs, const double *h, const int32_t *kernel,
                         const int32_t *bcl, const int32_t *bcr, T *c) {
  int eqs[CIB200_MAX_SWEEP_DIM];
  long long cs[CIB200_MAX_SWEEP_DIM], cstr[CIB200_MAX_SWEEP_DIM];
  long long ctot = 1, vtot = 1, nmax = 0;
  for (int d = 0; d < ndim; ++d) {
    eqs[d] = kernel_eqs(kernel[d]);
    if (eqs[d] < 0) {
      set_error("cib200_build_coefs: unknown kernel id %d on axis %d", kernel[d], d);
      return CIB200_EINVAL;
    }
    if (n[d] < 3) {
      set_error("cib200_build_coefs: axis %d has %lld points (< 3)", d, (long long)n[d]);
      return CIB200_EINVAL;
    }
    cs[d] = n[d] + 2 * (eqs[d] - 1);
    cstr[d] = ctot;
    ctot *= cs[d];
    vtot *= n[d];
    nmax = std::max<long long>(nmax, n[d]);
  }
  std::fill(c, c + ctot, (T)0);
  // interior copy, row of axis 0 at a time
  const long long rows = vtot / n[0];
  for (long long r = 0; r < rows; ++r) {
    long long rem = r, off = (long long)(eqs[0] - 1);
    for (int d = 1; d < ndim; ++d) {
      long long id = rem % n[d];
      rem /= n[d];
      off += (id + eqs[d] - 1) * cstr[d];
    }
    memcpy(c + off, vs + r * n[0], sizeof(T) * (size_t)n[0]);
  }
  unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
  for (int fd = 0; fd < ndim; ++fd) {
    const long long nlines = ctot / cs[fd];
    auto work = [&](int tid) {
      std::vector<double> slice((size_t)nmax);
      for (long long ln = tid; ln < nlines; ln += nthreads) {
        long long rem = ln, off = 0;
        for (int d = 0; d < ndim; ++d) {
          if (d == fd) continue;
          long long id = rem % cs[d];
          rem /= cs[d];
          off += id * cstr[d];
        }
        off += (long long)(eqs[fd] - 1) * cstr[fd];
        fill_line(c + off, cstr[fd], n[fd], h[fd], eqs[fd], kernel[fd], bcl[fd], bcr[fd], slice.data());
      }
    };
    if (nthreads == 1 || nlines < 64) {
      for (int t = 0; t < nthreads; ++t) work(t);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < nthreads; ++t) pool.emplace_back(work, t);
      for (auto &th : pool) th.join();
    }
  }
  return CIB200_OK;
}

// ---------------------------------------------------------------------------------------------
static int current_device(int *dev) {
  CIB_CUDA(cudaGetDevice(dev));
  return CIB200_OK;
}

static size_t esize(int dtype) { return dtype == CIB200_F32 ? 4 : 8; }

static int fetch_status(cib200_itp *h, cudaStream_t st, int64_t *first_bad, int *code) {
  EvalStatus hs;
  CIB_CUDA(cudaMemcpyAsync(&hs, h->d_status, sizeof hs, cudaMemcpyDeviceToHost, st));
  CIB_CUDA(cudaStreamSynchronize(st));
  const unsigned long long none = std::numeric_limits<unsigned long long>::max();
  *code = CIB200_OK;
  long long idx = -1;
  if (hs.first_oob != none || hs.first_nan != none) {
    if (hs.first_nan < hs.first_oob) {
      *code = CIB200_ENAN;
      idx = (long long)hs.first_nan;
    } else {
      *code = CIB200_EOOB;
      idx = (long long)hs.first_oob;
    }
  }
  if (first_bad) *first_bad = idx;
  return CIB200_OK;
}

static int reset_status(cib200_itp *h, cudaStream_t st) {
  CIB_CUDA(cudaMemsetAsync(h->d_status, 0xff, sizeof(EvalStatus), st));
  return CIB200_OK;
}

static int ensure_stage(cib200_itp *h, size_t bytes) {
  if (!h->streams[0]) {
    for (int i = 0; i < 3; ++i) CIB_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
    for (int i = 0; i < cib200_itp::NSTAGE; ++i) {
      CIB_CUDA(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      CIB_CUDA(cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
      CIB_CUDA(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
  }
  if (h->stage_bytes >= bytes) return CIB200_OK;
  for (int i = 0; i < cib200_itp::NSTAGE; ++i) {
    if (h->d_stage[i]) cudaFree(h->d_stage[i]);
    h->d_stage[i] = nullptr;
  }
  h->stage_bytes = 0;
  for (int i = 0; i < cib200_itp::NSTAGE; ++i) CIB_CUDA(cudaMalloc(&h->d_stage[i], bytes));
  h->stage_bytes = bytes;
  return CIB200_OK;
}

static int ensure_host_stage(cib200_itp *h, size_t bytes) {
  if (h->h_stage_bytes >= bytes) return CIB200_OK;
  for (int i = 0; i < cib200_itp::NSTAGE; ++i) {
    if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
    h->h_stage[i] = nullptr;
  }
  h->h_stage_bytes = 0;
  for (int i = 0; i < cib200_itp::NSTAGE; ++i) CIB_CUDA(cudaHostAlloc(&h->h_stage[i], bytes, cudaHostAllocDefault));
  h->h_stage_bytes = bytes;
  return CIB200_OK;
}

// is this host pointer pageable (neither cudaHostAlloc'ed nor cudaHostRegister'ed)?
static bool is_pageable(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

// memcpy split over a few host threads (one thread moves ~10 GB/s; a PCIe 5 x16 link wants ~50)
static void par_memcpy(void *dst, const void *src, size_t bytes) {
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned cap = 16;  // measured on a 16-core host: 8 threads 0.94e9, 16 threads 1.24e9 evals/s through pageable buffers
  if (const char *e = tune_env("CIB200_COPY_THREADS")) cap = (unsigned)std::max(1, atoi(e));
  const size_t nt = std::min<size_t>(std::min<unsigned>(cap, hw), bytes >> 22);  // >= 4 MB per thread
  if (nt <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
  for (size_t t = 0; t < nt; ++t) {
    const size_t o = t * per;
    if (o >= bytes) break;
    const size_t m = std::min(per, bytes - o);
    pool.emplace_back([=]() { memcpy((char *)dst + o, (const char *)src + o, m); });
  }
  for (auto &t : pool) t.join();
}

int prof_mark(cib200_itp *h, int i, cudaStream_t st) {
  if (!h->profiling) return CIB200_OK;
  if (!h->prof_ev[0])
    for (int k = 0; k < 5; ++k) CIB_CUDA(cudaEventCreate(&h->prof_ev[k]));
  CIB_CUDA(cudaEventRecord(h->prof_ev[i], st));
  if (i == 4) h->prof_valid = 1;
  return CIB200_OK;
}

static int eval_dispatch(cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st) {
  int mode = h->eval_mode;
  if (mode == CIB200_EVAL_AUTO) {
    // large point lists: bin by coefficient brick (HBM traffic ~130 B/query instead of ~18 kB for random
    // queries); small batches cannot amortise one brick load per tile and go straight to the gather kernel
    // measured crossover (tools/sweep_threshold.py, :b5): the binned pipeline pays one brick load per occupied tile and
    // five small launches, the gather kernel 18 kB of DRAM traffic per query - binned wins from ~100 queries per tile
    // (512^3: 2^21.5 queries, 2.8x at 2^22; 128^3: 2^17, 2.5x at 2^20)
    bool big = false;
    if (nq >= (1LL << 17) && !c.grid && binned_supported(h)) big = nq >= 96 * binned_tile_count(h);
    mode = big ? CIB200_EVAL_BINNED : CIB200_EVAL_GATHER;
  }
  if (mode == CIB200_EVAL_BINNED) {
    if (!binned_supported(h) || c.grid) {
      set_error("cib200_itp_eval: BINNED strategy needs a 3-D interpolant with one kernel on all axes and a point list");
      return CIB200_EUNSUPPORTED;
    }
    // records carry a 32-bit query index: batches of 2^32 queries or more are walked in 2^31-query slices on the
    // same stream (the slices share the handle's scratch, so they run back to back anyway)
    const long long slice = 1LL << 31;
    const size_t es = esize(h->dtype);
    for (long long q0 = 0; q0 < nq; q0 += slice) {
      Coords cs = c;
      for (int d = 0; d < 3; ++d)
        if (c.x[d]) cs.x[d] = (const char *)c.x[d] + es * (size_t)q0;
      cs.qoff = c.qoff + q0;
      const int rc = launch_eval_binned(h, cs, (char *)out + es * (size_t)q0, std::min(slice, nq - q0), st);
      if (rc) return rc;
    }
    return CIB200_OK;
  }
  for (int i = 0; i < 3; ++i) prof_mark(h, i, st);
  h->last_strategy = CIB200_STRATEGY_GATHER;
  const int rc = launch_eval(h, c, out, nq, st);
  prof_mark(h, 3, st);
  prof_mark(h, 4, st);
  return rc;
}

}  // namespace cib

using namespace cib;

extern "C" {

int32_t cib200_version(void) { return 100; }
const char *cib200_last_error(void) { return g_err; }
int64_t cib200_launch_count(void) { return g_launches.load(); }

const double *cib200_kernel_table(int32_t kernel, int32_t deriv, int32_t *eqs_out) {
  if (eqs_out) *eqs_out = kernel_eqs(kernel);
  if (deriv == -1) return host_table_i1(kernel);
  return host_table(kernel, deriv);
}

int32_t cib200_smooth_halfwidth(double B) {
  // convolution_gaussian_smoothing.jl:46
  return (int32_t)std::ceil(std::sqrt(-std::log(1e-12 / 2) / B));
}

int32_t cib200_extend_knots(const double *knots, int64_t n, int32_t n_extra, double *out) {
  if (!knots || !out || n < 2 || n_extra < 0) {
    set_error("cib200_extend_knots: bad arguments");
    return CIB200_EINVAL;
  }
  const double step_start = knots[1] - knots[0];
  const double step_end = knots[n - 1] - knots[n - 2];
  for (int i = n_extra; i >= 1; --i) out[n_extra - i] = knots[0] - (double)i * step_start;
  for (int64_t i = 0; i < n; ++i) out[n_extra + i] = knots[i];
  for (int i = 1; i <= n_extra; ++i) out[n_extra + n - 1 + i] = knots[n - 1] + (double)i * step_end;
  return CIB200_OK;
}

int32_t cib200_build_coefs(int32_t ndim, int32_t dtype, const int64_t *n, const void *values, const double *h,
                           const int32_t *kernel, const int32_t *bc_left, const int32_t *bc_right,
                           void *coefs_out) {
  if (ndim < 1 || ndim > CIB200_MAX_SWEEP_DIM || !n || !values || !h || !kernel || !bc_left || !bc_right ||
      !coefs_out) {
    set_error("cib200_build_coefs: bad arguments");
    return CIB200_EINVAL;
  }
  for (int d = 0; d < ndim; ++d)
    if (bc_left[d] < 0 || bc_left[d] > 3 || bc_right[d] < 0 || bc_right[d] > 3) {
      set_error("cib200_build_coefs: unknown boundary condition on axis %d", d);
      return CIB200_EINVAL;
    }
  if (dtype == CIB200_F32)
    return build_coefs_t<float>(ndim, n, (const float *)values, h, kernel, bc_left, bc_right, (float *)coefs_out);
  if (dtype == CIB200_F64)
    return build_coefs_t<double>(ndim, n, (const double *)values, h, kernel, bc_left, bc_right, (double *)coefs_out);
  set_error("cib200_build_coefs: unknown dtype %d", dtype);
  return CIB200_EINVAL;
}

static int32_t sweep_io(int ndim, int dtype, long long tot_in, long long tot_out, const void *values, void *out,
                        int memkind, cudaStream_t st, void **d_in, void **d_out, bool *owned);

int32_t cib200_build_coefs_device(int32_t ndim, int32_t dtype, const int64_t *n, const void *values, const double *h,
                                  const int32_t *kernel, const int32_t *bc_left, const int32_t *bc_right,
                                  void *coefs_out, int32_t memkind, void *stream) {
  if (ndim < 1 || ndim > CIB200_MAX_SWEEP_DIM || !n || !values || !h || !kernel || !bc_left || !bc_right ||
      !coefs_out) {
    set_error("cib200_build_coefs_device: bad arguments");
    return CIB200_EINVAL;
  }
  if (dtype != CIB200_F64 && dtype != CIB200_F32) {
    set_error("cib200_build_coefs_device: unknown dtype %d", dtype);
    return CIB200_EINVAL;
  }
  long long nn[CIB200_MAX_SWEEP_DIM], tin = 1, tout = 1;
  int k[CIB200_MAX_SWEEP_DIM], bl[CIB200_MAX_SWEEP_DIM], br[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) {
    const int eqs = kernel_eqs(kernel[d]);
    if (eqs < 0) {
      set_error("cib200_build_coefs_device: unknown kernel id %d on axis %d", kernel[d], d);
      return CIB200_EUNSUPPORTED;
    }
    if (n[d] < 3) {
      set_error("cib200_build_coefs_device: axis %d has %lld points (< 3)", d, (long long)n[d]);
      return CIB200_EINVAL;
    }
    if (bc_left[d] < 0 || bc_left[d] > 3 || bc_right[d] < 0 || bc_right[d] > 3) {
      set_error("cib200_build_coefs_device: unknown boundary condition on axis %d", d);
      return CIB200_EINVAL;
    }
    nn[d] = n[d];
    k[d] = kernel[d];
    bl[d] = bc_left[d];
    br[d] = bc_right[d];
    tin *= n[d];
    tout *= n[d] + 2 * (eqs - 1);
  }
  cudaStream_t st = (cudaStream_t)stream;
  void *d_in, *d_out;
  bool owned;
  int rc = sweep_io(ndim, dtype, tin, tout, values, coefs_out, memkind, st, &d_in, &d_out, &owned);
  if (rc) return rc;
  rc = sweep_build_coefs(ndim, dtype, nn, d_in, h, k, bl, br, d_out, st);
  if (owned) {
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(coefs_out, d_out, esize(dtype) * (size_t)tout, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        set_error("cib200_build_coefs_device: D2H failed: %s", cudaGetErrorString(e));
        rc = CIB200_ECUDA;
      }
    }
    cudaFree(d_in);
    cudaFree(d_out);
  }
  return rc;
}

static int32_t itp_create_impl(const cib200_itp_desc *desc, const double *const *tables, const int32_t *n_pre,
                               cib200_itp **handle);
int32_t cib200_itp_create(const cib200_itp_desc *desc, cib200_itp **handle) {
  return itp_create_impl(desc, nullptr, nullptr, handle);
}
int32_t cib200_itp_create_tables(const cib200_itp_desc *desc, const double *const *tables, const int32_t *n_pre,
                                 cib200_itp **handle) {
  return itp_create_impl(desc, tables, n_pre, handle);
}
static int32_t itp_create_impl(const cib200_itp_desc *desc, const double *const *tables, const int32_t *n_pre,
                               cib200_itp **handle) {
  if (!desc || !handle) {
    set_error("cib200_itp_create: null argument");
    return CIB200_EINVAL;
  }
  *handle = nullptr;
  if (desc->ndim < 1 || desc->ndim > CIB200_MAX_ITP_DIM) {
    set_error("cib200_itp_create: ndim=%d outside 1..3 (N>=4 uses the reference's linear-subgrid path)", desc->ndim);
    return CIB200_EUNSUPPORTED;
  }
  if (desc->dtype != CIB200_F64 && desc->dtype != CIB200_F32) {
    set_error("cib200_itp_create: unknown dtype %d", desc->dtype);
    return CIB200_EINVAL;
  }
  if (desc->extrap < CIB200_THROW || desc->extrap > CIB200_LINE) {
    set_error("cib200_itp_create: unknown extrapolation %d", desc->extrap);
    return CIB200_EINVAL;
  }
  if (!desc->coefs) {
    set_error("cib200_itp_create: coefs is null");
    return CIB200_EINVAL;
  }
  const int nd = desc->ndim;
  int eqs[3] = {1, 1, 1};
  for (int d = 0; d < nd; ++d) {
    eqs[d] = kernel_eqs(desc->kernel[d]);
    if (eqs[d] < 0) {
      set_error("cib200_itp_create: unknown kernel id %d on axis %d", desc->kernel[d], d);
      return CIB200_EUNSUPPORTED;
    }
    {
      // tables the subgrid needs: K^(d) (linear), + K^(d+1) (cubic), + K^(d+2) (quintic)
      const int sgd = desc->subgrid[d];
      const int need = desc->deriv[d] + (sgd == CIB200_SG_LINEAR ? 0 : (sgd == CIB200_SG_QUINTIC ? 2 : 1));
      if (desc->deriv[d] < 0 || need > kernel_max_deriv_table(desc->kernel[d])) {
        set_error("cib200_itp_create: derivative %d on axis %d needs K^(%d), which is not a shipped table", desc->deriv[d], d, need);
        return CIB200_EUNSUPPORTED;
      }
    }
    if (desc->subgrid[d] < CIB200_SG_CUBIC || desc->subgrid[d] > CIB200_SG_LINEAR) {
      set_error("cib200_itp_create: unknown subgrid %d on axis %d", desc->subgrid[d], d);
      return CIB200_EINVAL;
    }
    if (nd == 3 && (desc->subgrid[d] == CIB200_SG_LINEAR) != (desc->subgrid[0] == CIB200_SG_LINEAR)) {
      set_error("cib200_itp_create: a 3-D interpolant is all-cubic or all-linear (eager_3d.jl / perdim_3d.jl)");
      return CIB200_EUNSUPPORTED;
    }
    if (desc->subgrid[d] == CIB200_SG_QUINTIC) {
      if (nd > 2) {
        set_error("cib200_itp_create: the quintic subgrid exists for 1-D and 2-D interpolants only (3-D resolves to :linear)");
        return CIB200_EUNSUPPORTED;
      }
      if (desc->deriv[d] + 2 > kernel_max_deriv_table(desc->kernel[d])) {
        set_error("cib200_itp_create: quintic subgrid on axis %d needs K^(%d), which is not a shipped table", d, desc->deriv[d] + 2);
        return CIB200_EUNSUPPORTED;
      }
    }
    if (desc->coef_shape[d] < 2 * eqs[d] || desc->coef_shape[d] > 0x7fffffffLL || !desc->knots[d]) {
      set_error("cib200_itp_create: bad coef_shape / knots on axis %d", d);
      return CIB200_EINVAL;
    }
    if (!(desc->h[d] > 0)) {
      set_error("cib200_itp_create: h[%d] must be positive", d);
      return CIB200_EINVAL;
    }
  }
  int dev = 0;
  int rc = current_device(&dev);
  if (rc) return rc;
  // The stencil gather touches 80-byte rows at arbitrary offsets; a smaller L2->DRAM fetch granularity
  // avoids dragging in the rest of a 64/128-byte block for every partially used sector (a hint).
  if (const char *fg = tune_env("CIB200_L2_FETCH")) {
    size_t g = (size_t)atoi(fg);
    if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
  }

  cib200_itp *h = new cib200_itp();
  memset(h, 0, sizeof *h);
  h->mu = new std::mutex();
  h->device = dev;
  h->dtype = desc->dtype;
  h->eval_mode = CIB200_EVAL_AUTO;
  h->perdim = desc->perdim;
  ItpParams &p = h->p;
  p.ndim = nd;
  p.extrap = desc->extrap;
  long long total = 1;
  for (int d = 0; d < 3; ++d) {
    h->cshape[d] = d < nd ? desc->coef_shape[d] : 1;
    h->kernel[d] = d < nd ? desc->kernel[d] : 0;
    h->deriv[d] = d < nd ? desc->deriv[d] : 0;
    p.nk[d] = (int)h->cshape[d];
    p.eqs[d] = eqs[d];
    p.h[d] = d < nd ? desc->h[d] : 1.0;
    total *= h->cshape[d];
  }
  p.s1 = h->cshape[0];
  p.s2 = h->cshape[0] * h->cshape[1];
  double hmin = p.h[0];
  for (int d = 1; d < nd; ++d) hmin = std::min(hmin, p.h[d]);
  p.eps_line = hmin / 100.0;
  p.scale = 1.0;

#define CIB_CREATE_CUDA(expr)                                                            \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                         \
      cib200_itp_destroy(h);                                                             \
      return _e == cudaErrorMemoryAllocation ? CIB200_ENOMEM : CIB200_ECUDA;             \
    }                                                                                    \
  } while (0)

  const size_t es = esize(h->dtype);
  CIB_CREATE_CUDA(cudaMalloc(&h->d_coefs, es * (size_t)total));
  // cudaMemcpyDefault: desc->coefs may also be device memory (a grid built by cib200_build_coefs_device)
  CIB_CREATE_CUDA(cudaMemcpy(h->d_coefs, desc->coefs, es * (size_t)total, cudaMemcpyDefault));
  p.coefs = h->d_coefs;
  for (int d = 0; d < nd; ++d) {
    CIB_CREATE_CUDA(cudaMalloc(&h->d_knots[d], sizeof(double) * (size_t)h->cshape[d]));
    CIB_CREATE_CUDA(cudaMemcpy(h->d_knots[d], desc->knots[d], sizeof(double) * (size_t)h->cshape[d], cudaMemcpyHostToDevice));
    p.knots[d] = h->d_knots[d];
    p.knot0[d] = desc->knots[d][0];
    p.lo[d] = desc->knots[d][eqs[d] - 1];
    p.hi[d] = desc->knots[d][h->cshape[d] - 1 - (eqs[d] - 1)];
  }
  // distinct (kernel, derivative, rows) tables -> one packed device image.  rows = 101: the shipped tables; a :linear
  // axis may bring the generated table the reference builds for it (tables[d], n_pre[d] >= 10 000 rows, Julia memory
  // order n_pre x 2 eqs column-major = itp.kernel_pre[d]; constructor.jl:117-118), transposed here to row-major
  struct Slot { int kernel, d, rows; const double *src; };
  std::vector<Slot> slots;
  auto slot_of = [&](int kernel, int d, int rows, const double *src) {
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i].kernel == kernel && slots[i].d == d && slots[i].rows == rows) return (int)i;
    slots.push_back({kernel, d, rows, src});
    return (int)slots.size() - 1;
  };
  for (int d = 0; d < nd; ++d) {
    const bool gen = tables && n_pre && tables[d] && desc->subgrid[d] == CIB200_SG_LINEAR;
    if (gen && n_pre[d] < 2) {
      set_error("cib200_itp_create_tables: n_pre[%d] = %d", d, n_pre[d]);
      cib200_itp_destroy(h);
      return CIB200_EINVAL;
    }
    p.npre[d] = gen ? n_pre[d] : CIB_TAB_ROWS;
    p.tab_slot_k[d] = slot_of(h->kernel[d], h->deriv[d], p.npre[d], gen ? tables[d] : nullptr);
    // linear axes read K only: their "derivative" table aliases K (its terms carry zero factors)
    p.tab_slot_d[d] = desc->subgrid[d] == CIB200_SG_LINEAR ? p.tab_slot_k[d]
                                                           : slot_of(h->kernel[d], h->deriv[d] + 1, CIB_TAB_ROWS, nullptr);
  }
  for (int d = nd; d < 3; ++d) p.npre[d] = CIB_TAB_ROWS;
  p.prerange_pos = (nd == 2 && !desc->perdim && desc->subgrid[0] == CIB200_SG_LINEAR && desc->subgrid[1] == CIB200_SG_LINEAR) ? 1 : 0;
  for (int d = 0; d < 3; ++d) {
    p.subgrid[d] = d < nd ? desc->subgrid[d] : 0;
    p.tab_slot_e[d] = (d < nd && desc->subgrid[d] == CIB200_SG_QUINTIC) ? slot_of(h->kernel[d], h->deriv[d] + 2, CIB_TAB_ROWS, nullptr) : -1;
  }
  p.n_tab_slots = (int)slots.size();
  size_t tab_doubles = 0;
  for (auto &s : slots) tab_doubles += (size_t)s.rows * 2 * kernel_eqs(s.kernel);
  CIB_CREATE_CUDA(cudaMalloc(&h->d_tables, sizeof(double) * tab_doubles));
  size_t off = 0;
  for (size_t i = 0; i < slots.size(); ++i) {
    const int e = kernel_eqs(slots[i].kernel);
    const size_t nelem = (size_t)slots[i].rows * 2 * e;
    if (slots[i].src) {
      std::vector<double> rm(nelem);
      for (int r = 0; r < slots[i].rows; ++r)
        for (int c = 0; c < 2 * e; ++c) rm[(size_t)r * 2 * e + c] = slots[i].src[(size_t)c * slots[i].rows + r];
      CIB_CREATE_CUDA(cudaMemcpy(h->d_tables + off, rm.data(), sizeof(double) * nelem, cudaMemcpyHostToDevice));
    } else {
      CIB_CREATE_CUDA(cudaMemcpy(h->d_tables + off, host_table(slots[i].kernel, slots[i].d), sizeof(double) * nelem, cudaMemcpyHostToDevice));
    }
    p.tab_slot_eqs[i] = e;
    p.tab_slot_rows[i] = slots[i].rows;
    p.tab_slot_ptr[i] = h->d_tables + off;
    off += nelem;
  }
  for (int d = 0; d < nd; ++d) {
    p.tabk[d] = p.tab_slot_ptr[p.tab_slot_k[d]];
    p.tabd[d] = p.tab_slot_ptr[p.tab_slot_d[d]];
  }
  CIB_CREATE_CUDA(cudaMalloc(&h->d_status, sizeof(EvalStatus)));
  CIB_CREATE_CUDA(cudaMemset(h->d_status, 0xff, sizeof(EvalStatus)));
  CIB_CREATE_CUDA(cudaEventCreateWithFlags(&h->ev_busy, cudaEventDisableTiming));
#undef CIB_CREATE_CUDA
  *handle = h;
  return CIB200_OK;
}

int32_t cib200_itp_destroy(cib200_itp *h) {
  if (!h) return CIB200_OK;
  binned_destroy(h);
  if (h->d_coefs) cudaFree(h->d_coefs);
  for (int d = 0; d < 3; ++d)
    if (h->d_knots[d]) cudaFree(h->d_knots[d]);
  if (h->d_tables) cudaFree(h->d_tables);
  if (h->d_status) cudaFree(h->d_status);
  for (int i = 0; i < cib200_itp::NSTAGE; ++i) {
    if (h->d_stage[i]) cudaFree(h->d_stage[i]);
    if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
    if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
  }
  for (int i = 0; i < 3; ++i)
    if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
  if (h->d_bin_scratch) cudaFree(h->d_bin_scratch);
  for (int k = 0; k < 5; ++k)
    if (h->prof_ev[k]) cudaEventDestroy(h->prof_ev[k]);
  if (h->ev_busy) cudaEventDestroy(h->ev_busy);
  delete h->mu;
  delete h;
  return CIB200_OK;
}

int32_t cib200_itp_set_eval_mode(cib200_itp *h, int32_t mode) {
  if (!h || mode < CIB200_EVAL_AUTO || mode > CIB200_EVAL_BINNED) {
    set_error("cib200_itp_set_eval_mode: bad arguments");
    return CIB200_EINVAL;
  }
  if (mode == CIB200_EVAL_BINNED && !binned_supported(h)) {
    set_error("cib200_itp_set_eval_mode: BINNED needs a 3-D f64 interpolant with one kernel on all axes");
    return CIB200_EUNSUPPORTED;
  }
  h->eval_mode = mode;
  return CIB200_OK;
}

int32_t cib200_itp_set_profiling(cib200_itp *h, int32_t enable) {
  if (!h) {
    set_error("cib200_itp_set_profiling: null handle");
    return CIB200_EINVAL;
  }
  h->profiling = enable ? 1 : 0;
  h->prof_valid = 0;
  return CIB200_OK;
}

int32_t cib200_itp_last_pass_ms(cib200_itp *h, float *ms) {
  if (!h || !ms || !h->prof_valid) {
    set_error("cib200_itp_last_pass_ms: no profiled evaluation (enable cib200_itp_set_profiling first)");
    return CIB200_EINVAL;
  }
  CIB_CUDA(cudaEventSynchronize(h->prof_ev[4]));
  for (int i = 0; i < 4; ++i) CIB_CUDA(cudaEventElapsedTime(&ms[i], h->prof_ev[i], h->prof_ev[i + 1]));
  return CIB200_OK;
}

int32_t cib200_itp_last_strategy(cib200_itp *h) { return h ? h->last_strategy : 0; }

int64_t cib200_itp_scratch_bytes(cib200_itp *h, int64_t nq) {
  if (!h) return 0;
  return h->eval_mode == CIB200_EVAL_BINNED ? binned_scratch_bytes(h, nq) : 0;
}

static int32_t eval_common_locked(cib200_itp *h, const Coords &c_in, long long nq, void *out, int32_t memkind,
                                  void *stream, int64_t *first_bad, const char *who);

// Calls on ONE handle share its scratch (binning records, staging ring, status words): they are serialised here - the
// host side by a mutex, the device side by an event recorded after a call's last kernel and awaited by the next call's
// stream - so that `Threads.@threads` over chunks of a query vector (the reference's threading pattern) is safe.
static int32_t eval_common(cib200_itp *h, const Coords &c_in, long long nq, void *out, int32_t memkind,
                           void *stream, int64_t *first_bad, const char *who) {
  std::lock_guard<std::mutex> guard(*h->mu);
  int dev = -1;
  CIB_CUDA(cudaGetDevice(&dev));
  if (dev != h->device) CIB_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (h->busy_valid) {
    // an earlier DEVICE call may still be running on the caller's stream: it shares the status words and the binning
    // scratch with this call, whatever stream this one uses (a HOST call runs on the handle's own three streams)
    if (memkind == CIB200_DEVICE) {
      CIB_CUDA(cudaStreamWaitEvent(st, h->ev_busy, 0));
    } else if (h->streams[0]) {
      for (int i = 0; i < 3; ++i) CIB_CUDA(cudaStreamWaitEvent(h->streams[i], h->ev_busy, 0));
    } else {
      CIB_CUDA(cudaEventSynchronize(h->ev_busy));  // the streams do not exist yet (first HOST call)
    }
  }
  const int32_t rc = eval_common_locked(h, c_in, nq, out, memkind, stream, first_bad, who);
  if (memkind == CIB200_DEVICE && nq > 0 && cudaEventRecord(h->ev_busy, st) == cudaSuccess) h->busy_valid = 1;
  if (memkind == CIB200_HOST) {
    // a HOST call returns only after its streams drained (also on the error paths: no copy into or out of the caller's
    // buffers may still be in flight when the call returns)
    if (rc != CIB200_OK && h->streams[0])
      for (int i = 0; i < 3; ++i) cudaStreamSynchronize(h->streams[i]);
    h->busy_valid = 0;
  }
  if (dev != h->device) cudaSetDevice(dev);
  return rc;
}

static int32_t eval_common_locked(cib200_itp *h, const Coords &c_in, long long nq, void *out, int32_t memkind,
                                  void *stream, int64_t *first_bad, const char *who) {
  cudaStream_t st = (cudaStream_t)stream;
  const int nd = h->p.ndim;
  const size_t es = esize(h->dtype);
  if (first_bad) *first_bad = -1;
  if (nq == 0) return CIB200_OK;
  int rc;
  if (memkind == CIB200_DEVICE) {
    if ((rc = reset_status(h, st))) return rc;
    if ((rc = eval_dispatch(h, c_in, out, nq, st))) return rc;
    if (first_bad || h->p.extrap == CIB200_THROW) {
      int code;
      if ((rc = fetch_status(h, st, first_bad, &code))) return rc;
      if (code == CIB200_EOOB) set_error("%s: query %lld is outside the interpolation domain (extrap=Throw)", who, first_bad ? (long long)*first_bad : -1LL);
      if (code == CIB200_ENAN) set_error("%s: query %lld has a NaN/Inf coordinate (InexactError in the reference)", who, first_bad ? (long long)*first_bad : -1LL);
      return code;
    }
    return CIB200_OK;
  }
  if (memkind != CIB200_HOST) {
    set_error("%s: unknown memkind %d", who, memkind);
    return CIB200_EINVAL;
  }
  // host buffers: chunked pipeline H2D (stream 0) -> kernels (stream 1) -> D2H (stream 2) over a ring of device buffers
  cudaStream_t s_in, s_comp, s_out;
  if (c_in.grid) {
    // coordinate vectors are tiny: upload them whole, keep the result pipeline chunk-free
    size_t xbytes = 0;
    for (int d = 0; d < nd; ++d) xbytes += es * (size_t)c_in.n[d];
    if ((rc = ensure_stage(h, std::max(xbytes, es * (size_t)nq)))) return rc;
    s_comp = h->streams[1];
    Coords c = c_in;
    char *dst = (char *)h->d_stage[0];
    for (int d = 0; d < nd; ++d) {
      CIB_CUDA(cudaMemcpyAsync(dst, c_in.x[d], es * (size_t)c_in.n[d], cudaMemcpyHostToDevice, s_comp));
      c.x[d] = dst;
      dst += es * (size_t)c_in.n[d];
    }
    if ((rc = reset_status(h, s_comp))) return rc;
    if ((rc = eval_dispatch(h, c, h->d_stage[1], nq, s_comp))) return rc;
    CIB_CUDA(cudaMemcpyAsync(out, h->d_stage[1], es * (size_t)nq, cudaMemcpyDeviceToHost, s_comp));
    int code;
    if ((rc = fetch_status(h, s_comp, first_bad, &code))) return rc;
    if (code) set_error("%s: query %lld %s", who, first_bad ? (long long)*first_bad : -1LL, code == CIB200_EOOB ? "is outside the interpolation domain (extrap=Throw)" : "has a NaN/Inf coordinate");
    return code;
  }
  // uniform chunks: 2^26 queries for large batches (dense enough for the cell-sorted strategy: 0.5 query per cell on a
  // 512^3 grid; the pipeline is then H2D-bound, 24 B/query over PCIe), smaller for small batches so that at least four
  // chunks overlap.  Measured on B200 (512^3 b5, pinned buffers): 10^9 queries 1.99e9 evals/s with 2^26 (1.66e9 with
  // 2^24); geometric ramp-up / ramp-down schedules were tried and lost to uniform chunks (small chunks bin poorly).
  // Pageable caller buffers (what a caller that never pins hands over): cudaMemcpyAsync from pageable memory is staged
  // by the runtime through one small internal buffer, synchronously (measured: 0.30e9 evals/s against 1.9e9 with pinned
  // buffers).  Instead the coordinates are packed into the handle's own pinned ring by a few host threads while the
  // device works on the previous chunk, and results are unpacked from it the same way.
  bool stage_in = false;
  for (int d = 0; d < nd; ++d) stage_in = stage_in || is_pageable(c_in.x[d]);
  const bool stage_out = is_pageable(out);
  const bool staged = stage_in || stage_out;
  int chunk_log2 = nq >= (1LL << 28) ? 26 : (nq >= (1LL << 27) ? 25 : 24);
  // pageable buffers go through the handle's pinned ring: 2^24-query chunks (1.5 GB of pinned memory), 2^26 for batches of
  // 2^27 queries and more (6.4 GB: the chunks are then dense enough for the cell-sorted kernels - measured on 512^3 b5,
  // 5 10^8 queries: 1.10e9 -> 1.42e9 evals/s, tools/ab_pageable_chunk.py); falls back to 2^24 when that much cannot be pinned
  if (staged) chunk_log2 = nq >= (1LL << 27) ? 26 : 24;
  if (const char *e = tune_env("CIB200_HOST_CHUNK_LOG2")) chunk_log2 = std::max(16, std::min(30, atoi(e)));
  std::vector<long long> sizes;
  long long chunk = 0;
  for (;;) {
    sizes.clear();
    for (long long rem = nq; rem > 0;) {
      const long long m = std::min(rem, 1LL << chunk_log2);
      sizes.push_back(m);
      rem -= m;
    }
    chunk = 0;
    for (long long m : sizes) chunk = std::max(chunk, m);
    // per buffer: nd coordinate chunks + 1 result chunk
    if ((rc = ensure_stage(h, es * (size_t)chunk * (size_t)(nd + 1)))) return rc;
    if (!staged) break;
    rc = ensure_host_stage(h, es * (size_t)chunk * (size_t)(nd + 1));
    if (rc == CIB200_OK) break;
    if (chunk_log2 <= 24) return rc;
    cudaGetLastError();
    chunk_log2 = 24;
  }
  s_in = h->streams[0];
  s_comp = h->streams[1];
  s_out = h->streams[2];
  if ((rc = reset_status(h, s_comp))) return rc;
  constexpr int NS = cib200_itp::NSTAGE;
  long long q0 = 0;
  long long slot_q0[NS] = {0}, slot_m[NS] = {0};
  // results of the chunk that last used slot b: pinned ring -> caller's array, on a helper thread so that it overlaps
  // the packing of the next chunk; joined before the slot's result area is handed to the next D2H
  std::thread unpacker;
  auto join_unpacker = [&]() {
    if (unpacker.joinable()) unpacker.join();
  };
  auto unpack = [&](int b) -> int {
    CIB_CUDA(cudaEventSynchronize(h->ev_out[b]));
    join_unpacker();
    char *dst = (char *)out + es * (size_t)slot_q0[b];
    const char *src = (const char *)h->h_stage[b] + es * (size_t)chunk * nd;
    const size_t bytes = es * (size_t)slot_m[b];
    unpacker = std::thread([=]() { par_memcpy(dst, src, bytes); });
    return CIB200_OK;
  };
  struct JoinGuard {
    std::thread &t;
    ~JoinGuard() {
      if (t.joinable()) t.join();
    }
  } join_guard{unpacker};
  for (size_t ichunk = 0; ichunk < sizes.size(); ++ichunk) {
    const int b = (int)(ichunk % NS);
    const long long m = sizes[ichunk];
    char *base = (char *)h->d_stage[b];
    Coords c = c_in;
    if (ichunk >= (size_t)NS) {
      CIB_CUDA(cudaStreamWaitEvent(s_in, h->ev_out[b], 0));  // device buffer b drained by its last D2H
      if (stage_out) {
        if ((rc = unpack(b))) return rc;
      } else if (stage_in) {
        CIB_CUDA(cudaEventSynchronize(h->ev_in[b]));  // the pinned slot's last H2D has left it
      }
    }
    for (int d = 0; d < nd; ++d) {
      char *dst = base + es * (size_t)chunk * d;
      const char *src = (const char *)c_in.x[d] + es * (size_t)q0;
      if (stage_in) {
        char *hp = (char *)h->h_stage[b] + es * (size_t)chunk * d;
        par_memcpy(hp, src, es * (size_t)m);
        src = hp;
      }
      CIB_CUDA(cudaMemcpyAsync(dst, src, es * (size_t)m, cudaMemcpyHostToDevice, s_in));
      c.x[d] = dst;
    }
    CIB_CUDA(cudaEventRecord(h->ev_in[b], s_in));
    void *dout = base + es * (size_t)chunk * nd;
    c.qoff = q0;  // failing queries are reported in absolute numbering
    CIB_CUDA(cudaStreamWaitEvent(s_comp, h->ev_in[b], 0));
    if ((rc = eval_dispatch(h, c, dout, m, s_comp))) return rc;
    CIB_CUDA(cudaEventRecord(h->ev_comp[b], s_comp));
    CIB_CUDA(cudaStreamWaitEvent(s_out, h->ev_comp[b], 0));
    join_unpacker();  // the slot's previous results have left the pinned ring
    void *hdst = stage_out ? (void *)((char *)h->h_stage[b] + es * (size_t)chunk * nd) : (void *)((char *)out + es * (size_t)q0);
    CIB_CUDA(cudaMemcpyAsync(hdst, dout, es * (size_t)m, cudaMemcpyDeviceToHost, s_out));
    CIB_CUDA(cudaEventRecord(h->ev_out[b], s_out));
    slot_q0[b] = q0;
    slot_m[b] = m;
    q0 += m;
  }
  if (stage_out) {  // the last chunks, oldest first
    const size_t nch = sizes.size();
    for (size_t ichunk = nch > (size_t)NS ? nch - NS : 0; ichunk < nch; ++ichunk)
      if ((rc = unpack((int)(ichunk % NS)))) return rc;
    join_unpacker();
  }
  CIB_CUDA(cudaStreamSynchronize(s_out));
  int code;
  if ((rc = fetch_status(h, s_comp, first_bad, &code))) return rc;
  if (code)
    set_error("%s: query %lld %s", who, first_bad ? (long long)*first_bad : -1LL,
              code == CIB200_EOOB ? "is outside the interpolation domain (extrap=Throw)" : "has a NaN/Inf coordinate");
  return code;
}

int32_t cib200_itp_get_coefs(cib200_itp *h, void *coefs_out) {
  if (!h || !coefs_out) {
    set_error("cib200_itp_get_coefs: bad arguments");
    return CIB200_EINVAL;
  }
  size_t total = 1;
  for (int d = 0; d < h->p.ndim; ++d) total *= (size_t)h->cshape[d];
  CIB_CUDA(cudaMemcpy(coefs_out, h->d_coefs, esize(h->dtype) * total, cudaMemcpyDeviceToHost));
  return CIB200_OK;
}

int32_t cib200_itp_eval(cib200_itp *h, int64_t nq, const void *const *x, void *out, int32_t memkind,
                        void *stream, int64_t *first_bad) {
  if (!h || nq < 0 || (nq > 0 && (!x || !out))) {
    set_error("cib200_itp_eval: bad arguments");
    return CIB200_EINVAL;
  }
  Coords c;
  memset(&c, 0, sizeof c);
  for (int d = 0; d < h->p.ndim; ++d) {
    if (nq > 0 && !x[d]) {
      set_error("cib200_itp_eval: x[%d] is null", d);
      return CIB200_EINVAL;
    }
    c.x[d] = nq > 0 ? x[d] : nullptr;
  }
  c.grid = 0;
  return eval_common(h, c, nq, out, memkind, stream, first_bad, "cib200_itp_eval");
}

// Single-process multi-GPU evaluation (replicated grid, split queries, no collective): one handle per
// device, one host thread per handle, each slice through its handle's own H2D / kernel / D2H pipeline.
int32_t cib200_itp_eval_multi(cib200_itp *const *handles, int32_t n_handles, int64_t nq, const void *const *x, void *out,
                              int64_t *first_bad) {
  if (!handles || n_handles < 1 || nq < 0 || (nq > 0 && (!x || !out))) {
    set_error("cib200_itp_eval_multi: bad arguments");
    return CIB200_EINVAL;
  }
  for (int i = 0; i < n_handles; ++i)
    if (!handles[i] || handles[i]->p.ndim != handles[0]->p.ndim || handles[i]->dtype != handles[0]->dtype) {
      set_error("cib200_itp_eval_multi: handle %d is null or describes a different interpolant", i);
      return CIB200_EINVAL;
    }
  if (first_bad) *first_bad = -1;
  if (n_handles == 1 || nq < 2 * (int64_t)n_handles)
    return cib200_itp_eval(handles[0], nq, x, out, CIB200_HOST, nullptr, first_bad);
  const int nd = handles[0]->p.ndim;
  const size_t es = esize(handles[0]->dtype);
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  std::vector<int32_t> rcs((size_t)n_handles, CIB200_OK);
  std::vector<int64_t> bads((size_t)n_handles, -1);
  std::vector<std::string> msgs((size_t)n_handles);
  std::vector<std::thread> pool;
  const int64_t per = (nq + n_handles - 1) / n_handles;
  for (int i = 0; i < n_handles; ++i) {
    pool.emplace_back([&, i]() {
      const int64_t q0 = std::min<int64_t>(nq, per * i), q1 = std::min<int64_t>(nq, q0 + per);
      if (q1 <= q0) return;
      const void *xs[3] = {nullptr, nullptr, nullptr};
      for (int d = 0; d < nd; ++d) xs[d] = (const char *)x[d] + es * (size_t)q0;
      int64_t bad = -1;
      rcs[(size_t)i] = cib200_itp_eval(handles[i], q1 - q0, xs, (char *)out + es * (size_t)q0, CIB200_HOST, nullptr, &bad);
      if (bad >= 0) bads[(size_t)i] = q0 + bad;
      if (rcs[(size_t)i]) msgs[(size_t)i] = cib200_last_error();
    });
  }
  for (auto &t : pool) t.join();
  cudaSetDevice(prev_dev);
  // the lowest offending query decides, as in a sequential pass over the batch
  int32_t rc = CIB200_OK;
  int64_t bad = -1;
  for (int i = 0; i < n_handles; ++i) {
    if (rcs[(size_t)i] == CIB200_OK) continue;
    if (rc == CIB200_OK || (bads[(size_t)i] >= 0 && (bad < 0 || bads[(size_t)i] < bad))) {
      rc = rcs[(size_t)i];
      bad = bads[(size_t)i];
      // slice-relative index in the message -> absolute
      set_error("cib200_itp_eval_multi (device slice %d): %s [absolute query index %lld]", i, msgs[(size_t)i].c_str(), (long long)bad);
    }
  }
  if (first_bad) *first_bad = bad;
  return rc;
}

int32_t cib200_itp_eval_grid(cib200_itp *h, const int64_t *n, const void *const *x, void *out, int32_t memkind,
                             void *stream, int64_t *first_bad) {
  if (!h || !n || !x || !out) {
    set_error("cib200_itp_eval_grid: bad arguments");
    return CIB200_EINVAL;
  }
  Coords c;
  memset(&c, 0, sizeof c);
  long long nq = 1;
  for (int d = 0; d < 3; ++d) c.n[d] = 1;
  for (int d = 0; d < h->p.ndim; ++d) {
    if (n[d] < 0 || (n[d] > 0 && !x[d])) {
      set_error("cib200_itp_eval_grid: bad axis %d", d);
      return CIB200_EINVAL;
    }
    c.x[d] = x[d];
    c.n[d] = n[d];
    nq *= n[d];
  }
  c.grid = 1;
  return eval_common(h, c, nq, out, memkind, stream, first_bad, "cib200_itp_eval_grid");
}

// ---- sweeps -----------------------------------------------------------------------------------
static int32_t sweep_io(int ndim, int dtype, long long tot_in, long long tot_out, const void *values, void *out,
                        int memkind, cudaStream_t st, void **d_in, void **d_out, bool *owned) {
  const size_t es = esize(dtype);
  *owned = false;
  if (memkind == CIB200_DEVICE) {
    *d_in = const_cast<void *>(values);
    *d_out = out;
    return CIB200_OK;
  }
  if (memkind != CIB200_HOST) {
    set_error("sweep: unknown memkind %d", memkind);
    return CIB200_EINVAL;
  }
  CIB_CUDA(cudaMalloc(d_in, es * (size_t)tot_in));
  cudaError_t e = cudaMalloc(d_out, es * (size_t)tot_out);
  if (e != cudaSuccess) {
    cudaFree(*d_in);
    set_error("cudaMalloc failed: %s", cudaGetErrorString(e));
    return CIB200_ENOMEM;
  }
  *owned = true;
  CIB_CUDA(cudaMemcpyAsync(*d_in, values, es * (size_t)tot_in, cudaMemcpyHostToDevice, st));
  (void)ndim;
  return CIB200_OK;
}

int32_t cib200_resample(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out, const void *values,
                        void *out, const double *const *knots_in, const double *const *knots_out,
                        const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                        const int32_t *bc_right, int32_t memkind, void *stream) {
  return cib200_resample_sg(ndim, dtype, n_in, n_out, values, out, knots_in, knots_out, kernel, deriv, bc_left, bc_right,
                            nullptr, memkind, stream);
}

int32_t cib200_resample_sg(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out, const void *values,
                           void *out, const double *const *knots_in, const double *const *knots_out,
                           const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                           const int32_t *bc_right, const int32_t *subgrid, int32_t memkind, void *stream) {
  return cib200_resample_tables(ndim, dtype, n_in, n_out, values, out, knots_in, knots_out, kernel, deriv, bc_left,
                                bc_right, subgrid, nullptr, nullptr, memkind, stream);
}

int32_t cib200_resample_tables(int32_t ndim, int32_t dtype, const int64_t *n_in, const int64_t *n_out, const void *values,
                               void *out, const double *const *knots_in, const double *const *knots_out,
                               const int32_t *kernel, const int32_t *deriv, const int32_t *bc_left,
                               const int32_t *bc_right, const int32_t *subgrid, const double *const *tables,
                               const int32_t *n_pre, int32_t memkind, void *stream) {
  if (ndim < 1 || ndim > CIB200_MAX_SWEEP_DIM || !n_in || !n_out || !values || !out || !knots_in || !knots_out ||
      !kernel || !deriv || !bc_left || !bc_right) {
    set_error("cib200_resample: bad arguments");
    return CIB200_EINVAL;
  }
  if (dtype != CIB200_F64 && dtype != CIB200_F32) {
    set_error("cib200_resample: unknown dtype %d", dtype);
    return CIB200_EINVAL;
  }
  long long nin[CIB200_MAX_SWEEP_DIM], nout[CIB200_MAX_SWEEP_DIM], tin = 1, tout = 1;
  int k[CIB200_MAX_SWEEP_DIM], dv[CIB200_MAX_SWEEP_DIM], bl[CIB200_MAX_SWEEP_DIM], br[CIB200_MAX_SWEEP_DIM];
  int sgv[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) {
    sgv[d] = subgrid ? subgrid[d] : CIB200_SG_CUBIC;
    if (sgv[d] < CIB200_SG_CUBIC || sgv[d] > CIB200_SG_LINEAR) {
      set_error("cib200_resample: unknown subgrid %d on axis %d", sgv[d], d);
      return CIB200_EINVAL;
    }
    if (kernel_eqs(kernel[d]) < 0) {
      set_error("cib200_resample: unknown kernel id %d on axis %d", kernel[d], d);
      return CIB200_EUNSUPPORTED;
    }
    if (deriv[d] < 0) {
      set_error("convolution_resample does not support antiderivatives (derivative=-1)");
      return CIB200_EINVAL;
    }
    {
      const int need = deriv[d] + (sgv[d] == CIB200_SG_LINEAR ? 0 : (sgv[d] == CIB200_SG_QUINTIC ? 2 : 1));
      if (need > kernel_max_deriv_table(kernel[d])) {
        set_error("cib200_resample: derivative %d on axis %d needs K^(%d), which is not a shipped table", deriv[d], d, need);
        return CIB200_EUNSUPPORTED;
      }
    }
    if (n_in[d] < 3 || n_out[d] < 1 || !knots_in[d] || !knots_out[d] || bc_left[d] < 0 || bc_left[d] > 3 ||
        bc_right[d] < 0 || bc_right[d] > 3) {
      set_error("cib200_resample: bad axis %d", d);
      return CIB200_EINVAL;
    }
    nin[d] = n_in[d];
    nout[d] = n_out[d];
    tin *= nin[d];
    tout *= nout[d];
    k[d] = kernel[d];
    dv[d] = deriv[d];
    bl[d] = bc_left[d];
    br[d] = bc_right[d];
  }
  cudaStream_t st = (cudaStream_t)stream;
  void *d_in, *d_out;
  bool owned;
  int rc = sweep_io(ndim, dtype, tin, tout, values, out, memkind, st, &d_in, &d_out, &owned);
  if (rc) return rc;
  int rows[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) rows[d] = (tables && n_pre && tables[d]) ? n_pre[d] : 0;
  rc = sweep_resample(ndim, dtype, nin, nout, d_in, d_out, knots_in, knots_out, k, dv, bl, br, sgv, tables, rows, st);
  if (owned) {
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(out, d_out, esize(dtype) * (size_t)tout, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        set_error("cib200_resample: D2H failed: %s", cudaGetErrorString(e));
        rc = CIB200_ECUDA;
      }
    }
    cudaFree(d_in);
    cudaFree(d_out);
  }
  return rc;
}

int32_t cib200_smooth(int32_t ndim, int32_t dtype, const int64_t *n, const void *values, void *out, const double *h,
                      double B, int32_t memkind, void *stream) {
  if (ndim < 1 || ndim > CIB200_MAX_SWEEP_DIM || !n || !values || !out || !h || !(B > 0)) {
    set_error("cib200_smooth: bad arguments");
    return CIB200_EINVAL;
  }
  if (dtype != CIB200_F64 && dtype != CIB200_F32) {
    set_error("cib200_smooth: unknown dtype %d", dtype);
    return CIB200_EINVAL;
  }
  const int E = cib200_smooth_halfwidth(B);
  long long nn[CIB200_MAX_SWEEP_DIM], tot = 1;
  for (int d = 0; d < ndim; ++d) {
    if (E >= n[d]) {
      set_error("Gaussian stencil half-width (%d points) exceeds grid size (n=%lld) in dimension %d. Use a larger B (currently B=%g) or more grid points (need n > %d).", E, (long long)n[d], d + 1, B, E);
      return CIB200_EINVAL;
    }
    nn[d] = n[d];
    tot *= n[d];
  }
  cudaStream_t st = (cudaStream_t)stream;
  void *d_in, *d_out;
  bool owned;
  int rc = sweep_io(ndim, dtype, tot, tot, values, out, memkind, st, &d_in, &d_out, &owned);
  if (rc) return rc;
  rc = sweep_smooth(ndim, dtype, nn, d_in, d_out, h, B, st);
  if (owned) {
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(out, d_out, esize(dtype) * (size_t)tot, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        set_error("cib200_smooth: D2H failed: %s", cudaGetErrorString(e));
        rc = CIB200_ECUDA;
      }
    }
    cudaFree(d_in);
    cudaFree(d_out);
  }
  return rc;
}

int32_t cib200_device_count(void) {
  int n = 0;
  return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
int32_t cib200_set_device(int32_t device) {
  CIB_CUDA(cudaSetDevice(device));
  return CIB200_OK;
}

int32_t cib200_host_register(void *ptr, size_t bytes) {
  CIB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));  // pinned for every device of the process
  return CIB200_OK;
}
int32_t cib200_host_unregister(void *ptr) {
  CIB_CUDA(cudaHostUnregister(ptr));
  return CIB200_OK;
}

}  // extern "C"
