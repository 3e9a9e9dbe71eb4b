// cib200_sweep.cu - K2 / K3: separable per-axis sweeps behind convolution_resample and convolution_smooth.
//
// Replaces the per-line Julia loops of
//   convolution_resample_internal   convolution/convolution_resample.jl:240-386 (hot loop :297-381)
//   convolution_smooth              convolution/convolution_gaussian_smoothing.jl:45-126 (:70-124)
// including the per-line 1-D ghost fill they call
//   apply_boundary_conditions_for_dim!  convolution_coefs/create_convolutional_coefs.jl:200-370
//   fill_ghost_points_polynomial! :408-441, fill_ghost_points_gaussian! :443-494.
//
// One pass = one axis.  The array is viewed as [inner][n][outer] (inner contiguous).  A CTA owns a
// tile of TI lines: for axis >= 1 these are TI consecutive `inner` positions (global loads/stores
// are coalesced along inner); for axis 0 they are TI consecutive lines, transposed through shared
// memory on the way in so that loads AND stores stay coalesced along the line.  Per tile:
//   1. stage the lines in shared memory as buf[pad + j][ti] (double, whatever T is),
//   2. one thread per line fills the ghosts: sequential mean (same summation order as a scalar
//      loop), :detect thresholds, Lagrange / least-squares extrapolation - plain mul/add via
//      __d*_rn so the branch decisions and ghost values do not depend on fma contraction,
//   3. all threads produce the outputs of the tile: resample uses per-output-knot blended weights
//      W[l][k] + first-tap index (computed once per pass by resample_weights_kernel: they depend on
//      (axis, k) only, the reference recomputes them for every line); smooth uses the (2E+1)-tap
//      normalised Gaussian held in shared memory.
#include <math.h>

#include <vector>

#include "cib200_device.cuh"

namespace cib {

struct GhostMats {
  double poly[9][8];  // rows 0..eqs-2 used; columns 0..m-1
  double lin[9][2];
  double quad[9][3];
  int poly_nodes;
};

struct SweepParams {
  long long inner, n_in, n_out, outer;
  int pad;        // ghost slots per side in the line buffer: eqs-1 (resample) or E (smooth)
  int eqs;        // resample: kernel eqs ; smooth: E
  int taps;       // 2 eqs (resample) / 2E+1 (smooth)
  int ti;         // lines per tile
  int pitch;      // buffer pitch (doubles)
  int bcl, bcr;
  double h;
  double scale;   // (-1/h)^derivative (resample) ; 1 (smooth)
  const double *W;     // resample: [taps][n_out] blended weights ; smooth: kw[taps]
  const int *base;     // resample: first tap index into the padded line (0-based) per output
};

// ---- per-output-knot weights: convolution_resample.jl:320-361 minus the data ------------------
__global__ void resample_weights_kernel(const double *__restrict__ knots_out, long long n_out, double x_expanded_1,
                                        double h, int eqs, long long n_coef, const double *__restrict__ tk,
                                        const double *__restrict__ td, double *__restrict__ W, int *__restrict__ base) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  AxisPos p;
  locate_computed(knots_out[k], x_expanded_1, h, eqs, n_coef, p);
  base[k] = p.base;
  Hermite hb = hermite_basis(p.t);
  const int w = 2 * eqs;
  for (int l = 0; l < w; ++l) W[(long long)l * n_out + k] = blended_weight(hb, tk, td, p.row, l, w);
}

// ---- ghost fill of one line held in shared memory ----------------------------------------------
// line(j) = buf[(pad + j) * pitch + ti], j = 0..n-1
__device__ __forceinline__ double line_mean(const double *buf, int pad, int pitch, int ti, long long n) {
  double sum = 0.0;
  for (long long j = 0; j < n; ++j) sum = __dadd_rn(sum, buf[(pad + j) * pitch + ti]);
  return __ddiv_rn(sum, (double)n);
}

__device__ void ghost_poly_side(double *buf, int pad, int pitch, int ti, long long n, double mean, double h, int eqs,
                                int bc, bool right, const GhostMats &gm) {
  const int ns = gm.poly_nodes;
  double y[8];  // centred boundary values, y[0] = the boundary point itself, then inward
  const int navail = n < 8 ? (int)n : 8;
  for (int k = 0; k < navail; ++k) {
    long long j = right ? (n - 1 - k) : k;
    y[k] = __dsub_rn(buf[(pad + j) * pitch + ti], mean);
  }
  bool use_poly = false;
  if (n >= ns && (bc == CIB200_BC_DETECT || bc == CIB200_BC_POLY)) {
    if (bc == CIB200_BC_POLY) {
      use_poly = true;
    } else {
      // natural-order window p[0..ns-1]: left p[i] = y[i]; right p[i] = y[ns-1-i]
      double d1[8], d2[8];
      double mn = INFINITY, mx = -INFINITY;
      for (int d = 0; d < ns - 1; ++d) {
        double a = right ? y[ns - 2 - d] : y[d + 1];
        double b = right ? y[ns - 1 - d] : y[d];
        d1[d] = __dsub_rn(a, b);
        mn = fmin(mn, d1[d]);
        mx = fmax(mx, d1[d]);
      }
      bool periodic = __ddiv_rn(__dmul_rn(mn, mx), __dmul_rn(h, h)) < -1.0 / 10;
      mn = INFINITY, mx = -INFINITY;
      for (int d = 0; d < ns - 2; ++d) {
        d2[d] = __dsub_rn(d1[d + 1], d1[d]);
        mn = fmin(mn, d2[d]);
        mx = fmax(mx, d2[d]);
      }
      bool high_curv = __ddiv_rn(fabs(__dsub_rn(mx, mn)), h) > 1.0 / 2;
      bool high_d3 = false;
      if (ns > 4) {
        mn = INFINITY, mx = -INFINITY;
        for (int d = 0; d < ns - 3; ++d) {
          double v = __dsub_rn(d2[d + 1], d2[d]);
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        high_d3 = __ddiv_rn(fabs(__dsub_rn(mx, mn)), h) > 1.0 / 2;
      }
      use_poly = !periodic && !high_curv && !high_d3;
    }
  }
  const int m = use_poly ? ns : (bc == CIB200_BC_QUADRATIC ? 3 : 2);
  for (int j = 1; j <= eqs - 1; ++j) {
    double acc = 0.0;
    for (int k = 0; k < m; ++k) {
      double mv = use_poly ? gm.poly[j - 1][k] : (m == 3 ? gm.quad[j - 1][k] : gm.lin[j - 1][k]);
      acc = __dadd_rn(acc, __dmul_rn(mv, y[k]));
    }
    double gv = __dadd_rn(mean, acc);
    long long slot = right ? (pad + n - 1 + j) : (pad - j);
    buf[slot * pitch + ti] = gv;
  }
}

// fill_ghost_points_gaussian! with convolution_smooth's buffer geometry: padding E, boundary index from
// get_boundary_indices(size = n + 2E, eqs = E) -> ghosts land one slot further out, slot next to the
// interior stays 0.  Values are the centred least-squares line (mean NOT added back in
// exact arithmetic: y_mean + (a + b(1-k) - y_mean)).
__device__ void ghost_gauss_side(double *buf, int pad, int pitch, int ti, long long n, double mean, int E, bool right) {
  long long n_fit = E / 3 > 6 ? E / 3 : 6;
  if (n_fit > n) n_fit = n;
  double sx = 0, sy = 0, sxx = 0, sxy = 0;
  for (long long k = 1; k <= n_fit; ++k) {
    double xk = (double)k;
    long long j = right ? (n - k) : (k - 1);
    double yk = __dsub_rn(buf[(pad + j) * pitch + ti], mean);
    sx = __dadd_rn(sx, xk);
    sy = __dadd_rn(sy, yk);
    sxx = __dadd_rn(sxx, __dmul_rn(xk, xk));
    sxy = __dadd_rn(sxy, __dmul_rn(xk, yk));
  }
  double nf = (double)n_fit;
  double denom = __dsub_rn(__dmul_rn(nf, sxx), __dmul_rn(sx, sx));
  double b = __ddiv_rn(__dsub_rn(__dmul_rn(nf, sxy), __dmul_rn(sx, sy)), denom);
  double a = __ddiv_rn(__dsub_rn(sy, __dmul_rn(b, sx)), nf);
  for (int k = 1; k <= E - 1; ++k) {
    double y_temp = __dsub_rn(__dadd_rn(a, __dmul_rn(b, (double)(1 - k))), mean);
    double gv = __dadd_rn(mean, y_temp);
    long long slot = right ? (pad + n + k) : (pad - 1 - k);
    buf[slot * pitch + ti] = gv;
  }
}

static constexpr int SMOOTH_R = 8;  // consecutive outputs per thread in the Gaussian FIR (sliding register window)

// TAPS = 2 eqs (compile time) for resample, 0 for smooth (run-time 2E+1 taps)
// CPITCH = compile-time buffer pitch (shared-memory loads get immediate offsets), 0 = run-time pitch
template <typename T, bool SMOOTH, bool AXIS0, int TAPS, int CPITCH>
__global__ void __launch_bounds__(256) sweep_kernel(SweepParams sp, GhostMats gm, const T *__restrict__ in,
                                                    T *__restrict__ out) {
  extern __shared__ double sm[];
  const int pad = sp.pad, pitch = CPITCH > 0 ? CPITCH : sp.pitch, TI = sp.ti;
  const long long n = sp.n_in, nout = sp.n_out;
  const long long nrow = n + 2 * pad;
  double *buf = sm;
  double *kw = sm + (nrow + (SMOOTH ? 2 * SMOOTH_R : 0)) * pitch;  // smooth only
  const int tid = threadIdx.x, nth = blockDim.x;

  // tile coordinates
  long long tiles_inner, tile_i, o;
  long long line0;  // AXIS0: first line of the tile
  int valid;        // lines of this tile that exist
  if (AXIS0) {
    line0 = (long long)blockIdx.x * TI;
    long long rem = sp.outer - line0;
    valid = rem < TI ? (int)rem : TI;
    tiles_inner = 1; tile_i = 0; o = 0;
  } else {
    tiles_inner = (sp.inner + TI - 1) / TI;
    tile_i = blockIdx.x % tiles_inner;
    o = blockIdx.x / tiles_inner;
    long long rem = sp.inner - tile_i * TI;
    valid = rem < TI ? (int)rem : TI;
    line0 = 0;
  }
  (void)tiles_inner;

  // (all tile loops are warp-per-row / lane-per-column: no integer division in the hot path)
  const int warp_ = tid >> 5, lane_ = tid & 31, nwarp_ = nth >> 5;
  const int n_i = (int)n;
  // 0. clear ghost slots (the reference zero-fills the line buffer: resample :309, smooth :96)
  for (int r = warp_; r < 2 * pad; r += nwarp_) {
    const int slot = r < pad ? r : (n_i + r);
    for (int ti = lane_; ti < TI; ti += 32) buf[slot * pitch + ti] = 0.0;
  }
  if (SMOOTH)
    for (int e = tid; e < sp.taps; e += nth) kw[e] = sp.W[e];
  // 1. stage the lines
  if (AXIS0) {
    const T *src = in + line0 * n;
    for (int ti = warp_; ti < valid; ti += nwarp_) {
      const T *ln = src + (long long)ti * n;
      for (int j = lane_; j < n_i; j += 32) buf[(pad + j) * pitch + ti] = (double)ln[j];
    }
  } else {
    const T *src = in + tile_i * TI + o * sp.inner * n;
    for (int j = warp_; j < n_i; j += nwarp_) {
      const T *row = src + (long long)j * sp.inner;
      for (int ti = lane_; ti < valid; ti += 32) buf[(pad + j) * pitch + ti] = (double)row[ti];
    }
  }
  __syncthreads();
  // 2. ghosts: one thread per line
  if (tid < valid) {
    double mean = line_mean(buf, pad, pitch, tid, n);
    if (SMOOTH) {
      ghost_gauss_side(buf, pad, pitch, tid, n, mean, sp.eqs, false);
      ghost_gauss_side(buf, pad, pitch, tid, n, mean, sp.eqs, true);
    } else {
      ghost_poly_side(buf, pad, pitch, tid, n, mean, sp.h, sp.eqs, sp.bcl, false, gm);
      ghost_poly_side(buf, pad, pitch, tid, n, mean, sp.h, sp.eqs, sp.bcr, true, gm);
    }
  }
  __syncthreads();
  // 3. outputs
  if (SMOOTH) {
    // thread = (line, block of R consecutive outputs): the R outputs share a sliding window of inputs kept in
    // registers, so one LDS feeds R DFMAs; taps are accumulated in the reference's order l = -E..E
    constexpr int R = SMOOTH_R;
    const int taps = sp.taps;
    const int nblk = (int)((nout + R - 1) / R);
    for (int kb = warp_; kb < nblk; kb += nwarp_)
    for (int ti = lane_; ti < valid; ti += 32) {
      const long long k0 = (long long)kb * R;
      double acc[R], win[R];
      int off = (int)k0 * pitch + ti;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r] = 0.0;
        win[r] = buf[off + r * pitch];
      }
      off += R * pitch;
      for (int l = 0; l < taps; l += R) {
#pragma unroll
        for (int u = 0; u < R; ++u) {
          if (l + u < taps) {  // uniform
            const double wv = kw[l + u];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(win[(u + r) % R], wv, acc[r]);
            win[u] = buf[off];  // element k0 + (l+u) + R replaces k0 + (l+u), which no output needs any more
            off += pitch;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const long long k = k0 + r;
        if (k < nout) {
          if (AXIS0) out[(line0 + ti) * nout + k] = (T)acc[r];
          else out[tile_i * TI + ti + k * sp.inner + o * sp.inner * nout] = (T)acc[r];
        }
      }
    }
  } else if (AXIS0) {
    // lane = output knot k (stores coalesced along the line); its 2 eqs blended weights stay in registers while
    // the thread walks the lines of the tile
    const int kthreads = nout >= nth ? nth : (int)((nout + 31) / 32 * 32);
    const int lgroups = nth / kthreads;
    const int kk = tid % kthreads, lg = tid / kthreads;
    if (lg < lgroups) {
      for (long long k = kk; k < nout; k += kthreads) {
        const int base = __ldg(sp.base + k);
        double w[TAPS > 0 ? TAPS : 1];
#pragma unroll
        for (int l = 0; l < TAPS; ++l) w[l] = __ldg(sp.W + (long long)l * nout + k);
        for (int ti = lg; ti < valid; ti += lgroups) {
          int off = base * pitch + ti;
          double acc = 0.0;
#pragma unroll
          for (int l = 0; l < TAPS; ++l) {
            acc = fma(buf[off], w[l], acc);
            off += pitch;
          }
          out[(line0 + ti) * nout + k] = (T)__dmul_rn(acc, sp.scale);
        }
      }
    }
  } else {
    // lane = line (loads from the tile and global stores coalesced along `inner`); a warp owns one output knot at
    // a time: its weights are warp-uniform registers reused for every line of the tile
    const int warp = tid >> 5, lane = tid & 31, nwarp = nth >> 5;
    for (long long k = warp; k < nout; k += nwarp) {
      const int base = __ldg(sp.base + k);
      double w[TAPS > 0 ? TAPS : 1];
#pragma unroll
      for (int l = 0; l < TAPS; ++l) w[l] = __ldg(sp.W + (long long)l * nout + k);
      T *orow = out + tile_i * TI + k * sp.inner + o * sp.inner * nout;
      for (int ti = lane; ti < valid; ti += 32) {
        int off = base * pitch + ti;
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < TAPS; ++l) {
          acc = fma(buf[off], w[l], acc);
          off += pitch;
        }
        orow[ti] = (T)__dmul_rn(acc, sp.scale);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
static void lagrange_rows(int m, int rows, double *out, int ld) {
  for (int j = 1; j <= rows; ++j)
    for (int k = 0; k < m; ++k) {
      long long num = 1, den = 1;
      for (int i = 0; i < m; ++i)
        if (i != k) {
          num *= (long long)(-j - i);
          den *= (long long)(k - i);
        }
      out[(j - 1) * ld + k] = (double)(num / den);
    }
}

static double pow_int(double x, int n) {  // same restatement as cib200_eval.cu (Julia pow_body)
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

static int max_dyn_smem() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (v <= 0) v = 48 * 1024;
  }
  return v;
}

template <typename T, bool SMOOTH, bool AXIS0, int TAPS>
static int launch_kernel(const SweepParams &sp, const GhostMats &gm, const T *in, T *out, long long blocks, size_t smem,
                         cudaStream_t st) {
  auto k = sweep_kernel<T, SMOOTH, AXIS0, TAPS, 0>;
  if (sp.pitch == (AXIS0 ? 33 : 32)) k = sweep_kernel<T, SMOOTH, AXIS0, TAPS, AXIS0 ? 33 : 32>;
  if (smem > 48 * 1024) CIB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<(unsigned)blocks, 256, smem, st>>>(sp, gm, in, out);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

template <typename T, bool SMOOTH, bool AXIS0>
static int launch_taps(const SweepParams &sp, const GhostMats &gm, const T *in, T *out, long long blocks, size_t smem,
                       cudaStream_t st) {
  if (SMOOTH) return launch_kernel<T, SMOOTH, AXIS0, 0>(sp, gm, in, out, blocks, smem, st);
  switch (sp.taps) {
    case 10: return launch_kernel<T, false, AXIS0, 10>(sp, gm, in, out, blocks, smem, st);
    case 12: return launch_kernel<T, false, AXIS0, 12>(sp, gm, in, out, blocks, smem, st);
    case 14: return launch_kernel<T, false, AXIS0, 14>(sp, gm, in, out, blocks, smem, st);
    case 16: return launch_kernel<T, false, AXIS0, 16>(sp, gm, in, out, blocks, smem, st);
    case 18: return launch_kernel<T, false, AXIS0, 18>(sp, gm, in, out, blocks, smem, st);
  }
  set_error("sweep: unsupported tap count %d", sp.taps);
  return CIB200_EUNSUPPORTED;
}

template <typename T, bool SMOOTH>
static int launch_pass(SweepParams sp, const GhostMats &gm, const T *in, T *out, bool axis0, cudaStream_t st) {
  // tile width: as many lines as fit, a power of two, at most 32 (the common full tile gets a compile-time pitch)
  const long long lines = axis0 ? sp.outer : sp.inner;
  const long long nrow = sp.n_in + 2 * sp.pad + (SMOOTH ? 2 * SMOOTH_R : 0);
  const size_t budget = (size_t)std::min(max_dyn_smem(), 100 * 1024);
  int ti = 32;
  while (ti > 1 && (long long)ti / 2 >= lines) ti /= 2;
  auto bytes_for = [&](int t) {
    int pitch = axis0 ? (t > 1 ? t + 1 : 1) : t;
    return sizeof(double) * ((size_t)nrow * pitch + (SMOOTH ? sp.taps : 0));
  };
  while (ti > 1 && bytes_for(ti) > budget) ti /= 2;
  if (bytes_for(ti) > (size_t)max_dyn_smem()) {
    set_error("sweep: a line of %lld points (+%d ghosts per side) does not fit in shared memory; lines up to ~%d points are supported",
              sp.n_in, sp.pad, (int)(max_dyn_smem() / 8) - 2 * sp.pad - sp.taps);
    return CIB200_EUNSUPPORTED;
  }
  sp.ti = ti;
  sp.pitch = axis0 ? (ti > 1 ? ti + 1 : 1) : ti;
  const size_t smem = bytes_for(ti);
  long long blocks = axis0 ? (sp.outer + ti - 1) / ti : ((sp.inner + ti - 1) / ti) * sp.outer;
  if (blocks > 0x7fffffffLL) {
    set_error("sweep: grid too large");
    return CIB200_EUNSUPPORTED;
  }
  if (axis0) return launch_taps<T, SMOOTH, true>(sp, gm, in, out, blocks, smem, st);
  return launch_taps<T, SMOOTH, false>(sp, gm, in, out, blocks, smem, st);
}

struct TempBufs {
  std::vector<void *> ptrs;
  cudaStream_t st;
  ~TempBufs() {
    for (void *p : ptrs) cudaFreeAsync(p, st);
  }
  int alloc(void **p, size_t bytes) {
    CIB_CUDA(cudaMallocAsync(p, bytes, st));
    ptrs.push_back(*p);
    return CIB200_OK;
  }
};

template <typename T>
static int resample_t(int ndim, const long long *n_in, const long long *n_out, const T *d_in, T *d_out,
                      const double *const *knots_in, const double *const *knots_out, const int *kernel,
                      const int *deriv, const int *bcl, const int *bcr, cudaStream_t st) {
  TempBufs tmp;
  tmp.st = st;
  long long shape[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) shape[d] = n_in[d];
  const T *cur = d_in;
  int rc;
  for (int d = 0; d < ndim; ++d) {
    const int eqs = kernel_eqs(kernel[d]);
    const int taps = 2 * eqs;
    const long long nin = shape[d], nout = n_out[d];
    const double h = knots_in[d][1] - knots_in[d][0];                 // convolution_resample.jl:270
    const double x_expanded_1 = knots_in[d][0] - (double)(eqs - 1) * h;  // :285
    const long long n_coef = nin + 2 * (eqs - 1);
    // tables + output knots + weights on the device
    double *d_tk, *d_td, *d_ko, *d_W;
    int *d_base;
    const size_t tab_bytes = sizeof(double) * CIB_TAB_ROWS * taps;
    if ((rc = tmp.alloc((void **)&d_tk, tab_bytes))) return rc;
    if ((rc = tmp.alloc((void **)&d_td, tab_bytes))) return rc;
    if ((rc = tmp.alloc((void **)&d_ko, sizeof(double) * (size_t)nout))) return rc;
    if ((rc = tmp.alloc((void **)&d_W, sizeof(double) * (size_t)nout * taps))) return rc;
    if ((rc = tmp.alloc((void **)&d_base, sizeof(int) * (size_t)nout))) return rc;
    CIB_CUDA(cudaMemcpyAsync(d_tk, host_table(kernel[d], deriv[d]), tab_bytes, cudaMemcpyHostToDevice, st));
    CIB_CUDA(cudaMemcpyAsync(d_td, host_table(kernel[d], deriv[d] + 1), tab_bytes, cudaMemcpyHostToDevice, st));
    CIB_CUDA(cudaMemcpyAsync(d_ko, knots_out[d], sizeof(double) * (size_t)nout, cudaMemcpyHostToDevice, st));
    resample_weights_kernel<<<(unsigned)((nout + 127) / 128), 128, 0, st>>>(d_ko, nout, x_expanded_1, h, eqs, n_coef,
                                                                           d_tk, d_td, d_W, d_base);
    CIB_LAUNCHED();
    CIB_CUDA(cudaGetLastError());

    SweepParams sp;
    memset(&sp, 0, sizeof sp);
    sp.inner = 1;
    for (int j = 0; j < d; ++j) sp.inner *= shape[j];
    sp.outer = 1;
    for (int j = d + 1; j < ndim; ++j) sp.outer *= shape[j];
    sp.n_in = nin;
    sp.n_out = nout;
    sp.pad = eqs - 1;
    sp.eqs = eqs;
    sp.taps = taps;
    sp.bcl = bcl[d];
    sp.bcr = bcr[d];
    sp.h = h;
    sp.scale = pow_int(-1.0 / h, deriv[d]);  // :379
    sp.W = d_W;
    sp.base = d_base;
    GhostMats gm;
    memset(&gm, 0, sizeof gm);
    gm.poly_nodes = kernel[d] == CIB200_B5 ? 6 : 8;
    lagrange_rows(gm.poly_nodes, eqs - 1, &gm.poly[0][0], 8);
    lagrange_rows(2, eqs - 1, &gm.lin[0][0], 2);
    lagrange_rows(3, eqs - 1, &gm.quad[0][0], 3);

    T *next;
    if (d == ndim - 1) {
      next = d_out;
    } else {
      long long tot = 1;
      for (int j = 0; j < ndim; ++j) tot *= (j == d ? nout : shape[j]);
      if ((rc = tmp.alloc((void **)&next, sizeof(T) * (size_t)tot))) return rc;
    }
    const bool axis0 = (sp.inner == 1);
    if (axis0) {
      // axis 0 (or leading singleton axes): lines are contiguous; fold outer
      if ((rc = launch_pass<T, false>(sp, gm, cur, next, true, st))) return rc;
    } else {
      if ((rc = launch_pass<T, false>(sp, gm, cur, next, false, st))) return rc;
    }
    cur = next;
    shape[d] = nout;
  }
  return CIB200_OK;
}

int sweep_resample(int ndim, int dtype, const long long *n_in, const long long *n_out, const void *d_in, void *d_out,
                   const double *const *knots_in, const double *const *knots_out, const int *kernel, const int *deriv,
                   const int *bcl, const int *bcr, cudaStream_t st) {
  if (dtype == CIB200_F32)
    return resample_t<float>(ndim, n_in, n_out, (const float *)d_in, (float *)d_out, knots_in, knots_out, kernel, deriv, bcl, bcr, st);
  return resample_t<double>(ndim, n_in, n_out, (const double *)d_in, (double *)d_out, knots_in, knots_out, kernel, deriv, bcl, bcr, st);
}

// convolution_kernels/convolution_kernel_gaussian.jl:23-31
static double theta(double B, int terms) {
  double q = exp(-B);
  double sum = 1.0;
  for (int n = 1; n <= terms; ++n) sum += 2 * pow_int(q, n * n);
  return sum;
}

template <typename T>
static int smooth_t(int ndim, const long long *n, const T *d_in, T *d_out, const double *h, double B, cudaStream_t st) {
  TempBufs tmp;
  tmp.st = st;
  const int E = cib200_smooth_halfwidth(B);
  const int taps = 2 * E + 1;
  std::vector<double> kw((size_t)taps);
  const double norm = theta(B, E);
  for (int l = -E; l <= E; ++l) kw[(size_t)(l + E)] = exp(-B * ((double)l * (double)l)) / norm;  // :59-63
  double *d_kw;
  int rc;
  if ((rc = tmp.alloc((void **)&d_kw, sizeof(double) * (size_t)taps))) return rc;
  CIB_CUDA(cudaMemcpyAsync(d_kw, kw.data(), sizeof(double) * (size_t)taps, cudaMemcpyHostToDevice, st));
  // (pageable source: the runtime stages the copy before returning, so kw may go out of scope)
  long long tot = 1;
  for (int d = 0; d < ndim; ++d) tot *= n[d];
  T *scratch[2] = {nullptr, nullptr};
  if (ndim > 1 && (rc = tmp.alloc((void **)&scratch[0], sizeof(T) * (size_t)tot))) return rc;
  if (ndim > 2 && (rc = tmp.alloc((void **)&scratch[1], sizeof(T) * (size_t)tot))) return rc;
  const T *cur = d_in;
  GhostMats gm;
  memset(&gm, 0, sizeof gm);
  for (int d = 0; d < ndim; ++d) {
    SweepParams sp;
    memset(&sp, 0, sizeof sp);
    sp.inner = 1;
    for (int j = 0; j < d; ++j) sp.inner *= n[j];
    sp.outer = 1;
    for (int j = d + 1; j < ndim; ++j) sp.outer *= n[j];
    sp.n_in = sp.n_out = n[d];
    sp.pad = E;
    sp.eqs = E;
    sp.taps = taps;
    sp.h = h[d];
    sp.scale = 1.0;
    sp.W = d_kw;
    T *next = (d == ndim - 1) ? d_out : scratch[d & 1];
    if ((rc = launch_pass<T, true>(sp, gm, cur, next, sp.inner == 1, st))) return rc;
    cur = next;
  }
  return CIB200_OK;
}

int sweep_smooth(int ndim, int dtype, const long long *n, const void *d_in, void *d_out, const double *h, double B,
                 cudaStream_t st) {
  if (dtype == CIB200_F32) return smooth_t<float>(ndim, n, (const float *)d_in, (float *)d_out, h, B, st);
  return smooth_t<double>(ndim, n, (const double *)d_in, (double *)d_out, h, B, st);
}

}  // namespace cib
