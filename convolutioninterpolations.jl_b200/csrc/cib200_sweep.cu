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
  int w_smem;     // resample, axis >= 1: blended weights + first-tap indices are staged in shared memory
  double h;
  double scale;   // (-1/h)^derivative (resample) ; 1 (smooth)
  const double *W;     // resample: [taps][n_out] blended weights ; smooth: kw[taps]
  const double *Wk;    // resample: the same weights as [n_out][taps] (16-byte aligned rows: taps is even)
  const int *base;     // resample: first tap index into the padded line (0-based) per output
};

// ---- per-output-knot weights: convolution_resample.jl:320-361 minus the data ------------------
__global__ void resample_weights_kernel(const double *__restrict__ knots_out, long long n_out, double x_expanded_1,
                                        double h, int eqs, long long n_coef, const double *__restrict__ tk,
                                        const double *__restrict__ td, double *__restrict__ W, double *__restrict__ Wk,
                                        int *__restrict__ base) {
  long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_out) return;
  AxisPos p;
  locate_computed(knots_out[k], x_expanded_1, h, eqs, n_coef, p);
  base[k] = p.base;
  Hermite hb = hermite_basis(p.t);
  const int w = 2 * eqs;
  for (int l = 0; l < w; ++l) {
    const double v = blended_weight(hb, tk, td, p.row, l, w);
    W[(long long)l * n_out + k] = v;
    Wk[k * w + l] = v;
  }
}

// ---- ghost fill of one line held in shared memory ----------------------------------------------
// bl = buf + ti; line(j) = bl[(pad + j) * pitch], j = 0..n-1.  Everything below is lane-private (lane = line),
// fully unrolled with compile-time node counts so that no array is indexed dynamically (no local memory), and
// plain mul/add via __d*_rn so that the :detect branch decisions and the ghost values do not depend on fma
// contraction.
// The reference computes sum(slice)/n with Julia's `sum` (create_convolutional_coefs.jl:242,312): a @simd loop whose
// lanes are partial sums in an unspecified (vector-width dependent) association.  Here: 8 interleaved partial sums
// combined pairwise - the dependent chain is n/8 additions instead of n, and the result is within a few ulp of any
// other association (the oracle's sequential loop included).
__device__ __forceinline__ double line_mean(const double *bl, int pad, int pitch, int n) {
  double s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.0;
  const double *p = bl + pad * pitch;
  int j = 0;
  for (; j + 8 <= n; j += 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = __dadd_rn(s[i], p[(j + i) * pitch]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (j + i < n) s[i] = __dadd_rn(s[i], p[(j + i) * pitch]);
  const double sum = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                               __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
  return __ddiv_rn(sum, (double)n);
}

// NS = polynomial node count of the kernel (6 for b5, 8 otherwise), EQS-1 ghosts per side
template <int NS, int EQS>
__device__ __forceinline__ void ghost_poly_side(double *bl, int pad, int pitch, int n, double mean, double h, int bc,
                                                bool right, const GhostMats &gm) {
  double y[8];  // centred boundary values, y[0] = the boundary point itself, then inward
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = right ? (n - 1 - k) : k;
    y[k] = k < n ? __dsub_rn(bl[(pad + j) * pitch], mean) : 0.0;
  }
  bool use_poly = false;
  if (n >= NS && (bc == CIB200_BC_DETECT || bc == CIB200_BC_POLY)) {
    if (bc == CIB200_BC_POLY) {
      use_poly = true;
    } else {
      // natural-order window p[0..NS-1]: left p[i] = y[i]; right p[i] = y[NS-1-i]
      double d1[NS - 1], d2[NS - 2];
      double mn = INFINITY, mx = -INFINITY;
#pragma unroll
      for (int d = 0; d < NS - 1; ++d) {
        const double a = right ? y[NS - 2 - d] : y[d + 1];
        const double b = right ? y[NS - 1 - d] : y[d];
        d1[d] = __dsub_rn(a, b);
        mn = fmin(mn, d1[d]);
        mx = fmax(mx, d1[d]);
      }
      const bool periodic = __ddiv_rn(__dmul_rn(mn, mx), __dmul_rn(h, h)) < -1.0 / 10;
      mn = INFINITY, mx = -INFINITY;
#pragma unroll
      for (int d = 0; d < NS - 2; ++d) {
        d2[d] = __dsub_rn(d1[d + 1], d1[d]);
        mn = fmin(mn, d2[d]);
        mx = fmax(mx, d2[d]);
      }
      const bool high_curv = __ddiv_rn(fabs(__dsub_rn(mx, mn)), h) > 1.0 / 2;
      bool high_d3 = false;
      if (NS > 4) {
        mn = INFINITY, mx = -INFINITY;
#pragma unroll
        for (int d = 0; d < NS - 3; ++d) {
          const double v = __dsub_rn(d2[d + 1], d2[d]);
          mn = fmin(mn, v);
          mx = fmax(mx, v);
        }
        high_d3 = __ddiv_rn(fabs(__dsub_rn(mx, mn)), h) > 1.0 / 2;
      }
      use_poly = !periodic && !high_curv && !high_d3;
    }
  }
  const int row0 = right ? (pad + n) : (pad - 1);  // slot of ghost j = 1
  const int step = right ? pitch : -pitch;
  double *g = bl + row0 * pitch;
  if (use_poly) {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NS; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.poly[j - 1][k], y[k]));
      g[(j - 1) * step] = __dadd_rn(mean, acc);
    }
  } else if (bc == CIB200_BC_QUADRATIC) {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.quad[j - 1][k], y[k]));
      g[(j - 1) * step] = __dadd_rn(mean, acc);
    }
  } else {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.lin[j - 1][k], y[k]));
      g[(j - 1) * step] = __dadd_rn(mean, acc);
    }
  }
}

// fill_ghost_points_gaussian! with convolution_smooth's buffer geometry: padding E, boundary index from
// get_boundary_indices(size = n + 2E, eqs = E) -> ghosts land one slot further out, slot next to the
// interior stays 0.  Values are the centred least-squares line (mean NOT added back in
// exact arithmetic: y_mean + (a + b(1-k) - y_mean)).
__device__ __forceinline__ void ghost_gauss_side(double *bl, int pad, int pitch, int n, double mean, int E, bool right) {
  int n_fit = E / 3 > 6 ? E / 3 : 6;
  if (n_fit > n) n_fit = n;
  // sx = sum k and sxx = sum k^2 are sums of small integers: exact in any order, so closed forms give the same bits
  const double nfd = (double)n_fit;
  const double sx = nfd * (nfd + 1.0) * 0.5, sxx = nfd * (nfd + 1.0) * (2.0 * nfd + 1.0) / 6.0;
  double sy = 0, sxy = 0, xk = 1.0;
  const double *q = bl + (pad + (right ? n - 1 : 0)) * pitch;
  const int qs = right ? -pitch : pitch;
#pragma unroll 2
  for (int k = 0; k < n_fit; ++k) {
    const double yk = __dsub_rn(q[k * qs], mean);
    sy = __dadd_rn(sy, yk);
    sxy = __dadd_rn(sxy, __dmul_rn(xk, yk));
    xk += 1.0;
  }
  const double nf = (double)n_fit;
  const double denom = __dsub_rn(__dmul_rn(nf, sxx), __dmul_rn(sx, sx));
  const double b = __ddiv_rn(__dsub_rn(__dmul_rn(nf, sxy), __dmul_rn(sx, sy)), denom);
  const double a = __ddiv_rn(__dsub_rn(sy, __dmul_rn(b, sx)), nf);
  double *g = bl + (right ? (pad + n + 1) : (pad - 2)) * pitch;  // slot of ghost k = 1
  const int step = right ? pitch : -pitch;
  double c = 0.0;  // 1 - k
#pragma unroll 4
  for (int k = 1; k <= E - 1; ++k) {
    const double y_temp = __dsub_rn(__dadd_rn(a, __dmul_rn(b, c)), mean);
    g[(k - 1) * step] = __dadd_rn(mean, y_temp);
    c -= 1.0;
  }
}

static constexpr int SMOOTH_R = 8;    // consecutive outputs per thread in the Gaussian FIR (sliding register window)
static constexpr int RESAMPLE_R = 8;  // anchor offsets a register window of the resample pass can serve

// sum_l win[O + l] * w[l] with static register indices; three partial sums keep the FP64 pipe busy from one warp
template <int TAPS, int WIN, int O>
__device__ __forceinline__ double window_dot(const double (&win)[WIN], const double (&w)[TAPS]) {
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
  for (int l = 0; l < TAPS; l += 3) {
    a0 = fma(win[O + l], w[l], a0);
    if (l + 1 < TAPS) a1 = fma(win[O + l + 1], w[l + 1], a1);
    if (l + 2 < TAPS) a2 = fma(win[O + l + 2], w[l + 2], a2);
  }
  return (a0 + a1) + a2;
}

// Phase 3 of a resample pass along an axis >= 1, one lane = one line.  A warp walks a contiguous range of output
// knots.  Its inputs sit in a register window of TAPS + R - 1 consecutive rows anchored at row a: an output whose
// first tap is row a + o, 0 <= o < R, is served from registers (one shared-memory load feeds ~TAPS/ratio DFMAs when
// upsampling); the window is re-anchored when o leaves that range.  The weights of an output knot are warp-uniform:
// they are fetched with uniform 16-byte loads one knot ahead (from shared memory when they fit next to the tile).
template <bool WS>
__device__ __forceinline__ double2 load_w2(const double2 *p) {
  if (WS) {  // explicit ld.shared: a pointer that may be shared or global would compile to generic loads
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
  }
  return __ldg(p);
}
template <bool WS>
__device__ __forceinline__ int load_b(const int *p) {
  if (WS) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
  }
  return __ldg(p);
}

template <typename T, int TP, bool WS>
__device__ __forceinline__ void resample_lines(const double *bl, const int pitch, const double2 *__restrict__ wsrc,
                                               const int *__restrict__ bsrc, int kbeg, int kend, T *optr,
                                               long long ostride, double scale) {
  constexpr int R = RESAMPLE_R, WIN = TP + R - 1;
  double win[WIN];
  int a = load_b<WS>(bsrc + kbeg);
#pragma unroll
  for (int j = 0; j < WIN; ++j) win[j] = bl[(a + j) * pitch];
  double wA[TP], wB[TP];
  auto fetch = [&](int k, double(&w)[TP]) {
    const double2 *src = wsrc + k * (TP / 2);
#pragma unroll
    for (int l = 0; l < TP; l += 2) {
      const double2 v = load_w2<WS>(src + (l >> 1));
      w[l] = v.x;
      w[l + 1] = v.y;
    }
  };
  auto emit = [&](int b, const double(&w)[TP]) {
    int oo = b - a;
    if ((unsigned)oo >= (unsigned)R) {  // warp-uniform
      a = b;
      oo = 0;
#pragma unroll
      for (int j = 0; j < WIN; ++j) win[j] = bl[(a + j) * pitch];
    }
    double acc;
    switch (oo) {
      case 0: acc = window_dot<TP, WIN, 0>(win, w); break;
      case 1: acc = window_dot<TP, WIN, 1>(win, w); break;
      case 2: acc = window_dot<TP, WIN, 2>(win, w); break;
      case 3: acc = window_dot<TP, WIN, 3>(win, w); break;
      case 4: acc = window_dot<TP, WIN, 4>(win, w); break;
      case 5: acc = window_dot<TP, WIN, 5>(win, w); break;
      case 6: acc = window_dot<TP, WIN, 6>(win, w); break;
      default: acc = window_dot<TP, WIN, 7>(win, w); break;
    }
    *optr = (T)__dmul_rn(acc, scale);
    optr += ostride;
  };
  fetch(kbeg, wA);
  int b0 = a;
  for (int k = kbeg; k < kend; k += 2) {
    const int k1 = k + 1 < kend ? k + 1 : k;
    const int b1 = load_b<WS>(bsrc + k1);
    fetch(k1, wB);
    emit(b0, wA);
    if (k + 1 < kend) {
      const int k2 = k + 2 < kend ? k + 2 : k1;
      b0 = load_b<WS>(bsrc + k2);
      fetch(k2, wA);
      emit(b1, wB);
    }
  }
}

// TAPS = 2 eqs (compile time) for resample, 0 for smooth (run-time 2E+1 taps)
// CPITCH = compile-time buffer pitch (shared-memory loads get immediate offsets), 0 = run-time pitch
template <typename T, bool SMOOTH, bool AXIS0, int TAPS, int CPITCH>
__global__ void __launch_bounds__(256) sweep_kernel(SweepParams sp, GhostMats gm, const T *__restrict__ in,
                                                    T *__restrict__ out) {
  extern __shared__ double sm[];
  const int pad = sp.pad, pitch = CPITCH > 0 ? CPITCH : sp.pitch, TI = sp.ti;
  const int n = (int)sp.n_in, nout = (int)sp.n_out;
  const int nrow = n + 2 * pad;
  double *buf = sm;
  // rows past the line: the register windows of phase 3 may read (never use) up to R-1 rows beyond it
  double *kw = sm + (nrow + (SMOOTH ? 2 * SMOOTH_R : RESAMPLE_R)) * pitch;  // smooth only
  const int tid = threadIdx.x, nth = blockDim.x;
  const int warp_ = tid >> 5, lane_ = tid & 31, nwarp_ = nth >> 5;

  // tile coordinates
  long long tile_i = 0, o = 0;
  long long line0 = 0;  // AXIS0: first line of the tile
  int valid;            // lines of this tile that exist
  if (AXIS0) {
    line0 = (long long)blockIdx.x * TI;
    const long long rem = sp.outer - line0;
    valid = rem < TI ? (int)rem : TI;
  } else {
    const long long tiles_inner = (sp.inner + TI - 1) / TI;
    tile_i = blockIdx.x % tiles_inner;
    o = blockIdx.x / tiles_inner;
    const long long rem = sp.inner - tile_i * TI;
    valid = rem < TI ? (int)rem : TI;
  }

  // (all tile loops are warp-per-row / lane-per-column: no integer division in the hot path)
  // 0. clear ghost slots (the reference zero-fills the line buffer: resample :309, smooth :96)
  for (int r = warp_; r < 2 * pad; r += nwarp_) {
    const int slot = r < pad ? r : (n + r);
    for (int ti = lane_; ti < TI; ti += 32) buf[slot * pitch + ti] = 0.0;
  }
  if (SMOOTH)
    for (int e = tid; e < sp.taps; e += nth) kw[e] = sp.W[e];
  // resample, axis >= 1: the per-output-knot weights (warp-uniform in phase 3) next to the tile
  double *wsm = kw;
  int *bsm = (int *)(wsm + (TAPS > 0 ? TAPS : 2) * nout);
  if (!SMOOTH && !AXIS0 && sp.w_smem) {
    const double2 *src = (const double2 *)sp.Wk;
    double2 *dst = (double2 *)wsm;
    const int nv = (TAPS > 0 ? TAPS : 2) * nout / 2;
    for (int e = tid; e < nv; e += nth) dst[e] = __ldg(src + e);
    for (int e = tid; e < nout; e += nth) bsm[e] = __ldg(sp.base + e);
  }
  // 1. stage the lines: explicit batches of independent loads (a whole batch is in flight before its first use)
  if (AXIS0) {
    const T *src = in + line0 * n;
    constexpr int LB = 4, JB = 2;
    for (int t0 = warp_; t0 < valid; t0 += nwarp_ * LB)
      for (int j0 = lane_; j0 < n; j0 += 32 * JB) {
        double v[LB][JB];
#pragma unroll
        for (int u = 0; u < LB; ++u)
#pragma unroll
          for (int w = 0; w < JB; ++w) {
            const int ti = t0 + u * nwarp_, j = j0 + 32 * w;
            v[u][w] = (ti < valid && j < n) ? (double)src[(long long)ti * n + j] : 0.0;
          }
#pragma unroll
        for (int u = 0; u < LB; ++u)
#pragma unroll
          for (int w = 0; w < JB; ++w) {
            const int ti = t0 + u * nwarp_, j = j0 + 32 * w;
            if (ti < valid && j < n) buf[(pad + j) * pitch + ti] = v[u][w];
          }
      }
  } else {
    const T *src = in + tile_i * TI + o * sp.inner * n;
    if (lane_ < valid) {
      const T *col = src + lane_;
      double *dst = buf + pad * pitch + lane_;
      constexpr int RB = 16;
      for (int j0 = warp_; j0 < n; j0 += nwarp_ * RB) {
        double v[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int j = j0 + u * nwarp_;
          v[u] = j < n ? (double)col[(long long)j * sp.inner] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int j = j0 + u * nwarp_;
          if (j < n) dst[j * pitch] = v[u];
        }
      }
    }
  }
  __syncthreads();
  // 2. ghosts: lane = line, warp 0 fills the left side and warp 1 the right side (each sums the line itself, in
  //    the same order, so both use the same mean)
  for (int side = warp_; side < 2; side += nwarp_) {
    if (lane_ < valid) {
      double *bl = buf + lane_;
      const double mean = line_mean(bl, pad, pitch, n);
      if (SMOOTH) {
        ghost_gauss_side(bl, pad, pitch, n, mean, sp.eqs, side == 1);
      } else {
        constexpr int EQS = TAPS > 0 ? TAPS / 2 : 5;
        constexpr int NS = EQS == 5 ? 6 : 8;
        ghost_poly_side<NS, EQS>(bl, pad, pitch, n, mean, sp.h, side ? sp.bcr : sp.bcl, side == 1, gm);
      }
    }
  }
  __syncthreads();
  // 3. outputs
  if (SMOOTH) {
    // thread = (line, block of R consecutive outputs): the R outputs share a sliding window of inputs kept in
    // registers, so one LDS feeds R DFMAs; taps are accumulated in the reference's order l = -E..E
    constexpr int R = SMOOTH_R;
    const int taps = sp.taps;
    const int nblk = (nout + R - 1) / R;
    for (int kb = warp_; kb < nblk; kb += nwarp_)
      for (int ti = lane_; ti < valid; ti += 32) {
        const int k0 = kb * R;
        double acc[R], win[R];
        const double *p = buf + k0 * pitch + ti;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          acc[r] = 0.0;
          win[r] = p[r * pitch];
        }
        p += R * pitch;
        int l = 0;
        for (; l + R <= taps; l += R) {  // full groups: no predicates
#pragma unroll
          for (int u = 0; u < R; ++u) {
            const double wv = kw[l + u];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(win[(u + r) % R], wv, acc[r]);
            win[u] = p[u * pitch];  // element k0 + (l+u) + R replaces k0 + (l+u), which no output needs any more
          }
          p += R * pitch;
        }
#pragma unroll
        for (int u = 0; u < R; ++u) {
          if (l + u < taps) {  // uniform
            const double wv = kw[l + u];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fma(win[(u + r) % R], wv, acc[r]);
            win[u] = p[u * pitch];
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int k = k0 + r;
          if (k < nout) {
            if (AXIS0) out[(line0 + ti) * nout + k] = (T)acc[r];
            else out[tile_i * TI + ti + (long long)k * sp.inner + o * sp.inner * nout] = (T)acc[r];
          }
        }
      }
  } else if (AXIS0) {
    // lane = output knot k (stores coalesced along the line); its 2 eqs blended weights stay in registers while
    // the thread walks the lines of the tile
    const int kthreads = nout >= nth ? nth : ((nout + 31) / 32 * 32);
    const int lgroups = nth / kthreads;
    const int kk = tid % kthreads, lg = tid / kthreads;
    if (lg < lgroups) {
      for (int k = kk; k < nout; k += kthreads) {
        const int base = __ldg(sp.base + k);
        double w[TAPS > 0 ? TAPS : 1];
#pragma unroll
        for (int l = 0; l < TAPS; ++l) w[l] = __ldg(sp.W + (long long)l * nout + k);
        T *orow = out + line0 * nout + k;
#pragma unroll 2
        for (int ti = lg; ti < valid; ti += lgroups) {
          const double *p = buf + base * pitch + ti;
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int l = 0; l < TAPS; l += 2) {
            a0 = fma(p[l * pitch], w[l], a0);
            a1 = fma(p[(l + 1) * pitch], w[l + 1], a1);
          }
          orow[(long long)ti * nout] = (T)__dmul_rn(a0 + a1, sp.scale);
        }
      }
    }
  } else {
    // lane = line (loads from the tile and global stores coalesced along `inner`): resample_lines
    constexpr int TPX = (TAPS > 0) ? TAPS : 2;
    const int kper = (nout + nwarp_ - 1) / nwarp_;
    const int kbeg = warp_ * kper, kend = min(nout, kbeg + kper);
    if (kbeg < kend && lane_ < valid) {
      T *optr = out + tile_i * TI + lane_ + o * sp.inner * nout + (long long)kbeg * sp.inner;
      if (sp.w_smem)
        resample_lines<T, TPX, true>(buf + lane_, pitch, (const double2 *)wsm, bsm, kbeg, kend, optr, sp.inner, sp.scale);
      else
        resample_lines<T, TPX, false>(buf + lane_, pitch, (const double2 *)sp.Wk, sp.base, kbeg, kend, optr, sp.inner, sp.scale);
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
  int threads = 256;
  if (SMOOTH) threads = 32 * (int)std::max<long long>(2, std::min<long long>(8, (sp.n_out + SMOOTH_R - 1) / SMOOTH_R));
  else if (!AXIS0) threads = 128;
  k<<<(unsigned)blocks, threads, smem, st>>>(sp, gm, in, out);
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
  const long long nrow = sp.n_in + 2 * sp.pad + (SMOOTH ? 2 * SMOOTH_R : RESAMPLE_R);
  const size_t budget = (size_t)std::min(max_dyn_smem(), 100 * 1024);
  int ti = 32;
  while (ti > 1 && (long long)ti / 2 >= lines) ti /= 2;
  auto bytes_for = [&](int t) {
    int pitch = axis0 ? (t > 1 ? t + 1 : 1) : t;
    return sizeof(double) * ((size_t)nrow * pitch + (SMOOTH ? sp.taps : 0));
  };
  // resample along axis >= 1: weights [n_out][taps] + first-tap indices next to the tile when three CTAs still fit
  const size_t w_bytes = (!SMOOTH && !axis0) ? sizeof(double) * (size_t)sp.taps * sp.n_out + sizeof(int) * (size_t)sp.n_out : 0;
  while (ti > 1 && bytes_for(ti) > budget) ti /= 2;
  if (bytes_for(ti) > (size_t)max_dyn_smem()) {
    set_error("sweep: a line of %lld points (+%d ghosts per side) does not fit in shared memory; lines up to ~%d points are supported",
              sp.n_in, sp.pad, (int)(max_dyn_smem() / 8) - 2 * sp.pad - sp.taps);
    return CIB200_EUNSUPPORTED;
  }
  sp.ti = ti;
  sp.pitch = axis0 ? (ti > 1 ? ti + 1 : 1) : ti;
  size_t smem = bytes_for(ti);
  sp.w_smem = 0;
  if (w_bytes && sp.pitch == 32 && smem + w_bytes <= 75 * 1024) {
    sp.w_smem = 1;
    smem += w_bytes;
  }
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
    double *d_tk, *d_td, *d_ko, *d_W, *d_Wk;
    int *d_base;
    const size_t tab_bytes = sizeof(double) * CIB_TAB_ROWS * taps;
    if ((rc = tmp.alloc((void **)&d_tk, tab_bytes))) return rc;
    if ((rc = tmp.alloc((void **)&d_td, tab_bytes))) return rc;
    if ((rc = tmp.alloc((void **)&d_ko, sizeof(double) * (size_t)nout))) return rc;
    if ((rc = tmp.alloc((void **)&d_W, sizeof(double) * (size_t)nout * taps))) return rc;
    if ((rc = tmp.alloc((void **)&d_Wk, sizeof(double) * (size_t)nout * taps))) return rc;
    if ((rc = tmp.alloc((void **)&d_base, sizeof(int) * (size_t)nout))) return rc;
    CIB_CUDA(cudaMemcpyAsync(d_tk, host_table(kernel[d], deriv[d]), tab_bytes, cudaMemcpyHostToDevice, st));
    CIB_CUDA(cudaMemcpyAsync(d_td, host_table(kernel[d], deriv[d] + 1), tab_bytes, cudaMemcpyHostToDevice, st));
    CIB_CUDA(cudaMemcpyAsync(d_ko, knots_out[d], sizeof(double) * (size_t)nout, cudaMemcpyHostToDevice, st));
    resample_weights_kernel<<<(unsigned)((nout + 127) / 128), 128, 0, st>>>(d_ko, nout, x_expanded_1, h, eqs, n_coef,
                                                                           d_tk, d_td, d_W, d_Wk, d_base);
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
    sp.Wk = d_Wk;
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
