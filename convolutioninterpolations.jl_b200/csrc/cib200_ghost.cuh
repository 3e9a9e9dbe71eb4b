// cib200_ghost.cuh - per-line ghost fill shared by the sweep kernels (cib200_sweep.cu, cib200_resample.cu):
// apply_boundary_conditions_for_dim! (convolution_coefs/create_convolutional_coefs.jl:200-370) and
// fill_ghost_points_polynomial! (:408-441) on a line held in shared memory or registers.
#pragma once
#include <math.h>

#include "cib200_device.cuh"

namespace cib {

struct GhostMats {
  double poly[9][8];  // rows 0..eqs-2 used; columns 0..m-1
  double lin[9][2];
  double quad[9][3];
  int poly_nodes;
};

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

// sum of a line in the association of line_mean (cib200_sweep.cu): 8 interleaved partial sums, pairwise combine
template <int ROW>
__device__ __forceinline__ double line_sum_rows(const double *__restrict__ x0, int n) {
  double s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = 0.0;
  int j = 0;
  for (; j + 8 <= n; j += 8) {
    double v[8];
    if (ROW == 1) {  // axis 0: 16-byte loads (x0 is 16-byte aligned: E and the pitch are even)
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        const double2 t = *(const double2 *)(x0 + j + i);
        v[i] = t.x;
        v[i + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x0[(j + i) * ROW];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = __dadd_rn(s[i], v[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (j + i < n) s[i] = __dadd_rn(s[i], x0[(j + i) * ROW]);
  const double sum = __dadd_rn(__dadd_rn(__dadd_rn(s[0], s[1]), __dadd_rn(s[2], s[3])),
                               __dadd_rn(__dadd_rn(s[4], s[5]), __dadd_rn(s[6], s[7])));
  return sum;
}

// NS = polynomial node count of the kernel (6 for b5, 8 otherwise), EQS-1 ghosts per side
// core on registers: y[0..7] = centred boundary values (y[0] = the boundary point itself, then inward; 0 past the line),
// gh[j-1] = ghost j = 1..EQS-1 counted outward
template <int NS, int EQS>
__device__ __forceinline__ void ghost_poly_core(const double (&y)[8], int n, double mean, double h, int bc, bool right,
                                                const GhostMats &gm, double (&gh)[EQS > 1 ? EQS - 1 : 1]) {
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
  if (use_poly) {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < NS; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.poly[j - 1][k], y[k]));
      gh[j - 1] = __dadd_rn(mean, acc);
    }
  } else if (bc == CIB200_BC_QUADRATIC) {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.quad[j - 1][k], y[k]));
      gh[j - 1] = __dadd_rn(mean, acc);
    }
  } else {
#pragma unroll
    for (int j = 1; j <= EQS - 1; ++j) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 2; ++k) acc = __dadd_rn(acc, __dmul_rn(gm.lin[j - 1][k], y[k]));
      gh[j - 1] = __dadd_rn(mean, acc);
    }
  }
}

// the same on a line held in shared memory: bl = buf + ti, line(j) = bl[(pad + j) * pitch]
template <int NS, int EQS>
__device__ __forceinline__ void ghost_poly_side(double *bl, int pad, int pitch, int n, double mean, double h, int bc,
                                                bool right, const GhostMats &gm) {
  double y[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int j = right ? (n - 1 - k) : k;
    y[k] = k < n ? __dsub_rn(bl[(pad + j) * pitch], mean) : 0.0;
  }
  double gh[EQS > 1 ? EQS - 1 : 1];
  ghost_poly_core<NS, EQS>(y, n, mean, h, bc, right, gm, gh);
  const int row0 = right ? (pad + n) : (pad - 1);  // slot of ghost j = 1
  const int step = right ? pitch : -pitch;
  double *g = bl + row0 * pitch;
#pragma unroll
  for (int j = 1; j <= EQS - 1; ++j) g[(j - 1) * step] = gh[j - 1];
}


// host: integer Lagrange extrapolation rows (polynomial_boundary.jl:29-179 are these numbers)
void ghost_mats_for(int kernel, int eqs, GhostMats *gm);

}  // namespace cib
