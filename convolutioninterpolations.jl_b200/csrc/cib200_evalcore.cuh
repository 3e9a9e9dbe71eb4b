// cib200_evalcore.cuh - device code shared by the direct (cib200_eval.cu) and binned (cib200_binned.cu)
// point-evaluation strategies: query classification, per-axis positions, the scalar and the
// warp-cooperative stencil contraction.
#pragma once
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
      ok &= locate_axis(p, d, x[d], pa[d]);
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
      ok &= locate_axis(p, d, cl[d], pa[d]);
    pl.ntask = 1;
    if (p.extrap == CIB200_LINE && !(dist_sq < 2.220446049250313e-16)) {
      double dist = __dsqrt_rn(dist_sq);
#pragma unroll
      for (int d = 0; d < NDIM; ++d) {
        double xin = __dsub_rn(cl[d], __ddiv_rn(__dmul_rn(p.eps_line, dx[d]), dist));
        ok &= locate_axis(p, d, xin, pb[d]);
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
// per-axis Hermite basis / blended weight honouring the axis' subgrid (quintic exists for NDIM <= 2 only)
template <int NDIM>
__device__ __forceinline__ Hermite axis_basis(const ItpParams &p, int d, double t) {
  if (p.subgrid[d] == CIB200_SG_LINEAR) return linear_basis(t);
  if (NDIM <= 2 && p.subgrid[d] == CIB200_SG_QUINTIC) return hermite_basis_quintic(t);
  return hermite_basis(t);
}
template <int NDIM>
__device__ __forceinline__ double axis_weight(const ItpParams &p, int d, const Hermite &hb, const double *tk,
                                              const double *td, const double *te, int row, int col, int w) {
  if (NDIM <= 2 && p.subgrid[d] == CIB200_SG_QUINTIC) return blended_weight_quintic(hb, tk, td, te, row, col, w);
  return blended_weight(hb, tk, td, row, col, w);  // linear axes: td aliases tk, its terms carry zero factors
}


template <typename T, int NDIM, int E>
__device__ __forceinline__ double core_thread(const ItpParams &p, const double *const *tk, const double *const *td,
                                              const AxisPos *pos, const double *const *te = nullptr) {
  const T *__restrict__ coefs = (const T *)p.coefs;
  const double *tq[3] = {te ? te[0] : nullptr, te ? te[1] : nullptr, te ? te[2] : nullptr};
  if (E > 0) {
    constexpr int W = 2 * (E > 0 ? E : 1);
    double w[NDIM][W];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      Hermite hb = axis_basis<NDIM>(p, d, pos[d].t);
#pragma unroll
      for (int l = 0; l < W; ++l) w[d][l] = axis_weight<NDIM>(p, d, hb, tk[d], td[d], tq[d], pos[d].row, l, W);
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
  for (int d = 0; d < NDIM; ++d) hb[d] = axis_basis<NDIM>(p, d, pos[d].t);
  const int w0 = 2 * p.eqs[0];
  if (NDIM == 1) {
    double acc = 0.0;
    for (int l = 0; l < w0; ++l)
      acc = fma((double)coefs[pos[0].base + l], axis_weight<NDIM>(p, 0, hb[0], tk[0], td[0], tq[0], pos[0].row, l, w0), acc);
    return acc;
  }
  const int w1 = 2 * p.eqs[NDIM > 1 ? 1 : 0];
  if (NDIM == 2) {
    double acc = 0.0;
    const T *c0 = coefs + pos[0].base + (long long)pos[NDIM > 1 ? 1 : 0].base * p.s1;
    for (int m = 0; m < w1; ++m) {
      double a = 0.0;
      for (int l = 0; l < w0; ++l)
        a = fma((double)c0[l + (long long)m * p.s1], axis_weight<NDIM>(p, 0, hb[0], tk[0], td[0], tq[0], pos[0].row, l, w0), a);
      acc = fma(a, axis_weight<NDIM>(p, 1, hb[NDIM > 1 ? 1 : 0], tk[1], td[1], tq[1], pos[NDIM > 1 ? 1 : 0].row, m, w1), acc);
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


// ---------------------------------------------------------------------------------------------
// warp-cooperative stencil of ONE evaluation point (3-D, one kernel on all axes)
//   lanes = (tap l along axis 0) x (G stencil rows); one load instruction fetches G complete rows.
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

template <int E>
struct WarpGeom {
  static constexpr int W = 2 * E;        // taps per axis
  static constexpr int G = 32 / W;       // stencil rows fetched per load instruction
  static constexpr int ACTIVE = W * G;   // lanes that carry a tap
  static constexpr int ROWS = W * W;     // (m, n) rows of the brick
  static constexpr int ITERS = (ROWS + G - 1) / G;
  static constexpr int CHUNK = ITERS < 36 ? ITERS : 32;
  static constexpr int SMEM_DOUBLES = ROWS + 3 * W;  // per warp: wyz[ROWS] + wax[3][W]
};

// fills wax[3][W] (Hermite-blended per-axis weights) and wyz[m + W n] = wy[m] wz[n]; returns wx of this lane
template <int E>
__device__ __forceinline__ double warp_weights(double *wyz, double *wax, const double *const *tk, const double *const *td,
                                               const int *row, const double *tt, int lane, bool linear = false) {
  using GM = WarpGeom<E>;
  constexpr int W = GM::W;
  for (int idx = lane; idx < 3 * W; idx += 32) {
    const int a = idx / W, col = idx - a * W;
    const double ta = a == 0 ? tt[0] : (a == 1 ? tt[1] : tt[2]);
    const int ra = a == 0 ? row[0] : (a == 1 ? row[1] : row[2]);
    Hermite hb = linear ? linear_basis(ta) : hermite_basis(ta);
    wax[idx] = blended_weight(hb, a == 0 ? tk[0] : (a == 1 ? tk[1] : tk[2]), a == 0 ? td[0] : (a == 1 ? td[1] : td[2]),
                              ra, col, W);
  }
  __syncwarp();
  for (int r = lane; r < GM::ROWS; r += 32) wyz[r] = wax[W + r % W] * wax[2 * W + r / W];
  const double wx = lane < GM::ACTIVE ? wax[lane % W] : 0.0;
  __syncwarp();
  return wx;
}

// sum over the (2E)^3 brick; ptr already points at this lane's (l, g) element of the brick's first plane.
// LOAD: functor pointer -> double.  Strides in elements.  Returns the warp-wide total in every lane.
template <int E, typename P, typename S, typename LOAD>
__device__ __forceinline__ double warp_stencil(P ptr, S s1, S s2, const double *wyz, double wx, int lane, LOAD load) {
  using GM = WarpGeom<E>;
  constexpr int W = GM::W, G = GM::G;
  const int g = lane / W;
  const bool active = lane < GM::ACTIVE;
  const S wrap_adj = s2 - (S)W * s1;
  double acc = 0.0;
#pragma unroll 1
  for (int it0 = 0; it0 < GM::ITERS; it0 += GM::CHUNK) {
    double cv[GM::CHUNK];
#pragma unroll
    for (int u = 0; u < GM::CHUNK; ++u) {
      const int r0 = (it0 + u) * G;  // row of lane group 0 in this step
      const int m0 = r0 % W, n0 = r0 / W;
      const bool wrap = m0 + g >= W;
      const bool valid = active && (it0 + u) < GM::ITERS && (r0 + g) < GM::ROWS;
      const S off = (S)m0 * s1 + (S)n0 * s2 + (wrap ? wrap_adj : (S)0);
      cv[u] = valid ? load(ptr + off) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < GM::CHUNK; ++u) {
      const int r = (it0 + u) * G + g;
      const double wv = (active && (it0 + u) < GM::ITERS && r < GM::ROWS) ? wyz[r] : 0.0;
      acc = fma(cv[u], wv, acc);
    }
  }
  return warp_sum(acc * wx);
}

}  // namespace cib
