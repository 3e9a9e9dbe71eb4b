// cib200_device.cuh - device helpers shared by the point-evaluation and sweep kernels.
#pragma once
#include "cib200_internal.cuh"

namespace cib {

// Position of x on one axis.  The index arithmetic restates the reference exactly, operation by
// operation and WITHOUT fma contraction (__d*_rn intrinsics are never contracted), so that cell
// and table-row indices are the reference's own for every finite input
// (convolution_fast_interpolation_eager_3d.jl:115-137; same lines in eager_1d/2d, perdim_2d.jl:17-49).
struct AxisPos {
  int base;  // 0-based index of the first stencil coefficient: (i - eqs) with Julia's 1-based i
  int row;   // 0-based lower table row: idx - 1
  double t;  // fractional position between table rows
};

__device__ __forceinline__ bool locate_from_knot(double x, double knot0, double h, int eqs, int nk,
                                                 const double *__restrict__ knots, AxisPos &p, int npre = CIB_N_PRE) {
  double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(x, knot0), h), 1.0);
  bool ok = fabs(i_float) < 9.0e18;  // floor(Int, .) throws InexactError otherwise (NaN / Inf)
  long long i = ok ? (long long)floor(i_float) : (long long)eqs;
  i = i < eqs ? eqs : i;
  i = i > nk - eqs ? nk - eqs : i;
  double knot_i = knots[i - 1];  // the STORED knot, not knot0 + (i-1) h
  double x_diff_right = __dsub_rn(1.0, __ddiv_rn(__dsub_rn(x, knot_i), h));
  double cidx = __dadd_rn(__dmul_rn(x_diff_right, (double)(npre - 1)), 1.0);
  ok = ok && (fabs(cidx) < 9.0e18);
  int idx = ok ? (int)floor(cidx) : 1;
  idx = idx < 1 ? 1 : idx;
  idx = idx > npre - 1 ? npre - 1 : idx;
  p.base = (int)(i - eqs);
  p.row = idx - 1;
  p.t = __dsub_rn(cidx, (double)idx);
  return ok;
}

// (linear, linear) on the single-kernel 2-D path (eager_2d.jl:93-107): the table row comes from
// floor(diff_right (n_pre - 1)) + 1 and t from the stored pre_range values, pre_range[r] = Float64((r-1)//(n_pre-1))
__device__ __forceinline__ bool locate_prerange(double x, double knot0, double h, int eqs, int nk,
                                                const double *__restrict__ knots, AxisPos &p, int npre = CIB_N_PRE) {
  double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(x, knot0), h), 1.0);
  bool ok = fabs(i_float) < 9.0e18;
  long long i = ok ? (long long)floor(i_float) : (long long)eqs;
  i = i < eqs ? eqs : i;
  i = i > nk - eqs ? nk - eqs : i;
  const double knot_i = knots[i - 1];
  const double diff_right = __dsub_rn(1.0, __ddiv_rn(__dsub_rn(x, knot_i), h));
  const double f = __dmul_rn(diff_right, (double)(npre - 1));
  ok = ok && (fabs(f) < 9.0e18);
  int idx = ok ? (int)floor(f) + 1 : 1;
  idx = idx < 1 ? 1 : idx;
  idx = idx > npre - 1 ? npre - 1 : idx;
  const double r0 = __ddiv_rn((double)(idx - 1), (double)(npre - 1)), r1 = __ddiv_rn((double)idx, (double)(npre - 1));
  p.base = (int)(i - eqs);
  p.row = idx - 1;
  p.t = __ddiv_rn(__dsub_rn(diff_right, r0), __dsub_rn(r1, r0));
  return ok;
}

// position of x on axis d in the variant the interpolant's evaluator uses
__device__ __forceinline__ bool locate_axis(const ItpParams &p, int d, double x, AxisPos &pos) {
  if (p.prerange_pos) return locate_prerange(x, p.knot0[d], p.h[d], p.eqs[d], p.nk[d], p.knots[d], pos, p.npre[d]);
  return locate_from_knot(x, p.knot0[d], p.h[d], p.eqs[d], p.nk[d], p.knots[d], pos, p.npre[d]);
}

// resample variant: the cell's left knot is COMPUTED (convolution_resample.jl:322-326)
__device__ __forceinline__ void locate_computed(double x, double x_expanded_1, double h, int eqs,
                                                long long n_coef, AxisPos &p, int npre = CIB_N_PRE) {
  double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(x, x_expanded_1), h), 1.0);
  bool ok = fabs(i_float) < 9.0e18;
  long long i = ok ? (long long)floor(i_float) : (long long)eqs;
  i = i < eqs ? eqs : i;
  i = i > n_coef - eqs ? n_coef - eqs : i;
  double knot_i = __dadd_rn(x_expanded_1, __dmul_rn((double)(i - 1), h));
  double x_diff_left = __ddiv_rn(__dsub_rn(x, knot_i), h);
  double x_diff_right = __dsub_rn(1.0, x_diff_left);
  double cidx = __dadd_rn(__dmul_rn(x_diff_right, (double)(npre - 1)), 1.0);
  int idx = (fabs(cidx) < 9.0e18) ? (int)floor(cidx) : 1;
  idx = idx < 1 ? 1 : idx;
  idx = idx > npre - 1 ? npre - 1 : idx;
  p.base = (int)(i - eqs);
  p.row = idx - 1;
  p.t = __dsub_rn(cidx, (double)idx);
}

// cubic Hermite basis (cubic_hermite_interpolation.jl:15-28), h_pre folded into the derivative terms
struct Hermite {
  double h00, h01, h10, h11;
  double h20, h21;  // quintic only: second-derivative basis, (h_pre^2 / 2) folded in
};
__device__ __forceinline__ Hermite hermite_basis(double t) {
  const double h_pre = 1.0 / (double)(CIB_N_PRE - 1);
  double t2 = t * t;
  Hermite b;
  b.h00 = 1.0 + t2 * (-3.0 + 2.0 * t);
  b.h01 = t2 * (3.0 - 2.0 * t);
  b.h10 = h_pre * (t + t2 * (-2.0 + t));
  b.h11 = h_pre * (t2 * (t - 1.0));
  b.h20 = 0.0;
  b.h21 = 0.0;
  return b;
}

// linear subgrid: (1 - t) K[row] + t K[row+1] (eager_1d.jl:88, eager_3d.jl:167-171, perdim_*.jl) expressed with the same
// four-term weight formula: the derivative-table terms get zero factors (the kernels then read K instead of K')
__device__ __forceinline__ Hermite linear_basis(double t) {
  Hermite b;
  b.h00 = 1.0 - t;
  b.h01 = t;
  b.h10 = b.h11 = b.h20 = b.h21 = 0.0;
  return b;
}

// quintic Hermite basis (quintic_hermite_interpolation.jl:16-37), h_pre and h_pre^2/2 folded into the derivative terms
__device__ __forceinline__ Hermite hermite_basis_quintic(double t) {
  const double h_pre = 1.0 / (double)(CIB_N_PRE - 1);
  const double t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
  Hermite b;
  b.h00 = 1.0 - 10 * t3 + 15 * t4 - 6 * t5;
  b.h01 = 10 * t3 - 15 * t4 + 6 * t5;
  b.h10 = h_pre * (t - 6 * t3 + 8 * t4 - 3 * t5);
  b.h11 = h_pre * (-4 * t3 + 7 * t4 - 3 * t5);
  b.h20 = (h_pre * h_pre / 2) * (t2 - 3 * t3 + 3 * t4 - t5);
  b.h21 = (h_pre * h_pre / 2) * (t3 - 2 * t4 + t5);
  return b;
}

// One stencil weight.  By linearity of cubic_hermite in (f0, f1, d0, d1) the reference's
// "4 / 16 / 64 convolved sums, then nested Hermite" equals "Hermite-blend the table rows per axis,
// then one tensor contraction" (SURVEY.md 3.2); this is that per-axis blended weight.
__device__ __forceinline__ double blended_weight(const Hermite &b, const double *__restrict__ tk,
                                                 const double *__restrict__ td, int row, int col, int w) {
  const double *k0 = tk + row * w + col;
  const double *d0 = td + row * w + col;
  return b.h00 * k0[0] + b.h01 * k0[w] + (b.h10 * d0[0] + b.h11 * d0[w]);
}

// quintic subgrid: the same plus the second-derivative rows (te = K^(d+2))
__device__ __forceinline__ double blended_weight_quintic(const Hermite &b, const double *__restrict__ tk,
                                                         const double *__restrict__ td, const double *__restrict__ te,
                                                         int row, int col, int w) {
  const double *e0 = te + row * w + col;
  return blended_weight(b, tk, td, row, col, w) + (b.h20 * e0[0] + b.h21 * e0[w]);
}

// copy the distinct 101-row tables of an interpolant into shared memory and fill the per-axis pointers; the generated
// tables of :linear axes (>= 10 000 rows, constructor.jl:117-118) are too large to stage and are read from global
// memory (they are L2-resident: 0.8 - 1.4 MB)
__device__ __forceinline__ void stage_tables(const ItpParams &p, double *smem, const double *tk[3],
                                             const double *td[3]) {
  const double *at[9];
  int o = 0;
  for (int s = 0; s < p.n_tab_slots; ++s) {
    const double *src = p.tab_slot_ptr[s];
    if (p.tab_slot_rows[s] != CIB_TAB_ROWS) {
      at[s] = src;
      continue;
    }
    at[s] = smem + o;
    int nelem = CIB_TAB_ROWS * 2 * p.tab_slot_eqs[s];
    for (int i = threadIdx.x; i < nelem; i += blockDim.x) smem[o + i] = src[i];
    o += nelem;
  }
  for (int d = 0; d < p.ndim; ++d) {
    tk[d] = at[p.tab_slot_k[d]];
    td[d] = at[p.tab_slot_d[d]];
  }
}

// shared-memory address of the staged K^(d+2) table of axis d (quintic subgrid), nullptr for a cubic axis
__device__ __forceinline__ const double *staged_e_table(const ItpParams &p, const double *smem, int d) {
  if (p.tab_slot_e[d] < 0) return nullptr;
  int o = 0;
  for (int s = 0; s < p.tab_slot_e[d]; ++s)
    if (p.tab_slot_rows[s] == CIB_TAB_ROWS) o += CIB_TAB_ROWS * 2 * p.tab_slot_eqs[s];
  return smem + o;
}

__host__ __device__ inline int tables_smem_doubles(const ItpParams &p) {
  int o = 0;
  for (int s = 0; s < p.n_tab_slots; ++s)
    if (p.tab_slot_rows[s] == CIB_TAB_ROWS) o += CIB_TAB_ROWS * 2 * p.tab_slot_eqs[s];
  return o;
}

template <typename T>
__device__ __forceinline__ double load_coord(const Coords &c, int d, long long q, long long qd) {
  const T *x = (const T *)c.x[d];
  return (double)x[c.grid ? qd : q];
}

__device__ __forceinline__ void atomic_min_u64(unsigned long long *addr, unsigned long long v) {
  atomicMin(addr, v);
}

}  // namespace cib
