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
// deviation from the literal 64-accumulator form is a few ulp of sum|c w| (tests/test_gpu_eval.py).
#include "cib200_evalcore.cuh"

namespace cib {

template <typename T, int NDIM, int E>
__global__ void __launch_bounds__(256) eval_thread_kernel(ItpParams p, Coords c, Scale sc, T *__restrict__ out,
                                                          long long nq, EvalStatus *status) {
  extern __shared__ double smem[];
  const double *tk[3];
  const double *td[3];
  stage_tables(p, smem, tk, td);
  const double *te[3] = {nullptr, nullptr, nullptr};
  if (NDIM <= 2)
    for (int d = 0; d < NDIM; ++d) te[d] = staged_e_table(p, smem, d);
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
      double fb = apply_scale<NDIM>(core_thread<T, NDIM, E>(p, tk, td, pa, te), sc);
      r = fb;
      if (pl.ntask == 2) {
        double fin = apply_scale<NDIM>(core_thread<T, NDIM, E>(p, tk, td, pb, te), sc);
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
template <typename T, int E>
__global__ void __launch_bounds__(256, 2) eval_warp3d_kernel(ItpParams p, Coords c, Scale sc, T *__restrict__ out,
                                                             long long nq, EvalStatus *status) {
  using GM = WarpGeom<E>;
  constexpr int W = GM::W;
  extern __shared__ double smem[];
  const double *tk[3];
  const double *td[3];
  stage_tables(p, smem, tk, td);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *wyz = smem + tables_smem_doubles(p) + warp * GM::SMEM_DOUBLES;
  double *wax = wyz + GM::ROWS;  // [3][W] blended weights of the current evaluation point
  __syncthreads();

  const int l = lane % W, g = lane / W;
  const T *__restrict__ coefs = (const T *)p.coefs;
  const long long s1 = p.s1, s2 = p.s2;
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
    // warp-uniform walk over the 32 lanes (a ballot-driven loop after the divergent planning phase produced
    // lane-inconsistent trip counts in the binned kernel on sm_100a; the shuffle-broadcast count cannot)
    for (int j = 0; j < 32; ++j) {
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
        const double wx = warp_weights<E>(wyz, wax, tk, td, row, tt, lane, p.subgrid[0] == CIB200_SG_LINEAR);
        const T *ptr = coefs + base[0] + l + ((long long)base[1] + g) * s1 + (long long)base[2] * s2;
        const double total = warp_stencil<E>(ptr, s1, s2, wyz, wx, lane, [](const T *a) { return ldg_as_double(a); });
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
double jl_pow_int(double x, long long n) {  // Julia Base pow_body (math.jl), restated
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

int sm_count() {
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
  size_t smem = sizeof(double) * ((size_t)tables_smem_doubles(p) + 8 * WarpGeom<E>::SMEM_DOUBLES);
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
    case 2: return launch_warp3d<T, 2>(p, c, sc, out, nq, h->d_status, st);
    case 3: return launch_warp3d<T, 3>(p, c, sc, out, nq, h->d_status, st);
    case 4: return launch_warp3d<T, 4>(p, c, sc, out, nq, h->d_status, st);
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
