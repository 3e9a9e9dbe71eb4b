// cib200_smooth.cu - convolution_smooth, Float64, as a chain of persistent TMA-fed FIR passes.
//
// Replaces the per-line loop of convolution_smooth (convolution/convolution_gaussian_smoothing.jl:70-124) and the
// ghost fill it calls (fill_ghost_points_gaussian!, convolution_coefs/create_convolutional_coefs.jl:443-494) for the
// shapes BASELINE config 4 names; every other case stays on sweep_pipe_kernel / sweep_kernel (cib200_sweep.cu).
//
// One pass = one axis; the array is viewed as [inner][n][outer].  What bounds a pass is the FP64 pipe (2E+1 DFMAs per
// point: E = 24 for B = 0.05), so the design removes every FP64 instruction and every stall that is not one of those:
//
//  * zero padding instead of ghost rows.  The reference extends a line by E ghosts per side (an affine extrapolation
//    a + b (1 - k) of the mean-centred line, least-squares fit over its first / last points) and convolves the
//    extended line.  A ghost's contribution to output i is linear in (a, b):  out[i] += a S0[i] + b S1[i] with
//    S0[i] = sum_k kw[.], S1[i] = sum_k kw[.] (1 - k) over the ghosts output i can see - two host-computed tables of
//    E - 1 entries (the reference's slot quirk, SURVEY.md 3.5, is in the tables: ghost k sits k + 1 slots outside, the
//    slot next to the line stays 0).  The FIR then runs over the line padded with ZEROS, which the TMA unit produces for
//    free (box coordinates start at -E, out-of-bounds elements are zero-filled): no ring clearing, no ghost stores,
//    and every window row that lies outside the line is skipped as a whole (19 % of the DFMAs at n = 64, 31 % at 40).
//  * a "stats" stage (lane = line) computes the line mean and the two fits: n + 64 FP64 instructions per line instead
//    of ~540 for mean (twice) + fit + 46 ghost evaluations.
//  * consumers pull (tile, block of 8 outputs) items from a shared counter, so clipped boundary blocks and interior
//    blocks balance by themselves; taps are literal offsets into the kernel-parameter bank (DFMA R, R, c[0][imm], R).
//  * axis >= 1: one 3-D TMA box {BI, n + 2E + 7, 32 / BI} per tile, lane = (inner index, outer index); BI = 8 keeps
//    all 32 lanes busy when `inner` is 40.  axis 0: one 2-D box of 32 lines, line-major with pitch P = 2 (mod 4); a
//    lane = line reads its window two elements at a time (LDS.128: 8 lanes x 16 bytes per wavefront, and P / 2 odd
//    puts the 8 lanes into 8 distinct 16-byte banks: conflict-free without padding tricks; a TMA box can only start
//    on a 16-byte boundary, so odd column shifts are not available).  Results go back through a per-warp transposing
//    stage so that global stores are full 32-byte sectors.
//  * passes are chained with programmatic dependent launch: the next pass's prologue overlaps this pass's tail.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "cib200_ghost.cuh"
#include "cib200_pipe.cuh"

namespace cib {
using namespace pipe;

struct SmoothPipeParams {
  long long inner, outer, ntiles;
  long long line_o_stride;  // axis >= 1: inner * n (elements between consecutive `outer` slabs)
  int n, nblk;
  int H;            // axis >= 1: rows of a tile buffer; axis 0: pitch P of a line
  int tiles_inner;  // axis >= 1: tiles along `inner`
  int nb, nstats, ncons;
  unsigned magic_nblk, magic_nb, magic_ti;  // floor(2^32 / d) + 1: exact division of the small counters below by d
  double nf, sx, inv_denom, inv_nf, inv_n;
  double kw[64];          // normalised Gaussian taps, l = -E..E
  double s0[32], s1[32];  // ghost tables, index = distance of the output from the boundary (0..E-2)
};

// per-tile header written by the stats warp: where the tile's outputs go (consumers stay in 32-bit arithmetic)
struct TileHdr {
  long long obase;  // element offset of (first line of the tile, output 0)
  int vi, vo;       // valid lines along inner / outer (axis 0: vo = valid lines)
};

// R consecutive outputs of one line from a window of TAPS + R - 1 input rows; chunks of 8 rows outside [clo, chi) lie in
// the zero padding and are skipped (warp-uniform); taps are literal offsets into the kernel-parameter bank
template <int TAPS, int ROW, int R>
__device__ __forceinline__ void fir_block(const double *__restrict__ p, int clo, int chi, const SmoothPipeParams &sp,
                                          double (&acc)[R]) {
  constexpr int W = TAPS + R - 1, NCH = (W + 7) / 8;
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    if (c >= clo && c < chi) {
      double x[8];
      if (ROW == 1) {  // axis 0: the window is contiguous and 16-byte aligned (W is even)
#pragma unroll
        for (int i = 0; i < 8; i += 2)
          if (8 * c + i < W) {
            const double2 v = *(const double2 *)(p + 8 * c + i);
            x[i] = v.x;
            x[i + 1] = v.y;
          }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * c + i < W) x[i] = p[(8 * c + i) * ROW];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = 8 * c + i;
        if (j < W) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int l = j - r;
            if (l >= 0 && l < TAPS) acc[r] = fma(x[i], sp.kw[l], acc[r]);
          }
        }
      }
    }
  }
}

// least-squares lines through the first NFIT mean-centred points counted from either end (fill_ghost_points_gaussian!,
// create_convolutional_coefs.jl:455-478): x_k = k = 1..NFIT; returns (a, b) = centred intercept and slope, left and
// right.  Both fits are unrolled together (four independent accumulation chains), and the three divisions of the
// reference (mean = sum / n, b = ... / denom, a = ... / n_fit) are multiplications by host-computed reciprocals: one
// more rounding each (<= 1e-16 relative), off the tile's critical path by ~400 cycles of dependent DDIV sequences.
template <int ROW, int NFIT>
__device__ __forceinline__ void fit_both(const double *__restrict__ x0, int n, double mean, const SmoothPipeParams &sp,
                                         double2 &L, double2 &Rr) {
  double yl[NFIT], yr[NFIT];
#pragma unroll
  for (int k = 0; k < NFIT; ++k) {
    yl[k] = __dsub_rn(x0[k * ROW], mean);
    yr[k] = __dsub_rn(x0[(n - 1 - k) * ROW], mean);
  }
  double syl = 0.0, sxyl = 0.0, syr = 0.0, sxyr = 0.0;
#pragma unroll
  for (int k = 0; k < NFIT; ++k) {
    syl = __dadd_rn(syl, yl[k]);
    sxyl = __dadd_rn(sxyl, __dmul_rn((double)(k + 1), yl[k]));
    syr = __dadd_rn(syr, yr[k]);
    sxyr = __dadd_rn(sxyr, __dmul_rn((double)(k + 1), yr[k]));
  }
  const double bl = __dmul_rn(__dsub_rn(__dmul_rn(sp.nf, sxyl), __dmul_rn(sp.sx, syl)), sp.inv_denom);
  const double br = __dmul_rn(__dsub_rn(__dmul_rn(sp.nf, sxyr), __dmul_rn(sp.sx, syr)), sp.inv_denom);
  L = make_double2(__dmul_rn(__dsub_rn(syl, __dmul_rn(bl, sp.sx)), sp.inv_nf), bl);
  Rr = make_double2(__dmul_rn(__dsub_rn(syr, __dmul_rn(br, sp.sx)), sp.inv_nf), br);
}

static constexpr int SM_MMA_PITCH = 36;  // doubles between two rows of a tile for the tensor-core consumer (axis >= 1)

__device__ __forceinline__ void sm_mma_f64_884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// BI: lanes along `inner` in a tile of an axis >= 1 pass (32 or 8); BI = 0: axis 0.  R: outputs per work item.
// MMA: the FIR of a block runs on the FP64 tensor cores (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): rows = the block's 8
// outputs, K = the TAPS + 7 window rows, columns = 8 lines.  The A operand is the Toeplitz band of the taps - the same for
// every block, so a lane keeps its (TAPS + 10) / 4 fragments in registers for the whole kernel - and the B fragments are
// one conflict-free LDS.64 per DMMA (tile rows at a pitch of 36 doubles, axis 0: line pitch = 4 mod 16).
template <int TAPS, int BI, int R, bool MMA>
__global__ void __launch_bounds__(640, 1)
    smooth_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ SmoothPipeParams sp,
                       double *__restrict__ out) {
  constexpr int E = (TAPS - 1) / 2;
  constexpr bool AXIS0 = BI == 0;
  constexpr int BIX = BI > 0 ? BI : 32;
  constexpr int ROW = AXIS0 ? 1 : (MMA ? SM_MMA_PITCH : BI);  // doubles between two rows of a tile
  static_assert(!MMA || (R == 8 && BI != 8), "tensor-core consumer: blocks of 8 outputs, 32 lines along inner or axis 0");
  constexpr int BO = AXIS0 ? 1 : 32 / BIX;
  constexpr int W = TAPS + R - 1;
  constexpr int SPITCH = R + 1;  // transposing stage of axis 0
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NB = sp.nb, n = sp.n, nblk = sp.nblk;
  const int TILE = ((MMA && !AXIS0) ? SM_MMA_PITCH : 32) * sp.H;  // doubles
  double *tiles = (double *)smem_raw;
  double *lpar = tiles + (size_t)NB * TILE;  // [NB][32][4]: aL bL aR bR of every line of the tile
  double *stab = lpar + NB * 128;            // s0[32] s1[32]
  TileHdr *hdr = (TileHdr *)(stab + 64);     // [NB]
  double *stage = (double *)(hdr + NB);      // axis 0: [ncons][32 * SPITCH] transposing store stage
  unsigned long long *bars = (unsigned long long *)(stage + (AXIS0 ? sp.ncons * 32 * SPITCH : 0));
  unsigned long long *loaded = bars, *ready = bars + NB, *empty = bars + 2 * NB;
  unsigned *counter = (unsigned *)(bars + 3 * NB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 64) stab[tid] = tid < 32 ? sp.s0[tid] : sp.s1[tid - 32];
  if (tid == 0) {
    for (int b = 0; b < NB; ++b) {
      mbar_init(loaded + b, 1);
      mbar_init(ready + b, 1);
      mbar_init(empty + b, (unsigned)nblk);
    }
    *counter = 0;
    mbar_init_fence();
  }
  __syncthreads();
  pdl_launch_dependents();
  const unsigned tstep = gridDim.x;
  const unsigned ntiles = (unsigned)sp.ntiles;
  const unsigned my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + tstep - 1) / tstep : 0u;

  // this lane's line: offset of its row -E (axis 0: column -E) in a tile buffer, and of its outputs from the tile's
  const int li = AXIS0 ? 0 : lane % BIX, lo = AXIS0 ? lane : lane / BIX;
  const int lbase = AXIS0 ? lane * sp.H : lo * sp.H * ROW + li;

  if (warp == 0) {
    // ------------------------------- producer: one elected lane feeds the ring with TMA boxes -----------------
    if (lane == 0) {
      pdl_wait();  // the input of this pass is the output of the previous launch
      unsigned b = 0, rev = 0;
      for (unsigned seq = 0; seq < my_tiles; ++seq) {
        const unsigned t = blockIdx.x + seq * tstep;
        if (rev > 0) mbar_wait(empty + b, (rev - 1) & 1u, 1000 + (int)seq);
        fence_async_proxy();
        mbar_expect_tx(loaded + b, (unsigned)(TILE * 8));
        double *dst = tiles + (size_t)b * TILE;
        if (AXIS0) {
          tma_load_2d(dst, &tmap, -E, (int)(t * 32), loaded + b);
        } else {
          const unsigned tile_o = div_u(t, sp.magic_ti), tile_i = t - tile_o * sp.tiles_inner;
          tma_load_3d(dst, &tmap, (int)tile_i * BIX, -E, (int)tile_o * BO, loaded + b);
        }
        if (++b == (unsigned)NB) {
          b = 0;
          ++rev;
        }
      }
    }
  } else if (warp <= sp.nstats) {
    // ------------------------------- stats: line mean + the two boundary fits (lane = line) -------------------
    for (unsigned seq = warp - 1; seq < my_tiles; seq += sp.nstats) {
      const unsigned rev = div_u(seq, sp.magic_nb), b = seq - rev * NB;
      const unsigned t = blockIdx.x + seq * tstep;
      mbar_wait(loaded + b, rev & 1u, 2000 + (int)seq);  // also: the slot's previous tile has been consumed
      if (lane == 0) {
        TileHdr h;
        if (AXIS0) {
          h.obase = (long long)t * 32 * n;
          const long long rem = sp.outer - (long long)t * 32;
          h.vi = 1;
          h.vo = rem < 32 ? (int)rem : 32;
        } else {
          const unsigned tile_o = div_u(t, sp.magic_ti), tile_i = t - tile_o * sp.tiles_inner;
          h.obase = (long long)tile_i * BIX + sp.line_o_stride * ((long long)tile_o * BO);
          const long long ri = sp.inner - (long long)tile_i * BIX, ro = sp.outer - (long long)tile_o * BO;
          h.vi = ri < BIX ? (int)ri : BIX;
          h.vo = ro < BO ? (int)ro : BO;
        }
        hdr[b] = h;
      }
      const double *x0 = tiles + (size_t)b * TILE + lbase + E * ROW;  // element 0 of the line
      const double mean = __dmul_rn(line_sum_rows<ROW>(x0, n), sp.inv_n);
      constexpr int NFIT = E / 3 > 6 ? E / 3 : 6;  // n_fit = max(E / 3, 6) (the host checks n >= NFIT)
      double2 L, Rr;
      fit_both<ROW, NFIT>(x0, n, mean, sp, L, Rr);
      double2 *lp = (double2 *)(lpar + b * 128 + lane * 4);
      lp[0] = L;
      lp[1] = Rr;
      __syncwarp();
      if (lane == 0) mbar_arrive(ready + b);
    }
  } else if (warp - 1 - sp.nstats < sp.ncons) {
    // ------------------------------- consumers: (tile, block of R outputs) items from a shared counter ---------
    const int cw = warp - 1 - sp.nstats;
    double *mystage = stage + cw * 32 * SPITCH;
    // per-lane output offset from the tile's base (axis >= 1) / transposed store roles (axis 0)
    const long long lane_off = AXIS0 ? 0 : (long long)li + sp.line_o_stride * lo;
    const int sr = lane % R, sl = lane / R;  // axis 0: this lane stores output sr of lines sl, sl + 32 / R, ...
    // tensor-core consumer: this lane's fragments of the Toeplitz band, A[output mq][window row j] = kw[j - mq]
    double afr[MMA ? (W + 3) / 4 : 1];
    if (MMA) {
#pragma unroll
      for (int s = 0; s < (W + 3) / 4; ++s) {
        const int l = 4 * s + (lane & 3) - (lane >> 2);
        afr[s] = (l >= 0 && l < TAPS) ? sp.kw[l] : 0.0;
      }
    }
    for (;;) {
      unsigned item = 0;
      if (lane == 0) item = atomicAdd(counter, 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      const unsigned seq = div_u(item, sp.magic_nblk);
      if (seq >= my_tiles) break;
      const int k0 = (int)(item - seq * (unsigned)nblk) * R;
      const unsigned rev = div_u(seq, sp.magic_nb), b = seq - rev * NB;
      // window row j <-> input index k0 - E + j; rows outside [0, n) are zero padding
      const int jlo = E - k0 > 0 ? E - k0 : 0;
      const int jhi = n + E - k0 < W ? n + E - k0 : W;
      const int clo = jlo >> 3, chi = (jhi + 7) >> 3;
      mbar_wait(ready + b, rev & 1u, 3000 + (int)seq);
      if (MMA) {
        constexpr int KS = (W + 3) / 4;
        const int mq = lane >> 2, mr = lane & 3;
        const double *xb = tiles + (size_t)b * TILE + (AXIS0 ? mq * sp.H + k0 + mr : (k0 + mr) * ROW + mq);  // B[row 4 s + mr][line mq]
        constexpr int XS = AXIS0 ? 4 : 4 * ROW;      // stride of a K step (4 window rows)
        const int XG = AXIS0 ? 8 * sp.H : 8;         // stride of a column group (8 lines)
        const int slo = jlo >> 2, shi = (jhi + 3) >> 2;  // K steps outside [slo, shi) lie in the zero padding
        double ac[4][2];
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) ac[cg][0] = ac[cg][1] = 0.0;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          if (s >= slo && s < shi) {  // warp-uniform
            double xv[4];
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) xv[cg] = xb[s * XS + cg * XG];
#pragma unroll
            for (int cg = 0; cg < 4; ++cg) sm_mma_f64_884(ac[cg][0], ac[cg][1], afr[s], xv[cg]);
          }
        }
        const int ko = k0 + mq;  // this lane's output
        if (ko < E - 1) {  // sees left ghosts
          const double t0 = stab[ko], t1 = stab[32 + ko];
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            const double *lp = lpar + b * 128 + (8 * cg + 2 * mr) * 4;
            const double2 a0 = *(const double2 *)lp, a1 = *(const double2 *)(lp + 4);
            ac[cg][0] = fma(a0.y, t1, fma(a0.x, t0, ac[cg][0]));
            ac[cg][1] = fma(a1.y, t1, fma(a1.x, t0, ac[cg][1]));
          }
        }
        const int dr = n - 1 - ko;
        if (dr >= 0 && dr < E - 1) {  // sees right ghosts
          const double t0 = stab[dr], t1 = stab[32 + dr];
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            const double *lp = lpar + b * 128 + (8 * cg + 2 * mr) * 4 + 2;
            const double2 a0 = *(const double2 *)lp, a1 = *(const double2 *)(lp + 4);
            ac[cg][0] = fma(a0.y, t1, fma(a0.x, t0, ac[cg][0]));
            ac[cg][1] = fma(a1.y, t1, fma(a1.x, t0, ac[cg][1]));
          }
        }
        const TileHdr h = hdr[b];
        if (ko < n) {
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            const int line = 8 * cg + 2 * mr;
            if (AXIS0) {
              double *op = out + h.obase + (long long)line * n + ko;
              if (line < h.vo) op[0] = ac[cg][0];
              if (line + 1 < h.vo) op[n] = ac[cg][1];
            } else if (h.vo > 0) {
              double *op = out + h.obase + (long long)ko * sp.inner + line;
              if (line + 1 < h.vi) *(double2 *)op = make_double2(ac[cg][0], ac[cg][1]);  // inner is even: 16-byte aligned
              else if (line < h.vi) op[0] = ac[cg][0];
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + b);
        continue;
      }
      const double *p = tiles + (size_t)b * TILE + lbase + k0 * ROW;
      double acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = 0.0;
      fir_block<TAPS, ROW, R>(p, clo, chi, sp, acc);
      if (k0 < E - 1) {  // outputs that see left ghosts
        const double2 ab = *(const double2 *)(lpar + b * 128 + lane * 4);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (k0 + r < E - 1) acc[r] = fma(ab.y, stab[32 + k0 + r], fma(ab.x, stab[k0 + r], acc[r]));
      }
      if (k0 + R - 1 > n - E) {  // outputs that see right ghosts: distance n - 1 - k < E - 1
        const double2 ab = *(const double2 *)(lpar + b * 128 + lane * 4 + 2);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int d = n - 1 - (k0 + r);
          if (d >= 0 && d < E - 1) acc[r] = fma(ab.y, stab[32 + d], fma(ab.x, stab[d], acc[r]));
        }
      }
      const TileHdr h = hdr[b];
      if (AXIS0) {
        // transposing stage: lane = line writes its R outputs, then lane = (line % (32 / R), output) stores whole
        // sectors: R consecutive doubles per line
#pragma unroll
        for (int r = 0; r < R; ++r) mystage[lane * SPITCH + r] = acc[r];
        __syncwarp();
        double *op = out + h.obase + (long long)sl * n + k0 + sr;
        const bool kok = k0 + sr < n;
#pragma unroll
        for (int it = 0; it < R; ++it) {
          const int tl2 = it * (32 / R) + sl;
          const double v = mystage[tl2 * SPITCH + sr];
          if (kok && tl2 < h.vo) op[(long long)it * (32 / R) * n] = v;
        }
      } else if (li < h.vi && lo < h.vo) {
        double *op = out + h.obase + lane_off + (long long)k0 * sp.inner;
        if (k0 + R <= n) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            *op = acc[r];
            op += sp.inner;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (k0 + r < n) *op = acc[r];
            op += sp.inner;
          }
        }
      }
      __syncwarp();  // every lane is done with the tile (and with the stage) before the slot is handed back
      if (lane == 0) mbar_arrive(empty + b);
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------
static int pipe_smem_budget() {
  static int cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int &v = cached[dev & 63];
  if (!v) {
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (v <= 0) v = 48 * 1024;
  }
  return v;
}

template <int TAPS, int BI, int R, bool MMA = false>
static int launch_smooth_pass(const CUtensorMap &tmap, const SmoothPipeParams &sp, double *out, int threads, size_t smem,
                              cudaStream_t st) {
  auto k = smooth_pipe_kernel<TAPS, BI, R, MMA>;
  {
    // the attribute is per (function, device): set it once to the device's limit instead of before every launch
    static std::atomic<unsigned long long> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      CIB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem_budget()));
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)std::min<long long>(sp.ntiles, (long long)sm_count()));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (tune_env("CIB200_NO_PDL")) cfg.numAttrs = 0;
  CIB_CUDA(cudaLaunchKernelEx(&cfg, k, tmap, sp, out));
  CIB_LAUNCHED();
  return CIB200_OK;
}

template <int TAPS>
static int launch_smooth_bi(int bi, int R, bool mma, const CUtensorMap &tmap, const SmoothPipeParams &sp, double *out, int threads,
                            size_t smem, cudaStream_t st) {
  if (mma) return bi == 0 ? launch_smooth_pass<TAPS, 0, 8, true>(tmap, sp, out, threads, smem, st)
                          : launch_smooth_pass<TAPS, 32, 8, true>(tmap, sp, out, threads, smem, st);
  (void)R;  // R = 16 was measured (64^4: 0.279 vs 0.272 ms, 40^4: 0.077 vs 0.069 ms) and is not instantiated
  if (bi == 0) return launch_smooth_pass<TAPS, 0, 8>(tmap, sp, out, threads, smem, st);
  if (bi == 8) return launch_smooth_pass<TAPS, 8, 8>(tmap, sp, out, threads, smem, st);
  return launch_smooth_pass<TAPS, 32, 8>(tmap, sp, out, threads, smem, st);
}

// One pass of convolution_smooth along the axis that splits the array into [inner][n][outer].  Returns -1 when the
// pass does not qualify (the caller falls back to the generic sweep kernels), a cib200 status otherwise.
int smooth_pipe_pass(const double *in, double *out, long long inner, long long n_ll, long long outer, int E,
                     const double *kw, cudaStream_t st) {
  const int taps = 2 * E + 1;
  if (taps != 13 && taps != 17 && taps != 25 && taps != 35 && taps != 49) return -1;
  if (n_ll <= E || n_ll + 2 * E + 16 > 250) return -1;  // TMA box extents are at most 256 elements
  if (((uintptr_t)in & 15) || !tma_encode_tiled_fn()) return -1;
  const int n = (int)n_ll;
  const bool axis0 = inner == 1;
  if (const char *e = tune_env("CIB200_SMOOTH_NEW")) {  // 1: axis 0 only, 2: axis >= 1 only
    if ((e[0] == '1' && !axis0) || (e[0] == '2' && axis0)) return -1;
  }
  SmoothPipeParams sp;
  memset(&sp, 0, sizeof sp);
  sp.inner = inner;
  sp.outer = outer;
  sp.n = n;
  // outputs per work item
  const int R = 8;
  sp.nblk = (n + R - 1) / R;
  sp.line_o_stride = inner * n_ll;
  int bi = 0;
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof tmap);
  const int rows = sp.nblk * R + 2 * E + 3;  // -E .. nblk R + E - 1: every window row a block can touch (+ the K-step padding)
  // tensor-core consumer (DMMA): measured SLOWER than the DFMA consumer here (64^4: 0.320 against 0.259 ms, 40^4: 0.066
  // against 0.061 ms) - the FIR pass is bound by the FP64 pipe itself, and the Toeplitz band pads 49 taps to 56 K rows
  // while the DFMA loop touches real taps only - so it is off unless an experiments build asks for it (CIB200_SMOOTH_MMA=1)
  bool mma = (tune_env("CIB200_SMOOTH_MMA") && tune_env("CIB200_SMOOTH_MMA")[0] == '1') &&
             (axis0 || inner % 32 == 0 || !(inner % 8 == 0 && inner < 256));
  if (axis0) {
    if ((n & 1) || (E & 1)) return -1;  // line stride and box start (-E elements) must be multiples of 16 bytes
    int P = rows;
    if (mma) {
      while ((P & 15) != 4) ++P;  // line pitch = 4 mod 16 doubles: the B fragments (4 rows x 8 lines) hit distinct banks
    } else {
      while ((P & 3) != 2) ++P;  // P / 2 odd: the 8 lanes of an LDS.128 wavefront fall into 8 distinct 16-byte banks
    }
    if (P > 256 || outer >= (1LL << 31)) return -1;
    sp.H = P;
    sp.ntiles = (outer + 31) / 32;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)outer};
    cuuint64_t gstr[1] = {(cuuint64_t)n * 8};
    cuuint32_t box[2] = {(cuuint32_t)P, 32};
    if (pipe::encode_f64(&tmap, 2, in, gdim, gstr, box)) return -1;
  } else {
    if ((inner & 1) || inner >= (1LL << 31) || outer >= (1LL << 31)) return -1;
    bi = (inner % 32 == 0) ? 32 : ((inner % 8 == 0 && inner < 256) ? 8 : 32);
    int H = rows;
    if (bi < 32 && !(H & 1)) ++H;  // odd: the 32 / BI outer slabs of a tile start in different banks
    if (bi != 32) mma = false;
    if (mma) H = (H + 3) & ~3;     // 36-double rows: a tile is a multiple of 128 bytes (TMA destination alignment)
    if (H > 256) return -1;
    sp.H = H;
    sp.tiles_inner = (int)((inner + bi - 1) / bi);
    const int bo = 32 / bi;
    sp.ntiles = (long long)sp.tiles_inner * ((outer + bo - 1) / bo);
    cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)n, (cuuint64_t)outer};
    cuuint64_t gstr[2] = {(cuuint64_t)inner * 8, (cuuint64_t)inner * (cuuint64_t)n * 8};
    cuuint32_t box[3] = {(cuuint32_t)(mma ? SM_MMA_PITCH : bi), (cuuint32_t)H, (cuuint32_t)bo};
    if (pipe::encode_f64(&tmap, 3, in, gdim, gstr, box)) return -1;
  }
  if (sp.ntiles * (long long)std::max(sp.nblk, 32) >= (1LL << 26)) return -1;  // 32-bit tile / item counters with exact magic division
  // roles: 1 producer warp, stats warps, consumer warps
  sp.nstats = 2;  // a tile's stats are a ~600-cycle dependent chain, its FIR ~1 300 cycles of the FP64 pipe
  sp.ncons = 16;
  if (const char *e = tune_env("CIB200_SMOOTH_NSTATS")) sp.nstats = std::max(1, std::min(4, atoi(e)));
  if (const char *e = tune_env("CIB200_SMOOTH_NCONS")) sp.ncons = std::max(1, std::min(17 - sp.nstats + 2, atoi(e)));
  const int threads = 32 * (1 + sp.nstats + sp.ncons);
  if (threads > 640) return -1;
  const size_t tile_bytes = (size_t)((mma && !axis0) ? SM_MMA_PITCH : 32) * sp.H * 8;
  const size_t fixed = 64 * 8 + 8 * sizeof(TileHdr) + (axis0 ? (size_t)sp.ncons * 32 * (R + 1) * 8 : 0) + 8 * 3 * 8 + 16 +
                       1024 /* alignment slack */;
  const size_t per_slot = tile_bytes + 128 * 8;
  const size_t budget = (size_t)pipe_smem_budget();
  if (fixed + 3 * per_slot > budget) return -1;
  sp.nb = (int)std::min<size_t>(8, (budget - fixed) / per_slot);
  if (const char *e = tune_env("CIB200_SMOOTH_NB")) sp.nb = std::max(2, std::min(sp.nb, atoi(e)));
  auto magic = [](unsigned d) { return pipe::div_magic_for(d); };
  sp.magic_nblk = magic((unsigned)sp.nblk);
  sp.magic_nb = magic((unsigned)sp.nb);
  sp.magic_ti = magic((unsigned)std::max(1, sp.tiles_inner));
  // least-squares constants (create_convolutional_coefs.jl:455-470): sums of small integers, exact
  const int nfit = std::max(E / 3, 6);
  if (n < nfit) return -1;
  sp.nf = (double)nfit;
  sp.sx = sp.nf * (sp.nf + 1.0) * 0.5;
  const double sxx = sp.nf * (sp.nf + 1.0) * (2.0 * sp.nf + 1.0) / 6.0;
  sp.inv_denom = 1.0 / (sp.nf * sxx - sp.sx * sp.sx);
  sp.inv_nf = 1.0 / sp.nf;
  sp.inv_n = 1.0 / (double)n;
  for (int l = 0; l < taps; ++l) sp.kw[l] = kw[l];
  // ghost k = 1..E-1 sits k + 1 slots outside the line and holds a + b (1 - k); output at distance i from the boundary
  // sees the ghosts with k + 1 + i <= E
  for (int i = 0; i + 1 < E && i < 32; ++i) {
    double a0 = 0.0, a1 = 0.0;
    for (int k = 1; k + 1 + i <= E; ++k) {
      const double w = kw[E - (k + 1 + i)];
      a0 += w;
      a1 += w * (double)(1 - k);
    }
    sp.s0[i] = a0;
    sp.s1[i] = a1;
  }
  const size_t smem = fixed + (size_t)sp.nb * per_slot;
  switch (taps) {
    case 13: return launch_smooth_bi<13>(bi, R, mma, tmap, sp, out, threads, smem, st);
    case 17: return launch_smooth_bi<17>(bi, R, mma, tmap, sp, out, threads, smem, st);
    case 25: return launch_smooth_bi<25>(bi, R, mma, tmap, sp, out, threads, smem, st);
    case 35: return launch_smooth_bi<35>(bi, R, mma, tmap, sp, out, threads, smem, st);
    default: return launch_smooth_bi<49>(bi, R, mma, tmap, sp, out, threads, smem, st);
  }
}

}  // namespace cib
