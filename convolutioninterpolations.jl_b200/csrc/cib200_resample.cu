// cib200_resample.cu - convolution_resample, Float64, as a chain of persistent TMA-fed passes.
//
// Replaces the hot loop of convolution_resample_internal (convolution/convolution_resample.jl:297-381) - per line:
// ghost fill (apply_boundary_conditions_for_dim!, create_convolutional_coefs.jl:200-370), then for every output knot
// the 2 eqs-tap sum with Hermite-blended table weights (:320-379) - for the shapes BASELINE config 3 names.  Anything
// that does not qualify (Float32, lines longer than a TMA box, odd strides) stays on cib200_sweep.cu's kernels.
//
// One pass = one axis, array viewed as [inner][n_in][outer] -> [inner][n_out][outer].  What bounds a pass is HBM (an
// output point costs 18 DFMAs for :b13, 21 bytes of traffic), so the kernel is organised to keep the copy engines
// busy and everything else off the critical path:
//
//  * producer warp: one TMA box per tile of 32 lines (axis >= 1: {BI, n_in + 2 (eqs - 1), 32 / BI}, box origin at row
//    -(eqs - 1), so the ghost rows arrive zero-filled; axis 0: line-major {P, 32}), ring of tiles, mbarrier hand-offs;
//  * ghost teams (lane = line, one warp per side): line mean, :detect thresholds, Lagrange ghosts written into the
//    tile's margin rows - the reference's arithmetic, operation by operation (cib200_ghost.cuh);
//  * consumers pull (tile, block of 8 output knots) items from a shared counter.  The weights of a block are a DENSE
//    [window row][8 knots] table with zeros where a tap does not reach (resample_blocks_kernel builds it once per pass:
//    positions and blended weights depend on the output knot only, the reference recomputes them for every line), so
//    the inner loop has no per-knot offset logic: per window row one load of the lane's input and four broadcast
//    16-byte loads of the row's 8 weights feed 8 DFMAs with static register indices.  At 2x upsampling a block's
//    window is 18 + 4 rows: 22 % more DFMAs than taps, on a pass that has FP64 to spare;
//  * axis 0: lane = line reads two window rows per 16-byte load (pitch / 2 odd: conflict-free), results go through a
//    per-warp transposing stage so that global stores are whole sectors;
//  * the passes of a call are chained with programmatic dependent launch; the block tables of all axes are built
//    before the first pass so that nothing sits between two passes.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "cib200_ghost.cuh"
#include "cib200_pipe.cuh"

namespace cib {
using namespace pipe;

static constexpr int RS_R = 8;  // output knots per work item
static constexpr int RS_MMA_PITCH = 36;  // doubles between two rows of a tile for the tensor-core consumer (axis >= 1)

__device__ __forceinline__ void mma_f64_884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct ResamplePipeParams {
  long long inner, outer, ntiles;
  long long out_o_stride;  // axis >= 1: inner * n_out
  int n_in, n_out, nblk;
  int nitem, bpi;   // work items per tile slot, blocks per item (tensor-core consumer: several blocks per item)
  int H;            // axis >= 1: rows of a tile buffer; axis 0: pitch P of a line
  int tiles_inner;
  int nb, nteams, ncons;
  int pad;          // eqs - 1 ghosts per side
  int padx;         // physical left margin of a line in the tile (axis 0: pad rounded up to even)
  int winmax;       // rows per block in the dense weight table
  int bcl, bcr;
  unsigned magic_nblk, magic_nb, magic_ti, magic_teams, magic_nitem;
  double h, scale;
  const double *Wd;  // [nblk][winmax][8]
  const int2 *binfo;  // per block: first physical window row, window rows
};

struct RsTileHdr {
  long long obase;
  int vi, vo;
};

// ---- per-pass dense block tables: convolution_resample.jl:320-361 minus the data -------------------------------
// One CTA per block of 8 output knots.  shift = padx - pad (physical row of padded index i is i + shift).
__global__ void __launch_bounds__(256) resample_blocks_kernel(const double *__restrict__ knots_out, int n_out, double x_expanded_1,
                                                              double h, int eqs, long long n_coef,
                                                              const double *__restrict__ tk, const double *__restrict__ td,
                                                              const double *__restrict__ te, int linear, int npre, int shift,
                                                              int even, int winmax, double *__restrict__ Wd,
                                                              int2 *__restrict__ binfo) {
  __shared__ AxisPos pos[RS_R];
  __shared__ int s_a, s_win;
  const int b = blockIdx.x, k0 = b * RS_R, tid = threadIdx.x;
  const int taps = 2 * eqs;
  if (tid < RS_R) {
    const int k = k0 + tid < n_out ? k0 + tid : n_out - 1;
    locate_computed(knots_out[k], x_expanded_1, h, eqs, n_coef, pos[tid], npre);
  }
  __syncthreads();
  if (tid == 0) {
    int lo = pos[0].base, hi = pos[0].base;
    for (int r = 1; r < RS_R; ++r) {
      lo = min(lo, pos[r].base);
      hi = max(hi, pos[r].base);
    }
    // axis 0 reads the window two rows per 16-byte load: even origin, even length
    const int a = even ? ((lo + shift) & ~1) : lo + shift;
    int win = hi + shift + taps - a;
    if (even) win = (win + 1) & ~1;
    s_a = a;
    s_win = win;
    binfo[b] = make_int2(a, win <= winmax ? win : -1);  // -1: host bound violated, the pass kernel traps
  }
  __syncthreads();
  const int a = s_a, win = s_win;
  for (int idx = tid; idx < winmax * RS_R; idx += blockDim.x) {
    const int j = idx / RS_R, r = idx - j * RS_R;
    const int l = a + j - (pos[r].base + shift);
    double w = 0.0;
    if (k0 + r < n_out && j < win && l >= 0 && l < taps) {
      const double t = pos[r].t;
      w = te ? blended_weight_quintic(hermite_basis_quintic(t), tk, td, te, pos[r].row, l, taps)
             : blended_weight(linear ? linear_basis(t) : hermite_basis(t), tk, td, pos[r].row, l, taps);
    }
    Wd[((size_t)b * winmax + j) * RS_R + r] = w;
  }
}

// host twin of locate_computed's cell index (same IEEE operations; the library is compiled without fp contraction):
// only used to bound the window length of a block before the device computes the real one
static int host_base(double x, double x1, double h, int eqs, long long n_coef) {
  const double i_float = (x - x1) / h + 1.0;
  long long i = fabs(i_float) < 9.0e18 ? (long long)floor(i_float) : (long long)eqs;
  i = i < eqs ? eqs : i;
  i = i > n_coef - eqs ? n_coef - eqs : i;
  return (int)(i - eqs);
}

// ---- one pass ---------------------------------------------------------------------------------------------------
// BI: lanes along `inner` in a tile (32 or 8); BI = 0: axis 0 (line-major tile).
// RL: tiles per ring slot = lines per consumer thread.  The 8 weights of a window row are warp-uniform 16-byte loads
// (2 shared-memory wavefronts each): with one line per thread a row costs 2 + 8 wavefronts for 8 DFMAs and the pass is
// bound by shared-memory bandwidth (measured: 124 k of 159 k cycles on the 256 x 256 x 128 -> 256^3 pass); two lines per
// thread share the weights, 12 wavefronts for 16 DFMAs.
// MMA: the consumers contract a block on the FP64 tensor cores instead (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4): rows of the
// product = the block's 8 output knots, K = the window rows, columns = 8 lines; the dense block table IS the A operand
// and a lane loads ONE double per operand and 8 FMAs, where the DFMA consumer needs a weight broadcast per FMA pair -
// 60 shared-memory wavefronts per (block, 32 lines) instead of 220.  Tile rows get a pitch of 36 doubles (axis >= 1) /
// lines a pitch = 4 mod 16 (axis 0) so that the four K rows x eight lines of a B fragment hit distinct banks.
template <int EQS, int BI, int RL, bool MMA>
__global__ void __maxnreg__(RL == 2 ? 128 : 96)
    resample_pipe_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ ResamplePipeParams sp,
                         const __grid_constant__ GhostMats gm, double *__restrict__ out) {
  constexpr bool AXIS0 = BI == 0;
  constexpr int BIX = BI > 0 ? BI : 32;
  constexpr int ROW = AXIS0 ? 1 : (MMA ? RS_MMA_PITCH : BI);  // doubles between two rows of a tile
  static_assert(!MMA || (RL == 1 && BI != 8), "tensor-core consumer: one tile per slot, 32 lines along inner or axis 0");
  constexpr int BO = AXIS0 ? 1 : 32 / BIX;
  constexpr int NS = EQS < 5 ? 4 : (EQS == 5 ? 6 : 8);
  constexpr int SPITCH = RS_R + 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int NB = sp.nb, nblk = sp.nblk, n = sp.n_in;
  const int TILE = (AXIS0 ? 32 : (MMA ? RS_MMA_PITCH : 32)) * sp.H;  // doubles
  const int SLOT = RL * TILE;
  double *tiles = (double *)smem_raw;
  double *wd = tiles + (size_t)NB * SLOT;                       // [nblk][winmax][8]
  int2 *binfo = (int2 *)(wd + (size_t)nblk * sp.winmax * RS_R);  // [nblk]
  RsTileHdr *hdr = (RsTileHdr *)(binfo + ((nblk + 1) & ~1));    // [NB][RL]
  double *stage = (double *)(hdr + NB * RL);                    // axis 0: [ncons][32 * SPITCH]
  unsigned long long *bars = (unsigned long long *)(stage + (AXIS0 ? sp.ncons * 32 * SPITCH : 0));
  unsigned long long *loaded = bars, *ready = bars + NB, *empty = bars + 2 * NB;
  unsigned *counter = (unsigned *)(bars + 3 * NB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int b = 0; b < NB; ++b) {
      mbar_init(loaded + b, 1);
      mbar_init(ready + b, 2);  // the two side warps of a ghost team
      mbar_init(empty + b, (unsigned)sp.nitem);
    }
    *counter = 0;
    mbar_init_fence();
  }
  {
    // the pass's block tables (written by resample_blocks_kernel launches that completed before the chain of passes
    // started: safe to read ahead of griddepcontrol.wait)
    const double2 *src = (const double2 *)sp.Wd;
    double2 *dst = (double2 *)wd;
    const int nv = nblk * sp.winmax * (RS_R / 2);
    for (int e = tid; e < nv; e += blockDim.x) dst[e] = __ldg(src + e);
    for (int e = tid; e < nblk; e += blockDim.x) binfo[e] = __ldg(sp.binfo + e);
  }
  __syncthreads();
  pdl_launch_dependents();
  const unsigned tstep = gridDim.x;
  const unsigned nslots = (unsigned)((sp.ntiles + RL - 1) / RL);  // a slot = RL consecutive tiles
  const unsigned my_slots = blockIdx.x < nslots ? (nslots - blockIdx.x + tstep - 1) / tstep : 0u;
  const int li = AXIS0 ? 0 : lane % BIX, lo = AXIS0 ? lane : lane / BIX;
  const int lbase = AXIS0 ? lane * sp.H : lo * sp.H * ROW + li;  // this lane's line: physical row 0 in a tile buffer

  if (warp == 0) {
    // ------------------------------- producer -------------------------------------------------------------------
    if (lane == 0) {
      pdl_wait();
      unsigned b = 0, rev = 0;
      for (unsigned seq = 0; seq < my_slots; ++seq) {
        const unsigned t0 = (blockIdx.x + seq * tstep) * RL;
        if (rev > 0) mbar_wait(empty + b, (rev - 1) & 1u, 1000 + (int)seq);
        fence_async_proxy();
        mbar_expect_tx(loaded + b, (unsigned)(SLOT * 8));
#pragma unroll
        for (int q = 0; q < RL; ++q) {
          // tiles past the end of the array are boxes outside the tensor: they arrive zero-filled and are never stored
          const unsigned t = t0 + q;
          double *dst = tiles + (size_t)b * SLOT + q * TILE;
          if (AXIS0) {
            tma_load_2d(dst, &tmap, -sp.padx, (int)(t * 32), loaded + b);
          } else {
            const unsigned tile_o = div_u(t, sp.magic_ti), tile_i = t - tile_o * sp.tiles_inner;
            tma_load_3d(dst, &tmap, (int)tile_i * BIX, -sp.padx, (int)tile_o * BO, loaded + b);
          }
        }
        if (++b == (unsigned)NB) {
          b = 0;
          ++rev;
        }
      }
    }
  } else if (warp <= 2 * sp.nteams) {
    // ------------------------------- ghost teams: lane = line, one warp per side ----------------------------------
    const int team = (warp - 1) >> 1, side = (warp - 1) & 1;
    for (unsigned seq = team; seq < my_slots; seq += sp.nteams) {
      const unsigned rev = div_u(seq, sp.magic_nb), b = seq - rev * NB;
      const unsigned t0 = (blockIdx.x + seq * tstep) * RL;
      mbar_wait(loaded + b, rev & 1u, 2000 + (int)seq);  // also: the slot's previous tiles have been consumed
      if (side == 0 && lane < RL) {
        const unsigned t = t0 + lane;
        RsTileHdr hd;
        hd.obase = 0;
        hd.vi = hd.vo = 0;
        if (t < (unsigned)sp.ntiles) {
          if (AXIS0) {
            hd.obase = (long long)t * 32 * sp.n_out;
            const long long rem = sp.outer - (long long)t * 32;
            hd.vi = 1;
            hd.vo = rem < 32 ? (int)rem : 32;
          } else {
            const unsigned tile_o = div_u(t, sp.magic_ti), tile_i = t - tile_o * sp.tiles_inner;
            hd.obase = (long long)tile_i * BIX + sp.out_o_stride * ((long long)tile_o * BO);
            const long long ri = sp.inner - (long long)tile_i * BIX, ro = sp.outer - (long long)tile_o * BO;
            hd.vi = ri < BIX ? (int)ri : BIX;
            hd.vo = ro < BO ? (int)ro : BO;
          }
        }
        hdr[b * RL + lane] = hd;
      }
#pragma unroll
      for (int q = 0; q < RL; ++q) {
        double *bl = tiles + (size_t)b * SLOT + q * TILE + lbase;
        // both side warps sum the line themselves, in the same order, so both use the same mean
        const double mean = __ddiv_rn(line_sum_rows<ROW>(bl + sp.padx * ROW, n), (double)n);
        ghost_poly_side<NS, EQS>(bl, sp.padx, ROW, n, mean, sp.h, side ? sp.bcr : sp.bcl, side == 1, gm);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ready + b);
    }
  } else if (warp - 1 - 2 * sp.nteams < sp.ncons) {
    // ------------------------------- consumers ---------------------------------------------------------------------
    const int cw = warp - 1 - 2 * sp.nteams;
    double *mystage = stage + cw * 32 * SPITCH;
    const long long lane_off = AXIS0 ? 0 : (long long)li + sp.out_o_stride * lo;
    const int sr = lane % RS_R, sl = lane / RS_R;
    const int nout = sp.n_out;
    for (;;) {
      unsigned item = 0;
      if (lane == 0) item = atomicAdd(counter, 1u);
      item = __shfl_sync(0xffffffffu, item, 0);
      const unsigned seq = div_u(item, sp.magic_nitem);
      if (seq >= my_slots) break;
      const unsigned rev = div_u(seq, sp.magic_nb), b = seq - rev * NB;
      const int blk_first = (int)(item - seq * (unsigned)sp.nitem) * sp.bpi;
      mbar_wait(ready + b, rev & 1u, 3000 + (int)seq);
      if (MMA) {
       const RsTileHdr hd = hdr[b * RL];
       for (int blk = blk_first; blk < min(blk_first + sp.bpi, nblk); ++blk) {
        const int k0 = blk * RS_R;
        const int2 bw = binfo[blk];
        if (bw.y < 0) __trap();  // a block window longer than the host's bound: never silently drop taps
        // rows = the 8 knots of the block, K = window rows (4 per step), columns = 8 lines; 4 column groups per tile
        const int mq = lane >> 2, mr = lane & 3;
        const double *wblk = wd + (size_t)blk * sp.winmax * RS_R + mq;           // A[knot mq][row j] = wblk[j * 8]
        const double *xb = tiles + (size_t)b * SLOT + (AXIS0 ? mq * sp.H + bw.x : bw.x * ROW + mq);  // B[row j][line mq]
        constexpr int XJ = AXIS0 ? 1 : ROW;              // stride of a window row
        const int XG = AXIS0 ? 8 * sp.H : 8;             // stride of a column group (8 lines)
        double ac[4][2];
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) ac[cg][0] = ac[cg][1] = 0.0;
#pragma unroll 2
        for (int j0 = 0; j0 < bw.y; j0 += 4) {
          const int j = j0 + mr;
          const bool jv = j < bw.y;  // (rows past the window may lie outside the tile buffer: never read them)
          const double a = jv ? wblk[j * RS_R] : 0.0;
          double xv[4];
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) xv[cg] = jv ? xb[j * XJ + cg * XG] : 0.0;
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) mma_f64_884(ac[cg][0], ac[cg][1], a, xv[cg]);
        }
        if (k0 + mq < nout) {
#pragma unroll
          for (int cg = 0; cg < 4; ++cg) {
            const int line = 8 * cg + 2 * mr;
            const double v0 = __dmul_rn(ac[cg][0], sp.scale), v1 = __dmul_rn(ac[cg][1], sp.scale);  // (-1/h)^derivative, :379
            if (AXIS0) {
              double *op = out + hd.obase + (long long)line * nout + k0 + mq;
              if (line < hd.vo) op[0] = v0;
              if (line + 1 < hd.vo) op[nout] = v1;
            } else if (hd.vo > 0) {
              double *op = out + hd.obase + (long long)(k0 + mq) * sp.inner + line;
              if (line + 1 < hd.vi) *(double2 *)op = make_double2(v0, v1);  // inner is even: 16-byte aligned
              else if (line < hd.vi) op[0] = v0;
            }
          }
        }
       }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + b);
        continue;
      }
      const int blk = blk_first;  // (bpi = 1)
      const int k0 = blk * RS_R;
      const int2 bw = binfo[blk];
      if (bw.y < 0) __trap();  // a block window longer than the host's bound: never silently drop taps
      const double2 *w = (const double2 *)(wd + (size_t)blk * sp.winmax * RS_R);
      const double *p = tiles + (size_t)b * SLOT + lbase + bw.x * ROW;
      double acc[RL][RS_R];
#pragma unroll
      for (int q = 0; q < RL; ++q)
#pragma unroll
        for (int r = 0; r < RS_R; ++r) acc[q][r] = 0.0;
      if (AXIS0) {
#pragma unroll 2
        for (int j = 0; j < bw.y; j += 2) {
          double2 x[RL];
#pragma unroll
          for (int q = 0; q < RL; ++q) x[q] = *(const double2 *)(p + q * TILE + j);
          double2 wa[4], wb[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            wa[u] = w[4 * j + u];
            wb[u] = w[4 * j + 4 + u];
          }
#pragma unroll
          for (int q = 0; q < RL; ++q) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[q][2 * u] = fma(x[q].x, wa[u].x, acc[q][2 * u]);
              acc[q][2 * u + 1] = fma(x[q].x, wa[u].y, acc[q][2 * u + 1]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[q][2 * u] = fma(x[q].y, wb[u].x, acc[q][2 * u]);
              acc[q][2 * u + 1] = fma(x[q].y, wb[u].y, acc[q][2 * u + 1]);
            }
          }
        }
      } else {
#pragma unroll 4
        for (int j = 0; j < bw.y; ++j) {
          double x[RL];
#pragma unroll
          for (int q = 0; q < RL; ++q) x[q] = p[q * TILE + j * ROW];
          double2 wa[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) wa[u] = w[4 * j + u];
#pragma unroll
          for (int q = 0; q < RL; ++q)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[q][2 * u] = fma(x[q], wa[u].x, acc[q][2 * u]);
              acc[q][2 * u + 1] = fma(x[q], wa[u].y, acc[q][2 * u + 1]);
            }
        }
      }
#pragma unroll
      for (int q = 0; q < RL; ++q) {
#pragma unroll
        for (int r = 0; r < RS_R; ++r) acc[q][r] = __dmul_rn(acc[q][r], sp.scale);  // (-1/h)^derivative, :379
        const RsTileHdr hd = hdr[b * RL + q];
        if (AXIS0) {
          if (q > 0) __syncwarp();  // the stage is reused by the slot's next tile
#pragma unroll
          for (int r = 0; r < RS_R; ++r) mystage[lane * SPITCH + r] = acc[q][r];
          __syncwarp();
          double *op = out + hd.obase + (long long)sl * nout + k0 + sr;
          const bool kok = k0 + sr < nout;
#pragma unroll
          for (int it = 0; it < RS_R; ++it) {
            const int tl2 = it * (32 / RS_R) + sl;
            const double v = mystage[tl2 * SPITCH + sr];
            if (kok && tl2 < hd.vo) op[(long long)it * (32 / RS_R) * nout] = v;
          }
        } else if (li < hd.vi && lo < hd.vo) {
          double *op = out + hd.obase + lane_off + (long long)k0 * sp.inner;
          if (k0 + RS_R <= nout) {
#pragma unroll
            for (int r = 0; r < RS_R; ++r) {
              *op = acc[q][r];
              op += sp.inner;
            }
          } else {
#pragma unroll
            for (int r = 0; r < RS_R; ++r) {
              if (k0 + r < nout) *op = acc[q][r];
              op += sp.inner;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + b);
    }
  }
}

// ---- host -------------------------------------------------------------------------------------------------------
static int rs_smem_budget() {
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

struct RsPass {
  CUtensorMap tmap;
  ResamplePipeParams sp;
  GhostMats gm;
  int eqs, bi, rl, threads, mma;
  size_t smem;
  double *out;
};

template <int EQS, int BI, int RL, bool MMA = false>
static int launch_rs(const RsPass &ps, cudaStream_t st) {
  auto k = resample_pipe_kernel<EQS, BI, RL, MMA>;
  {
    static std::atomic<unsigned long long> done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      CIB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_smem_budget()));
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)std::min<long long>((ps.sp.ntiles + RL - 1) / RL, (long long)sm_count()));
  cfg.blockDim = dim3((unsigned)ps.threads);
  cfg.dynamicSmemBytes = ps.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tune_env("CIB200_NO_PDL") ? 0 : 1;
  CIB_CUDA(cudaLaunchKernelEx(&cfg, k, ps.tmap, ps.sp, ps.gm, ps.out));
  CIB_LAUNCHED();
  return CIB200_OK;
}

template <int EQS>
static int launch_rs_bi(const RsPass &ps, cudaStream_t st) {
  if (ps.mma) return ps.bi == 0 ? launch_rs<EQS, 0, 1, true>(ps, st) : launch_rs<EQS, 32, 1, true>(ps, st);
  if (ps.rl == 2) {
    if (ps.bi == 0) return launch_rs<EQS, 0, 2>(ps, st);
    if (ps.bi == 8) return launch_rs<EQS, 8, 2>(ps, st);
    return launch_rs<EQS, 32, 2>(ps, st);
  }
  if (ps.bi == 0) return launch_rs<EQS, 0, 1>(ps, st);
  if (ps.bi == 8) return launch_rs<EQS, 8, 1>(ps, st);
  return launch_rs<EQS, 32, 1>(ps, st);
}

static int launch_rs_eqs(const RsPass &ps, cudaStream_t st) {
  switch (ps.eqs) {
    case 2: return launch_rs_bi<2>(ps, st);
    case 3: return launch_rs_bi<3>(ps, st);
    case 4: return launch_rs_bi<4>(ps, st);
    case 5: return launch_rs_bi<5>(ps, st);
    case 6: return launch_rs_bi<6>(ps, st);
    case 7: return launch_rs_bi<7>(ps, st);
    case 8: return launch_rs_bi<8>(ps, st);
    case 9: return launch_rs_bi<9>(ps, st);
  }
  return -1;
}

// Plans one pass.  Returns false when it does not qualify.  `knots_out_h`: host knots (bounds the block windows).
static bool plan_pass(RsPass &ps, const double *in, double *out, long long inner, long long n_in, long long n_out,
                      long long outer, int kernel, int eqs, int bcl, int bcr, double h, double x_expanded_1, double scale,
                      const double *knots_out_h) {
  memset(&ps, 0, sizeof ps);
  const int pad = eqs - 1;
  const long long n_coef = n_in + 2 * pad;
  if (n_coef + 2 > 250 || n_in < 2 || n_out < 1 || n_out > 4096) return false;
  if (((uintptr_t)in & 15) || !tma_encode_tiled_fn()) return false;
  const bool axis0 = inner == 1;
  // tensor-core consumer (DMMA): axis 0 and passes with 32 lines along `inner`
  const bool mma = !(tune_env("CIB200_RS_MMA") && tune_env("CIB200_RS_MMA")[0] == '0') && (axis0 || inner % 32 == 0 || !(inner % 8 == 0 && inner < 256));
  ResamplePipeParams &sp = ps.sp;
  sp.inner = inner;
  sp.outer = outer;
  sp.n_in = (int)n_in;
  sp.n_out = (int)n_out;
  sp.nblk = (int)((n_out + RS_R - 1) / RS_R);
  sp.out_o_stride = inner * n_out;
  sp.pad = pad;
  sp.bcl = bcl;
  sp.bcr = bcr;
  sp.h = h;
  sp.scale = scale;
  ps.eqs = eqs;
  ps.out = out;
  ghost_mats_for(kernel, eqs, &ps.gm);
  if (axis0) {
    if (n_in & 1) return false;  // global line stride must be a multiple of 16 bytes
    sp.padx = pad + (pad & 1);   // the box origin (-padx elements) must be a multiple of 16 bytes
    int P = (int)n_coef + (sp.padx - pad) + 1;
    if (mma) {
      while ((P & 15) != 4) ++P;  // line pitch = 4 mod 16 doubles: the B fragments (4 rows x 8 lines) hit distinct banks
    } else {
      while ((P & 3) != 2) ++P;  // P / 2 odd: conflict-free 16-byte reads with lane = line
    }
    if (P > 256 || outer >= (1LL << 31)) return false;
    sp.H = P;
    sp.ntiles = (outer + 31) / 32;
    ps.bi = 0;
    cuuint64_t gdim[2] = {(cuuint64_t)n_in, (cuuint64_t)outer};
    cuuint64_t gstr[1] = {(cuuint64_t)n_in * 8};
    cuuint32_t box[2] = {(cuuint32_t)P, 32};
    if (pipe::encode_f64(&ps.tmap, 2, in, gdim, gstr, box)) return false;
  } else {
    if ((inner & 1) || inner >= (1LL << 31) || outer >= (1LL << 31)) return false;
    sp.padx = pad;
    ps.bi = (inner % 32 == 0) ? 32 : ((inner % 8 == 0 && inner < 256) ? 8 : 32);
    int H = (int)n_coef + 1;  // + 1: a block window is rounded up to an even number of rows
    if (ps.bi < 32 && !(H & 1)) ++H;
    if (mma && ps.bi == 32) H = (H + 3) & ~3;  // 36-double rows: a tile is a multiple of 128 bytes (TMA destination alignment)
    if (H > 256) return false;
    sp.H = H;
    sp.tiles_inner = (int)((inner + ps.bi - 1) / ps.bi);
    const int bo = 32 / ps.bi;
    sp.ntiles = (long long)sp.tiles_inner * ((outer + bo - 1) / bo);
    cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)n_in, (cuuint64_t)outer};
    cuuint64_t gstr[2] = {(cuuint64_t)inner * 8, (cuuint64_t)inner * (cuuint64_t)n_in * 8};
    cuuint32_t box[3] = {(cuuint32_t)(mma ? RS_MMA_PITCH : ps.bi), (cuuint32_t)H, (cuuint32_t)bo};
    if (pipe::encode_f64(&ps.tmap, 3, in, gdim, gstr, box)) return false;
  }
  if (sp.ntiles * (long long)std::max(sp.nblk, 32) >= (1LL << 26)) return false;
  // bound of the block windows from the host twin of the cell index (+ 2 rows of margin)
  const int shift = sp.padx - pad;
  int winmax = 0;
  for (int b = 0; b < sp.nblk; ++b) {
    int lo = 1 << 30, hi = -1;
    for (int r = 0; r < RS_R; ++r) {
      const long long k = std::min<long long>((long long)b * RS_R + r, n_out - 1);
      const int base = host_base(knots_out_h[k], x_expanded_1, h, eqs, n_coef);
      lo = std::min(lo, base);
      hi = std::max(hi, base);
    }
    const int a = axis0 ? ((lo + shift) & ~1) : lo + shift;
    const int win = hi + shift + 2 * eqs - a;
    winmax = std::max(winmax, axis0 ? ((win + 1) & ~1) : win);
  }
  // (the device recomputes the windows with the same IEEE operations; a longer one would make the pass trap, not drop taps)
  sp.winmax = winmax;
  // roles: producer, ghost teams (2 warps each), consumers.  Two lines per consumer thread (ring slots of two tiles:
  // the 8 broadcast weight loads of a window row then feed 16 DFMAs) when two such slots fit and every SM gets at least
  // one of them (measured on BASELINE config 3: 0.144 ms against 0.161 ms with one line per thread everywhere); else one
  // line per thread and a deeper ring.
  ps.mma = (mma && (axis0 || ps.bi == 32)) ? 1 : 0;
  const size_t tile_bytes = (size_t)((ps.mma && !axis0) ? RS_MMA_PITCH : 32) * sp.H * 8;
  const size_t wd_bytes = (size_t)sp.nblk * winmax * RS_R * 8 + (size_t)((sp.nblk + 1) & ~1) * 8;
  const size_t budget = (size_t)rs_smem_budget();
  auto fixed_for = [&](int ncons) {
    return wd_bytes + 16 * sizeof(RsTileHdr) + (axis0 ? (size_t)ncons * 32 * (RS_R + 1) * 8 : 0) + 8 * 3 * 8 + 16 +
           1024 /* alignment slack */;
  };
  sp.nteams = 2;
  if (const char *e = tune_env("CIB200_RS_NTEAMS")) sp.nteams = std::max(1, std::min(3, atoi(e)));
  ps.rl = (fixed_for(11) + 4 * tile_bytes <= budget && sp.ntiles >= 2 * (long long)sm_count()) ? 2 : 1;
  if (const char *e = tune_env("CIB200_RS_RL")) ps.rl = (atoi(e) == 2 && fixed_for(11) + 4 * tile_bytes <= budget) ? 2 : 1;
  if (ps.mma) ps.rl = 1;
  // 128 registers x 16 warps (two lines per thread) or 96 x 20 warps
  sp.ncons = (ps.rl == 2 ? 15 : 19) - 2 * sp.nteams;
  if (const char *e = tune_env("CIB200_RS_NCONS")) sp.ncons = std::max(1, std::min(sp.ncons, atoi(e)));
  ps.threads = 32 * (1 + 2 * sp.nteams + sp.ncons);
  const size_t fixed = fixed_for(sp.ncons);
  if (fixed + 3 * tile_bytes > budget) return false;
  sp.nb = (int)std::min<size_t>(8, (budget - fixed) / (tile_bytes * ps.rl));
  if (const char *e = tune_env("CIB200_RS_NB")) sp.nb = std::max(2, std::min(sp.nb, atoi(e)));
  ps.smem = fixed + (size_t)sp.nb * tile_bytes * ps.rl;
  auto magic = [](unsigned d) { return pipe::div_magic_for(d); };
  sp.bpi = ps.mma ? 2 : 1;
  if (const char *e = tune_env("CIB200_RS_BPI")) sp.bpi = ps.mma ? std::max(1, std::min(8, atoi(e))) : 1;
  sp.nitem = (sp.nblk + sp.bpi - 1) / sp.bpi;
  sp.magic_nblk = magic((unsigned)sp.nblk);
  sp.magic_nitem = magic((unsigned)sp.nitem);
  sp.magic_nb = magic((unsigned)sp.nb);
  sp.magic_ti = magic((unsigned)std::max(1, sp.tiles_inner));
  sp.magic_teams = magic((unsigned)sp.nteams);
  return true;
}

// ---- cache of the per-pass block tables ---------------------------------------------------------------------------
// The dense block tables of a pass depend on the output knots and the kernel only - not on the data - and building them
// costs an upload of the knots and one small launch per axis (3 x 4 us of the 0.12 ms of BASELINE config 3).  A repeated
// resample onto the same knots (the usual case: every time step, every field) finds them here.  Entries are keyed on every
// input of resample_blocks_kernel (the table pointers are the library's own device copies of the shipped tables) and on
// the knot values themselves; an event orders a first use on another stream after the build.
struct BlockTabEntry {
  int dev, n_out, eqs, linear, npre, shift, even, winmax;
  long long n_coef;
  double x1, h;
  const double *tk, *td, *te;
  std::vector<double> knots;
  double *wd;
  int2 *bi;
  cudaEvent_t ready;
  cudaStream_t stream;
  unsigned long long stamp;
};
static std::mutex g_bt_mu;
static std::vector<BlockTabEntry> g_bt;
static unsigned long long g_bt_clock = 0;
static constexpr size_t BT_MAX_ENTRIES = 48;

// returns 0 and the device tables, or -1 (the caller builds them into stream-ordered scratch as before)
static int block_tables_cached(const double *knots_out, int n_out, double x1, double h, int eqs, long long n_coef, const double *tk,
                               const double *td, const double *te, int linear, int npre, int shift, int even, int winmax,
                               cudaStream_t st, const double **wd_out, const int2 **bi_out) {
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) return -1;
  const bool capturing = cap != cudaStreamCaptureStatusNone;  // a hit is fine under capture; a miss would need cudaMalloc
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> guard(g_bt_mu);
  for (BlockTabEntry &e : g_bt) {
    if (e.dev == dev && e.n_out == n_out && e.eqs == eqs && e.linear == linear && e.npre == npre && e.shift == shift &&
        e.even == even && e.winmax == winmax && e.n_coef == n_coef && e.x1 == x1 && e.h == h && e.tk == tk && e.td == td &&
        e.te == te && memcmp(e.knots.data(), knots_out, sizeof(double) * (size_t)n_out) == 0) {
      e.stamp = ++g_bt_clock;
      if (st != e.stream && cudaEventQuery(e.ready) != cudaSuccess) {  // built on another stream and not finished yet
        if (capturing || cudaStreamWaitEvent(st, e.ready, 0) != cudaSuccess) return -1;
      }
      *wd_out = e.wd;
      *bi_out = e.bi;
      return 0;
    }
  }
  if (capturing) return -1;
  if (g_bt.size() >= BT_MAX_ENTRIES) {  // evict the least recently used entry
    size_t lru = 0;
    for (size_t i = 1; i < g_bt.size(); ++i)
      if (g_bt[i].stamp < g_bt[lru].stamp) lru = i;
    cudaEventSynchronize(g_bt[lru].ready);
    cudaDeviceSynchronize();  // (rare) nothing in flight may still read the tables
    cudaFree(g_bt[lru].wd);
    cudaFree(g_bt[lru].bi);
    cudaEventDestroy(g_bt[lru].ready);
    g_bt.erase(g_bt.begin() + (long)lru);
  }
  BlockTabEntry e;
  e.dev = dev; e.n_out = n_out; e.eqs = eqs; e.linear = linear; e.npre = npre; e.shift = shift; e.even = even; e.winmax = winmax;
  e.n_coef = n_coef; e.x1 = x1; e.h = h; e.tk = tk; e.td = td; e.te = te;
  e.knots.assign(knots_out, knots_out + n_out);
  e.wd = nullptr; e.bi = nullptr; e.stream = st; e.stamp = ++g_bt_clock;
  const int nblk = (n_out + RS_R - 1) / RS_R;
  double *d_ko = nullptr;
  if (cudaMalloc((void **)&e.wd, sizeof(double) * (size_t)nblk * winmax * RS_R) != cudaSuccess ||
      cudaMalloc((void **)&e.bi, sizeof(int2) * (size_t)nblk) != cudaSuccess ||
      cudaMallocAsync((void **)&d_ko, sizeof(double) * (size_t)n_out, st) != cudaSuccess ||
      cudaEventCreateWithFlags(&e.ready, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    if (e.wd) cudaFree(e.wd);
    if (e.bi) cudaFree(e.bi);
    if (d_ko) cudaFreeAsync(d_ko, st);
    return -1;
  }
  // (e.knots is the library's own copy: the asynchronous upload may outlive the caller's array)
  if (cudaMemcpyAsync(d_ko, e.knots.data(), sizeof(double) * (size_t)n_out, cudaMemcpyHostToDevice, st) != cudaSuccess) {
    cudaFree(e.wd);
    cudaFree(e.bi);
    cudaFreeAsync(d_ko, st);
    cudaEventDestroy(e.ready);
    return -1;
  }
  resample_blocks_kernel<<<nblk, 256, 0, st>>>(d_ko, n_out, x1, h, eqs, n_coef, tk, td, te, linear, npre, shift, even, winmax, e.wd, e.bi);
  CIB_LAUNCHED();
  cudaEventRecord(e.ready, st);
  cudaFreeAsync(d_ko, st);
  *wd_out = e.wd;
  *bi_out = e.bi;
  g_bt.push_back(std::move(e));
  // the vector may have reallocated: the upload reads the moved-from buffer's storage, which std::vector's move keeps alive
  return 0;
}

// The whole call: -1 when any axis does not qualify (the caller then runs the generic path for the whole call).
// tk / td / te[d]: device tables K^(d), K^(d+1) (K again for a linear axis), K^(d+2) (quintic axes, else nullptr).
int resample_pipe_all(int ndim, const long long *n_in, const long long *n_out, const double *d_in, double *d_out,
                      const double *const *knots_in, const double *const *knots_out, const int *kernel, const int *eqs_a,
                      const double *const *tk, const double *const *td, const double *const *te, const int *linear,
                      const int *npre, const double *scale, const int *bcl, const int *bcr, cudaStream_t st) {
  if (ndim < 1 || ndim > CIB200_MAX_SWEEP_DIM || tune_env("CIB200_RESAMPLE_OLD")) return -1;
  std::vector<RsPass> passes((size_t)ndim);
  long long shape[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) shape[d] = n_in[d];
  // plan with placeholder addresses first (qualification does not depend on them beyond alignment), then allocate
  size_t inter_bytes[CIB200_MAX_SWEEP_DIM];
  for (int d = 0; d < ndim; ++d) {
    long long tot = 1;
    for (int j = 0; j < ndim; ++j) tot *= (j <= d ? n_out[j] : n_in[j]);
    inter_bytes[d] = sizeof(double) * (size_t)tot;
  }
  std::vector<void *> allocs;
  auto cleanup = [&]() {
    for (void *p : allocs) cudaFreeAsync(p, st);
  };
  auto dalloc = [&](void **p, size_t bytes) {
    if (cudaMallocAsync(p, bytes, st) != cudaSuccess) return false;
    allocs.push_back(*p);
    return true;
  };
  const double *cur = d_in;
  for (int d = 0; d < ndim; ++d) {
    long long inner = 1, outer = 1;
    for (int j = 0; j < d; ++j) inner *= shape[j];
    for (int j = d + 1; j < ndim; ++j) outer *= shape[j];
    const double h = knots_in[d][1] - knots_in[d][0];                          // convolution_resample.jl:270
    const double x1 = knots_in[d][0] - (double)(eqs_a[d] - 1) * h;             // :285
    double *next = d_out;
    if (d < ndim - 1) {
      void *p = nullptr;
      if (!dalloc(&p, inter_bytes[d])) {
        cleanup();
        return -1;
      }
      next = (double *)p;
    }
    if (!plan_pass(passes[(size_t)d], cur, next, inner, shape[d], n_out[d], outer, kernel[d], eqs_a[d], bcl[d], bcr[d], h, x1,
                   scale[d], knots_out[d])) {
      cleanup();
      return -1;
    }
    cur = next;
    shape[d] = n_out[d];
  }
  // block tables of every axis first (nothing but the passes themselves then sits between two passes)
  for (int d = 0; d < ndim; ++d) {
    RsPass &ps = passes[(size_t)d];
    ResamplePipeParams &sp = ps.sp;
    const double h = sp.h;
    const double x1 = knots_in[d][0] - (double)(ps.eqs - 1) * h;
    {
      const double *c_wd = nullptr;
      const int2 *c_bi = nullptr;
      if (!tune_env("CIB200_RS_NO_TABLE_CACHE") &&
          block_tables_cached(knots_out[d], sp.n_out, x1, h, ps.eqs, (long long)sp.n_in + 2 * sp.pad, tk[d], td[d], te[d], linear[d],
                              npre[d], sp.padx - sp.pad, ps.bi == 0 ? 1 : 0, sp.winmax, st, &c_wd, &c_bi) == 0) {
        sp.Wd = c_wd;
        sp.binfo = c_bi;
        continue;
      }
    }
    double *d_ko = nullptr, *d_wd = nullptr;
    int2 *d_bi = nullptr;
    if (!dalloc((void **)&d_ko, sizeof(double) * (size_t)sp.n_out) ||
        !dalloc((void **)&d_wd, sizeof(double) * (size_t)sp.nblk * sp.winmax * RS_R) ||
        !dalloc((void **)&d_bi, sizeof(int2) * (size_t)sp.nblk)) {
      cleanup();
      return -1;
    }
    if (cudaMemcpyAsync(d_ko, knots_out[d], sizeof(double) * (size_t)sp.n_out, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      cleanup();
      set_error("resample: knot upload failed");
      return CIB200_ECUDA;
    }
    resample_blocks_kernel<<<sp.nblk, 256, 0, st>>>(d_ko, sp.n_out, x1, h, ps.eqs, (long long)sp.n_in + 2 * sp.pad, tk[d], td[d],
                                                    te[d], linear[d], npre[d], sp.padx - sp.pad, ps.bi == 0 ? 1 : 0, sp.winmax, d_wd, d_bi);
    CIB_LAUNCHED();
    sp.Wd = d_wd;
    sp.binfo = d_bi;
  }
  int rc = CIB200_OK;
  for (int d = 0; d < ndim && rc == CIB200_OK; ++d) rc = launch_rs_eqs(passes[(size_t)d], st);
  if (rc == CIB200_OK && cudaGetLastError() != cudaSuccess) {
    set_error("resample: launch failed");
    rc = CIB200_ECUDA;
  }
  cleanup();
  return rc < 0 ? CIB200_ECUDA : rc;
}

}  // namespace cib
