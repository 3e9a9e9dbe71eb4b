// cib200_lanequery.cu - pass C of the tile-binned strategy, "lane per query" variant (round 2).
//
// eval_cellrun_kernel (cib200_binned.cu) spreads ONE query's (2E)^3 stencil over the 32 lanes of a warp and keeps the
// cell's coefficients in registers; every lane then needs 15 per-query weights from shared memory and the warp a
// reduce-scatter per 8 queries: 54 shared-memory wavefronts per query, the bound of that kernel (profiles/r02_eval512).
//
// Here the roles are swapped.  The records arrive sorted by (tile, cell) (pass B2), and
//   * a LANE owns one query - or, QPL = 2, two queries of the SAME cell (the slot list of pass B2) - with all of its
//     3 x 2E blended weights in registers;
//   * the coefficients are read straight from the TMA-staged brick.  Neighbouring lanes sit in the same or in
//     neighbouring cells, so one LDS.64 of the warp touches a handful of distinct words: the hardware multicasts them
//     in a single wavefront.  With two queries per lane a coefficient load feeds two DFMAs.
// Per query: (2E)^3 / 32 / QPL wavefronts for the stencil (15.6 for :b5, QPL = 2) plus the lane's own table rows
// (random rows: the part that still conflicts), no scratch, no warp synchronisation, no shuffles, and no idle lanes in
// the contraction (1110 DFMAs per query instead of 48 x 32).
//
// Arithmetic: positions and weights exactly as in the other strategies (locate_from_knot, the four-term blended weight);
// the stencil sum is associated as  sum_n wz[n] (sum_m wy[m] (sum_l c[l,m,n] wx[l])),  each a plain left-to-right fma
// chain - the association of the scalar evaluator core_thread.  A query's result does not depend on the batch it arrives
// in, nor on which of the two places of a slot it takes.
#include "cib200_binned.cuh"

namespace cib {

// blended weights of one axis: table rows r, r + 1 of K and K' fetched as 16-byte loads (rows are 2E doubles: 16-byte
// multiples), combined with the expression of blended_weight (cib200_device.cuh)
template <int W>
__device__ __forceinline__ void lq_axis_weights(const double *__restrict__ kr, const double *__restrict__ dr,
                                                const Hermite &hb, double *w) {
  const double2 *k = (const double2 *)kr;
  const double2 *d = (const double2 *)dr;
#pragma unroll
  for (int c = 0; c < W / 2; ++c) {
    const double2 k0 = k[c], k1 = k[W / 2 + c], d0 = d[c], d1 = d[W / 2 + c];
    w[2 * c] = hb.h00 * k0.x + hb.h01 * k1.x + (hb.h10 * d0.x + hb.h11 * d1.x);
    w[2 * c + 1] = hb.h00 * k0.y + hb.h01 * k1.y + (hb.h10 * d0.y + hb.h11 * d1.y);
  }
}

template <int E, int QPL, int NT, bool ZLAZY, typename TC>
__global__ void __launch_bounds__(NT, 1)
    eval_lanequery_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                          const Record *__restrict__ rec, const unsigned *__restrict__ slots,
                          const unsigned *__restrict__ tile_start, const unsigned *__restrict__ tile_nslots,
                          const WorkItem *__restrict__ items, unsigned *__restrict__ n_items, TC *__restrict__ out,
                          int use_tma) {
  using BG = BrickGeom<TC>;
  constexpr int W = 2 * E;
  constexpr int PITCH = BG::PITCH, PLANE = BG::PLANE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TC *brick = (TC *)smem_raw;        // [BRICK_SIDE][BRICK_SIDE][PITCH], axis 0 contiguous
  double *smd = (double *)smem_raw;  // the same memory in doubles
  double *tab = smd + BG::DOUBLES;
  const double *tk[3];
  const double *td[3];
  stage_tables(p, tab, tk, td);
  const int nwarps = NT >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tbase = BG::DOUBLES + ((tables_smem_doubles(p) + 1) & ~1);
  unsigned long long *bar = (unsigned long long *)(smd + tbase);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]
  double *kt = (double *)(bar + 2);        // [3][32]: the stored knots the tile's queries can select
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double k0[3] = {p.knot0[0], p.knot0[1], p.knot0[2]};
  // every table of the interpolant is staged (the launcher checks): offsets from `smd` keep the loads in the shared window
  const int tko[3] = {(int)(tk[0] - smd), (int)(tk[1] - smd), (int)(tk[2] - smd)};
  const int tdo[3] = {(int)(td[0] - smd), (int)(td[1] - smd), (int)(td[2] - smd)};
  const bool lin[3] = {p.subgrid[0] == CIB200_SG_LINEAR, p.subgrid[1] == CIB200_SG_LINEAR, p.subgrid[2] == CIB200_SG_LINEAR};
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
    // work units of the item: records (QPL = 1), or slots = pairs of records of one cell (QPL = 2); item j of a tile
    // covers the tile's units [j CHUNK, (j + 1) CHUNK) - a tile has at most as many slots as records
    const unsigned r0 = tile_start[it.bin];
    unsigned first = it.start - r0, cnt = it.count;
    if (QPL == 2) {
      const unsigned ns = tile_nslots[it.bin];
      first = min(first, ns);
      cnt = min((unsigned)CHUNK_QUERIES, ns - first);
    }
    if (cnt == 0) continue;  // block-uniform
    const int tx = it.bin % g.nt[0], ty = (it.bin / g.nt[0]) % g.nt[1], tz = it.bin / (g.nt[0] * g.nt[1]);
    const int org[3] = {tx * g.T, ty * g.T, tz * g.T};
    if (threadIdx.x < 96) {
      const int d = threadIdx.x >> 5, r = threadIdx.x & 31;
      const int o = d == 0 ? org[0] : (d == 1 ? org[1] : org[2]);
      kt[threadIdx.x] = __ldg(p.knots[d] + min(o + r, p.nk[d] - 1));
    }
    if (use_tma) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, BG::ELEMS * (int)sizeof(TC));
        tma_load_3d(brick, &tmap, org[0], org[1], org[2], bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
      __syncthreads();  // kt
    } else {
      const TC *__restrict__ cg = (const TC *)p.coefs;
      for (int e = threadIdx.x; e < BG::ELEMS; e += NT) {
        const int bx = e % PITCH, by = (e / PITCH) % BRICK_SIDE, bz = e / PLANE;
        const int gx = org[0] + bx, gy = org[1] + by, gz = org[2] + bz;
        brick[e] = (gx < p.nk[0] && gy < p.nk[1] && gz < p.nk[2]) ? __ldg(cg + gx + (long long)gy * p.s1 + (long long)gz * p.s2) : (TC)0;
      }
      __syncthreads();
    }
    // contiguous range of this warp, in whole batches of 32 units
    const unsigned per = ((cnt + nwarps - 1) / nwarps + 31u) & ~31u;
    const unsigned w0 = warp * per, w1 = min(cnt, w0 + per);
    const unsigned *sl = slots + r0 + first;
    const Record *rc = rec + r0 + (QPL == 2 ? 0u : first);
    // unit of the NEXT batch: its slot entry is fetched one batch ahead, its records are pulled towards L2 / L1
    unsigned e_n = 0;
    if (w0 < w1) {
      const unsigned u = min(w0 + lane, w1 - 1);  // lanes past the end repeat the last unit (result discarded)
      e_n = QPL == 2 ? __ldg(sl + u) : u;
    }
    for (unsigned b0 = w0; b0 < w1; b0 += 32) {
      const bool have = b0 + lane < w1;
      const unsigned e = e_n;
      if (b0 + 32 < w1) {
        const unsigned u = min(b0 + 32 + lane, w1 - 1);
        e_n = QPL == 2 ? __ldg(sl + u) : u;
      }
      const unsigned ia = QPL == 2 ? (e >> 1) : e;
      const bool has2 = QPL == 2 && (e & 1u);
      double2 qa[QPL], qb[QPL];
      ld_record(rc + ia, qa[0], qb[0]);
      if (QPL == 2) ld_record(rc + ia + (has2 ? 1u : 0u), qa[QPL - 1], qb[QPL - 1]);
      if (b0 + 32 < w1) {
        const unsigned ian = QPL == 2 ? (e_n >> 1) : e_n;
        prefetch_l2(rc + ian);
        if (QPL == 2) prefetch_l2(rc + ian + 1);
      }
      double wx[QPL][W], wy[QPL][W], wz[ZLAZY ? 1 : QPL][ZLAZY ? 1 : W];
      Hermite hz[QPL];
      int rowz[QPL];
      unsigned idx[QPL];
      int m = 0;
#pragma unroll
      for (int s = 0; s < QPL; ++s) {
        const double x[3] = {qa[s].x, qa[s].y, qb[s].x};
        idx[s] = (unsigned)__double2loint(qb[s].y);
        AxisPos pa[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
          locate_from_knot(x[d], k0[d], p.h[d], p.eqs[d], p.nk[d], kt + 32 * d - org[d], pa[d]);
        if (s == 0) m = (pa[0].base - org[0]) + (pa[1].base - org[1]) * PITCH + (pa[2].base - org[2]) * PLANE;
        {
          const Hermite hb = lin[0] ? linear_basis(pa[0].t) : hermite_basis(pa[0].t);
          lq_axis_weights<W>(smd + tko[0] + pa[0].row * W, smd + tdo[0] + pa[0].row * W, hb, wx[s]);
        }
        {
          const Hermite hb = lin[1] ? linear_basis(pa[1].t) : hermite_basis(pa[1].t);
          lq_axis_weights<W>(smd + tko[1] + pa[1].row * W, smd + tdo[1] + pa[1].row * W, hb, wy[s]);
        }
        hz[s] = lin[2] ? linear_basis(pa[2].t) : hermite_basis(pa[2].t);
        rowz[s] = pa[2].row * W;
        if (!ZLAZY) lq_axis_weights<W>(smd + tko[2] + rowz[s], smd + tdo[2] + rowz[s], hz[s], wz[ZLAZY ? 0 : s]);
      }
      const TC *cb = brick + m;
      double r[QPL];
#pragma unroll
      for (int s = 0; s < QPL; ++s) r[s] = 0.0;
#pragma unroll
      for (int n = 0; n < W; ++n) {
        double u[QPL];
#pragma unroll
        for (int s = 0; s < QPL; ++s) u[s] = 0.0;
#pragma unroll
        for (int mm = 0; mm < W; ++mm) {
          const TC *row = cb + n * PLANE + mm * PITCH;
          double c[W];
#pragma unroll
          for (int l = 0; l < W; ++l) c[l] = (double)row[l];
#pragma unroll
          for (int s = 0; s < QPL; ++s) {
            double t = c[0] * wx[s][0];
#pragma unroll
            for (int l = 1; l < W; ++l) t = fma(c[l], wx[s][l], t);
            u[s] = fma(t, wy[s][mm], u[s]);
          }
        }
#pragma unroll
        for (int s = 0; s < QPL; ++s) {
          const double wzn = ZLAZY ? blended_weight(hz[s], smd + tko[2], smd + tdo[2], 0, rowz[s] + n, W) : wz[ZLAZY ? 0 : s][ZLAZY ? 0 : n];
          r[s] = fma(u[s], wzn, r[s]);
        }
      }
      if (have) {
        out[idx[0]] = (TC)apply_scale<3>(r[0], sc);
        if (QPL == 2 && has2) out[idx[QPL - 1]] = (TC)apply_scale<3>(r[QPL - 1], sc);
      }
    }
  }
}

size_t lanequery_smem_bytes(const ItpParams &p, bool f32) {
  const size_t brick = f32 ? BrickGeom<float>::BYTES : BrickGeom<double>::BYTES;
  return brick + 8 * ((size_t)((tables_smem_doubles(p) + 1) & ~1) + 2 + 96);
}

template <int E, int QPL, int NT, bool ZLAZY, typename TC>
static int launch_lq(cib200_itp *h, BinState *bs, const Scale &sc, TC *out, cudaStream_t st) {
  const ItpParams &p = h->p;
  const size_t smem = lanequery_smem_bytes(p, sizeof(TC) == 4);
  auto kern = eval_lanequery_kernel<E, QPL, NT, ZLAZY, TC>;
  // (the attribute is per function AND device: set on every launch, a host-side call of about a microsecond)
  CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<bs->nblk, NT, smem, st>>>(bs->tmap, p, bs->g, sc, bs->rec2, bs->slots, bs->tile_start, bs->tile_nslots, bs->items,
                                   bs->n_items, out, 1);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

// variant: 0 = default of the build; experiments builds select (threads, lazy z weights) through CIB200_LQ_VARIANT
int launch_eval_lanequery(cib200_itp *h, BinState *bs, const Scale &sc, void *out, int qpl, cudaStream_t st) {
  int variant = 0;
  if (const char *e = tune_env("CIB200_LQ_VARIANT")) variant = atoi(e);
  double *o = (double *)out;
  if (qpl == 2) {
    switch (variant) {
      case 1: return launch_lq<5, 2, 384, true, double>(h, bs, sc, o, st);
      case 2: return launch_lq<5, 2, 256, true, double>(h, bs, sc, o, st);
      default: return launch_lq<5, 2, 256, false, double>(h, bs, sc, o, st);
    }
  }
  switch (variant) {
    case 1: return launch_lq<5, 1, 384, true, double>(h, bs, sc, o, st);
    case 2: return launch_lq<5, 1, 512, true, double>(h, bs, sc, o, st);
    default: return launch_lq<5, 1, 384, false, double>(h, bs, sc, o, st);
  }
}

}  // namespace cib
