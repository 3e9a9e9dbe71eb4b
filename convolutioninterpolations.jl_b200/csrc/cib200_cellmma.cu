// cib200_cellmma.cu - pass C of the tile-binned strategy on the FP64 tensor cores (round 2).
//
// What bounds the cell-sorted kernels is not arithmetic but operand delivery: the shared-memory return path gives a
// warp 128 bytes per clock whatever the access pattern (a multicast LDS.64 still takes one wavefront per half-warp:
// measured, profiles/r02_lanequery.txt), so every scheme that feeds DFMAs from shared memory pays ~8 bytes per one or
// two FMAs - eval_cellrun_kernel 54 wavefronts per query (per-query weights to every lane), the lane-per-query kernel
// 53 (the coefficients to every lane).
//
// For queries that share a cell the axis-0 contraction is a small dense matrix product,
//     U[q, (m, n)] = sum_l  wx_q[l] * c[l, m, n]          (queries x taps) x (taps x 100 columns),
// and B200's FP64 tensor-core instruction (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4: the full 64 FMA/clk/SM of the FP64
// pipe, tools/micro/dmma_rate.cu) takes ONE double per lane for each operand and does the broadcasting inside the
// tensor core: 8 FMAs per lane for 16 operand bytes instead of 1.  Two adjacent cells along axis 0 share an 11-wide
// window of every coefficient row, so the rows of one product are up to 8 queries of an aligned PAIR of cells (pass B2
// emits the group list), K = 12 (3 DMMA steps), N = 100 columns in 13 tiles: 39 DMMAs per group.  The remaining two
// contractions are per-query dot products on the accumulator fragments (lane (q, c) holds two columns of every tile:
// 39 DFMAs per group) and one quad reduction.
//
// Column layout of the product (chosen so that the epilogue needs few distinct weights per lane):
//   tile t = 0..9 :  column n = 0..7  ->  (m = n, n_z = t)
//   tile 10 + u   :  column n         ->  (m = 8 + (n >> 2), n_z = 4 u + (n & 3)),  u = 0, 1, 2 (n_z >= 10: zero)
// (with brick rows of 28 doubles and planes of 700 every B-fragment load of a half-warp is bank-conflict-free)
//
// Arithmetic: positions and blended weights as everywhere else; the association of the stencil sum is the tensor
// core's (4-term dot products accumulated over the three K steps) followed by the epilogue chains.  A query's K
// position depends only on the parity of its cell index, so its result does not depend on the batch it arrives in.
#include "cib200_binned.cuh"

namespace cib {

static constexpr int MP = CELLMMA_PITCH, MPL = MP * BRICK_SIDE, MELEMS = MPL * BRICK_SIDE;  // 28 x 25 x 25 doubles
static constexpr int WQS = 36;  // doubles per query in the weight scratch: WX[14] (2 zeros, 10 weights, 2 zeros) | wy[10] | wz[10] 0 0
static constexpr int PTAB_ROWS = CIB_TAB_ROWS - 1, PTAB_ROW = 40, PTAB_DOUBLES = PTAB_ROWS * PTAB_ROW;  // per distinct (K, K') pair

struct MmaWarpSmem {
  static constexpr int TQ = 32 * 3;    // t of (record, axis)
  static constexpr int MISC = 32 * 2;  // int4 per record: {2 * brick offset + cell parity, table rows 10 bits each, idx, -}
  static constexpr int WQ = 8 * WQS;
  static constexpr int DOUBLES = TQ + MISC + WQ;
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// axes with the same (K, K') tables share one staged polynomial table
__host__ __device__ inline int mma_ptab_rep(const ItpParams &p, int d) {  // first axis with the tables of axis d
  for (int e = 0; e < d; ++e)
    if (p.tab_slot_k[e] == p.tab_slot_k[d] && p.tab_slot_d[e] == p.tab_slot_d[d] && p.subgrid[e] == p.subgrid[d]) return e;
  return d;
}
__host__ __device__ inline int mma_ptab_index(const ItpParams &p, int d) {
  const int r = mma_ptab_rep(p, d);
  int n = 0;
  for (int e = 0; e < r; ++e)
    if (mma_ptab_rep(p, e) == e) ++n;
  return n;
}
__host__ __device__ inline int mma_ptab_count(const ItpParams &p) {
  int n = 0;
  for (int d = 0; d < 3; ++d) n = max(n, mma_ptab_index(p, d) + 1);
  return n;
}

// The blended weight  K0 h00(t) + K1 h01(t) + h_pre (D0 h10(t) + D1 h11(t))  of a table row is a cubic in t; the kernel
// stages its four coefficients per (row, tap) - c0 = K0, c1 = h D0, c2 = 3 (K1 - K0) - 2 h D0 - h D1,
// c3 = 2 (K0 - K1) + h D0 + h D1 (linear subgrid: c1 = K1 - K0) - so a weight costs three FMAs and no basis evaluation.
__device__ __forceinline__ void stage_poly_tables(const ItpParams &p, double *ptab, int nthreads) {
  const double h_pre = 1.0 / (double)(CIB_N_PRE - 1);
  for (int d = 0; d < 3; ++d) {
    if (mma_ptab_rep(p, d) != d) continue;
    const int pi = mma_ptab_index(p, d);
    const double *__restrict__ K = p.tabk[d], *__restrict__ D = p.tabd[d];
    const bool lin = p.subgrid[d] == CIB200_SG_LINEAR;
    double *dst = ptab + pi * PTAB_DOUBLES;
    const int W = 2 * p.eqs[d];
    for (int i = threadIdx.x; i < PTAB_ROWS * W; i += nthreads) {
      const int r = i / W, c = i - r * W;
      const double k0 = K[r * W + c], k1 = K[(r + 1) * W + c];
      double c1, c2, c3;
      if (lin) {
        c1 = k1 - k0;
        c2 = 0.0;
        c3 = 0.0;
      } else {
        const double d0 = h_pre * D[r * W + c], d1 = h_pre * D[(r + 1) * W + c];
        c1 = d0;
        c2 = 3.0 * (k1 - k0) - 2.0 * d0 - d1;
        c3 = 2.0 * (k0 - k1) + d0 + d1;
      }
      dst[r * PTAB_ROW + c] = k0;
      dst[r * PTAB_ROW + 10 + c] = c1;
      dst[r * PTAB_ROW + 20 + c] = c2;
      dst[r * PTAB_ROW + 30 + c] = c3;
    }
  }
}

// TC = float: the opt-in Float32 mode - the brick is staged as floats (same 28-element rows: the B-fragment loads stay
// conflict-free) and converted when a fragment is loaded; positions, weights and accumulation are Float64 as everywhere
template <int NT, bool BREG, typename TC>
__global__ void __launch_bounds__(NT, 1)
    eval_cellmma_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                        const Record *__restrict__ rec, const unsigned *__restrict__ perm, const uint2 *__restrict__ groups,
                        const unsigned *__restrict__ tile_start, const unsigned *__restrict__ tile_ngroups,
                        const WorkItem *__restrict__ items, unsigned *__restrict__ n_items, TC *__restrict__ out) {
  using WS = MmaWarpSmem;
  constexpr int BRICKD = MELEMS * (int)sizeof(TC) / 8;  // doubles the brick occupies
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *smd = (double *)smem_raw;
  TC *brick = (TC *)smem_raw;  // [BRICK_SIDE][BRICK_SIDE][MP], axis 0 contiguous
  double *ptab = smd + BRICKD;
  stage_poly_tables(p, ptab, NT);
  constexpr int nwarps = NT >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbase = BRICKD + mma_ptab_count(p) * PTAB_DOUBLES;
  double *wsm = smd + wbase + warp * WS::DOUBLES;
  double *tq = wsm;                      // [32][3]
  int4 *misc = (int4 *)(wsm + WS::TQ);   // [32]
  double *wq = wsm + WS::TQ + WS::MISC;  // [8][WQS]
  unsigned long long *bar = (unsigned long long *)(smd + wbase + nwarps * WS::DOUBLES);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < WS::WQ; i += 32) wq[i] = 0.0;  // the zero pads of WX / wz are never written again
  __syncthreads();
  // weight role of this lane: quarter-warp = one (query, axis) pair per pass, lane = pair of taps (lanes 5..7 idle)
  const int wquarter = lane >> 3;
  const bool wrole = (lane & 7) < 5;
  const int wpp = wrole ? (lane & 7) : 4;
  const int wtab0 = BRICKD + mma_ptab_index(p, 0) * PTAB_DOUBLES + 2 * wpp;  // from smd
  const int wtab1 = BRICKD + mma_ptab_index(p, 1) * PTAB_DOUBLES + 2 * wpp;
  const int wtab2 = BRICKD + mma_ptab_index(p, 2) * PTAB_DOUBLES + 2 * wpp;
  // MMA roles: A row / accumulator row q, K index / column pair r
  const int mq = lane >> 2, mr = lane & 3;
  // B-fragment offsets of this lane inside the window: tiles 0..9 and tiles 10..12
  const int boff = mr + mq * MP;
  const int boff2 = mr + (8 + (mq >> 2)) * MP + (mq & 3) * MPL;
  const bool bk3 = mr == 3;   // K index 11 (step 2): outside the 11-wide window
  const bool bt12 = (mq & 3) >= 2;  // tile 12: columns with n_z >= 10
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
    // item j of a tile covers the tile's groups [j CHUNK, (j + 1) CHUNK): a tile has at most as many groups as records
    const unsigned r0 = tile_start[it.bin];
    const unsigned ng = tile_ngroups[it.bin];
    const unsigned first = min(it.start - r0, ng);
    const unsigned cnt = min((unsigned)CHUNK_QUERIES, ng - first);
    if (cnt == 0) continue;  // block-uniform
    const int tx = it.bin % g.nt[0], ty = (it.bin / g.nt[0]) % g.nt[1], tz = it.bin / (g.nt[0] * g.nt[1]);
    const int org[3] = {tx * g.T, ty * g.T, tz * g.T};
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, MELEMS * (int)sizeof(TC));
      tma_load_3d(brick, &tmap, org[0], org[1], org[2], bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    // contiguous range of this warp, in batches of 4 groups (32 records: lane = (group, row))
    const unsigned per = ((cnt + nwarps - 1) / nwarps + 3u) & ~3u;
    const unsigned w0 = warp * per, w1 = min(cnt, w0 + per);
    const uint2 *gl = groups + r0 + first;
    const Record *rc = rec + r0;
    const int pg = lane >> 3, prow = lane & 7;
    // group entries are fetched two batches ahead, the permutation entry of a record one batch ahead, and the record itself
    // is pulled into L2 while the previous batch is being evaluated: three dependent loads, none of them waited for
    uint2 ge_n = make_uint2(0u, 0u), ge_nn = make_uint2(0u, 0u);
    if (w0 + pg < w1) ge_n = __ldg(gl + w0 + pg);
    if (w0 + 4 + pg < w1) ge_nn = __ldg(gl + w0 + 4 + pg);
    unsigned pv_n = 0;
    if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
    int curwin = -1;
    double breg[39];
#pragma unroll
    for (int i = 0; i < 39; ++i) breg[i] = 0.0;
    for (unsigned b0 = w0; b0 < w1; b0 += 4) {
      const int ngb = (int)min(4u, w1 - b0);
      const uint2 ge = ge_n;
      const unsigned pv = pv_n;
      ge_n = ge_nn;
      ge_nn = make_uint2(0u, 0u);
      if (b0 + 8 + pg < w1) ge_nn = __ldg(gl + b0 + 8 + pg);
      pv_n = 0;
      if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
      // ---- phase 1: one record per lane: positions into the warp's scratch ----------------------------------------
      __syncwarp();  // the previous batch is done with tq / misc
      if ((unsigned)prow < ge.y) {
        // position record (bin_scatter_kernel<double, true>): {t_0, t_1, t_2, index, rows << 12 | cell in tile}
        double2 ra, rb;
        // (perm: the cell sort left the records where the scatter pass put them and wrote the sorted order as indices)
        ld_record(rc + (perm ? pv : ge.x + prow), ra, rb);
        tq[lane * 3 + 0] = ra.x;
        tq[lane * 3 + 1] = ra.y;
        tq[lane * 3 + 2] = rb.x;
        const unsigned code = (unsigned)__double2hiint(rb.y);
        const unsigned lc = code & 4095u, rcode = code >> 12;
        const unsigned rz = rcode / 10000u, ry = (rcode - rz * 10000u) / 100u, rx = rcode - rz * 10000u - ry * 100u;
        int bx, by, bz;
        lcell_decode(lc, 16u, bx, by, bz);  // T = 16 cells per tile side (2E = 10)
        const int m = bx + by * MP + bz * MPL;
        misc[lane] = make_int4(2 * m + (bx & 1), (int)(rx | (ry << 10) | (rz << 20)), __double2loint(rb.y), 0);
      }
      if (!perm && ge_n.y) prefetch_l2(rc + ge_n.x + min((unsigned)prow, ge_n.y - 1));
      __syncwarp();
      for (int gi = 0; gi < ngb; ++gi) {
        const int cntg = (int)__shfl_sync(0xffffffffu, ge.y, gi * 8);
        const int sb = gi * 8;  // first scratch slot of the group
        // ---- weights of the group's queries, two queries per pass: Horner on the staged cubics ----------------------
        // one (query, axis) pair per QUARTER-warp and pass, lane = pair of taps (5 of 8 lanes): the four 16-byte loads of
        // a quarter read one contiguous 80-byte run of a table row each - conflict-free whatever rows the other quarters
        // need.  All loads of two passes are issued before their first store (the compiler cannot prove that the wq
        // stores do not alias the scratch it reads, and would serialise the passes).
        const int ncombo = 3 * cntg;
#pragma unroll 1
        for (int c8 = 0; c8 < ncombo; c8 += 8) {
          double tt[2];
          double2 c0[2], c1[2], c2[2], c3[2];
          int dst[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ci = min(c8 + 4 * u + wquarter, ncombo - 1);  // past the end: recompute the last pair, not stored
            const int qq = (ci * 11) >> 5, ax = ci - 3 * qq;        // ci / 3, ci % 3 (ci < 24)
            tt[u] = tq[(sb + qq) * 3 + ax];
            const int row = (misc[sb + qq].y >> (10 * ax)) & 1023;
            const double2 *pr = (const double2 *)(smd + (ax == 0 ? wtab0 : (ax == 1 ? wtab1 : wtab2)) + row * PTAB_ROW);
            c0[u] = pr[0];
            c1[u] = pr[5];
            c2[u] = pr[10];
            c3[u] = pr[15];
            dst[u] = qq * WQS + (ax == 0 ? 2 : (ax == 1 ? 14 : 24)) + 2 * wpp;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            double2 wv;
            wv.x = fma(tt[u], fma(tt[u], fma(tt[u], c3[u].x, c2[u].x), c1[u].x), c0[u].x);
            wv.y = fma(tt[u], fma(tt[u], fma(tt[u], c3[u].y, c2[u].y), c1[u].y), c0[u].y);
            if (wrole && c8 + 4 * u + wquarter < ncombo) *(double2 *)(wq + dst[u]) = wv;
          }
        }
        __syncwarp();
        // ---- the axis-0 contraction on the tensor cores ------------------------------------------------------------
        const int mf = misc[sb].x;
        const int win = (mf >> 1) - (mf & 1);  // brick offset of the pair's first cell
        if (BREG && win != curwin) {  // warp-uniform: a new cell pair - fetch its B fragments
          curwin = win;
          const TC *bb = brick + win + boff, *bc = brick + win + boff2;
#pragma unroll
          for (int t = 0; t < 10; ++t) {
            breg[t * 3 + 0] = (double)bb[t * MPL];
            breg[t * 3 + 1] = (double)bb[t * MPL + 4];
            breg[t * 3 + 2] = bk3 ? 0.0 : (double)bb[t * MPL + 8];
          }
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const bool z = u == 2 && bt12;
            breg[(10 + u) * 3 + 0] = z ? 0.0 : (double)bc[4 * u * MPL];
            breg[(10 + u) * 3 + 1] = z ? 0.0 : (double)bc[4 * u * MPL + 4];
            breg[(10 + u) * 3 + 2] = (z || bk3) ? 0.0 : (double)bc[4 * u * MPL + 8];
          }
        }
        const bool qv = mq < cntg;
        const int4 mi = misc[sb + mq];
        const double *wrow = wq + mq * WQS;
        double a[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) a[s] = qv ? wrow[2 - (mi.x & 1) + 4 * s + mr] : 0.0;
        double acc[13][2];
#pragma unroll
        for (int t = 0; t < 13; ++t) {
          acc[t][0] = 0.0;
          acc[t][1] = 0.0;
        }
        if (BREG) {
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int t = 0; t < 13; ++t) dmma884(acc[t][0], acc[t][1], a[s], breg[t * 3 + s]);
        } else {  // B fragments straight from the brick, one LDS.64 per DMMA
          const TC *bb = brick + win + boff, *bc = brick + win + boff2;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int t = 0; t < 10; ++t) dmma884(acc[t][0], acc[t][1], a[s], (s == 2 && bk3) ? 0.0 : (double)bb[t * MPL + 4 * s]);
#pragma unroll
            for (int u = 0; u < 3; ++u)
              dmma884(acc[10 + u][0], acc[10 + u][1], a[s], ((u == 2 && bt12) || (s == 2 && bk3)) ? 0.0 : (double)bc[4 * u * MPL + 4 * s]);
          }
        }
        // ---- epilogue: the other two contractions on the accumulator fragments, quad reduction -----------------------
        //   sum_t wz[t] (U[t][0] wy[2c] + U[t][1] wy[2c+1])  =  wy[2c] sum_t wz[t] U[t][0]  +  wy[2c+1] sum_t wz[t] U[t][1]
        const double2 wy01 = *(const double2 *)(wrow + 14 + 2 * mr);
        const double wy89 = wrow[22 + (mr >> 1)];
        double s0, s1, v;
        {
          const double2 wz = *(const double2 *)(wrow + 24);
          s0 = acc[0][0] * wz.x;
          s1 = acc[0][1] * wz.x;
          s0 = fma(acc[1][0], wz.y, s0);
          s1 = fma(acc[1][1], wz.y, s1);
        }
#pragma unroll
        for (int t = 2; t < 10; t += 2) {
          const double2 wz = *(const double2 *)(wrow + 24 + t);
          s0 = fma(acc[t][0], wz.x, s0);
          s1 = fma(acc[t][1], wz.x, s1);
          s0 = fma(acc[t + 1][0], wz.y, s0);
          s1 = fma(acc[t + 1][1], wz.y, s1);
        }
        {
          const double2 wz = *(const double2 *)(wrow + 24 + 2 * (mr & 1));
          v = fma(acc[10][1], wz.y, acc[10][0] * wz.x);
        }
#pragma unroll
        for (int u = 1; u < 3; ++u) {
          const double2 wz = *(const double2 *)(wrow + 24 + 4 * u + 2 * (mr & 1));
          v = fma(acc[10 + u][0], wz.x, v);
          v = fma(acc[10 + u][1], wz.y, v);
        }
        double e = fma(s0, wy01.x, s1 * wy01.y) + v * wy89;
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        if (qv && mr == 0) out[(unsigned)mi.z] = (TC)apply_scale<3>(e, sc);
        __syncwarp();  // wq is rewritten by the next group
      }
      if (perm && (unsigned)prow < ge_n.y) prefetch_l2(rc + pv_n);  // the next batch's record (its index arrived long ago)
    }
  }
}

// ---- sparse batches: rows of a product from a 2 x 2 x 2 BLOCK of cells ----------------------------------------------------
// Below ~4 queries per cell an aligned cell pair no longer fills the 8 rows of a DMMA product.  The eight cells of a block
// of the block-major numbering share an 11 x 11 x 11 window of the brick, so the rows of one product can come from all of
// them: K = the 11 axis-0 rows (3 steps, as before), columns = (m', n') in 0..10 x 0..10 - 121 columns in 17 tiles, 51
// DMMAs per group, two chunks of accumulators.  A query's own stencil starts at (dx, dy, dz) in {0, 1}^3 inside the window:
// the axis-0 shift is in its A fragments (as in the pair kernel), the other two in which weight multiplies which column in
// the epilogue (a column outside the query's 10 x 10 range gets weight zero).  Column layout:
//   tile t = 0..10  :  column c -> (m' = c, n' = t)
//   tile 11 + u     :  column c -> (m' = 8 + (c & 3), n' = 2 u + (c >> 2)),  u = 0..5  (m' = 11 and n' = 11: zero)
template <int NT, typename TC>
__global__ void __launch_bounds__(NT, 1)
    eval_celloct_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                        const Record *__restrict__ rec, const unsigned *__restrict__ perm, const uint2 *__restrict__ groups,
                        const unsigned *__restrict__ tile_start, const unsigned *__restrict__ tile_ngroups,
                        const WorkItem *__restrict__ items, unsigned *__restrict__ n_items, TC *__restrict__ out) {
  using WS = MmaWarpSmem;
  constexpr int BRICKD = MELEMS * (int)sizeof(TC) / 8;  // doubles the brick occupies
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *smd = (double *)smem_raw;
  TC *brick = (TC *)smem_raw;  // [BRICK_SIDE][BRICK_SIDE][MP], axis 0 contiguous
  double *ptab = smd + BRICKD;
  stage_poly_tables(p, ptab, NT);
  constexpr int nwarps = NT >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbase = BRICKD + mma_ptab_count(p) * PTAB_DOUBLES;
  double *wsm = smd + wbase + warp * WS::DOUBLES;
  double *tq = wsm;                      // [32][3]
  int4 *misc = (int4 *)(wsm + WS::TQ);   // [32]
  double *wq = wsm + WS::TQ + WS::MISC;  // [8][WQS]
  unsigned long long *bar = (unsigned long long *)(smd + wbase + nwarps * WS::DOUBLES);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < WS::WQ; i += 32) wq[i] = 0.0;  // the zero pads of WX / wz are never written again
  __syncthreads();
  // weight role of this lane: quarter-warp = one (query, axis) pair per pass, lane = pair of taps (lanes 5..7 idle)
  const int wquarter = lane >> 3;
  const bool wrole = (lane & 7) < 5;
  const int wpp = wrole ? (lane & 7) : 4;
  const int wtab0 = BRICKD + mma_ptab_index(p, 0) * PTAB_DOUBLES + 2 * wpp;  // from smd
  const int wtab1 = BRICKD + mma_ptab_index(p, 1) * PTAB_DOUBLES + 2 * wpp;
  const int wtab2 = BRICKD + mma_ptab_index(p, 2) * PTAB_DOUBLES + 2 * wpp;
  // MMA roles: A row / accumulator row q, K index / column pair r
  const int mq = lane >> 2, mr = lane & 3;
  // B-fragment offsets of this lane inside the window: tiles 0..9 and tiles 10..12
  const int boff = mr + mq * MP;
  const int boff2 = mr + (8 + (mq & 3)) * MP + (mq >> 2) * MPL;
  const bool bk3 = mr == 3;   // K index 11 (step 2): outside the 11-wide window
  const bool bm11 = (mq & 3) == 3;  // leftover tiles: the column slot of m' = 11
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
    // item j of a tile covers the tile's groups [j CHUNK, (j + 1) CHUNK): a tile has at most as many groups as records
    const unsigned r0 = tile_start[it.bin];
    const unsigned ng = tile_ngroups[it.bin];
    const unsigned first = min(it.start - r0, ng);
    const unsigned cnt = min((unsigned)CHUNK_QUERIES, ng - first);
    if (cnt == 0) continue;  // block-uniform
    const int tx = it.bin % g.nt[0], ty = (it.bin / g.nt[0]) % g.nt[1], tz = it.bin / (g.nt[0] * g.nt[1]);
    const int org[3] = {tx * g.T, ty * g.T, tz * g.T};
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, MELEMS * (int)sizeof(TC));
      tma_load_3d(brick, &tmap, org[0], org[1], org[2], bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    // contiguous range of this warp, in batches of 4 groups (32 records: lane = (group, row))
    const unsigned per = ((cnt + nwarps - 1) / nwarps + 3u) & ~3u;
    const unsigned w0 = warp * per, w1 = min(cnt, w0 + per);
    const uint2 *gl = groups + r0 + first;
    const Record *rc = rec + r0;
    const int pg = lane >> 3, prow = lane & 7;
    // group entries are fetched two batches ahead, the permutation entry of a record one batch ahead, and the record itself
    // is pulled into L2 while the previous batch is being evaluated: three dependent loads, none of them waited for
    uint2 ge_n = make_uint2(0u, 0u), ge_nn = make_uint2(0u, 0u);
    if (w0 + pg < w1) ge_n = __ldg(gl + w0 + pg);
    if (w0 + 4 + pg < w1) ge_nn = __ldg(gl + w0 + 4 + pg);
    unsigned pv_n = 0;
    if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
    for (unsigned b0 = w0; b0 < w1; b0 += 4) {
      const int ngb = (int)min(4u, w1 - b0);
      const uint2 ge = ge_n;
      const unsigned pv = pv_n;
      ge_n = ge_nn;
      ge_nn = make_uint2(0u, 0u);
      if (b0 + 8 + pg < w1) ge_nn = __ldg(gl + b0 + 8 + pg);
      pv_n = 0;
      if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
      // ---- phase 1: one record per lane: positions into the warp's scratch ----------------------------------------
      __syncwarp();  // the previous batch is done with tq / misc
      if ((unsigned)prow < ge.y) {
        // position record (bin_scatter_kernel<double, true>): {t_0, t_1, t_2, index, rows << 12 | cell in tile}
        double2 ra, rb;
        // (perm: the cell sort left the records where the scatter pass put them and wrote the sorted order as indices)
        ld_record(rc + (perm ? pv : ge.x + prow), ra, rb);
        tq[lane * 3 + 0] = ra.x;
        tq[lane * 3 + 1] = ra.y;
        tq[lane * 3 + 2] = rb.x;
        const unsigned code = (unsigned)__double2hiint(rb.y);
        const unsigned lc = code & 4095u, rcode = code >> 12;
        const unsigned rz = rcode / 10000u, ry = (rcode - rz * 10000u) / 100u, rx = rcode - rz * 10000u - ry * 100u;
        int bx, by, bz;
        lcell_decode(lc, 16u, bx, by, bz);  // T = 16 cells per tile side (2E = 10)
        const int m = bx + by * MP + bz * MPL;
        misc[lane] = make_int4(m - (bx & 1) - (by & 1) * MP - (bz & 1) * MPL, (int)(rx | (ry << 10) | (rz << 20)), __double2loint(rb.y), (int)(lc & 7u));
      }
      if (!perm && ge_n.y) prefetch_l2(rc + ge_n.x + min((unsigned)prow, ge_n.y - 1));
      __syncwarp();
      for (int gi = 0; gi < ngb; ++gi) {
        const int cntg = (int)__shfl_sync(0xffffffffu, ge.y, gi * 8);
        const int sb = gi * 8;  // first scratch slot of the group
        // ---- weights of the group's queries, two queries per pass: Horner on the staged cubics ----------------------
        // one (query, axis) pair per QUARTER-warp and pass, lane = pair of taps (5 of 8 lanes): the four 16-byte loads of
        // a quarter read one contiguous 80-byte run of a table row each - conflict-free whatever rows the other quarters
        // need.  All loads of two passes are issued before their first store (the compiler cannot prove that the wq
        // stores do not alias the scratch it reads, and would serialise the passes).
        const int ncombo = 3 * cntg;
#pragma unroll 1
        for (int c8 = 0; c8 < ncombo; c8 += 8) {
          double tt[2];
          double2 c0[2], c1[2], c2[2], c3[2];
          int dst[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ci = min(c8 + 4 * u + wquarter, ncombo - 1);  // past the end: recompute the last pair, not stored
            const int qq = (ci * 11) >> 5, ax = ci - 3 * qq;        // ci / 3, ci % 3 (ci < 24)
            tt[u] = tq[(sb + qq) * 3 + ax];
            const int row = (misc[sb + qq].y >> (10 * ax)) & 1023;
            const double2 *pr = (const double2 *)(smd + (ax == 0 ? wtab0 : (ax == 1 ? wtab1 : wtab2)) + row * PTAB_ROW);
            c0[u] = pr[0];
            c1[u] = pr[5];
            c2[u] = pr[10];
            c3[u] = pr[15];
            dst[u] = qq * WQS + (ax == 0 ? 2 : (ax == 1 ? 14 : 24)) + 2 * wpp;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            double2 wv;
            wv.x = fma(tt[u], fma(tt[u], fma(tt[u], c3[u].x, c2[u].x), c1[u].x), c0[u].x);
            wv.y = fma(tt[u], fma(tt[u], fma(tt[u], c3[u].y, c2[u].y), c1[u].y), c0[u].y);
            if (wrole && c8 + 4 * u + wquarter < ncombo) *(double2 *)(wq + dst[u]) = wv;
          }
        }
        __syncwarp();
        // ---- the axis-0 contraction on the tensor cores ------------------------------------------------------------
        const int win = misc[sb].x;  // brick offset of the block's first cell (the same for every record of the group)
        const bool qv = mq < cntg;
        const int4 mi = misc[sb + mq];
        const int dx = mi.w & 1, dy = (mi.w >> 1) & 1, dz = (mi.w >> 2) & 1;
        const double *wrow = wq + mq * WQS;
        double a[3];
#pragma unroll
        for (int s = 0; s < 3; ++s) a[s] = qv ? wrow[2 - dx + 4 * s + mr] : 0.0;
        const TC *bb = brick + win + boff, *bc = brick + win + boff2;
        // weights of this lane's columns: m' = 2 mr, 2 mr + 1 (regular tiles) and m' = 8 + 2 (mr & 1), + 1 (leftover tiles)
        auto wyv = [&](int mp) { const unsigned l = (unsigned)(mp - dy); return l < 10u ? wrow[14 + l] : 0.0; };
        auto wzv = [&](int np) { const unsigned l = (unsigned)(np - dz); return l < 10u ? wrow[24 + l] : 0.0; };
        const double wyA = wyv(2 * mr), wyB = wyv(2 * mr + 1);
        const double wyC = wyv(8 + 2 * (mr & 1)), wyD = wyv(9 + 2 * (mr & 1));
        double s0 = 0.0, s1 = 0.0, v0 = 0.0, v1 = 0.0;
        {  // chunk 1: regular tiles n' = 0..8
          double acc[9][2];
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int t = 0; t < 9; ++t) dmma884(acc[t][0], acc[t][1], a[s], (s == 2 && bk3) ? 0.0 : (double)bb[t * MPL + 4 * s]);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const double wz = wzv(t);
            s0 = fma(acc[t][0], wz, s0);
            s1 = fma(acc[t][1], wz, s1);
          }
        }
        {  // chunk 2: regular tiles n' = 9, 10 and the six leftover tiles
          double acc[8][2];
#pragma unroll
          for (int t = 0; t < 8; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const bool kz = s == 2 && bk3;
#pragma unroll
            for (int t = 0; t < 2; ++t) dmma884(acc[t][0], acc[t][1], a[s], kz ? 0.0 : (double)bb[(9 + t) * MPL + 4 * s]);
#pragma unroll
            for (int u = 0; u < 6; ++u)
              dmma884(acc[2 + u][0], acc[2 + u][1], a[s], (kz || bm11 || (u == 5 && (mq >> 2))) ? 0.0 : (double)bc[2 * u * MPL + 4 * s]);
          }
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const double wz = wzv(9 + t);
            s0 = fma(acc[t][0], wz, s0);
            s1 = fma(acc[t][1], wz, s1);
          }
#pragma unroll
          for (int u = 0; u < 6; ++u) {  // columns 2 mr, 2 mr + 1 of tile 11 + u: m' = 8 + 2 (mr & 1), + 1; n' = 2 u + (mr >> 1)
            const double wz = wzv(2 * u + (mr >> 1));
            v0 = fma(acc[2 + u][0], wz, v0);
            v1 = fma(acc[2 + u][1], wz, v1);
          }
        }
        double e = fma(s0, wyA, s1 * wyB) + fma(v0, wyC, v1 * wyD);
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        if (qv && mr == 0) out[(unsigned)mi.z] = (TC)apply_scale<3>(e, sc);
        __syncwarp();  // wq is rewritten by the next group
      }
      if (perm && (unsigned)prow < ge_n.y) prefetch_l2(rc + pv_n);  // the next batch's record (its index arrived long ago)
    }
  }
}

// ---- the same kernel for every :b-series width (2E = 10 .. 18) and for interpolants with several table pairs ---------------
// Generalisation of eval_cellmma_kernel: E is a template parameter (tile side T = 26 - 2E, K = 2E + 1 padded to 4 KS,
// (2E)^2 columns), the weights are blended from the staged K / K' tables (no cubic tables: they would not fit next to the
// brick for wide kernels or several derivative orders), and the column tiles are processed in chunks of 13 so that the
// accumulators stay at 26 registers whatever the width.  Column layout:
//   regular tiles  (h, n), h < MH = 2E / 8:  column c -> (m = 8 h + c, n_z = n)
//   leftover tiles u (the ML = 2E mod 8 remaining m):  column c -> (m = 8 MH + c / NPT, n_z = NPT u + c % NPT)
// BLK: rows of a product from a 2 x 2 x 2 block of cells instead of an aligned pair (sparser batches): the window of the
// brick is 2E + 1 wide along every axis, so the columns are (m', n') in [0, 2E + 1)^2 and a query's own 2E x 2E range starts at
// (dy, dz) in {0, 1}^2 - its weights for columns outside that range are zero
template <int E, bool BLK = false>
struct MmaGeom {
  static constexpr int W = 2 * E, T = BRICK_SIDE + 1 - W;
  static constexpr int WC = BLK ? W + 1 : W;  // extent of the column indices m, n
  static constexpr int KS = (W + 1 + 3) / 4;
  static constexpr int MH = WC / 8, ML = WC % 8;
  static constexpr int NPT = ML == 0 ? 1 : (ML <= 1 ? 8 : (ML <= 2 ? 4 : (ML <= 4 ? 2 : 1)));
  static constexpr int NREG = WC * MH, NLEFT = ML ? (WC + NPT - 1) / NPT : 0, NTILES = NREG + NLEFT;
  static constexpr int WX_LEN = 4 * KS + 2;
  static constexpr int WY_OFF = WX_LEN, WY_LEN = (ML && NPT == 1) ? 8 * (MH + 1) : W;
  static constexpr int WZ_OFF = WY_OFF + WY_LEN, WZ_LEN = ML ? ((NPT * NLEFT + 1) & ~1) : W;
  static constexpr int RAW = WX_LEN + WY_LEN + WZ_LEN;
  static constexpr int WQS = RAW + ((4 - RAW % 16) + 16) % 16;  // = 4 mod 16 doubles: A-fragment loads of a half-warp hit distinct banks
  static constexpr int WARP_DOUBLES = 32 * 3 + 32 * 2 + 8 * WQS;
  static constexpr int CHUNK = 13;
};

template <int E, int NT, bool BLK>
__global__ void __launch_bounds__(NT, 1)
    eval_cellmma_gen_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                            const Record *__restrict__ rec, const unsigned *__restrict__ perm, const uint2 *__restrict__ groups,
                            const unsigned *__restrict__ tile_start, const unsigned *__restrict__ tile_ngroups,
                            const WorkItem *__restrict__ items, unsigned *__restrict__ n_items, double *__restrict__ out) {
  using G = MmaGeom<E, BLK>;
  constexpr int W = G::W, WC = G::WC, T = G::T, KS = G::KS, MH = G::MH, NPT = G::NPT, WQS_ = G::WQS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *smd = (double *)smem_raw;
  double *brick = smd;  // [BRICK_SIDE][BRICK_SIDE][MP], axis 0 contiguous
  double *tab = smd + MELEMS;
  const double *tk[3];
  const double *td[3];
  stage_tables(p, tab, tk, td);
  constexpr int nwarps = NT >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbase = MELEMS + ((tables_smem_doubles(p) + 1) & ~1);
  double *wsm = smd + wbase + warp * G::WARP_DOUBLES;
  double *tq = wsm;                  // [32][3]
  int4 *misc = (int4 *)(wsm + 96);   // [32]
  double *wq = wsm + 96 + 64;        // [8][WQS]
  unsigned long long *bar = (unsigned long long *)(smd + wbase + nwarps * G::WARP_DOUBLES);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = lane; i < 8 * WQS_; i += 32) wq[i] = 0.0;  // the zero pads are never written again
  __syncthreads();
  // weight role: LPC lanes per (query, axis) pair, lane = pair of taps; 32 / LPC pairs per pass
  constexpr int LPC = W / 2 <= 8 ? 8 : 16, CPP = 32 / LPC;
  const int wsub = lane / LPC;
  const bool wrole = (lane % LPC) < W / 2;
  const int wpp = wrole ? (lane % LPC) : W / 2 - 1;
  const int tko0 = (int)(tk[0] - smd) + 2 * wpp, tko1 = (int)(tk[1] - smd) + 2 * wpp, tko2 = (int)(tk[2] - smd) + 2 * wpp;
  const int tdo0 = (int)(td[0] - smd) + 2 * wpp, tdo1 = (int)(td[1] - smd) + 2 * wpp, tdo2 = (int)(td[2] - smd) + 2 * wpp;
  const bool lin0 = p.subgrid[0] == CIB200_SG_LINEAR, lin1 = p.subgrid[1] == CIB200_SG_LINEAR, lin2 = p.subgrid[2] == CIB200_SG_LINEAR;
  const int mq = lane >> 2, mr = lane & 3;
  const int boff = mr + mq * MP;                                             // regular tiles
  const int boff2 = mr + (8 * MH + mq / NPT) * MP + (mq % NPT) * MPL;        // leftover tiles
  const bool lm_ok = 8 * MH + mq / NPT < WC;                                 // leftover column has a real m
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
    const unsigned r0 = tile_start[it.bin];
    const unsigned ng = tile_ngroups[it.bin];
    const unsigned first = min(it.start - r0, ng);
    const unsigned cnt = min((unsigned)CHUNK_QUERIES, ng - first);
    if (cnt == 0) continue;  // block-uniform
    const int tx = it.bin % g.nt[0], ty = (it.bin / g.nt[0]) % g.nt[1], tz = it.bin / (g.nt[0] * g.nt[1]);
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, MELEMS * 8);
      tma_load_3d(brick, &tmap, tx * T, ty * T, tz * T, bar);
    }
    mbar_wait(bar, parity);
    parity ^= 1;
    const unsigned per = ((cnt + nwarps - 1) / nwarps + 3u) & ~3u;
    const unsigned w0 = warp * per, w1 = min(cnt, w0 + per);
    const uint2 *gl = groups + r0 + first;
    const Record *rc = rec + r0;
    const int pg = lane >> 3, prow = lane & 7;
    // group entries are fetched two batches ahead, the permutation entry of a record one batch ahead, and the record itself
    // is pulled into L2 while the previous batch is being evaluated: three dependent loads, none of them waited for
    uint2 ge_n = make_uint2(0u, 0u), ge_nn = make_uint2(0u, 0u);
    if (w0 + pg < w1) ge_n = __ldg(gl + w0 + pg);
    if (w0 + 4 + pg < w1) ge_nn = __ldg(gl + w0 + 4 + pg);
    unsigned pv_n = 0;
    if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
    for (unsigned b0 = w0; b0 < w1; b0 += 4) {
      const int ngb = (int)min(4u, w1 - b0);
      const uint2 ge = ge_n;
      const unsigned pv = pv_n;
      ge_n = ge_nn;
      ge_nn = make_uint2(0u, 0u);
      if (b0 + 8 + pg < w1) ge_nn = __ldg(gl + b0 + 8 + pg);
      pv_n = 0;
      if (perm && (unsigned)prow < ge_n.y) pv_n = __ldg(perm + r0 + ge_n.x + prow);
      __syncwarp();  // the previous batch is done with tq / misc
      if ((unsigned)prow < ge.y) {
        double2 ra, rb;
        // (perm: the cell sort left the records where the scatter pass put them and wrote the sorted order as indices)
        ld_record(rc + (perm ? pv : ge.x + prow), ra, rb);
        tq[lane * 3 + 0] = ra.x;
        tq[lane * 3 + 1] = ra.y;
        tq[lane * 3 + 2] = rb.x;
        const unsigned code = (unsigned)__double2hiint(rb.y);
        const unsigned lc = code & 4095u, rcode = code >> 12;
        const unsigned rz = rcode / 10000u, ry = (rcode - rz * 10000u) / 100u, rx = rcode - rz * 10000u - ry * 100u;
        int cx, cy, cz;
        lcell_decode(lc, (unsigned)T, cx, cy, cz);
        const int m = cx + cy * MP + cz * MPL;
        misc[lane] = BLK ? make_int4(m - (cx & 1) - (cy & 1) * MP - (cz & 1) * MPL, (int)(rx | (ry << 10) | (rz << 20)), __double2loint(rb.y), (int)(lc & 7u))
                         : make_int4(2 * m + (cx & 1), (int)(rx | (ry << 10) | (rz << 20)), __double2loint(rb.y), 0);
      }
      if (!perm && ge_n.y) prefetch_l2(rc + ge_n.x + min((unsigned)prow, ge_n.y - 1));
      __syncwarp();
      for (int gi = 0; gi < ngb; ++gi) {
        const int cntg = (int)__shfl_sync(0xffffffffu, ge.y, gi * 8);
        const int sb = gi * 8;
        // ---- weights: blended from the staged K / K' rows, CPP (query, axis) pairs per pass ------------------------------
        const int ncombo = 3 * cntg;
#pragma unroll 1
        for (int c8 = 0; c8 < ncombo; c8 += 2 * CPP) {
          double tt[2];
          double2 ka[2], kb[2], da[2], db[2];
          int dst[2];
          bool li[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int ci = min(c8 + CPP * u + wsub, ncombo - 1);  // past the end: recompute the last pair, not stored
            const int qq = (ci * 11) >> 5, ax = ci - 3 * qq;      // ci / 3, ci % 3 (ci < 24)
            tt[u] = tq[(sb + qq) * 3 + ax];
            const int row = ((misc[sb + qq].y >> (10 * ax)) & 1023) * W;
            const double2 *kp = (const double2 *)(smd + (ax == 0 ? tko0 : (ax == 1 ? tko1 : tko2)) + row);
            const double2 *dp = (const double2 *)(smd + (ax == 0 ? tdo0 : (ax == 1 ? tdo1 : tdo2)) + row);
            ka[u] = kp[0];
            kb[u] = kp[W / 2];
            da[u] = dp[0];
            db[u] = dp[W / 2];
            li[u] = ax == 0 ? lin0 : (ax == 1 ? lin1 : lin2);
            dst[u] = qq * WQS_ + (ax == 0 ? 2 : (ax == 1 ? G::WY_OFF : G::WZ_OFF)) + 2 * wpp;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const Hermite hb = li[u] ? linear_basis(tt[u]) : hermite_basis(tt[u]);
            double2 wv;
            wv.x = hb.h00 * ka[u].x + hb.h01 * kb[u].x + (hb.h10 * da[u].x + hb.h11 * db[u].x);
            wv.y = hb.h00 * ka[u].y + hb.h01 * kb[u].y + (hb.h10 * da[u].y + hb.h11 * db[u].y);
            if (wrole && c8 + CPP * u + wsub < ncombo) *(double2 *)(wq + dst[u]) = wv;
          }
        }
        __syncwarp();
        // ---- the axis-0 contraction on the tensor cores, 13 column tiles at a time ----------------------------------------
        const int mf = misc[sb].x;
        const int win = BLK ? mf : (mf >> 1) - (mf & 1);  // brick offset of the pair's / block's first cell
        const bool qv = mq < cntg;
        const int4 mi = misc[sb + mq];
        const int dx = BLK ? (mi.w & 1) : (mi.x & 1), dy = BLK ? ((mi.w >> 1) & 1) : 0, dz = BLK ? ((mi.w >> 2) & 1) : 0;
        const double *wrow = wq + mq * WQS_;
        // weight of column index m' / n' for this lane's query: its own range starts at dy / dz inside the window
        auto wyv = [&](int mp) { const unsigned l = (unsigned)(mp - dy); return l < (unsigned)W ? wrow[G::WY_OFF + l] : 0.0; };
        double a[KS];
#pragma unroll
        for (int s = 0; s < KS; ++s) a[s] = qv ? wrow[2 - dx + 4 * s + mr] : 0.0;
        double wz[WC];
        if (BLK) {
#pragma unroll
          for (int n = 0; n < WC; ++n) {
            const unsigned l = (unsigned)(n - dz);
            wz[n] = l < (unsigned)W ? wrow[G::WZ_OFF + l] : 0.0;
          }
        } else {
#pragma unroll
          for (int n = 0; n + 1 < WC; n += 2) {
            const double2 v = *(const double2 *)(wrow + G::WZ_OFF + n);
            wz[n] = v.x;
            wz[n + 1] = v.y;
          }
        }
        double sreg[MH > 0 ? MH : 1][2], sl0 = 0.0, sl1 = 0.0;
#pragma unroll
        for (int h = 0; h < MH; ++h) sreg[h][0] = sreg[h][1] = 0.0;
        const double *bb = brick + win + boff, *bc = brick + win + boff2;
#pragma unroll
        for (int c0 = 0; c0 < G::NTILES; c0 += G::CHUNK) {
          constexpr int CH = G::CHUNK;
          double acc[CH][2];
#pragma unroll
          for (int t = 0; t < CH; ++t) acc[t][0] = acc[t][1] = 0.0;
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            const bool kz = 4 * s + mr > W;  // K index beyond the (2E + 1)-wide window
#pragma unroll
            for (int t = 0; t < CH; ++t) {
              const int ti = c0 + t;
              if (ti < G::NREG) {
                const int h = ti / WC, n = ti - h * WC;
                dmma884(acc[t][0], acc[t][1], a[s], kz ? 0.0 : bb[8 * h * MP + n * MPL + 4 * s]);
              } else if (ti < G::NTILES) {
                const int u = ti - G::NREG;
                const bool z = kz || !lm_ok || NPT * u + (mq % NPT) >= WC;
                dmma884(acc[t][0], acc[t][1], a[s], z ? 0.0 : bc[NPT * u * MPL + 4 * s]);
              }
            }
          }
          // partial epilogue of the chunk
#pragma unroll
          for (int t = 0; t < CH; ++t) {
            const int ti = c0 + t;
            if (ti < G::NREG) {
              const int h = ti / WC, n = ti - h * WC;
              sreg[h][0] = fma(acc[t][0], wz[n], sreg[h][0]);
              sreg[h][1] = fma(acc[t][1], wz[n], sreg[h][1]);
            } else if (ti < G::NTILES) {
              const int u = ti - G::NREG;
              if (NPT == 1) {  // columns (m = 8 MH + 2c, 2c + 1; n_z = u)
                sl0 = fma(acc[t][0], wz[u], sl0);
                sl1 = fma(acc[t][1], wz[u], sl1);
              } else if (BLK) {  // columns (m = 8 MH + 2c / NPT; n_z = NPT u + (2c) % NPT, + 1): dynamic n_z, predicated
                const int n0 = NPT * u + ((2 * mr) % NPT);
                const unsigned l0 = (unsigned)(n0 - dz), l1 = (unsigned)(n0 + 1 - dz);
                sl0 = fma(acc[t][0], l0 < (unsigned)W ? wrow[G::WZ_OFF + l0] : 0.0, sl0);
                sl0 = fma(acc[t][1], l1 < (unsigned)W ? wrow[G::WZ_OFF + l1] : 0.0, sl0);
              } else {  // columns (m = 8 MH + 2c / NPT; n_z = NPT u + (2c) % NPT, + 1)
                const double2 v = *(const double2 *)(wrow + G::WZ_OFF + NPT * u + ((2 * mr) % NPT));
                sl0 = fma(acc[t][0], v.x, sl0);
                sl0 = fma(acc[t][1], v.y, sl0);
              }
            }
          }
        }
        double e = 0.0;
#pragma unroll
        for (int h = 0; h < MH; ++h) {
          if (BLK) {
            e += fma(sreg[h][0], wyv(8 * h + 2 * mr), sreg[h][1] * wyv(8 * h + 2 * mr + 1));
          } else {
            const double2 wy = *(const double2 *)(wrow + G::WY_OFF + 8 * h + 2 * mr);
            e += fma(sreg[h][0], wy.x, sreg[h][1] * wy.y);
          }
        }
        if (G::ML) {
          if (BLK) {
            if (NPT == 1) e += fma(sl0, wyv(8 * MH + 2 * mr), sl1 * wyv(8 * MH + 2 * mr + 1));
            else e += sl0 * wyv(8 * MH + (2 * mr) / NPT);
          } else if (NPT == 1) {
            const double2 wy = *(const double2 *)(wrow + G::WY_OFF + 8 * MH + 2 * mr);
            e += fma(sl0, wy.x, sl1 * wy.y);
          } else {
            e += sl0 * wrow[G::WY_OFF + 8 * MH + (2 * mr) / NPT];
          }
        }
        e += __shfl_xor_sync(0xffffffffu, e, 1);
        e += __shfl_xor_sync(0xffffffffu, e, 2);
        if (qv && mr == 0) out[(unsigned)mi.z] = apply_scale<3>(e, sc);
        __syncwarp();  // wq is rewritten by the next group
      }
      if (perm && (unsigned)prow < ge_n.y) prefetch_l2(rc + pv_n);  // the next batch's record (its index arrived long ago)
    }
  }
}

template <int E>
static size_t mma_gen_smem(const ItpParams &p, int nwarps) {
  constexpr int wd = MmaGeom<E, false>::WARP_DOUBLES > MmaGeom<E, true>::WARP_DOUBLES ? MmaGeom<E, false>::WARP_DOUBLES : MmaGeom<E, true>::WARP_DOUBLES;
  return 8 * ((size_t)MELEMS + (size_t)((tables_smem_doubles(p) + 1) & ~1) + (size_t)nwarps * wd + 2 + 96);
}

template <int E, int NT, bool BLK>
static int launch_mma_gen(cib200_itp *h, BinState *bs, const Scale &sc, double *out, cudaStream_t st) {
  const ItpParams &p = h->p;
  auto kern = eval_cellmma_gen_kernel<E, NT, BLK>;
  // (the attribute is per function AND device: set on every launch, a host-side call of about a microsecond)
  CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<bs->nblk, NT, mma_gen_smem<E>(p, NT / 32), st>>>(bs->tmap28, p, bs->g, sc, bs->perm ? bs->rec : bs->rec2, bs->perm, (const uint2 *)bs->slots, bs->tile_start,
                                                         bs->tile_nslots, bs->items, bs->n_items, out);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

// threads per CTA of the generic kernel: the most warps whose scratch fits next to the brick and the staged tables
template <int E>
static int mma_gen_threads(const ItpParams &p) {
  for (int nt : {512, 384, 256})
    if (mma_gen_smem<E>(p, nt / 32) <= 227 * 1024) return nt;
  return 0;
}
template <int E, bool BLK>
static int launch_mma_gen_e(cib200_itp *h, BinState *bs, const Scale &sc, double *out, cudaStream_t st) {
  switch (mma_gen_threads<E>(h->p)) {
    case 512: return launch_mma_gen<E, 512, BLK>(h, bs, sc, out, st);
    case 384: return launch_mma_gen<E, 384, BLK>(h, bs, sc, out, st);
    case 256: return launch_mma_gen<E, 256, BLK>(h, bs, sc, out, st);
  }
  set_error("tensor-core strategy: shared memory budget exceeded");
  return CIB200_EUNSUPPORTED;
}

static size_t mma_smem(const ItpParams &p, int nwarps, size_t esize = 8) {
  return esize * (size_t)MELEMS + 8 * ((size_t)mma_ptab_count(p) * PTAB_DOUBLES + (size_t)nwarps * MmaWarpSmem::DOUBLES + 2 + 96);
}
size_t cellmma_smem_bytes(const ItpParams &p) { return mma_smem(p, 16); }

// 1: the tuned :b5 kernel (one table pair, cubic tables); 2: the generic kernel; 0: this interpolant does not fit
int cellmma_kind(const ItpParams &p, bool f32) {
  const int E = p.eqs[0];
  if (f32) return (E == 5 && mma_ptab_count(p) == 1) ? 1 : 0;  // Float32: the tuned :b5 kernel only
  if (E == 5 && mma_ptab_count(p) == 1 && mma_smem(p, 16) <= 227 * 1024) return 1;
  switch (E) {
    case 5: return mma_gen_threads<5>(p) ? 2 : 0;
    case 6: return mma_gen_threads<6>(p) ? 2 : 0;
    case 7: return mma_gen_threads<7>(p) ? 2 : 0;
    case 8: return mma_gen_threads<8>(p) ? 2 : 0;
    case 9: return mma_gen_threads<9>(p) ? 2 : 0;
  }
  return 0;
}

template <int NT, bool BREG, typename TC = double>
static int launch_mma(cib200_itp *h, BinState *bs, const Scale &sc, TC *out, cudaStream_t st) {
  const ItpParams &p = h->p;
  const size_t smem = mma_smem(p, NT / 32, sizeof(TC));
  auto kern = eval_cellmma_kernel<NT, BREG, TC>;
  // (the attribute is per function AND device: set on every launch, a host-side call of about a microsecond)
  CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // (Float32: the handle's own tensor map already stages rows of 28 floats)
  kern<<<bs->nblk, NT, smem, st>>>(sizeof(TC) == 8 ? bs->tmap28 : bs->tmap, p, bs->g, sc, bs->perm ? bs->rec : bs->rec2, bs->perm, (const uint2 *)bs->slots,
                                   bs->tile_start, bs->tile_nslots, bs->items, bs->n_items, out);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

template <int NT, typename TC>
static int launch_oct(cib200_itp *h, BinState *bs, const Scale &sc, TC *out, cudaStream_t st) {
  const ItpParams &p = h->p;
  const size_t smem = mma_smem(p, NT / 32, sizeof(TC));
  auto kern = eval_celloct_kernel<NT, TC>;
  CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kern<<<bs->nblk, NT, smem, st>>>(sizeof(TC) == 8 ? bs->tmap28 : bs->tmap, p, bs->g, sc, bs->perm ? bs->rec : bs->rec2, bs->perm, (const uint2 *)bs->slots,
                                   bs->tile_start, bs->tile_nslots, bs->items, bs->n_items, out);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

// block variant (sparse batches): records sorted by (tile, block-major cell) + the group list of bin_cellsort_kernel<3>
int launch_eval_celloct(cib200_itp *h, BinState *bs, const Scale &sc, void *out, cudaStream_t st) {
  if (cellmma_kind(h->p, h->dtype == CIB200_F32) == 2) {  // generic kernel (any :b width, several table pairs), Float64
    switch (h->p.eqs[0]) {
      case 5: return launch_mma_gen_e<5, true>(h, bs, sc, (double *)out, st);
      case 6: return launch_mma_gen_e<6, true>(h, bs, sc, (double *)out, st);
      case 7: return launch_mma_gen_e<7, true>(h, bs, sc, (double *)out, st);
      case 8: return launch_mma_gen_e<8, true>(h, bs, sc, (double *)out, st);
      case 9: return launch_mma_gen_e<9, true>(h, bs, sc, (double *)out, st);
    }
    set_error("tensor-core strategy: unsupported kernel width");
    return CIB200_EUNSUPPORTED;
  }
  if (h->dtype == CIB200_F32) return launch_oct<512, float>(h, bs, sc, (float *)out, st);
  return launch_oct<512, double>(h, bs, sc, (double *)out, st);
}

int launch_eval_cellmma(cib200_itp *h, BinState *bs, const Scale &sc, void *out, cudaStream_t st) {
  if (h->dtype == CIB200_F32) return launch_mma<512, false, float>(h, bs, sc, (float *)out, st);
  if (cellmma_kind(h->p) == 2 || tune_env("CIB200_MMA_GENERIC")) {
    switch (h->p.eqs[0]) {
      case 5: return launch_mma_gen_e<5, false>(h, bs, sc, (double *)out, st);
      case 6: return launch_mma_gen_e<6, false>(h, bs, sc, (double *)out, st);
      case 7: return launch_mma_gen_e<7, false>(h, bs, sc, (double *)out, st);
      case 8: return launch_mma_gen_e<8, false>(h, bs, sc, (double *)out, st);
      case 9: return launch_mma_gen_e<9, false>(h, bs, sc, (double *)out, st);
    }
    set_error("tensor-core strategy: unsupported kernel width");
    return CIB200_EUNSUPPORTED;
  }
  int variant = 2;  // 16 warps, B fragments straight from the brick: measured best (tools/ab_lanequery.py)
  if (const char *e = tune_env("CIB200_MMA_VARIANT")) variant = atoi(e);
  switch (variant) {
    case 1: return launch_mma<384, false>(h, bs, sc, (double *)out, st);
    case 3: return launch_mma<512, true>(h, bs, sc, (double *)out, st);
    case 0: return launch_mma<384, true>(h, bs, sc, (double *)out, st);
    default: return launch_mma<512, false>(h, bs, sc, (double *)out, st);
  }
}

}  // namespace cib
