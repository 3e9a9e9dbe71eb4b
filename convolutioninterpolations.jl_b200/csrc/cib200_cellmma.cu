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

template <int NT, bool BREG>
__global__ void __launch_bounds__(NT, 1)
    eval_cellmma_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                        const Record *__restrict__ rec, const uint2 *__restrict__ groups,
                        const unsigned *__restrict__ tile_start, const unsigned *__restrict__ tile_ngroups,
                        const WorkItem *__restrict__ items, unsigned *__restrict__ n_items, double *__restrict__ out) {
  using WS = MmaWarpSmem;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double *smd = (double *)smem_raw;
  double *brick = smd;  // [BRICK_SIDE][BRICK_SIDE][MP], axis 0 contiguous
  double *ptab = smd + MELEMS;
  stage_poly_tables(p, ptab, NT);
  constexpr int nwarps = NT >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbase = MELEMS + mma_ptab_count(p) * PTAB_DOUBLES;
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
  const int wtab0 = MELEMS + mma_ptab_index(p, 0) * PTAB_DOUBLES + 2 * wpp;  // from smd
  const int wtab1 = MELEMS + mma_ptab_index(p, 1) * PTAB_DOUBLES + 2 * wpp;
  const int wtab2 = MELEMS + mma_ptab_index(p, 2) * PTAB_DOUBLES + 2 * wpp;
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
      mbar_expect_tx(bar, MELEMS * 8);
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
    uint2 ge_n = make_uint2(0u, 0u);
    if (w0 + pg < w1) ge_n = __ldg(gl + w0 + pg);
    int curwin = -1;
    double breg[39];
#pragma unroll
    for (int i = 0; i < 39; ++i) breg[i] = 0.0;
    for (unsigned b0 = w0; b0 < w1; b0 += 4) {
      const int ngb = (int)min(4u, w1 - b0);
      const uint2 ge = ge_n;
      ge_n = make_uint2(0u, 0u);
      if (b0 + 4 + pg < w1) ge_n = __ldg(gl + b0 + 4 + pg);
      // ---- phase 1: one record per lane: positions into the warp's scratch ----------------------------------------
      __syncwarp();  // the previous batch is done with tq / misc
      if ((unsigned)prow < ge.y) {
        // position record (bin_scatter_kernel<double, true>): {t_0, t_1, t_2, index, rows << 12 | cell in tile}
        double2 ra, rb;
        ld_record(rc + ge.x + prow, ra, rb);
        tq[lane * 3 + 0] = ra.x;
        tq[lane * 3 + 1] = ra.y;
        tq[lane * 3 + 2] = rb.x;
        const unsigned code = (unsigned)__double2hiint(rb.y);
        const unsigned lc = code & 4095u, rcode = code >> 12;
        const unsigned rz = rcode / 10000u, ry = (rcode - rz * 10000u) / 100u, rx = rcode - rz * 10000u - ry * 100u;
        const int bx = lc & 15u;  // T = 16 cells per tile side (2E = 10)
        const int m = bx + (int)((lc >> 4) & 15u) * MP + (int)(lc >> 8) * MPL;
        misc[lane] = make_int4(2 * m + (bx & 1), (int)(rx | (ry << 10) | (rz << 20)), __double2loint(rb.y), 0);
      }
      if (ge_n.y) prefetch_l2(rc + ge_n.x + min((unsigned)prow, ge_n.y - 1));
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
          const double *bb = brick + win + boff, *bc = brick + win + boff2;
#pragma unroll
          for (int t = 0; t < 10; ++t) {
            breg[t * 3 + 0] = bb[t * MPL];
            breg[t * 3 + 1] = bb[t * MPL + 4];
            breg[t * 3 + 2] = bk3 ? 0.0 : bb[t * MPL + 8];
          }
#pragma unroll
          for (int u = 0; u < 3; ++u) {
            const bool z = u == 2 && bt12;
            breg[(10 + u) * 3 + 0] = z ? 0.0 : bc[4 * u * MPL];
            breg[(10 + u) * 3 + 1] = z ? 0.0 : bc[4 * u * MPL + 4];
            breg[(10 + u) * 3 + 2] = (z || bk3) ? 0.0 : bc[4 * u * MPL + 8];
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
          const double *bb = brick + win + boff, *bc = brick + win + boff2;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int t = 0; t < 10; ++t) dmma884(acc[t][0], acc[t][1], a[s], (s == 2 && bk3) ? 0.0 : bb[t * MPL + 4 * s]);
#pragma unroll
            for (int u = 0; u < 3; ++u)
              dmma884(acc[10 + u][0], acc[10 + u][1], a[s], ((u == 2 && bt12) || (s == 2 && bk3)) ? 0.0 : bc[4 * u * MPL + 4 * s]);
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
        if (qv && mr == 0) out[(unsigned)mi.z] = apply_scale<3>(e, sc);
        __syncwarp();  // wq is rewritten by the next group
      }
    }
  }
}

static size_t mma_smem(const ItpParams &p, int nwarps) {
  return 8 * ((size_t)MELEMS + (size_t)mma_ptab_count(p) * PTAB_DOUBLES + (size_t)nwarps * MmaWarpSmem::DOUBLES + 2 + 96);
}
size_t cellmma_smem_bytes(const ItpParams &p) { return mma_smem(p, 16); }

template <int NT, bool BREG>
static int launch_mma(cib200_itp *h, BinState *bs, const Scale &sc, double *out, cudaStream_t st) {
  const ItpParams &p = h->p;
  const size_t smem = mma_smem(p, NT / 32);
  auto kern = eval_cellmma_kernel<NT, BREG>;
  static bool attr_set = false;
  if (!attr_set) {
    CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  kern<<<bs->nblk, NT, smem, st>>>(bs->tmap28, p, bs->g, sc, bs->rec2, (const uint2 *)bs->slots, bs->tile_start, bs->tile_nslots,
                                   bs->items, bs->n_items, out);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  return CIB200_OK;
}

int launch_eval_cellmma(cib200_itp *h, BinState *bs, const Scale &sc, void *out, cudaStream_t st) {
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
