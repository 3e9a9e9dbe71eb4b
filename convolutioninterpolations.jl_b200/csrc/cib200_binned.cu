// cib200_binned.cu - K1, "binned" strategy for large batches of 3-D queries (north_star subsystem 3:
// "TMA-staged coefficient bricks for spatially sorted queries" - here the library does the spatial
// sort itself, inside the timed region).
//
// Same arithmetic as the direct strategy (cib200_eval.cu): positions restated from the reference,
// factorised stencil sum; only WHERE the (2E)^3 coefficients come from changes.  A uniformly random
// batch on a 512^3 grid makes the direct kernel drag 18 kB per query through HBM (measured, profiles/
// r01_eval_gather_v0) because every query touches 100 rows at arbitrary 128-byte-line offsets.  Here:
//
//   pass A  bin_count_kernel     one persistent CTA per SM histograms its slice of the queries over
//                                T^3-cell tiles in shared memory (bin = tile of the stencil's first
//                                coefficient; queries that need extrapolation / raise go to a spill bin)
//   scan    bin_totals/scan/offsets   tile starts, per-(CTA, tile) write cursors, work list of
//                                (tile, <= CHUNK queries) items - all on the device, no host sync
//   pass B  bin_scatter_kernel   re-reads the queries and writes 32-byte records {x, y, z, index}
//                                (one full sector per query) to the tile-ordered record array
//   pass C  eval_binned_kernel   persistent CTAs pull work items; per item ONE 3-D TMA
//                                (cp.async.bulk.tensor, SASS UTMALDG) stages the (T+2E-1)^3 brick of the
//                                tile in shared memory; warps then run the same 32-query tiles as the
//                                direct kernel with the stencil rows read from the brick; results are
//                                scattered back to out[index]
//   pass D  eval_spill_kernel    the spill bin (Flat / Line / Throw / NaN queries) through the scalar
//                                evaluator straight from global memory
//
// HBM traffic per query drops from ~18 kB to ~130 B (coordinates twice, record write + read, result),
// the bound becomes shared-memory bandwidth (8 kB of stencil per query).
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "cib200_binned.cuh"

namespace cib {

// ---- classification -----------------------------------------------------------------------------
__device__ __forceinline__ int first_coef(double x, double knot0, double h, int eqs, int nk, bool &ok) {
  double i_float = __dadd_rn(__ddiv_rn(__dsub_rn(x, knot0), h), 1.0);
  ok = ok && fabs(i_float) < 9.0e18;
  long long i = ok ? (long long)floor(i_float) : (long long)eqs;
  i = i < eqs ? eqs : i;
  i = i > nk - eqs ? nk - eqs : i;
  return (int)(i - eqs);
}

// bin of a query; lcell = its cell inside the tile, block-major (lcell_encode; 0 for the spill bin)
__device__ __forceinline__ int classify(const ItpParams &p, const BinGeom &g, const double *x, const double *k0,
                                        unsigned &lcell) {
  lcell = 0;
  bool inside = true;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    if (x[d] > p.hi[d] || x[d] < p.lo[d]) inside = false;
  if (!inside) return g.ntiles;
  bool ok = true;
  int t[3], c[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const int fc = first_coef(x[d], k0[d], p.h[d], p.eqs[d], p.nk[d], ok);
    t[d] = g.T == 16 ? fc >> 4 : (g.T == 8 ? fc >> 3 : fc / g.T);  // (:b5 / :b13 tiles: no integer division)
    c[d] = fc - t[d] * g.T;
  }
  if (!ok) return g.ntiles;
  lcell = lcell_encode(c[0], c[1], c[2], g.T);
  return (t[2] * g.nt[1] + t[1]) * g.nt[0] + t[0];
}

template <typename T>
__global__ void __launch_bounds__(1024, 1) bin_count_kernel(ItpParams p, BinGeom g, Coords c, long long nq,
                                                            unsigned *__restrict__ hist, int global_hist) {
  extern __shared__ unsigned sh[];
  if (!global_hist)
    for (int i = threadIdx.x; i < g.nbin; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const double k0[3] = {p.knot0[0], p.knot0[1], p.knot0[2]};
  long long per = (nq + gridDim.x - 1) / gridDim.x;
  per = (per + 31) & ~31LL;
  const long long q0 = (long long)blockIdx.x * per;
  const long long q1 = q0 + per < nq ? q0 + per : nq;
  // BIN_U queries per thread and iteration: all their coordinate loads are in flight before the first is used
  for (long long qb = q0 + threadIdx.x; qb < q1; qb += (long long)BIN_U * blockDim.x) {
    double x[BIN_U][3];
#pragma unroll
    for (int u = 0; u < BIN_U; ++u) {
      const long long q = qb + (long long)u * blockDim.x;
#pragma unroll
      for (int d = 0; d < 3; ++d) x[u][d] = q < q1 ? (double)((const T *)c.x[d])[q] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < BIN_U; ++u) {
      const long long q = qb + (long long)u * blockDim.x;
      unsigned lc;
      if (q < q1) {
        const int bin_ = classify(p, g, x[u], k0, lc);
        if (global_hist) atomicAdd(hist + bin_, 1u);
        else atomicAdd(&sh[bin_], 1u);
      }
    }
  }
  if (global_hist) return;
  __syncthreads();
  unsigned *dst = hist + (size_t)blockIdx.x * g.nbin;
  for (int i = threadIdx.x; i < g.nbin; i += blockDim.x) dst[i] = sh[i];
}

// totals per bin (thread per bin, coalesced across bins)
__global__ void bin_totals_kernel(BinGeom g, int nblk, const unsigned *__restrict__ hist, unsigned *__restrict__ tile_start) {
  int bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= g.nbin) return;
  unsigned s = 0;
  for (int b = 0; b < nblk; ++b) s += hist[(size_t)b * g.nbin + bin];
  tile_start[bin] = s;
}

// single CTA: exclusive scan of the bin totals (in place) + work list
__global__ void __launch_bounds__(1024, 1) bin_scan_kernel(BinGeom g, unsigned *__restrict__ tile_start,
                                                           WorkItem *__restrict__ items, unsigned *__restrict__ n_items) {
  __shared__ unsigned part[1024];
  __shared__ unsigned part2[1024];
  const int tid = threadIdx.x;
  const int seg = (g.nbin + 1023) / 1024;
  const int b0 = tid * seg, b1 = min(b0 + seg, g.nbin);
  unsigned s = 0, w = 0;
  for (int b = b0; b < b1; ++b) {
    unsigned cnt = tile_start[b];
    s += cnt;
    if (b < g.ntiles) w += (cnt + CHUNK_QUERIES - 1) / CHUNK_QUERIES;
  }
  part[tid] = s;
  part2[tid] = w;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    unsigned a = tid >= o ? part[tid - o] : 0, b = tid >= o ? part2[tid - o] : 0;
    __syncthreads();
    part[tid] += a;
    part2[tid] += b;
    __syncthreads();
  }
  unsigned run = part[tid] - s, wrun = part2[tid] - w;
  for (int b = b0; b < b1; ++b) {
    unsigned cnt = tile_start[b];
    tile_start[b] = run;
    if (b < g.ntiles) {
      for (unsigned j = 0; j * CHUNK_QUERIES < cnt; ++j) {
        WorkItem it;
        it.bin = (unsigned)b;
        it.start = run + j * CHUNK_QUERIES;
        it.count = min((unsigned)CHUNK_QUERIES, cnt - j * CHUNK_QUERIES);
        items[wrun++] = it;
      }
    }
    run += cnt;
  }
  if (tid == 1023) {
    tile_start[g.nbin] = part[1023];
    n_items[0] = part2[1023];
    n_items[1] = 0;  // work counter of the evaluation kernel
    n_items[2] = 0;  // tile counter of the cell sort
  }
}

// counts -> write cursors: hist[b][bin] = tile_start[bin] + sum_{b' < b} count[b'][bin]
__global__ void bin_offsets_kernel(BinGeom g, int nblk, unsigned *__restrict__ hist, const unsigned *__restrict__ tile_start) {
  int bin = blockIdx.x * blockDim.x + threadIdx.x;
  if (bin >= g.nbin) return;
  unsigned run = tile_start[bin];
  for (int b = 0; b < nblk; ++b) {
    unsigned cnt = hist[(size_t)b * g.nbin + bin];
    hist[(size_t)b * g.nbin + bin] = run;
    run += cnt;
  }
}

// POSREC (tensor-core evaluation kernel): the record of a query inside the domain carries its POSITIONS instead of its
// coordinates - {t_0, t_1, t_2, index, (row_0 + 100 (row_1 + 100 row_2)) << 12 | cell in tile} - computed here once with
// the reference's arithmetic (locate_from_knot), so the evaluation kernel starts from table rows and fractions without a
// single division.  Records of the spill bin keep the coordinates.
template <typename T, bool POSREC>
__global__ void __launch_bounds__(1024, 1) bin_scatter_kernel(ItpParams p, BinGeom g, Coords c, long long nq,
                                                              unsigned *hist, Record *__restrict__ rec, int dbg) {
  extern __shared__ unsigned sh[];
  // dbg == 3: ONE write cursor per tile shared by all CTAs (row 0 of hist = tile starts, global atomics) instead of a
  // private cursor per (CTA, tile): 32 k open 128-byte lines (4 MB) instead of 148 x 32 k (600 MB), so a line collects
  // its four records in L2 before it is written back.  The order of the records inside a tile then varies from run to
  // run; the value returned for a query does not depend on it.
  const bool shared_cursors = (dbg & 3) == 3;
  // dbg & 4 (experiment): L2 eviction hints - coordinates are streamed (evict_first), records are written evict_last so
  // that a 128-byte line of four records stays in L2 until it is complete
  const bool hints = dbg & 4;
  unsigned long long pol_first = 0, pol_last = 0;
  if (hints) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  }
  if (!shared_cursors) {
    const unsigned *src = hist + (size_t)blockIdx.x * g.nbin;
    for (int i = threadIdx.x; i < g.nbin; i += blockDim.x) sh[i] = src[i];
  }
  __syncthreads();
  const double k0[3] = {p.knot0[0], p.knot0[1], p.knot0[2]};
  long long per = (nq + gridDim.x - 1) / gridDim.x;
  per = (per + 31) & ~31LL;
  const long long q0 = (long long)blockIdx.x * per;
  const long long q1 = q0 + per < nq ? q0 + per : nq;
  for (long long qb = q0 + threadIdx.x; qb < q1; qb += (long long)BIN_U * blockDim.x) {
    double x[BIN_U][3];
#pragma unroll
    for (int u = 0; u < BIN_U; ++u) {
      const long long q = qb + (long long)u * blockDim.x;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (hints && sizeof(T) == 8) {
          double v = 0.0;
          if (q < q1) asm volatile("ld.global.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"((const double *)c.x[d] + q), "l"(pol_first));
          x[u][d] = v;
        } else {
          x[u][d] = q < q1 ? (double)((const T *)c.x[d])[q] : 0.0;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BIN_U; ++u) {
      const long long q = qb + (long long)u * blockDim.x;
      if (q < q1) {
        unsigned lc;
        const int bin_ = classify(p, g, x[u], k0, lc);
        const unsigned slot = shared_cursors ? atomicAdd(hist + bin_, 1u) : atomicAdd(&sh[bin_], 1u);
        if (POSREC && bin_ != g.ntiles) {
          AxisPos pa[3];
#pragma unroll
          for (int d = 0; d < 3; ++d) locate_from_knot(x[u][d], k0[d], p.h[d], p.eqs[d], p.nk[d], p.knots[d], pa[d]);
          const unsigned code = ((unsigned)(pa[0].row + 100 * (pa[1].row + 100 * pa[2].row)) << 12) | lc;
          if (hints)
            asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1, %2, %3, %4}, %5;" ::"l"(rec + slot), "d"(pa[0].t), "d"(pa[1].t), "d"(pa[2].t),
                         "d"(__hiloint2double((int)code, (int)(unsigned)q)), "l"(pol_last)
                         : "memory");
          else
            st_record(rec + slot, pa[0].t, pa[1].t, pa[2].t, __hiloint2double((int)code, (int)(unsigned)q));
          continue;
        }
        // one 32-byte sector per query, written with one 32-byte store; the spare word carries the cell inside the tile
        st_record(rec + slot, x[u][0], x[u][1], x[u][2], __hiloint2double((int)lc, (int)(unsigned)q));
      }
    }
  }
}

// ---- pass B2: counting sort of every tile's records by cell ---------------------------------------
// One CTA per tile at a time (persistent CTAs pull tiles from a counter): histogram of the 12-bit cell ids carried
// by the records in shared memory, exclusive scan, then the records are re-read (L2) and written to their slot in
// the second record array.  Afterwards the queries of a cell are consecutive.
//
// MODE 1 (lane-per-query evaluation, two queries of one cell per lane): the pass also emits the tile's SLOT LIST - one
// u32 per pair of consecutive records of the same cell, (record offset inside the tile) * 2 + (the pair is complete) -
// at slots[tile_start[tile] ...], and the tile's slot count.  A cell with n records gets ceil(n / 2) slots, so a lane
// of the evaluation kernel never meets two different cells.
// MODE 2 (tensor-core evaluation, cib200_cellmma.cu): the GROUP LIST - one uint2 {record offset inside the tile, count}
// per run of up to 8 consecutive records of an aligned PAIR of cells along axis 0 (cells 2p, 2p + 1 of a row: their
// stencils share an 11-wide window), at ((uint2 *)slots)[tile_start[tile] ...], and the tile's group count.
// MODE 3: the same with the 2 x 2 x 2 BLOCKS of the block-major cell numbering (lcell >> 3) instead of pairs - sparse
// batches fill the 8 rows of a tensor-core product from eight cells.
static constexpr int CELLSORT_NIDS = 65536;  // cell ids kept in shared memory in the permutation mode (128 KB)
static size_t cellsort_smem_bytes(int mode, bool perm = false) {
  return 4 * (size_t)(mode ? 3 * CELLS_MAX + 2 : CELLS_MAX) + (perm ? 2 * (size_t)CELLSORT_NIDS : 0);
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bin_cellsort_kernel(BinGeom g, const Record *__restrict__ rec,
                                                               Record *__restrict__ rec2,
                                                               const unsigned *__restrict__ tile_start,
                                                               unsigned *__restrict__ counter,
                                                               unsigned *__restrict__ slots,
                                                               unsigned *__restrict__ tile_nslots, int l2hint,
                                                               unsigned *__restrict__ perm, int nids) {
  // perm != nullptr (tensor-core kernels): the records are NOT moved - the pass writes the permutation
  // perm[sorted position] = record offset inside the tile (4 bytes per record instead of 32 read + 32 written) and the
  // evaluation kernel gathers the 32-byte records of a group through it (whole sectors, inside the tile's 1-MB region).
  // The cell ids of up to `nids` records of the tile stay in shared memory between the two loops (u16), so a record's
  // sector is read once.
  // l2hint (experiment): the histogram loop reads the records' cell words with an evict_last policy, the scatter loop
  // re-reads the records evict_first - the tile's second read should then hit L2
  unsigned long long pol_first = 0, pol_last = 0;
  if (l2hint) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  }
  extern __shared__ unsigned cs_sh[];           // cellsort_smem_bytes(SLOTS)
  unsigned *hist = cs_sh;                         // [CELLS_MAX] counts, then write cursors
  unsigned *cstart = cs_sh + CELLS_MAX;           // [CELLS_MAX + 1] first record of each cell, relative to the tile (MODE > 0)
  unsigned *sstart = cs_sh + 2 * CELLS_MAX + 1;   // first slot of each cell (MODE 1) / first group of each cell pair (MODE 2)
  unsigned short *ids = (unsigned short *)(cs_sh + 3 * CELLS_MAX + 2);  // [nids] cell ids of the tile's records (perm mode)
  __shared__ unsigned wsum[32], wsum2[32];
  __shared__ unsigned s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (;;) {
    if (tid == 0) s_tile = atomicAdd(counter, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= (unsigned)g.ntiles) break;
    const unsigned r0 = tile_start[tile], r1 = tile_start[tile + 1];
    for (int i = tid; i < CELLS_MAX; i += 1024) hist[i] = 0;
    __syncthreads();
    for (unsigned rb = r0 + tid; rb < r1; rb += 4 * 1024) {
      int cell[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned r = rb + u * 1024;
        // {idx, cell (+ table rows above bit 12)}; hint: the sector is read again by the scatter loop below - keep it in L2
        int cw = -1;
        if (r < r1) {
          if (l2hint) asm volatile("ld.global.L2::cache_hint.b32 %0, [%1], %2;" : "=r"(cw) : "l"((const int *)(rec + r) + 7), "l"(pol_last));
          else cw = __ldg((const int2 *)(rec + r) + 3).y;
          cw &= CELLS_MAX - 1;
          if (perm && r - r0 < (unsigned)nids) ids[r - r0] = (unsigned short)cw;
        }
        cell[u] = cw;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (cell[u] >= 0) atomicAdd(&hist[cell[u]], 1u);
    }
    __syncthreads();
    // exclusive scans over hist[0..4095] (records) and over ceil(hist / 2) (slots): 4 consecutive bins per thread,
    // warp scan, scan of the 32 warp totals
    unsigned v[4], tsum = 0, tsum2 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = hist[tid * 4 + i];
      tsum += v[i];
      if (MODE == 1) tsum2 += (v[i] + 1) >> 1;
    }
    if (MODE == 2) tsum2 = ((v[0] + v[1] + 7) >> 3) + ((v[2] + v[3] + 7) >> 3);
    if (MODE == 3) {  // a block = the 4 cells of this thread and of its neighbour: the even lane owns it
      const unsigned other = __shfl_xor_sync(0xffffffffu, tsum, 1);
      tsum2 = (lane & 1) ? 0u : (tsum + other + 7) >> 3;
    }
    unsigned inc = tsum, inc2 = tsum2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
      const unsigned y2 = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) {
        inc += y;
        inc2 += y2;
      }
    }
    if (lane == 31) {
      wsum[warp] = inc;
      wsum2[warp] = inc2;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned w = wsum[lane], winc = w, w2 = wsum2[lane], winc2 = w2;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, winc, o);
        const unsigned y2 = __shfl_up_sync(0xffffffffu, winc2, o);
        if (lane >= o) {
          winc += y;
          winc2 += y2;
        }
      }
      wsum[lane] = winc - w;
      wsum2[lane] = winc2 - w2;
      if (MODE && lane == 31) tile_nslots[tile] = winc2;
    }
    __syncthreads();
    unsigned run = wsum[warp] + inc - tsum, run2 = wsum2[warp] + inc2 - tsum2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      hist[tid * 4 + i] = r0 + run;
      if (MODE) cstart[tid * 4 + i] = run;
      if (MODE == 1) {
        sstart[tid * 4 + i] = run2;
        run2 += (v[i] + 1) >> 1;
      }
      if (MODE == 2 && !(i & 1)) {
        sstart[tid * 2 + (i >> 1)] = run2;
        run2 += (v[i] + v[i + 1] + 7) >> 3;
      }
      if (MODE == 3 && i == 0 && !(lane & 1)) sstart[tid >> 1] = run2;
      run += v[i];
    }
    if (MODE && tid == 1023) cstart[CELLS_MAX] = run;
    __syncthreads();
    if (perm) {
      for (unsigned r = r0 + tid; r < r1; r += 1024) {
        const unsigned rel_in = r - r0;
        const int cell = rel_in < (unsigned)nids ? (int)ids[rel_in] : (__ldg((const int2 *)(rec + r) + 3).y & (CELLS_MAX - 1));
        const unsigned pos = atomicAdd(&hist[cell], 1u);
        perm[pos] = rel_in;
        if (MODE >= 2) {
          constexpr int BLK = MODE == 2 ? 2 : 8, BSH = MODE == 2 ? 1 : 3;
          const int c0 = cell & ~(BLK - 1);
          const unsigned rel = pos - r0, rank = rel - cstart[c0];
          if (!(rank & 7u))
            ((uint2 *)slots)[r0 + sstart[cell >> BSH] + (rank >> 3)] = make_uint2(rel, min(8u, cstart[c0 + BLK] - cstart[c0] - rank));
        }
      }
    } else
    for (unsigned rb = r0 + tid; rb < r1; rb += 4 * 1024) {
      double2 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned r = rb + u * 1024;
        if (r < r1) {
          if (l2hint)
            asm volatile("ld.global.L2::cache_hint.v4.f64 {%0, %1, %2, %3}, [%4], %5;"
                         : "=d"(a[u].x), "=d"(a[u].y), "=d"(b[u].x), "=d"(b[u].y) : "l"(rec + r), "l"(pol_first));
          else ld_record(rec + r, a[u], b[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned r = rb + u * 1024;
        if (r < r1) {
          const int cell = __double2hiint(b[u].y) & (CELLS_MAX - 1);
          const unsigned pos = atomicAdd(&hist[cell], 1u);
          st_record(rec2 + pos, a[u].x, a[u].y, b[u].x, b[u].y);
          if (MODE == 1) {
            const unsigned rel = pos - r0, rank = rel - cstart[cell];
            if (!(rank & 1u)) slots[r0 + sstart[cell] + (rank >> 1)] = (rel << 1) | (rel + 1 < cstart[cell + 1] ? 1u : 0u);
          }
          if (MODE >= 2) {
            constexpr int BLK = MODE == 2 ? 2 : 8, BSH = MODE == 2 ? 1 : 3;
            const int c0 = cell & ~(BLK - 1);
            const unsigned rel = pos - r0, rank = rel - cstart[c0];
            if (!(rank & 7u))
              ((uint2 *)slots)[r0 + sstart[cell >> BSH] + (rank >> 3)] = make_uint2(rel, min(8u, cstart[c0 + BLK] - cstart[c0] - rank));
          }
        }
      }
    }
    __syncthreads();  // hist / s_tile are reused by the next tile
  }
}

// per-warp shared-memory scratch of eval_binned_kernel (doubles)
template <int E>
struct BinnedWarpSmem {
  static constexpr int W = 2 * E;
  static constexpr int HB = 32 * 3 * 4;              // Hermite basis per (query, axis): h00 h01 h10' h11'
  static constexpr int MISC = 32 * 2;                // int4 per query: brick offset + 3 table rows
  static constexpr int WAX = 2 * ((3 * W + 7) / 8 * 8);  // double-buffered blended weights [3][W]
  static constexpr int DOUBLES = HB + MISC + WAX;
};

// Lane layout of the stencil contraction: lane = l + W*g, l = tap along axis 0 (contiguous in the brick),
// g = 0..G-1 owns the axis-1 taps m = g, g+G, g+2G, ... ; every lane walks all W taps of axis 2.
//   partial(lane) = wx[l] * sum_n wz[n] * sum_k wy[g+kG] * c[l, g+kG, n]
// so one LDS instruction (immediate offset, no index arithmetic) feeds one DFMA, and the per-query weights a
// lane needs are 1 + KM + W values.  The 32 lane partials of 8 consecutive queries are combined with a
// reduce-scatter butterfly (9 shuffle steps per 8 queries instead of 40).
template <int E, typename TC>
__global__ void __launch_bounds__(512, 1)
    eval_binned_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                       const Record *__restrict__ rec, const WorkItem *__restrict__ items,
                       unsigned *__restrict__ n_items, TC *__restrict__ out, int use_tma) {
  using WS = BinnedWarpSmem<E>;
  using BG = BrickGeom<TC>;
  constexpr int W = 2 * E, G = 32 / W, ACTIVE = W * G, KM = (W + G - 1) / G;
  constexpr int PITCH = BG::PITCH, PLANE = BG::PLANE;
  constexpr int WAXH = WS::WAX / 2;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TC *brick = (TC *)smem_raw;  // [BRICK_SIDE][BRICK_SIDE][PITCH], axis 0 contiguous
  double *tab = (double *)smem_raw + BG::DOUBLES;
  const double *tk[3];
  const double *td[3];
  stage_tables(p, tab, tk, td);
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *wsm = tab + ((tables_smem_doubles(p) + 1) & ~1) + warp * WS::DOUBLES;  // 16-byte aligned
  double *hbs = wsm;                         // [32][3][4]
  int4 *misc = (int4 *)(wsm + WS::HB);       // [32]
  double *wax = wsm + WS::HB + WS::MISC;     // [2][WAXH]
  unsigned long long *bar =
      (unsigned long long *)(tab + ((tables_smem_doubles(p) + 1) & ~1) + nwarps * WS::DOUBLES);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]: double-buffered so a fast thread 0 never overwrites an unread value
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double k0[3] = {p.knot0[0], p.knot0[1], p.knot0[2]};
  const int l = lane % W, gl = lane / W;
  const bool active = lane < ACTIVE;
  // axis-1 taps of this lane; a tap beyond the stencil (only the last k, only when W % G != 0) is predicated off
  // lanes without a tap (32 - W*G of them) alias the last active lane's addresses: a shared-memory broadcast,
  // not a bank conflict; their partial sum is discarded
  const int lane_off = active ? l + gl * PITCH : (W - 1) + (G - 1) * PITCH;
  const int gle = active ? gl : G - 1;
  bool kval[KM];
#pragma unroll
  for (int k = 0; k < KM; ++k) kval[k] = gle + k * G < W;
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
    const int tx = it.bin % g.nt[0], ty = (it.bin / g.nt[0]) % g.nt[1], tz = it.bin / (g.nt[0] * g.nt[1]);
    const int org[3] = {tx * g.T, ty * g.T, tz * g.T};
    if (use_tma) {
      if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of the old brick before the async write
        mbar_expect_tx(bar, BG::ELEMS * (int)sizeof(TC));
        tma_load_3d(brick, &tmap, org[0], org[1], org[2], bar);
      }
      mbar_wait(bar, parity);
      parity ^= 1;
    } else {
      // plain cooperative copy (driver without cuTensorMapEncodeTiled): same brick image, zero outside
      const TC *__restrict__ cg = (const TC *)p.coefs;
      for (int e = threadIdx.x; e < BG::ELEMS; e += blockDim.x) {
        const int bx = e % PITCH, by = (e / PITCH) % BRICK_SIDE, bz = e / PLANE;
        const int gx = org[0] + bx, gy = org[1] + by, gz = org[2] + bz;
        brick[e] = (gx < p.nk[0] && gy < p.nk[1] && gz < p.nk[2]) ? __ldg(cg + gx + (long long)gy * p.s1 + (long long)gz * p.s2) : (TC)0;
      }
      __syncthreads();
    }
    for (unsigned b = warp; b * 32 < it.count; b += nwarps) {
      // ---- phase 1: one query per lane: positions + Hermite bases into the warp's scratch ----------
      const unsigned qi = b * 32 + lane;
      const int nvalid = (int)min(32u, it.count - b * 32);
      const bool have = lane < nvalid;
      unsigned idx = 0;
      __syncwarp();  // the previous batch is done with hbs / misc
      if (have) {
        double2 ra, rb;
        ld_record(rec + it.start + qi, ra, rb);
        const double x[3] = {ra.x, ra.y, rb.x};
        idx = (unsigned)__double2loint(rb.y);
        AxisPos pa[3];
        int4 m;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          locate_from_knot(x[d], k0[d], p.h[d], p.eqs[d], p.nk[d], p.knots[d], pa[d], p.npre[d]);
          const Hermite hb = p.subgrid[d] == CIB200_SG_LINEAR ? linear_basis(pa[d].t) : hermite_basis(pa[d].t);
          double2 *dst = (double2 *)(hbs + (lane * 3 + d) * 4);
          dst[0] = make_double2(hb.h00, hb.h01);
          dst[1] = make_double2(hb.h10, hb.h11);
        }
        m.x = (pa[0].base - org[0]) + (pa[1].base - org[1]) * PITCH + (pa[2].base - org[2]) * PLANE;
        m.y = pa[0].row * W;
        m.z = pa[1].row * W;
        m.w = pa[2].row * W;
        misc[lane] = m;
      }
      __syncwarp();
      // ---- phase 2: the warp resolves the stencils, 8 queries per reduction group --------------------
      double res = 0.0;
      int buf = 0;
      for (int j0 = 0; j0 < nvalid; j0 += 8) {
        double part[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          part[jj] = 0.0;
          if (j < nvalid) {  // warp-uniform
            const int4 m = misc[j];
            double *wq = wax + buf * WAXH;
#pragma unroll
            for (int i0 = 0; i0 < 3 * W; i0 += 32) {
              const int i = i0 + lane;
              if (i < 3 * W) {
                const int a = i / W, col = i - a * W;
                const int rw = a == 0 ? m.y : (a == 1 ? m.z : m.w);
                const double2 *hp = (const double2 *)(hbs + (j * 3 + a) * 4);
                const double2 h0 = hp[0], h1 = hp[1];
                const double *kp = (a == 0 ? tk[0] : (a == 1 ? tk[1] : tk[2])) + rw + col;
                const double *dp = (a == 0 ? td[0] : (a == 1 ? td[1] : td[2])) + rw + col;
                wq[i] = h0.x * kp[0] + h0.y * kp[W] + (h1.x * dp[0] + h1.y * dp[W]);
              }
            }
            __syncwarp();
            const double wx = wq[l];
            double wy[KM];
#pragma unroll
            for (int k = 0; k < KM; ++k) wy[k] = ((G - 1) + k * G < W || kval[k]) ? wq[W + gle + k * G] : 0.0;
            const TC *base = brick + m.x + lane_off;
            // sum_k wy[k] sum_n c[k][n] wz[n], even / odd n in separate chains - the association of
            // eval_cellrun_kernel, so that both kernels return the same bits for the same query
            auto coef = [&](int k, int n) -> double {
              if ((G - 1) + k * G < W) return (double)base[k * G * PITCH + n * PLANE];  // compile-time: every lane group has this tap
              return kval[k] ? (double)base[k * G * PITCH + n * PLANE] : 0.0;          // last, partially populated tap
            };
            double be[KM], bo[KM];
            {
              const double z0 = wq[2 * W], z1 = wq[2 * W + 1];
#pragma unroll
              for (int k = 0; k < KM; ++k) {
                be[k] = coef(k, 0) * z0;
                bo[k] = coef(k, 1) * z1;
              }
            }
#pragma unroll
            for (int n = 2; n < W; n += 2) {
              const double z0 = wq[2 * W + n], z1 = wq[2 * W + n + 1];
#pragma unroll
              for (int k = 0; k < KM; ++k) {
                be[k] = fma(coef(k, n), z0, be[k]);
                bo[k] = fma(coef(k, n + 1), z1, bo[k]);
              }
            }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < KM; ++k) acc = fma(be[k] + bo[k], wy[k], acc);
            part[jj] = active ? acc * wx : 0.0;
            buf ^= 1;
          }
        }
        // reduce-scatter butterfly: after 3 steps lane i holds query (i & 7)'s sum over its 8-lane-class,
        // two more steps complete it; every lane of the warp ends with the total of query j0 + (lane & 7)
        double q4[4], q2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool hi = lane & 1;
          const double send = hi ? part[2 * i] : part[2 * i + 1];
          const double keep = hi ? part[2 * i + 1] : part[2 * i];
          q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool hi = lane & 2;
          const double send = hi ? q4[2 * i] : q4[2 * i + 1];
          const double keep = hi ? q4[2 * i + 1] : q4[2 * i];
          q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        double s1;
        {
          const bool hi = lane & 4;
          const double send = hi ? q2[0] : q2[1];
          const double keep = hi ? q2[1] : q2[0];
          s1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        if ((lane >> 3) == (j0 >> 3)) res = s1;
      }
      if (have) out[idx] = (TC)apply_scale<3>(res, sc);
    }
  }
}

// ---- pass C, cell-sorted variant ---------------------------------------------------------------
// Same lane layout and the same arithmetic as eval_binned_kernel, but the records of a work item arrive sorted by
// cell (pass B2) and every warp owns a CONTIGUOUS range of them: the KM x W coefficients a lane needs from the
// cell's stencil stay in registers for the whole run of queries that share the cell (10^9 queries on 512^3 cells:
// 7.5 per cell), so a query costs its weights and 50 DFMAs instead of 40 + 15 shared-memory loads feeding them.
// The blended weights of 8 queries are produced together (lane = one (axis, tap) pair, fixed for the whole kernel)
// into a per-warp [8][32] scratch, one __syncwarp per 8 queries.
template <int E>
struct CellrunWarpSmem {
  static constexpr int HB = 32 * 3 * 4;  // Hermite basis per (query, axis): h00 h01 h10' h11'
  static constexpr int MISC = 32 * 2;    // int4 per query: brick offset + 3 table rows
  static constexpr int WQ = 8 * 32;      // blended weights of 8 queries, 32 doubles per query ([3][W] used)
  static constexpr int DOUBLES = HB + MISC + WQ;
};

// POSREC: the records carry positions (bin_scatter_kernel<double, true>): fractions and table rows instead of coordinates
template <int E, typename TC, bool POSREC>
__global__ void __launch_bounds__(384, 1)
    eval_cellrun_kernel(const __grid_constant__ CUtensorMap tmap, ItpParams p, BinGeom g, Scale sc,
                        const Record *__restrict__ rec, const WorkItem *__restrict__ items,
                        unsigned *__restrict__ n_items, TC *__restrict__ out, int use_tma) {
  using WS = CellrunWarpSmem<E>;
  using BG = BrickGeom<TC>;
  constexpr int W = 2 * E, G = 32 / W, ACTIVE = W * G, KM = (W + G - 1) / G;
  static_assert(3 * W <= 32, "one lane per (axis, tap) pair");
  constexpr int PITCH = BG::PITCH, PLANE = BG::PLANE;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TC *brick = (TC *)smem_raw;  // [BRICK_SIDE][BRICK_SIDE][PITCH], axis 0 contiguous
  double *smd = (double *)smem_raw;  // the same memory in doubles: tables and per-warp scratch follow the brick
  double *tab = smd + BG::DOUBLES;
  const double *tk[3];
  const double *td[3];
  stage_tables(p, tab, tk, td);
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wbase = BG::DOUBLES + ((tables_smem_doubles(p) + 1) & ~1);  // doubles from the start of shared memory
  double *wsm = smd + wbase + warp * WS::DOUBLES;
  double *hbs = wsm;                        // [32][3][4]
  int *misc = (int *)(wsm + WS::HB);        // [32][4]
  double *wq = wsm + WS::HB + WS::MISC;     // [8][32]
  unsigned long long *bar = (unsigned long long *)(smd + wbase + nwarps * WS::DOUBLES);
  unsigned *slot = (unsigned *)(bar + 1);  // [2]
  double *kt = (double *)(bar + 2);         // [3][32]: the stored knots the tile's queries can select
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double k0[3] = {p.knot0[0], p.knot0[1], p.knot0[2]};
  const int l = lane % W, gl = lane / W;
  const bool active = lane < ACTIVE;
  const int lane_off = active ? l + gl * PITCH : (W - 1) + (G - 1) * PITCH;
  const int gle = active ? gl : G - 1;
  bool kval[KM];
#pragma unroll
  for (int k = 0; k < KM; ++k) kval[k] = gle + k * G < W;
  // weight role of this lane: axis wa, tap wcol; table offsets relative to `brick` keep the loads in the shared window
  const bool wrole = lane < 3 * W;
  const int wa = wrole ? lane / W : 2, wcol = wrole ? lane - wa * W : 0;
  const int tko = (int)((wa == 0 ? tk[0] : (wa == 1 ? tk[1] : tk[2])) - smd) + wcol;
  const int tdo = (int)((wa == 0 ? td[0] : (wa == 1 ? td[1] : td[2])) - smd) + wcol;
  const unsigned total_items = n_items[0];
  unsigned parity = 0;
  double creg[KM][W];  // this lane's share of the current cell's stencil: c[l, gle + k G, n]
#pragma unroll
  for (int k = 0; k < KM; ++k)
#pragma unroll
    for (int n = 0; n < W; ++n) creg[k][n] = 0.0;
  for (unsigned iter = 0;; ++iter) {
    if (threadIdx.x == 0) slot[iter & 1] = atomicAdd(&n_items[1], 1u);
    __syncthreads();  // also: every warp is done with the previous brick
    const unsigned item = slot[iter & 1];
    if (item >= total_items) break;
    const WorkItem it = items[item];
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
      for (int e = threadIdx.x; e < BG::ELEMS; e += blockDim.x) {
        const int bx = e % PITCH, by = (e / PITCH) % BRICK_SIDE, bz = e / PLANE;
        const int gx = org[0] + bx, gy = org[1] + by, gz = org[2] + bz;
        brick[e] = (gx < p.nk[0] && gy < p.nk[1] && gz < p.nk[2]) ? __ldg(cg + gx + (long long)gy * p.s1 + (long long)gz * p.s2) : (TC)0;
      }
      __syncthreads();
    }
    int cur = -1;  // brick offset of the cell held in creg (a new brick invalidates it)
    // contiguous range of this warp, in whole batches of 32 records
    const unsigned per = ((it.count + nwarps - 1) / nwarps + 31u) & ~31u;
    const unsigned w0 = warp * per, w1 = min(it.count, w0 + per);
    // the record of the NEXT batch is fetched while the current one is being evaluated
    double2 ra_n = make_double2(0.0, 0.0), rb_n = ra_n;
    if (w0 + lane < w1) ld_record(rec + it.start + w0 + lane, ra_n, rb_n);
    for (unsigned b0 = w0; b0 < w1; b0 += 32) {
      // ---- phase 1: one query per lane: positions + Hermite bases into the warp's scratch ----------
      const int nvalid = (int)min(32u, w1 - b0);
      const bool have = lane < nvalid;
      unsigned idx = 0;
      const double2 ra = ra_n, rb = rb_n;
      if (b0 + 32 + lane < w1) ld_record(rec + it.start + b0 + 32 + lane, ra_n, rb_n);
      __syncwarp();  // the previous batch is done with hbs / misc
      if (have && POSREC) {
        const double t[3] = {ra.x, ra.y, rb.x};
        idx = (unsigned)__double2loint(rb.y);
        const unsigned code = (unsigned)__double2hiint(rb.y);
        const unsigned lc = code & 4095u, rcode = code >> 12;
        const unsigned rz = rcode / 10000u, ry = (rcode - rz * 10000u) / 100u, rx = rcode - rz * 10000u - ry * 100u;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const Hermite hb = p.subgrid[d] == CIB200_SG_LINEAR ? linear_basis(t[d]) : hermite_basis(t[d]);
          double2 *dst = (double2 *)(hbs + (lane * 3 + d) * 4);
          dst[0] = make_double2(hb.h00, hb.h01);
          dst[1] = make_double2(hb.h10, hb.h11);
        }
        int cx, cy, cz;
        lcell_decode(lc, (unsigned)g.T, cx, cy, cz);
        int4 m;
        m.x = cx + cy * PITCH + cz * PLANE;
        m.y = (int)rx * W;
        m.z = (int)ry * W;
        m.w = (int)rz * W;
        ((int4 *)misc)[lane] = m;
      }
      if (have && !POSREC) {
        const double x[3] = {ra.x, ra.y, rb.x};
        idx = (unsigned)__double2loint(rb.y);
        AxisPos pa[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          locate_from_knot(x[d], k0[d], p.h[d], p.eqs[d], p.nk[d], kt + 32 * d - org[d], pa[d]);
          const Hermite hb = p.subgrid[d] == CIB200_SG_LINEAR ? linear_basis(pa[d].t) : hermite_basis(pa[d].t);
          double2 *dst = (double2 *)(hbs + (lane * 3 + d) * 4);
          dst[0] = make_double2(hb.h00, hb.h01);
          dst[1] = make_double2(hb.h10, hb.h11);
        }
        int4 m;
        m.x = (pa[0].base - org[0]) + (pa[1].base - org[1]) * PITCH + (pa[2].base - org[2]) * PLANE;
        m.y = pa[0].row * W;
        m.z = pa[1].row * W;
        m.w = pa[2].row * W;
        ((int4 *)misc)[lane] = m;
      }
      __syncwarp();
      // ---- phase 2: 8 queries per group: weights, then the contractions, then one reduce-scatter -----
      double res = 0.0;
      for (int j0 = 0; j0 < nvalid; j0 += 8) {
        if (wrole) {
          // two rounds of 4 queries: all shared-memory loads of a round are issued before its first store (the
          // compiler cannot prove that the wq stores do not alias the scratch it reads, and would serialise them)
#pragma unroll
          for (int r4 = 0; r4 < 8; r4 += 4) {
            double2 h0[4], h1[4];
            double ka[4], kb[4], da[4], db[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = min(j0 + r4 + u, nvalid - 1);  // past the end: recompute the last query, result unused
              const int rw = misc[j * 4 + 1 + wa];
              const double2 *hp = (const double2 *)(hbs + (j * 3 + wa) * 4);
              h0[u] = hp[0];
              h1[u] = hp[1];
              const double *kp = smd + tko + rw, *dp = smd + tdo + rw;
              ka[u] = kp[0];
              kb[u] = kp[W];
              da[u] = dp[0];
              db[u] = dp[W];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              wq[(r4 + u) * 32 + lane] = h0[u].x * ka[u] + h0[u].y * kb[u] + (h1[u].x * da[u] + h1[u].y * db[u]);
          }
        }
        __syncwarp();
        double part[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          const int j = j0 + jj;
          part[jj] = 0.0;
          if (j < nvalid) {  // warp-uniform
            const int mx = misc[j * 4];
            if (mx != cur) {  // warp-uniform: first query of a run - fetch the cell's stencil into registers
              cur = mx;
              const TC *base = brick + mx + lane_off;
#pragma unroll
              for (int n = 0; n < W; ++n)
#pragma unroll
                for (int k = 0; k < KM; ++k) {
                  if ((G - 1) + k * G < W) creg[k][n] = (double)base[k * G * PITCH + n * PLANE];
                  else creg[k][n] = kval[k] ? (double)base[k * G * PITCH + n * PLANE] : 0.0;
                }
            }
            const double *wj = wq + jj * 32;
            double wy[KM];
#pragma unroll
            for (int k = 0; k < KM; ++k) wy[k] = ((G - 1) + k * G < W || kval[k]) ? wj[W + gle + k * G] : 0.0;
            double wz[W];
#pragma unroll
            for (int n = 0; n < W; n += 2) {
              const double2 v = *(const double2 *)(wj + 2 * W + n);
              wz[n] = v.x;
              wz[n + 1] = v.y;
            }
            // sum_k wy[k] sum_n c[k][n] wz[n]: 2 KM independent chains (even / odd n) of W/2 DFMAs, then KM more
            double be[KM], bo[KM];
#pragma unroll
            for (int k = 0; k < KM; ++k) {
              be[k] = creg[k][0] * wz[0];
              bo[k] = creg[k][1] * wz[1];
            }
#pragma unroll
            for (int n = 2; n < W; n += 2)
#pragma unroll
              for (int k = 0; k < KM; ++k) {
                be[k] = fma(creg[k][n], wz[n], be[k]);
                bo[k] = fma(creg[k][n + 1], wz[n + 1], bo[k]);
              }
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < KM; ++k) acc = fma(be[k] + bo[k], wy[k], acc);
            part[jj] = active ? acc * wj[l] : 0.0;
          }
        }
        __syncwarp();  // wq is rewritten by the next group
        // reduce-scatter butterfly (see eval_binned_kernel)
        double q4[4], q2[2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool hi = lane & 1;
          const double send = hi ? part[2 * i] : part[2 * i + 1];
          const double keep = hi ? part[2 * i + 1] : part[2 * i];
          q4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool hi = lane & 2;
          const double send = hi ? q4[2 * i] : q4[2 * i + 1];
          const double keep = hi ? q4[2 * i + 1] : q4[2 * i];
          q2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        double s1;
        {
          const bool hi = lane & 4;
          const double send = hi ? q2[0] : q2[1];
          const double keep = hi ? q2[1] : q2[0];
          s1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        if ((lane >> 3) == (j0 >> 3)) res = s1;
      }
      if (have) out[idx] = (TC)apply_scale<3>(res, sc);
    }
  }
}

// spill bin: queries outside the domain (Flat / Line / Throw) or with NaN / Inf coordinates
template <int E, typename TC>
__global__ void __launch_bounds__(256) eval_spill_kernel(ItpParams p, BinGeom g, Scale sc, const Record *__restrict__ rec,
                                                         const unsigned *__restrict__ tile_start, TC *__restrict__ out,
                                                         EvalStatus *status, long long qoff) {
  extern __shared__ double smem[];
  const double *tk[3];
  const double *td[3];
  stage_tables(p, smem, tk, td);
  __syncthreads();
  const unsigned s0 = tile_start[g.ntiles], s1 = tile_start[g.nbin];
  for (unsigned r = s0 + blockIdx.x * blockDim.x + threadIdx.x; r < s1; r += gridDim.x * blockDim.x) {
    const Record rc = rec[r];
    const double x[3] = {rc.x, rc.y, rc.z};
    AxisPos pa[3], pb[3];
    Plan pl = make_plan<3>(p, x, pa, pb);
    double v = CUDART_NAN;
    if (pl.ntask >= 1) {
      double fb = apply_scale<3>(core_thread<TC, 3, E>(p, tk, td, pa), sc);
      v = fb;
      if (pl.ntask == 2) v = finish_line(fb, apply_scale<3>(core_thread<TC, 3, E>(p, tk, td, pb), sc), p.eps_line, pl.dist);
    } else {
      report(status, pl.err, (long long)rc.idx + qoff);
    }
    out[rc.idx] = (TC)v;
  }
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void *tma_encode_tiled_fn();
static EncodeTiledFn encode_fn() { return (EncodeTiledFn)tma_encode_tiled_fn(); }
void *tma_encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return (void *)fn;
}

bool binned_supported(const cib200_itp *h) {
  const ItpParams &p = h->p;
  if (p.ndim != 3) return false;
  const size_t es = h->dtype == CIB200_F32 ? 4 : 8;
  if (p.eqs[0] != p.eqs[1] || p.eqs[1] != p.eqs[2]) return false;
  // (generated >= 10 000-row tables of :linear axes stay in global memory - L2-resident - and are read from there by the
  // weight lanes of eval_binned_kernel; only the cell-sorted kernel needs every table in shared memory)
  if (p.prerange_pos) return false;
  if (p.eqs[0] < 2) return false;
  if (((size_t)p.s1 * es) % 16 || ((size_t)p.s2 * es) % 16) return false;  // TMA global strides: multiples of 16 bytes
  const int T = BRICK_SIDE + 1 - 2 * p.eqs[0];
  if (T < 4) return false;
  long long nt = 1;
  for (int d = 0; d < 3; ++d) nt *= (p.nk[d] - 2 * p.eqs[d] + 1 + T - 1) / T;
  // the evaluation kernels stage the brick, the interpolant's distinct tables and per-warp scratch (>= 4 warps): an
  // interpolant whose tables do not fit (e.g. :b13 with three different derivative orders: six 14.5 kB tables) stays
  // on the gather kernel instead of failing at launch time
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t need = (size_t)(es == 4 ? BrickGeom<float>::BYTES : BrickGeom<double>::BYTES) + 8 * ((size_t)((tables_smem_doubles(p) + 1) & ~1) + 4 * (size_t)(32 * 3 * 4 + 64 + 256)) + 32 + 8 * 96;
  if (need > (size_t)max_smem) return false;
  return nt + 1 <= MAX_BINS_GLOBAL && encode_fn() != nullptr;
}

long long binned_tile_count(const cib200_itp *h) {
  const ItpParams &p = h->p;
  const int T = BRICK_SIDE + 1 - 2 * p.eqs[0];
  long long nt = 1;
  for (int d = 0; d < 3; ++d) nt *= (p.nk[d] - 2 * p.eqs[d] + 1 + T - 1) / T;
  return nt;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

long long binned_scratch_bytes(const cib200_itp *h, long long nq) {
  if (!binned_supported(h)) return 0;
  const ItpParams &p = h->p;
  const int T = BRICK_SIDE + 1 - 2 * p.eqs[0];
  long long nt = 1;
  for (int d = 0; d < 3; ++d) nt *= (p.nk[d] - 2 * p.eqs[d] + 1 + T - 1) / T;
  const size_t nbin = (size_t)nt + 1, nblk = nbin > (size_t)MAX_BINS ? 1 : (size_t)sm_count();
  size_t b = 2 * align_up(sizeof(Record) * (size_t)nq, 256);
  b += align_up(4 * nblk * nbin, 256) + align_up(4 * (nbin + 1), 256);
  b += align_up(sizeof(WorkItem) * ((size_t)nt + (size_t)nq / CHUNK_QUERIES + 1), 256) + 256;
  b += align_up(8 * (size_t)nq, 256) + align_up(4 * nbin, 256);  // slot / group list + slots per tile (cell-sorted kernels)
  return (long long)b;
}

void binned_destroy(cib200_itp *h) {
  delete (BinState *)h->bin_state;
  h->bin_state = nullptr;
}

template <int E, typename TC>
static int launch_pass_cd(cib200_itp *h, BinState *bs, const Scale &sc, TC *out, long long qoff, bool cellsort, int lq,
                          cudaStream_t st) {
  const ItpParams &p = h->p;
  const int cs_hint = (tune_env("CIB200_CELLSORT_L2HINT") && tune_env("CIB200_CELLSORT_L2HINT")[0] == '1') ? 1 : 0;
  if (lq) {
    // lane-per-query kernel on (tile, cell)-sorted records (lq = 2: two queries of a cell per lane, slot list)
    CIB_CUDA(cudaFuncSetAttribute(bin_cellsort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cellsort_smem_bytes(1)));
    CIB_CUDA(cudaFuncSetAttribute(bin_cellsort_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cellsort_smem_bytes(2, true)));
    CIB_CUDA(cudaFuncSetAttribute(bin_cellsort_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cellsort_smem_bytes(3, true)));
    // (tensor-core kernels: the cell sort writes a permutation into the second record array's memory, records stay put)
    const bool use_perm = !(tune_env("CIB200_CELLSORT_PERM") && tune_env("CIB200_CELLSORT_PERM")[0] == '0');
    unsigned *permp = (use_perm && lq >= 3) ? (unsigned *)bs->rec2 : nullptr;
    bs->perm = permp;
    if (lq == 4)  // tensor-core kernel on 2 x 2 x 2 cell blocks: group list per block
      bin_cellsort_kernel<3><<<bs->nblk, 1024, cellsort_smem_bytes(3, use_perm), st>>>(bs->g, bs->rec, bs->rec2, bs->tile_start, bs->n_items + 2,
                                                                                      bs->slots, bs->tile_nslots, cs_hint, permp, CELLSORT_NIDS);
    else if (lq == 3)  // tensor-core kernel: group list
      bin_cellsort_kernel<2><<<bs->nblk, 1024, cellsort_smem_bytes(2, use_perm), st>>>(bs->g, bs->rec, bs->rec2, bs->tile_start, bs->n_items + 2,
                                                                                      bs->slots, bs->tile_nslots, cs_hint, permp, CELLSORT_NIDS);
    else if (lq == 2)
      bin_cellsort_kernel<1><<<bs->nblk, 1024, cellsort_smem_bytes(1), st>>>(bs->g, bs->rec, bs->rec2, bs->tile_start,
                                                                            bs->n_items + 2, bs->slots, bs->tile_nslots, cs_hint, nullptr, 0);
    else
      bin_cellsort_kernel<0><<<bs->nblk, 1024, cellsort_smem_bytes(0), st>>>(bs->g, bs->rec, bs->rec2, bs->tile_start,
                                                                            bs->n_items + 2, nullptr, nullptr, cs_hint, nullptr, 0);
    CIB_LAUNCHED();
    prof_mark(h, 2, st);
    int rc;
    if (lq == 4) rc = launch_eval_celloct(h, bs, sc, out, st);
    else if (lq == 3) rc = launch_eval_cellmma(h, bs, sc, out, st);
    else if constexpr (sizeof(TC) == 8) rc = launch_eval_lanequery(h, bs, sc, out, lq, st);
    else rc = CIB200_EUNSUPPORTED;
    if (rc != CIB200_OK) return rc;
  }
  using WS = BinnedWarpSmem<E>;
  int nwarps = cellsort ? 12 : 16;
  const size_t per_warp = cellsort ? (size_t)CellrunWarpSmem<5>::DOUBLES : (size_t)WS::DOUBLES;
  auto smem_for = [&](int nw) {
    return (size_t)BrickGeom<TC>::BYTES + 8 * ((size_t)((tables_smem_doubles(p) + 1) & ~1) + (size_t)nw * per_warp) + 32 + 8 * 96;
  };
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  while (nwarps > 4 && smem_for(nwarps) > (size_t)max_smem) nwarps -= 4;
  if (smem_for(nwarps) > (size_t)max_smem) {
    set_error("binned strategy: shared memory budget exceeded");
    return CIB200_EUNSUPPORTED;
  }
  const char *no_tma = tune_env("CIB200_BINNED_NO_TMA");
  const int use_tma = (no_tma && no_tma[0] == '1') ? 0 : 1;
  if constexpr (E == 5) {
   if (cellsort && !lq) {
    bin_cellsort_kernel<0><<<bs->nblk, 1024, cellsort_smem_bytes(0), st>>>(bs->g, bs->rec, bs->rec2, bs->tile_start, bs->n_items + 2, nullptr, nullptr, cs_hint, nullptr, 0);
    CIB_LAUNCHED();
    prof_mark(h, 2, st);
    auto kern = eval_cellrun_kernel<E, TC, true>;  // position records (launch_eval_binned picks the matching scatter)
    CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(nwarps)));
    kern<<<bs->nblk, nwarps * 32, smem_for(nwarps), st>>>(bs->tmap, p, bs->g, sc, bs->rec2, bs->items, bs->n_items, out,
                                                          use_tma);
   }
  }
  if (!(E == 5 && cellsort) && !lq) {
    prof_mark(h, 2, st);
    auto kern = eval_binned_kernel<E, TC>;
    CIB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(nwarps)));
    kern<<<bs->nblk, nwarps * 32, smem_for(nwarps), st>>>(bs->tmap, p, bs->g, sc, bs->rec, bs->items, bs->n_items, out,
                                                          use_tma);
  }
  if (!lq) CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  prof_mark(h, 3, st);
  auto spill = eval_spill_kernel<E, TC>;
  const size_t tsm = 8 * (size_t)tables_smem_doubles(p);
  if (tsm > 48 * 1024) CIB_CUDA(cudaFuncSetAttribute(spill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm));
  spill<<<bs->nblk, 256, tsm, st>>>(p, bs->g, sc, bs->rec, bs->tile_start, out, h->d_status, qoff);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  prof_mark(h, 4, st);
  return CIB200_OK;
}

int launch_eval_binned(cib200_itp *h, const Coords &c, void *out, long long nq, cudaStream_t st) {
  if (nq <= 0) return CIB200_OK;
  if (nq >= 0xffffffffLL) {
    set_error("binned strategy: at most 2^32-2 queries per call");
    return CIB200_EUNSUPPORTED;
  }
  const ItpParams &p = h->p;
  BinState *bs = (BinState *)h->bin_state;
  if (!bs) {
    bs = new BinState();
    memset(bs, 0, sizeof *bs);
    BinGeom &g = bs->g;
    g.T = BRICK_SIDE + 1 - 2 * p.eqs[0];
    g.ntiles = 1;
    for (int d = 0; d < 3; ++d) {
      g.nt[d] = (p.nk[d] - 2 * p.eqs[d] + 1 + g.T - 1) / g.T;
      g.ntiles *= g.nt[d];
    }
    g.nbin = g.ntiles + 1;
    bs->nblk = sm_count();
    cuuint64_t gdim[3] = {(cuuint64_t)p.nk[0], (cuuint64_t)p.nk[1], (cuuint64_t)p.nk[2]};
    const bool f32 = h->dtype == CIB200_F32;
    const cuuint64_t esz = f32 ? 4 : 8;
    cuuint64_t gstr[2] = {(cuuint64_t)p.s1 * esz, (cuuint64_t)p.s2 * esz};
    cuuint32_t box[3] = {(cuuint32_t)(f32 ? BrickGeom<float>::PITCH : BrickGeom<double>::PITCH), BRICK_SIDE, BRICK_SIDE};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(&bs->tmap, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(p.coefs), gdim, gstr, box,
                             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete bs;
      set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
      return CIB200_ECUDA;
    }
    if (!f32) {  // brick rows of 28 doubles: the B-fragment loads of the tensor-core kernel are conflict-free at that pitch
      cuuint32_t box28[3] = {CELLMMA_PITCH, BRICK_SIDE, BRICK_SIDE};
      bs->tmap28_valid = encode_fn()(&bs->tmap28, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(p.coefs), gdim, gstr, box28, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    h->bin_state = bs;
  }
  // scratch
  const size_t need = (size_t)binned_scratch_bytes(h, nq);
  if (h->bin_scratch_bytes < need) {
    if (h->d_bin_scratch) CIB_CUDA(cudaFree(h->d_bin_scratch));
    h->d_bin_scratch = nullptr;
    h->bin_scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&h->d_bin_scratch, need);
    if (e != cudaSuccess) {
      set_error("binned strategy: cannot allocate %zu bytes of scratch: %s", need, cudaGetErrorString(e));
      return CIB200_ENOMEM;
    }
    h->bin_scratch_bytes = need;
  }
  {
    const BinGeom &g = bs->g;
    char *base = (char *)h->d_bin_scratch;
    bs->rec = (Record *)base;
    base += align_up(sizeof(Record) * (size_t)nq, 256);
    bs->rec2 = (Record *)base;
    base += align_up(sizeof(Record) * (size_t)nq, 256);
    bs->hist = (unsigned *)base;
    base += align_up(4 * (size_t)(g.nbin > MAX_BINS ? 1 : bs->nblk) * g.nbin, 256);
    bs->tile_start = (unsigned *)base;
    base += align_up(4 * ((size_t)g.nbin + 1), 256);
    bs->items = (WorkItem *)base;
    base += align_up(sizeof(WorkItem) * ((size_t)g.ntiles + (size_t)nq / CHUNK_QUERIES + 1), 256);
    bs->n_items = (unsigned *)base;
    base += 256;
    bs->slots = (unsigned *)base;
    base += align_up(8 * (size_t)nq, 256);
    bs->tile_nslots = (unsigned *)base;
  }
  const BinGeom &g = bs->g;
  Scale sc;
  for (int d = 0; d < 3; ++d) sc.s[d] = jl_pow_int(-1.0 / p.h[d], h->deriv[d]);
  sc.perdim2d = 0;
  const bool f32 = h->dtype == CIB200_F32;
  // cell-sorted variant: pays one more streaming pass over the records (64 B/query) to keep each cell's stencil in
  // registers across the queries that share it - worth it from about one query per two cells upwards (measured)
  double ncell = 1.0;
  for (int d = 0; d < 3; ++d) ncell *= (double)(p.nk[d] - 2 * p.eqs[d] + 1);
  // (b5 only: a lane's share of a wider stencil, KM x W coefficients, does not fit in registers)
  bool cellsort = p.eqs[0] == 5 && (double)nq >= 0.5 * ncell && g.T * g.T * g.T <= CELLS_MAX;
  if (const char *e = tune_env("CIB200_BINNED_CELLSORT")) cellsort = (e[0] == '1') && p.eqs[0] == 5 && g.T * g.T * g.T <= CELLS_MAX;
  for (int d = 0; d < 3; ++d)
    if (p.npre[d] != CIB_N_PRE) cellsort = false;  // the cell-run kernel addresses its tables inside shared memory
  // cell-sorted strategies for :b5-width stencils (every table staged in shared memory):
  //   the tensor-core kernels (cib200_cellmma.cu), see below
  //   >= 0.5 query per cell and no tensor-core kernel for this interpolant: eval_cellrun_kernel
  // the lane-per-query kernel (cib200_lanequery.cu: 195 ms per 10^9 where the tensor-core kernel takes 166) is kept for
  // experiments builds (CIB200_LANEQUERY=1|2)
  int lq = 0;
  {
    // tensor-core kernel: any :b-series width (tile side 26 - 2E is even: aligned cell pairs), every table staged
    bool mma_ok = (f32 || bs->tmap28_valid) && g.T % 2 == 0 && g.T * g.T * g.T <= CELLS_MAX && p.eqs[0] >= 5;
    for (int d = 0; d < 3; ++d)
      if (p.npre[d] != CIB_N_PRE) mma_ok = false;
    const int kind = mma_ok ? cellmma_kind(p, f32) : 0;
    // Rows of a product from a 2 x 2 x 2 cell BLOCK from 0.4 query per cell, from an aligned cell PAIR where that is
    // faster: the window of a block is 2E + 1 wide in both column indices ((2E + 1)^2 / (2E)^2 more DMMAs per group: 21 %
    // for :b5, 11 % for :b13) but eight cells fill its 8 rows.  Measured on 512^3, whole call (tools/ab_pairs_vs_blocks.py,
    // tools/ab_cellmma_gen.py): :b5 2 per cell pairs 77.1 / blocks 67.0 ms, 3 per cell 96.5 / 96.0, 4 per cell 122.6 / 126.2;
    // :b7 1.1 per cell 104 / 73, 3.7 per cell 189 / 198; :b9 1.9 per cell 168 / 121; :b11 237 / 175; :b13 0.9 per cell
    // 253 / 148 (binned kernel 345), 1.9: 319 / 248, 3.7: 472 / 449, 7.5: 831 / 824.  Against the cell-run kernel (:b5):
    // 2^26 queries 25.6 against 29.8 ms, 2^28 67 against 90.
    const double pairs_from = p.eqs[0] <= 6 ? 3.0 : (p.eqs[0] == 7 ? 5.0 : 1e300);
    if (kind) lq = (double)nq >= pairs_from * ncell ? 3 : ((double)nq >= 0.4 * ncell ? 4 : 0);
    if (const char *e = tune_env("CIB200_LANEQUERY")) {
      lq = atoi(e);
      if (lq == 3 && !kind) lq = 0;
      if (lq == 4 && !kind) lq = 0;
      if ((lq == 1 || lq == 2) && !(cellsort && !f32)) lq = 0;
      if (f32 && lq != 3 && lq != 4) lq = 0;
      if (f32 && kind != 1) lq = 0;
    }
  }
  if ((lq == 1 || lq == 2) && !(nq < 0x7fffffffLL && lanequery_smem_bytes(p, false) <= 227 * 1024)) lq = 0;
  h->last_strategy = lq == 4 ? CIB200_STRATEGY_CELLMMA_BLOCK : lq == 3 ? CIB200_STRATEGY_CELLMMA : (lq ? CIB200_STRATEGY_LANEQUERY : (cellsort ? CIB200_STRATEGY_CELLRUN : CIB200_STRATEGY_BINNED));
  const bool global_hist = g.nbin > MAX_BINS;
  const int hrows = global_hist ? 1 : bs->nblk;
  const size_t hsm = global_hist ? 0 : 4 * (size_t)g.nbin;
  if (global_hist) CIB_CUDA(cudaMemsetAsync(bs->hist, 0, 4 * (size_t)g.nbin, st));
  auto kcount = f32 ? bin_count_kernel<float> : bin_count_kernel<double>;
  // (the tensor-core and the cell-run kernel read position records: table rows and fractions computed once, in the scatter pass)
  const bool posrec = lq == 3 || lq == 4 || (cellsort && lq == 0);
  auto kscatter = f32 ? (posrec ? bin_scatter_kernel<float, true> : bin_scatter_kernel<float, false>)
                      : (posrec ? bin_scatter_kernel<double, true> : bin_scatter_kernel<double, false>);
  CIB_CUDA(cudaFuncSetAttribute(kcount, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
  CIB_CUDA(cudaFuncSetAttribute(kscatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsm));
  prof_mark(h, 0, st);
  kcount<<<bs->nblk, 1024, hsm, st>>>(p, g, c, nq, bs->hist, global_hist ? 1 : 0);
  CIB_LAUNCHED();
  bin_totals_kernel<<<(g.nbin + 255) / 256, 256, 0, st>>>(g, hrows, bs->hist, bs->tile_start);
  CIB_LAUNCHED();
  bin_scan_kernel<<<1, 1024, 0, st>>>(g, bs->tile_start, bs->items, bs->n_items);
  CIB_LAUNCHED();
  bin_offsets_kernel<<<(g.nbin + 255) / 256, 256, 0, st>>>(g, hrows, bs->hist, bs->tile_start);
  CIB_LAUNCHED();
  int scatter_mode = (tune_env("CIB200_SCATTER_SHARED") && tune_env("CIB200_SCATTER_SHARED")[0] == '1') ? 3 : 0;
  if (tune_env("CIB200_SCATTER_HINTS") && tune_env("CIB200_SCATTER_HINTS")[0] == '1') scatter_mode |= 4;
  kscatter<<<bs->nblk, 1024, hsm, st>>>(p, g, c, nq, bs->hist, bs->rec, global_hist ? (3 | (scatter_mode & 4)) : scatter_mode);
  CIB_LAUNCHED();
  CIB_CUDA(cudaGetLastError());
  prof_mark(h, 1, st);
  switch (p.eqs[0]) {
    // a-series (:a3 eqs 2, :a4 / :a5 eqs 3, :a7 eqs 4): the same binned kernel on 4^3 .. 8^3 stencils
    case 2: return f32 ? launch_pass_cd<2, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<2, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 3: return f32 ? launch_pass_cd<3, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<3, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 4: return f32 ? launch_pass_cd<4, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<4, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 5: return f32 ? launch_pass_cd<5, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<5, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 6: return f32 ? launch_pass_cd<6, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<6, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 7: return f32 ? launch_pass_cd<7, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<7, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 8: return f32 ? launch_pass_cd<8, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<8, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
    case 9: return f32 ? launch_pass_cd<9, float>(h, bs, sc, (float *)out, c.qoff, cellsort, lq, st)
                   : launch_pass_cd<9, double>(h, bs, sc, (double *)out, c.qoff, cellsort, lq, st);
  }
  set_error("binned strategy: unsupported kernel");
  return CIB200_EUNSUPPORTED;
}

}  // namespace cib
