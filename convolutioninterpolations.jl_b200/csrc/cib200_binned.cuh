// cib200_binned.cuh - declarations shared by the tile-binned point-evaluation strategies (cib200_binned.cu: binning
// passes + eval_binned / eval_cellrun kernels; cib200_lanequery.cu: the lane-per-query evaluation kernel).
#pragma once
#include <cuda.h>

#include "cib200_evalcore.cuh"

namespace cib {

static constexpr int BRICK_SIDE = 25;        // T + 2E - 1
static constexpr int BRICK_PITCH = 26;       // inner TMA box extent: 26 * 8 B is a multiple of 16 B
static constexpr int BRICK_DOUBLES = BRICK_PITCH * BRICK_SIDE * BRICK_SIDE;
// Float32 mode (coefficients / queries / results in float, arithmetic in double): the brick is staged as floats
// (inner extent 28: 28 * 4 B is a multiple of 16 B) and converted when a lane loads its stencil share
template <typename T>
struct BrickGeom {
  static constexpr int PITCH = sizeof(T) == 8 ? 26 : 28;
  static constexpr int PLANE = PITCH * BRICK_SIDE;
  static constexpr int ELEMS = PLANE * BRICK_SIDE;
  static constexpr int BYTES = (ELEMS * (int)sizeof(T) + 15) & ~15;
  static constexpr int DOUBLES = BYTES / 8;  // where the staged tables start, in doubles from the brick
};
static constexpr int CHUNK_QUERIES = 16384;  // queries per work item
static constexpr int BIN_U = 4;              // queries per thread and iteration in the binning passes
static constexpr int MAX_BINS = 49152;       // u32 histogram must fit one CTA's shared memory (192 KB)
// more tiles than that (:b13 on 512^3: 8^3-cell tiles, 262 144 of them): ONE histogram / ONE write cursor per tile in
// global memory, shared by all CTAs (L2 atomics) instead of private ones per CTA in shared memory
static constexpr int MAX_BINS_GLOBAL = 1 << 21;
static constexpr int CELLMMA_PITCH = 28;       // brick row pitch of the tensor-core kernel (doubles)
static constexpr int CELLS_MAX = 4096;        // cells of a tile the cell sort handles: T^3 with T = 16 (E = 5)

struct BinGeom {
  int T;
  int nt[3];
  int ntiles;  // spill bin = ntiles
  int nbin;    // ntiles + 1
};

struct __align__(32) Record {
  double x, y, z;
  unsigned idx;
  unsigned pad;
};

struct WorkItem {
  unsigned bin, start, count;
};

// A record is one 32-byte sector: moved with ONE 256-bit access (sm_100: LDG/STG.E.ENL2.256).  Two 16-byte stores
// to a scattered sector reach L2 as two partial-sector writes, which is what limited the scatter pass.
__device__ __forceinline__ void st_record(Record *dst, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ld_record(const Record *src, double2 &a, double2 &b) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(src));
}

struct BinState {
  BinGeom g;
  CUtensorMap tmap;
  int nblk;           // CTAs of passes A / B (= SM count)
  size_t cap_q;       // queries the scratch can hold
  // scratch layout (device)
  Record *rec;
  Record *rec2;       // records sorted by (tile, cell) - cell-sorted variant only
  unsigned *hist;     // [nblk][nbin] counts -> write cursors
  unsigned *tile_start;  // [nbin + 1]
  WorkItem *items;
  unsigned *n_items;  // [0] = number of items, [1] = work counter
  size_t max_items;
  unsigned *slots;        // [nq] slot list (u32) of the lane-per-query kernel / group list (uint2) of the tensor-core kernel
  unsigned *tile_nslots;  // [nbin] slots / groups of each tile
  unsigned *perm;         // tensor-core kernels: sorted position -> record offset in the tile (nullptr: records were moved to rec2)
  CUtensorMap tmap28;     // the same coefficient array with brick rows of 28 elements (tensor-core kernel, Float64)
  int tmap28_valid;
};

// lane-per-query evaluation kernel (cib200_lanequery.cu): records sorted by (tile, cell) in bs->rec2; qpl = 2 also needs
// the slot list of bin_cellsort_kernel<true>
size_t lanequery_smem_bytes(const ItpParams &p, bool f32);
int launch_eval_lanequery(cib200_itp *h, BinState *bs, const Scale &sc, void *out, int qpl, cudaStream_t st);
// tensor-core evaluation kernel (cib200_cellmma.cu): records sorted by (tile, cell) + the group list of bin_cellsort_kernel<2>
size_t cellmma_smem_bytes(const ItpParams &p);
int cellmma_kind(const ItpParams &p, bool f32 = false);  // 0: not supported, 1: tuned :b5 kernel, 2: generic kernel (any :b width, several table pairs)
int launch_eval_cellmma(cib200_itp *h, BinState *bs, const Scale &sc, void *out, cudaStream_t st);
int launch_eval_celloct(cib200_itp *h, BinState *bs, const Scale &sc, void *out, cudaStream_t st);  // 2 x 2 x 2 cell blocks (:b5, one table pair)

// Cell of a query inside its tile.  Cells are numbered BLOCK-MAJOR: the tile (even side T) is cut into 2 x 2 x 2 blocks,
// lcell = block * 8 + (cz & 1) * 4 + (cy & 1) * 2 + (cx & 1), blocks in x-fastest order.  After the cell sort the records of
// an aligned cell pair along axis 0 (lcell >> 1) and of a whole block (lcell >> 3) are contiguous: the tensor-core kernels
// take the rows of one product from such a pair / block.
__device__ __forceinline__ unsigned lcell_encode(int cx, int cy, int cz, int T) {
  const int hb = T >> 1;
  return (unsigned)(((((cz >> 1) * hb + (cy >> 1)) * hb + (cx >> 1)) << 3) | ((cz & 1) << 2) | ((cy & 1) << 1) | (cx & 1));
}
__device__ __forceinline__ void lcell_decode(unsigned lc, unsigned T, int &cx, int &cy, int &cz) {
  const unsigned hb = T >> 1, blk = lc >> 3;
  const unsigned bz = blk / (hb * hb), by = (blk - bz * hb * hb) / hb, bx = blk - bz * hb * hb - by * hb;
  cx = (int)(2 * bx + (lc & 1u));
  cy = (int)(2 * by + ((lc >> 1) & 1u));
  cz = (int)(2 * bz + ((lc >> 2) & 1u));
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- TMA / mbarrier helpers (sm_90+ PTX; SASS: UTMALDG, SYNCS) -----------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned done;
  unsigned spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();  // a lost TMA must not hang the device
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace cib
