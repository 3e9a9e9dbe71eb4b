// cib200_pipe.cuh - mbarrier / TMA / programmatic-dependent-launch helpers shared by the persistent sweep kernels
// (sm_90+ PTX; SASS: SYNCS.*, UTMALDG, ACQBULK / griddepcontrol).
#pragma once
#include <cuda.h>
#include <stdio.h>

#include "cib200_internal.cuh"

namespace cib {
namespace pipe {

__device__ __forceinline__ unsigned saddr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(saddr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint parks the warp in hardware; the trip counter turns a broken pipeline into a trap
// instead of a hung device
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity, int tag = 0) {
  unsigned done, spins = 0;
  do {
    if (++spins > (1u << 22)) {
#ifdef CIB200_EXPERIMENTS
      if ((threadIdx.x & 31) == 0)
        printf("[cib200 pipe] barrier timeout: tag %d block %d warp %d parity %u\n", tag, (int)blockIdx.x, (int)(threadIdx.x >> 5), parity);
      return;  // experiments build: carry on with garbage so that the message is flushed
#endif
      __trap();
    }
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(saddr(bar)), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}
// x / d for the small counters of the pipelines: magic = floor(2^32 / d) + 1 (exact for x < 2^32 / d), 0 for d = 1
__device__ __forceinline__ unsigned div_u(unsigned x, unsigned magic) { return magic ? __umulhi(x, magic) : x; }
inline unsigned div_magic_for(unsigned d) { return d <= 1 ? 0u : (unsigned)((1ULL << 32) / d) + 1u; }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          saddr(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(saddr(bar))
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, int x, int y, int z, unsigned long long *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          saddr(dst)),
      "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(saddr(bar))
      : "memory");
}

// Programmatic dependent launch: a pass is launched while its predecessor still runs; its prologue (barrier init,
// constants) overlaps the predecessor's tail, and the first access to the predecessor's output waits here.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// f64 tensor map, no swizzle, zero fill outside the tensor (FLOAT_OOB_FILL_NONE = zeros)
inline int encode_f64(CUtensorMap *tm, int rank, const void *base, const cuuint64_t *gdim, const cuuint64_t *gstride_bytes,
                      const cuuint32_t *box) {
  EncodeTiledFn fn = (EncodeTiledFn)tma_encode_tiled_fn();
  if (!fn) return -1;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

}  // namespace pipe
}  // namespace cib
