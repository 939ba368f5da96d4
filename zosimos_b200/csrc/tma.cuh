// tma.cuh -- mbarrier / TMA (cp.async.bulk.tensor) primitives and the host-side tensor-map builder
// shared by the kernels that stage source tiles in shared memory (gather.cu, frame_pipeline.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "zos_internal.h"

namespace zos {

struct TensorMaps {  // passed as a __grid_constant__ kernel parameter; the TMA unit reads it in place
  CUtensorMap m0, m1, m2;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait with a suspend-time hint: the thread sleeps in hardware (up to ~hint_ns) instead of burning issue
// slots in a spin loop; still returns false on time-out so callers keep their run-away guard.
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

// ---------------- host: tensor maps ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode(zos_ctx* ctx) {
  if (!ctx->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      ctx->encode_tiled = fn;
  }
  return (EncodeTiledFn)ctx->encode_tiled;
}

static inline bool make_map(zos_ctx* ctx, CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, void* base, uint64_t w_elems, uint64_t h,
                     uint64_t pitch, uint64_t frames, uint64_t frame_stride, uint32_t box_w_elems, uint32_t box_h) {
  EncodeTiledFn enc = get_encode(ctx);
  if (!enc) return false;
  if (((uintptr_t)base & 15) || (pitch & 15) || (frame_stride & 15)) return false;
  if (box_w_elems > 256 || box_h > 256 || ((uint64_t)box_w_elems * elem_bytes) % 16) return false;
  cuuint64_t dims[3] = {w_elems, h, frames};
  cuuint64_t strides[2] = {pitch, frames > 1 ? frame_stride : pitch * h};
  cuuint32_t box[3] = {box_w_elems, box_h, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;  // none / 64 B / 256 B measured the same within noise
  CUresult r = enc(m, dt, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}


}  // namespace zos
