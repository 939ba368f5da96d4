// zos_internal.h -- shared between the host runtime and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <set>
#include <string>
#include <vector>

#include "../../include/zosimos_cuda.h"

struct zos_buf {
  void* ptr = nullptr;
  uint64_t size = 0;  // bytes the caller asked for
  uint64_t cap = 0;   // bytes of the arena block behind it (its size class)
};

struct zos_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  uint64_t launches = 0;
  uint32_t flags = 0;
  std::string err;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled, resolved lazily through the runtime
  std::vector<void*> scratch;    // device scratch owned by the ctx (tensor maps etc.)
  int* fault_host = nullptr;     // mapped pinned word: kernels set it when an mbarrier wait ran away; zos_sync reports it
  int* fault_dev = nullptr;      // device view of fault_host
  float* work_counter = nullptr; // device word: tile dispenser of dynamically scheduled kernels (zeroed before each launch)
  std::set<const void*> smem_configured;  // kernels whose dynamic shared memory limit was raised ON THIS CONTEXT'S DEVICE
  std::map<std::string, struct zos_dynamic*> dynamic_cache;  // NVRTC-compiled plugins by source text (dynamic.cu)
  // Device arena: freed blocks wait here, by size class, for the next allocation of that class.  All work of a context is
  // ordered on its one stream, so a block can be handed out again without waiting for the kernels that still read it
  // (they were enqueued before anything the next owner enqueues).  This is what the reference's pool cache does for
  // buffers and textures between runs (pool.rs:93-99, run.rs:1312-1347, 2876-2942).
  std::map<uint64_t, std::vector<void*>> arena_free;
  zos_arena_stats arena{};
  // affine_f16.cu: the last staged-box width worked out for (minimum width, per-lane source step): launches of one program repeat it
  struct { int min_w = 0; float dxl = 0.0f, dyl = 0.0f; int w = 0; } box_cache;
};

namespace zos {

// What a kernel needs to know about one image operand.
struct DevImage {
  uint8_t* p0;  // pixels / Y
  uint8_t* p1;  // U or UV
  uint8_t* p2;  // V
  uint64_t pitch, cpitch;
  uint64_t bstride, cbstride;  // per-frame strides
  int32_t w, h;
  int32_t bpp;      // bytes per texel of plane 0
  uint32_t block;   // ZOS_BLOCK_*
  zos_texfmt fmt;   // transfer of the colour for planar YUV
  float kr, kb;
  float yoff, ysc, csc, r_cr, b_cb, g_cr, g_cb;  // planar YUV: range scaling and matrix, prepared on the host
  uint32_t full_range, chroma_filter;
};

// n / d for any 32-bit n (Granlund-Montgomery round-up method), d >= 1
struct FastDiv {
  uint32_t d, m, l;
};
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) l++;
  f.l = l;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  return f;
}

zos_status fail(zos_ctx* ctx, zos_status code, const char* fmt, ...);
zos_status check_cuda(zos_ctx* ctx, cudaError_t e, const char* what);
zos_status make_dev_image(zos_ctx* ctx, const zos_image* img, DevImage* out, const char* name);
zos_status validate_steps(zos_ctx* ctx, const zos_step* steps, uint32_t n);
int grid_for(const zos_ctx* ctx, uint64_t work_items, int threads, int ctas_per_sm);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remembered per context, not per process
template <typename F>
inline cudaError_t ensure_dyn_smem(zos_ctx* ctx, F func, int bytes) {
  const void* key = reinterpret_cast<const void*>(func);
  if (ctx->smem_configured.count(key)) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) ctx->smem_configured.insert(key);
  return e;
}

// kernel launchers (one per .cu)
zos_status launch_rowwise(zos_ctx* ctx, const DevImage* below, const DevImage* above, const DevImage& dst,
                          const zos_compose_params* cp, const zos_step* steps, uint32_t nsteps, uint32_t batch);
bool rowwise_u8_eligible(const DevImage* below, const DevImage* above, const DevImage& dst, const zos_compose_params* cp,
                         const zos_step* steps, uint32_t nsteps);
zos_status launch_rowwise_u8(zos_ctx* ctx, const DevImage* below, const DevImage* above, const DevImage& dst,
                             const zos_compose_params* cp, const zos_step* steps, uint32_t nsteps, uint32_t batch);
bool rowwise_can_compose(const DevImage& below, const DevImage& above, const DevImage& dst, const zos_compose_params& cp);
zos_status launch_gather(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                         const zos_compose_params& cp, uint32_t batch);
bool frame_pipeline_eligible(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst, const zos_compose_params& cp);
void destroy_dynamic_cache(zos_ctx* ctx);
zos_status launch_yuv_fast(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch, bool* handled);
zos_status launch_yuv_chain(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch);
zos_status launch_affine_f16(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                             const zos_compose_params& cp, uint32_t batch, bool* handled);
zos_status launch_frame_pipeline(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                                 const zos_compose_params& cp, uint32_t batch, bool* handled);
zos_status launch_generate(zos_ctx* ctx, const DevImage& dst, const float* p, uint32_t batch, uint32_t kind, const float* dev = nullptr);
zos_status launch_box3(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const float* m, uint32_t batch);
zos_status launch_palette(zos_ctx* ctx, const DevImage& pal, const DevImage& idx, const DevImage& dst, const float* xc,
                          const float* yc, uint32_t batch);

}  // namespace zos
