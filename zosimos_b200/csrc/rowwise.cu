// rowwise.cu -- the streaming kernel: unpack -> per-pixel steps -> pack over 4-pixel groups with
// 16-byte vector accesses, optionally composing a second, axis-aligned and unscaled layer on top
// (inscribe / blend at an offset that is a multiple of 4 pixels).
//
// One kernel replaces, per operation of the reference, the decode pass (stage.frag decode_*),
// the PaintFullScreen / PaintToSelection draw(s) (linear.frag, oklab.frag, srlab2.frag,
// copy.frag + box.vert) and the encode pass (stage.frag encode_*), i.e.
// lib/zosimos/src/program.rs:1475-1533.  HBM traffic is the algorithmic minimum: every source
// texel is read once, every destination texel written once.
#include <stdlib.h>

// the generic codec is inlined here (one pixel per thread and iteration keeps the code size in check)
#define ZOS_SLOW_ATTR __forceinline__
#include "colorops.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_rowwise)

struct RowParams {
  DevImage below;  // the (only) source when !has_above
  DevImage above;
  DevImage dst;
  int32_t has_below, has_above;
  int32_t tx, ty, aw, ah;  // placement of `above` on the destination
  int32_t blend;           // ZOS_BLEND_*
  float inj[8];            // ZOS_BLEND_INJECT: mix, color
  uint32_t groups_per_row;
  uint32_t total_groups;   // groups_per_row * h * batch
  FastDiv div_gpr, div_h;
  StepList src_steps, dst_steps;
};


__device__ __forceinline__ uint4 load1(const int BPP, const uint8_t* p) {
  uint4 w = make_uint4(0, 0, 0, 0);
  if (BPP == 1) w.x = *p;
  else if (BPP == 2) w.x = *reinterpret_cast<const uint16_t*>(p);
  else if (BPP == 4) w.x = __ldcs(reinterpret_cast<const uint32_t*>(p));
  else if (BPP == 8) { uint2 t = __ldcs(reinterpret_cast<const uint2*>(p)); w.x = t.x; w.y = t.y; }
  else w = __ldcs(reinterpret_cast<const uint4*>(p));
  return w;
}
__device__ __forceinline__ void store1(const int BPP, uint8_t* p, const uint4& w) {
  if (BPP == 1) *p = (uint8_t)w.x;
  else if (BPP == 2) *reinterpret_cast<uint16_t*>(p) = (uint16_t)w.x;
  else if (BPP == 4) __stcs(reinterpret_cast<uint32_t*>(p), w.x);
  else if (BPP == 8) __stcs(reinterpret_cast<uint2*>(p), make_uint2(w.x, w.y));
  else __stcs(reinterpret_cast<uint4*>(p), w);
}

// The general streaming kernel: any texel pair, any step chain.  One pixel per thread and iteration: a
// warp touches 32 consecutive texels of a row (coalesced for every texel size) and the whole generic
// codec / step interpreter is inlined exactly once.  Texel sizes are kernel-uniform run-time values.
__global__ void __launch_bounds__(256) k_rowwise(const __grid_constant__ RowParams P) {
  __shared__ Tables T;
  load_tables(T);
  const int SB = P.below.bpp, DB = P.dst.bpp;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total_groups; idx += stride) {
    uint32_t rowid = fastdiv(idx, P.div_gpr);
    int x = (int)(idx - rowid * P.groups_per_row);
    uint32_t frame = fastdiv(rowid, P.div_h);
    int y = (int)(rowid - frame * (uint32_t)P.dst.h);
    const int ax = x - P.tx, ay = y - P.ty;
    const bool covered = P.has_above && ay >= 0 && ay < P.ah && ax >= 0 && ax < P.aw;
    float4 v = make_float4(0.0f, 0.0f, 1.0f, 1.0f);
    if (P.has_below && !(covered && P.blend == ZOS_BLEND_OVERWRITE))
      v = unpack_texel(P.below.fmt, load1(SB, P.below.p0 + frame * P.below.bstride + (uint64_t)y * P.below.pitch + (uint64_t)x * SB), T);
    if (covered) {
      float4 a = unpack_texel(P.above.fmt, load1(SB, P.above.p0 + frame * P.above.bstride + (uint64_t)ay * P.above.pitch + (uint64_t)ax * SB), T);
      apply_steps(P.src_steps, a, T);
      v = P.blend == ZOS_BLEND_OVERWRITE ? a : P.blend == ZOS_BLEND_INJECT ? inject_blend(P.inj, a, v) : porter_duff(P.blend, a, v);
    }
    apply_steps(P.dst_steps, v, T);
    store1(DB, P.dst.p0 + frame * P.dst.bstride + (uint64_t)y * P.dst.pitch + (uint64_t)x * DB, pack_texel(P.dst.fmt, v, T));
  }
}

static bool vec_ok(const DevImage& im) {
  int chunk = im.bpp >= 4 ? 16 : 4 * im.bpp;
  return ((uintptr_t)im.p0 % chunk) == 0 && (im.pitch % chunk) == 0 && (im.bstride % chunk) == 0 &&
         im.pitch >= (uint64_t)((im.w + 3) / 4) * 4 * im.bpp;
}

bool rowwise_can_compose(const DevImage& below, const DevImage& above, const DevImage& dst, const zos_compose_params& cp) {
  if (cp.map != ZOS_MAP_RECT || cp.sampling != ZOS_SAMPLE_NEAREST) return false;
  if (cp.dst_origin[0] || cp.dst_origin[1] || cp.src_origin[0] || cp.src_origin[1] || cp.src_full[0] || cp.src_full[1]) return false;
  if (below.block != ZOS_BLOCK_PIXEL || above.block != ZOS_BLOCK_PIXEL || dst.block != ZOS_BLOCK_PIXEL) return false;
  // unscaled: the selection is the whole layer and the target has its size
  if (cp.sel[0] != 0 || cp.sel[1] != 0 || cp.sel[2] != above.w || cp.sel[3] != above.h) return false;
  if (cp.tgt[2] != above.w || cp.tgt[3] != above.h) return false;
  if (cp.tgt[0] < 0 || cp.tgt[1] < 0 || (cp.tgt[0] & 3)) return false;
  if (below.bpp != above.bpp || below.w != dst.w || below.h != dst.h) return false;
  return vec_ok(below) && vec_ok(above) && vec_ok(dst);
}

zos_status launch_rowwise(zos_ctx* ctx, const DevImage* below, const DevImage* above, const DevImage& dst,
                          const zos_compose_params* cp, const zos_step* steps, uint32_t nsteps, uint32_t batch) {
  RowParams P;
  memset(&P, 0, sizeof P);
  P.dst = dst;
  P.has_below = below != nullptr;
  P.has_above = above != nullptr;
  if (below) P.below = *below;
  if (above) P.above = *above;
  if (!below && above) P.below = *above;  // SB comes from here
  if (cp) {
    P.tx = cp->tgt[0]; P.ty = cp->tgt[1]; P.aw = cp->tgt[2]; P.ah = cp->tgt[3];
    P.blend = cp->blend;
    memcpy(P.inj, cp->inject_mix, 16); memcpy(P.inj + 4, cp->inject_color, 16);
    P.src_steps.n = cp->n_src_steps;
    for (uint32_t i = 0; i < cp->n_src_steps; i++) P.src_steps.s[i] = cp->src_steps[i];
    P.dst_steps.n = cp->n_dst_steps;
    for (uint32_t i = 0; i < cp->n_dst_steps; i++) P.dst_steps.s[i] = cp->dst_steps[i];
  } else {
    P.blend = ZOS_BLEND_OVERWRITE;
    P.dst_steps.n = nsteps;
    for (uint32_t i = 0; i < nsteps; i++) P.dst_steps.s[i] = steps[i];
  }
  if (!vec_ok(P.below) || !vec_ok(dst))
    return fail(ctx, ZOS_ERR_INVALID, "rowwise: buffers must be 16-byte aligned with padded rows (use zos_aligned_row_stride)");
  // 8-bit -> Lab colour in an 8-bit register -> 8-bit: the specialised kernel (rowwise_lab.cu), same results
  if (!(ctx->flags & ZOS_CTX_NO_FAST_PATHS) && !cp && below && !above) {
    bool handled = false;
    cudaError_t e = launch_rowwise_lab(ctx, *below, dst, steps, nsteps, batch, &handled);
    if (handled) { ctx->launches++; return check_cuda(ctx, e, "k_rowwise_lab launch"); }
  }
  // native 8-bit / float texels with at most matrix steps: the specialised kernels (rowwise_fast.cu, rowwise_lut.cu), same results
  if (!(ctx->flags & ZOS_CTX_NO_FAST_PATHS) && rowwise_u8_eligible(below, above, dst, cp, steps, nsteps)) return launch_rowwise_u8(ctx, below, above, dst, cp, steps, nsteps, batch);
  uint64_t gpr = (uint64_t)dst.w;  // one pixel per work item
  uint64_t total = gpr * (uint64_t)dst.h * batch;
  if (total == 0) return ZOS_OK;
  if (total >= (1ull << 32)) return fail(ctx, ZOS_ERR_UNSUPPORTED, "rowwise: more than 2^32 pixels in one launch");
  P.groups_per_row = (uint32_t)gpr;
  P.total_groups = (uint32_t)total;
  P.div_gpr = make_fastdiv((uint32_t)gpr);
  P.div_h = make_fastdiv((uint32_t)dst.h);
  int grid = grid_for(ctx, total, 256, 32);  // more CTAs than resident (grid-size note)
  k_rowwise<<<grid, 256, 0, ctx->stream>>>(P);
  cudaError_t e = cudaGetLastError();
  ctx->launches++;
  return check_cuda(ctx, e, "k_rowwise launch");
}

}  // namespace zos
