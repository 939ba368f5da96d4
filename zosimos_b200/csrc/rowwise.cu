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

#include "colorops.cuh"
#include "zos_internal.h"

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

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}

// 4 consecutive texels starting at a (4*BPP)-aligned address
__device__ __forceinline__ void load4(const int BPP, const uint8_t* p, uint4 (&w)[4]) {
  if (BPP == 1) {
    uint32_t u = __ldcs(reinterpret_cast<const uint32_t*>(p));
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = make_uint4((u >> (8 * i)) & 255u, 0, 0, 0);
  } else if (BPP == 2) {
    uint2 u = __ldcs(reinterpret_cast<const uint2*>(p));
    w[0] = make_uint4(u.x & 65535u, 0, 0, 0); w[1] = make_uint4(u.x >> 16, 0, 0, 0);
    w[2] = make_uint4(u.y & 65535u, 0, 0, 0); w[3] = make_uint4(u.y >> 16, 0, 0, 0);
  } else if (BPP == 4) {
    uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
    w[0] = make_uint4(u.x, 0, 0, 0); w[1] = make_uint4(u.y, 0, 0, 0);
    w[2] = make_uint4(u.z, 0, 0, 0); w[3] = make_uint4(u.w, 0, 0, 0);
  } else if (BPP == 8) {
    uint4 a = __ldcs(reinterpret_cast<const uint4*>(p)), b = __ldcs(reinterpret_cast<const uint4*>(p) + 1);
    w[0] = make_uint4(a.x, a.y, 0, 0); w[1] = make_uint4(a.z, a.w, 0, 0);
    w[2] = make_uint4(b.x, b.y, 0, 0); w[3] = make_uint4(b.z, b.w, 0, 0);
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = __ldcs(reinterpret_cast<const uint4*>(p) + i);
  }
}
__device__ __forceinline__ void store1(const int BPP, uint8_t* p, const uint4& w) {
  if (BPP == 1) *p = (uint8_t)w.x;
  else if (BPP == 2) *reinterpret_cast<uint16_t*>(p) = (uint16_t)w.x;
  else if (BPP == 4) *reinterpret_cast<uint32_t*>(p) = w.x;
  else if (BPP == 8) *reinterpret_cast<uint2*>(p) = make_uint2(w.x, w.y);
  else *reinterpret_cast<uint4*>(p) = w;
}
__device__ __forceinline__ void store4(const int BPP, uint8_t* p, const uint4 (&w)[4]) {
  if (BPP == 1) {
    __stcs(reinterpret_cast<uint32_t*>(p), (w[0].x & 255u) | ((w[1].x & 255u) << 8) | ((w[2].x & 255u) << 16) | (w[3].x << 24));
  } else if (BPP == 2) {
    __stcs(reinterpret_cast<uint2*>(p), make_uint2((w[0].x & 65535u) | (w[1].x << 16), (w[2].x & 65535u) | (w[3].x << 16)));
  } else if (BPP == 4) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0].x, w[1].x, w[2].x, w[3].x));
  } else if (BPP == 8) {
    __stcs(reinterpret_cast<uint4*>(p), make_uint4(w[0].x, w[0].y, w[1].x, w[1].y));
    __stcs(reinterpret_cast<uint4*>(p) + 1, make_uint4(w[2].x, w[2].y, w[3].x, w[3].y));
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++) __stcs(reinterpret_cast<uint4*>(p) + i, w[i]);
  }
}

// Texel sizes are kernel-uniform run-time values (one binary for all formats); the branches on
// them are warp-uniform.
__global__ void __launch_bounds__(256) k_rowwise(const __grid_constant__ RowParams P) {
  __shared__ Tables T;
  load_tables(T);
  const int SB = P.below.bpp, DB = P.dst.bpp;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total_groups; idx += stride) {
    uint32_t rowid = fastdiv(idx, P.div_gpr);
    uint32_t g = idx - rowid * P.groups_per_row;
    uint32_t frame = fastdiv(rowid, P.div_h);
    int y = (int)(rowid - frame * (uint32_t)P.dst.h);
    int x0 = (int)g * 4;
    int npx = min(4, P.dst.w - x0);

    // does `above` cover this group?  (tx is a multiple of 4, so a group is covered from its start)
    int ax0 = x0 - P.tx, ay = y - P.ty;
    bool row_in = P.has_above && ay >= 0 && ay < P.ah && ax0 >= 0 && ax0 < P.aw;
    int ncov = row_in ? min(npx, P.aw - ax0) : 0;

    float4 v[4];
    const bool need_below = P.has_below && !(ncov == npx && P.blend == ZOS_BLEND_OVERWRITE);
    if (need_below) {
      uint4 w[4];
      load4(SB, P.below.p0 + frame * P.below.bstride + (uint64_t)y * P.below.pitch + (uint64_t)x0 * SB, w);
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = unpack_texel(P.below.fmt, w[i], T);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = make_float4(0.0f, 0.0f, 1.0f, 1.0f);
    }
    if (ncov > 0) {
      uint4 w[4];
      load4(SB, P.above.p0 + frame * P.above.bstride + (uint64_t)ay * P.above.pitch + (uint64_t)ax0 * SB, w);
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (i < ncov) {
          float4 a = unpack_texel(P.above.fmt, w[i], T);
          apply_steps(P.src_steps, a, T);
          v[i] = P.blend == ZOS_BLEND_OVERWRITE ? a : P.blend == ZOS_BLEND_INJECT ? inject_blend(P.inj, a, v[i]) : porter_duff(P.blend, a, v[i]);
        }
      }
    }
    uint4 o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      apply_steps(P.dst_steps, v[i], T);
      o[i] = pack_texel(P.dst.fmt, v[i], T);
    }
    uint8_t* dp = P.dst.p0 + frame * P.dst.bstride + (uint64_t)y * P.dst.pitch + (uint64_t)x0 * DB;
    if (npx == 4) {
      store4(DB, dp, o);
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i < npx) store1(DB, dp + i * DB, o[i]);
    }
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
  // native 8-bit texels with at most matrix steps: the specialised kernel (rowwise_u8.cu), same results
  if (!(ctx->flags & ZOS_CTX_NO_FAST_PATHS) && rowwise_u8_eligible(below, above, dst, cp, steps, nsteps)) return launch_rowwise_u8(ctx, below, above, dst, cp, steps, nsteps, batch);
  uint64_t gpr = (uint64_t)(dst.w + 3) / 4;
  uint64_t total = gpr * (uint64_t)dst.h * batch;
  if (total == 0) return ZOS_OK;
  if (total >= (1ull << 32)) return fail(ctx, ZOS_ERR_UNSUPPORTED, "rowwise: more than 2^34 pixels in one launch");
  P.groups_per_row = (uint32_t)gpr;
  P.total_groups = (uint32_t)total;
  P.div_gpr = make_fastdiv((uint32_t)gpr);
  P.div_h = make_fastdiv((uint32_t)dst.h);
  int grid = grid_for(ctx, total, 256, 8);
  k_rowwise<<<grid, 256, 0, ctx->stream>>>(P);
  cudaError_t e = cudaGetLastError();
  ctx->launches++;
  return check_cuda(ctx, e, "k_rowwise launch");
}

}  // namespace zos
