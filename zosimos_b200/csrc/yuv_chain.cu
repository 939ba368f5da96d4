// yuv_chain.cu -- per-pixel chains that END in a planar YUV 4:2:0 image (I420 / NV12), the
// YUV420 -> YUV420 and RGBA -> YUV420 rows of BASELINE config 5.  The reference has no planar texels
// at all (program.rs:794-938 lowers Block::Pixel only); semantics are ours and
// the oracle is zo_encode_yuv420 (oracle/zos_oracle.c): OETF of the colour, Y' = Kr R' + Kg G' + Kb B',
// Cb = (B' - Y') / (2 (1 - Kb)), Cr likewise, range scaling, round to nearest even; one chroma sample
// per 2x2 block = the mean of the block's (up to 4) chroma values.
//
// One thread owns one 2x2 block: 4 source texels in, 4 Y' bytes and one Cb / Cr pair out; a warp
// covers 64 x 2 pixels.  Sources: any pixel texel (generic codec) or planar YUV with nearest chroma.
#include "colorops.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_yuv_chain)

namespace {

struct YuvParams {
  DevImage src, dst;
  StepList steps;
  uint32_t bw, bh, total;  // 2x2 blocks per row, per column, in the launch
  FastDiv div_bw, div_bh;
};

__device__ __forceinline__ float yuv_eotf(uint32_t tr, float v) {  // the very code of gather.cu / frame_pipeline.cu
  if (tr == ZOS_TRANSFER_BT709 || tr == ZOS_TRANSFER_BT2020_10BIT || tr == ZOS_TRANSFER_BT2020_12BIT) {
    float lin = v * (1.0f / 4.5f);
    float l2, pw;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"((v + 0.099f) * (1.0f / 1.099f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(l2 * (1.0f / 0.45f)));
    return v >= 0.0812428582f ? pw : lin;
  }
  if (tr == ZOS_TRANSFER_LINEAR) return v;
  return eo_scalar(tr, v);
}

__device__ __forceinline__ float4 fetch(const YuvParams& P, uint32_t frame, int x, int y, float U, float V, const Tables& T) {
  const DevImage& A = P.src;
  if (A.block == ZOS_BLOCK_PIXEL) {
    const uint8_t* p = A.p0 + frame * A.bstride + (uint64_t)y * A.pitch + (uint64_t)x * A.bpp;
    uint4 w = make_uint4(0, 0, 0, 0);
    switch (A.bpp) {
      case 1: w.x = *p; break;
      case 2: w.x = *reinterpret_cast<const uint16_t*>(p); break;
      case 4: w.x = *reinterpret_cast<const uint32_t*>(p); break;
      case 8: { uint2 t = *reinterpret_cast<const uint2*>(p); w.x = t.x; w.y = t.y; break; }
      default: w = *reinterpret_cast<const uint4*>(p); break;
    }
    return unpack_texel(A.fmt, w, T);
  }
  const float Y = (float)A.p0[frame * A.bstride + (uint64_t)y * A.pitch + x];
  const float yy = (Y - A.yoff) * A.ysc, cb = (U - 128.0f) * A.csc, cr = (V - 128.0f) * A.csc;
  const float r = fmaf(A.r_cr, cr, yy), g = fmaf(-A.g_cb, cb, fmaf(-A.g_cr, cr, yy)), b = fmaf(A.b_cb, cb, yy);
  return make_float4(yuv_eotf(A.fmt.transfer, r), yuv_eotf(A.fmt.transfer, g), yuv_eotf(A.fmt.transfer, b), 1.0f);
}

__device__ __forceinline__ uint8_t q8(float v) { return (uint8_t)__float2int_rn(fminf(fmaxf(v, 0.0f), 255.0f)); }

__global__ void __launch_bounds__(256) k_yuv_chain(const __grid_constant__ YuvParams P) {
  __shared__ Tables T;
  load_tables(T);
  const DevImage& D = P.dst;
  const int cstep_s = P.src.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1, cstep_d = D.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1;
  const float kg = 1.0f - D.kr - D.kb;
  const float cbd = 2.0f * (1.0f - D.kb), crd = 2.0f * (1.0f - D.kr);
  const float cscale = D.full_range ? 255.0f : 224.0f;
  const uint32_t tr = D.fmt.transfer;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += stride) {
    const uint32_t rowid = fastdiv(idx, P.div_bw);
    const int ci = (int)(idx - rowid * P.bw);
    const uint32_t frame = fastdiv(rowid, P.div_bh);
    const int cj = (int)(rowid - frame * P.bh);
    float U = 0.0f, V = 0.0f;
    if (P.src.block != ZOS_BLOCK_PIXEL) {
      const uint64_t co = frame * P.src.cbstride + (uint64_t)cj * P.src.cpitch + (uint64_t)ci * cstep_s;
      U = (float)P.src.p1[co]; V = (float)P.src.p2[co];
    }
    float cbs = 0.0f, crs = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; dy++)
#pragma unroll
      for (int dx = 0; dx < 2; dx++) {
        const int x = 2 * ci + dx, y = 2 * cj + dy;
        if (x >= D.w || y >= D.h) continue;
        float4 t = fetch(P, frame, x, y, U, V, T);
        apply_steps(P.steps, t, T);
        const float r = oe_scalar(tr, t.x), g = oe_scalar(tr, t.y), b = oe_scalar(tr, t.z);
        const float yy = fmaf(D.kb, b, fmaf(kg, g, D.kr * r));
        const float cb = (b - yy) / cbd, cr = (r - yy) / crd;
        const float Yq = D.full_range ? yy * 255.0f : fmaf(yy, 219.0f, 16.0f);
        D.p0[frame * D.bstride + (uint64_t)y * D.pitch + x] = q8(Yq);
        cbs += cb; crs += cr; cnt++;
      }
    const float cbm = cbs / (float)cnt, crm = crs / (float)cnt;
    const uint64_t co = frame * D.cbstride + (uint64_t)cj * D.cpitch + (uint64_t)ci * cstep_d;
    D.p1[co] = q8(fmaf(cbm, cscale, 128.0f));
    D.p2[co] = q8(fmaf(crm, cscale, 128.0f));
  }
}
}  // namespace

zos_status launch_yuv_chain(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch) {
  if (dst.block == ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_INVALID, "yuv_chain: destination is not planar");
  if (src.block != ZOS_BLOCK_PIXEL && src.chroma_filter != 0)
    return fail(ctx, ZOS_ERR_UNSUPPORTED, "planar -> planar chains take nearest chroma sources (chroma_filter = 0)");
  if (src.w != dst.w || src.h != dst.h) return fail(ctx, ZOS_ERR_TYPE, "yuv_chain: size mismatch");
  YuvParams P;
  memset(&P, 0, sizeof P);
  P.src = src; P.dst = dst;
  P.steps.n = nsteps;
  for (uint32_t i = 0; i < nsteps; i++) P.steps.s[i] = steps[i];
  P.bw = (uint32_t)(dst.w + 1) / 2; P.bh = (uint32_t)(dst.h + 1) / 2;
  const uint64_t total = (uint64_t)P.bw * P.bh * batch;
  if (total == 0) return ZOS_OK;
  if (total >= (1ull << 32)) return fail(ctx, ZOS_ERR_UNSUPPORTED, "yuv_chain: more than 2^32 blocks in one launch");
  P.total = (uint32_t)total;
  P.div_bw = make_fastdiv(P.bw); P.div_bh = make_fastdiv(P.bh);
  const int grid = grid_for(ctx, total, 256, 8);
  k_yuv_chain<<<grid, 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_yuv_chain launch");
}

}  // namespace zos
