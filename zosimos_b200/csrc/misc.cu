// misc.cu -- constructors and the small neighbourhood / lookup operators.
//   generate : bilinear.frag:14-20 / solid_rgb.frag:9-11   (ConstructOp::{Bilinear,Solid}, command.rs:1524-1633)
//   box3     : box3.frag:16-52                             (derivative, command.rs:1493-1508)
//   palette  : palette.frag:21-32                          (command.rs:1442-1485)
// One destination pixel per thread; sources are unpacked on the fly (no intermediate texture).
#include "colorops.cuh"
#include "zos_internal.h"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_misc)

__device__ __forceinline__ uint4 load_word(const DevImage& im, uint32_t frame, int x, int y) {
  const uint8_t* p = im.p0 + frame * im.bstride + (uint64_t)y * im.pitch + (uint64_t)x * im.bpp;
  uint4 w = make_uint4(0, 0, 0, 0);
  switch (im.bpp) {
    case 1: w.x = *p; break;
    case 2: w.x = *reinterpret_cast<const uint16_t*>(p); break;
    case 4: w.x = *reinterpret_cast<const uint32_t*>(p); break;
    case 8: { uint2 t = *reinterpret_cast<const uint2*>(p); w.x = t.x; w.y = t.y; break; }
    default: w = *reinterpret_cast<const uint4*>(p); break;
  }
  return w;
}
__device__ __forceinline__ void store_word(const DevImage& im, uint32_t frame, int x, int y, const uint4& w) {
  uint8_t* p = im.p0 + frame * im.bstride + (uint64_t)y * im.pitch + (uint64_t)x * im.bpp;
  switch (im.bpp) {
    case 1: *p = (uint8_t)w.x; break;
    case 2: *reinterpret_cast<uint16_t*>(p) = (uint16_t)w.x; break;
    case 4: *reinterpret_cast<uint32_t*>(p) = w.x; break;
    case 8: *reinterpret_cast<uint2*>(p) = make_uint2(w.x, w.y); break;
    default: *reinterpret_cast<uint4*>(p) = w; break;
  }
}

struct GenParams { DevImage dst; float p[24]; uint32_t total; uint32_t solid; };
__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ GenParams P) {
  __shared__ Tables T;
  load_tables(T);
  const uint32_t wh = (uint32_t)P.dst.w * P.dst.h;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += gridDim.x * blockDim.x) {
    uint32_t frame = idx / wh, r = idx - frame * wh;
    int j = r / P.dst.w, i = r - j * P.dst.w;
    float u = ((float)i + 0.5f) / (float)P.dst.w, v = ((float)j + 0.5f) / (float)P.dst.h, uv = u * v;
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float a = P.p[k] * (1.0f - u) + P.p[4 + k] * u;
      float b = P.p[8 + k] * (1.0f - v) + P.p[12 + k] * v;
      float d = P.p[16 + k] * (1.0f - uv) + P.p[20 + k] * uv;
      c[k] = P.solid ? P.p[k] : a + b + d;  // solid_rgb.frag writes the colour as is
    }
    store_word(P.dst, frame, i, j, pack_texel(P.dst.fmt, make_float4(c[0], c[1], c[2], c[3]), T));
  }
}

struct Box3Params { DevImage src, dst; float m[9]; uint32_t total; };
__global__ void __launch_bounds__(256) k_box3(const __grid_constant__ Box3Params P) {
  __shared__ Tables T;
  load_tables(T);
  const uint32_t wh = (uint32_t)P.dst.w * P.dst.h;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += gridDim.x * blockDim.x) {
    uint32_t frame = idx / wh, r = idx - frame * wh;
    int j = r / P.dst.w, i = r - j * P.dst.w;
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dx = -1; dx <= 1; dx++) {
        int x = min(max(i + dx, 0), P.src.w - 1), y = min(max(j + dy, 0), P.src.h - 1);
        float4 s = unpack_texel(P.src.fmt, load_word(P.src, frame, x, y), T);
        float w = P.m[3 * (dy + 1) + (dx + 1)];
        ax = fmaf(w, s.x, ax); ay = fmaf(w, s.y, ay); az = fmaf(w, s.z, az);
      }
    store_word(P.dst, frame, i, j, pack_texel(P.dst.fmt, make_float4(ax, ay, az, 1.0f), T));
  }
}

struct PaletteParams { DevImage pal, idx, dst; float xc[4], yc[4]; uint32_t total; };
__global__ void __launch_bounds__(256) k_palette(const __grid_constant__ PaletteParams P) {
  __shared__ Tables T;
  load_tables(T);
  const uint32_t wh = (uint32_t)P.dst.w * P.dst.h;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += gridDim.x * blockDim.x) {
    uint32_t frame = idx / wh, r = idx - frame * wh;
    int j = r / P.dst.w, i = r - j * P.dst.w;
    float4 b = unpack_texel(P.idx.fmt, load_word(P.idx, frame, i, j), T);
    float pu = P.xc[0] * b.x + P.xc[1] * b.y + P.xc[2] * b.z + P.xc[3] * b.w + 0.5f / (float)P.dst.w;
    float pv = P.yc[0] * b.x + P.yc[1] * b.y + P.yc[2] * b.z + P.yc[3] * b.w + 0.5f / (float)P.dst.h;
    int x = min(max((int)floorf(pu * (float)P.pal.w), 0), P.pal.w - 1);
    int y = min(max((int)floorf(pv * (float)P.pal.h), 0), P.pal.h - 1);
    float4 c = unpack_texel(P.pal.fmt, load_word(P.pal, frame, x, y), T);
    store_word(P.dst, frame, i, j, pack_texel(P.dst.fmt, c, T));
  }
}

static zos_status total_px(zos_ctx* ctx, const DevImage& d, uint32_t batch, uint32_t* out) {
  uint64_t t = (uint64_t)d.w * d.h * batch;
  if (t >= (1ull << 32)) return fail(ctx, ZOS_ERR_UNSUPPORTED, "more than 2^32 pixels in one launch");
  *out = (uint32_t)t;
  return ZOS_OK;
}

zos_status launch_generate(zos_ctx* ctx, const DevImage& dst, const float* p, uint32_t batch, bool solid) {
  GenParams P;
  P.dst = dst;
  P.solid = solid ? 1u : 0u;
  memcpy(P.p, p, sizeof P.p);
  zos_status st = total_px(ctx, dst, batch, &P.total);
  if (st != ZOS_OK) return st;
  k_generate<<<grid_for(ctx, P.total, 256, 8), 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_generate launch");
}
zos_status launch_box3(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const float* m, uint32_t batch) {
  Box3Params P;
  P.src = src; P.dst = dst;
  memcpy(P.m, m, sizeof P.m);
  zos_status st = total_px(ctx, dst, batch, &P.total);
  if (st != ZOS_OK) return st;
  k_box3<<<grid_for(ctx, P.total, 256, 8), 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_box3 launch");
}
zos_status launch_palette(zos_ctx* ctx, const DevImage& pal, const DevImage& idx, const DevImage& dst, const float* xc,
                          const float* yc, uint32_t batch) {
  PaletteParams P;
  P.pal = pal; P.idx = idx; P.dst = dst;
  memcpy(P.xc, xc, 16); memcpy(P.yc, yc, 16);
  zos_status st = total_px(ctx, dst, batch, &P.total);
  if (st != ZOS_OK) return st;
  if (batch > 1 && pal.bstride == 0) { /* one palette shared by all frames: fine */ }
  k_palette<<<grid_for(ctx, P.total, 256, 8), 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_palette launch");
}

}  // namespace zos
