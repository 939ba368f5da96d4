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

struct GenParams { DevImage dst; float p[24]; uint32_t total; uint32_t kind; const float* dev; /* parameter block in device memory, or NULL */ };

// fractal_noise.frag: pcg4d (jcgt.org/published/0009/03/02) of a cell corner -> 4 uniform floats
__device__ __forceinline__ float4 noise_hash(uint32_t sx, uint32_t sy) {
  uint32_t x = sx * 1664525u + 1013904223u, y = sy * 1664525u + 1013904223u, z = 1013904223u, w = 1013904223u;
  x += y * w; y += z * x; z += x * y; w += y * z;
  x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
  x += y * w; y += z * x; z += x * y; w += y * z;
  return make_float4((float)x / 4294967296.0f, (float)y / 4294967296.0f, (float)z / 4294967296.0f, (float)w / 4294967296.0f);
}

__global__ void __launch_bounds__(256) k_generate(const __grid_constant__ GenParams P) {
  __shared__ Tables T;
  load_tables(T);
  __shared__ float pbuf[24];
  if (threadIdx.x < 24) pbuf[threadIdx.x] = P.dev ? P.dev[threadIdx.x] : P.p[threadIdx.x];
  __syncthreads();
  const uint32_t wh = (uint32_t)P.dst.w * P.dst.h;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += gridDim.x * blockDim.x) {
    uint32_t frame = idx / wh, r = idx - frame * wh;
    int j = r / P.dst.w, i = r - j * P.dst.w;
    float u = ((float)i + 0.5f) / (float)P.dst.w, v = ((float)j + 0.5f) / (float)P.dst.h, uv = u * v;
    float c[4];
    if (P.kind == ZOS_GEN_NORMAL2D) {  // distribution_normal2d.frag:43-55
      const float px = 2.0f * (u - 0.5f) - pbuf[0], py = 2.0f * (v - 0.5f) - pbuf[1];
      const float tx = pbuf[2] * px + pbuf[3] * py, ty = pbuf[4] * px + pbuf[5] * py;
      const float exponent = 0.5f * (px * tx + py * ty);
      c[0] = c[1] = c[2] = expf(-exponent) / sqrtf(pbuf[6]);
      c[3] = 1.0f;
    } else if (P.kind == ZOS_GEN_FRACTAL_NOISE) {  // fractal_noise.frag (same operation order as the oracle)
      const int octaves = (int)pbuf[4];
      const float c2 = 2.0f * 0.87758255f, s2 = 2.0f * 0.47942555f;
      float x = u, y = v, z = 1.0f, amp = pbuf[2];
      c[0] = c[1] = c[2] = c[3] = 0.0f;
      for (int o = 0; o < octaves; o++) {
        const float ptx = x * pbuf[0], pty = y * pbuf[1];
        const float flx = floorf(ptx), fly = floorf(pty);
        const float fx = ptx - flx, fy = pty - fly;
        const uint32_t sx = (uint32_t)(int32_t)flx, sy = (uint32_t)(int32_t)fly;
        const float4 a = noise_hash(sx, sy), b = noise_hash(sx + 1u, sy), cc = noise_hash(sx, sy + 1u), d = noise_hash(sx + 1u, sy + 1u);
        const float ux = fx * fx * (3.0f - 2.0f * fx), uy = fy * fy * (3.0f - 2.0f * fy);
#define ZOS_NOISE(k, f) c[k] = c[k] + amp * ((a.f * (1.0f - ux) + b.f * ux) + (cc.f - a.f) * uy * (1.0f - ux) + (d.f - b.f) * ux * uy);
        ZOS_NOISE(0, x) ZOS_NOISE(1, y) ZOS_NOISE(2, z) ZOS_NOISE(3, w)
#undef ZOS_NOISE
        const float nx = c2 * x - s2 * y, ny = s2 * x + c2 * y, nz = 2.0f * x + 2.0f * y + 2.0f * z;
        x = nx; y = ny; z = nz;
        amp = amp * pbuf[3];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float a = pbuf[k] * (1.0f - u) + pbuf[4 + k] * u;
        float b = pbuf[8 + k] * (1.0f - v) + pbuf[12 + k] * v;
        float d = pbuf[16 + k] * (1.0f - uv) + pbuf[20 + k] * uv;
        c[k] = P.kind == ZOS_GEN_SOLID ? pbuf[k] : a + b + d;  // solid_rgb.frag writes the colour as is
      }
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

zos_status launch_generate(zos_ctx* ctx, const DevImage& dst, const float* p, uint32_t batch, uint32_t kind, const float* dev) {
  GenParams P;
  P.dst = dst;
  P.kind = kind;
  P.dev = dev;
  if (p) memcpy(P.p, p, sizeof P.p); else memset(P.p, 0, sizeof P.p);
  zos_status st = total_px(ctx, dst, batch, &P.total);
  if (st != ZOS_OK) return st;
  k_generate<<<grid_for(ctx, P.total, 256, 32), 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_generate launch");
}
zos_status launch_box3(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const float* m, uint32_t batch) {
  Box3Params P;
  P.src = src; P.dst = dst;
  memcpy(P.m, m, sizeof P.m);
  zos_status st = total_px(ctx, dst, batch, &P.total);
  if (st != ZOS_OK) return st;
  k_box3<<<grid_for(ctx, P.total, 256, 32), 256, 0, ctx->stream>>>(P);
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
  k_palette<<<grid_for(ctx, P.total, 256, 32), 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_palette launch");
}

}  // namespace zos
