// runtime.cu -- context, device memory, host<->device transfers and the per-op C entry points.
// Replaces the device half of the reference's Pool (lib/zosimos/src/pool.rs:39-41,122-156) and the
// buffer/texture plumbing of its executor (lib/zosimos/src/run.rs:1896-2276, 3282-3309).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "colorops.cuh"
#include "zos_internal.h"

namespace zos {
cudaError_t upload_constants_rowwise(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_gather(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_misc(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_rowwise_u8(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_frame(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_affine(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_rowwise_lut(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_rowwise_lab(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_yuv_chain(const TablesGlobal*, const ColorConstants*, cudaStream_t);
cudaError_t upload_constants_rowwise_rgb10(const TablesGlobal*, const ColorConstants*, cudaStream_t);

static thread_local std::string g_create_error;

zos_status fail(zos_ctx* ctx, zos_status code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}
zos_status check_cuda(zos_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return ZOS_OK;
  return fail(ctx, e == cudaErrorMemoryAllocation ? ZOS_ERR_OOM : ZOS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
int grid_for(const zos_ctx* ctx, uint64_t work_items, int threads, int ctas_per_sm) {
  uint64_t need = (work_items + threads - 1) / threads;
  uint64_t cap = (uint64_t)ctx->sm_count * ctas_per_sm;  // a whole number of resident CTAs per SM
  if (need < cap) return (int)(need ? need : 1);
  return (int)cap;
}

// ---- constant tables (must equal the oracle's bit for bit: same formulas in double) ----
static double eotf_srgb_d(double v) { return v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4); }
static void inv3_d(const double* m, double* o) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  o[0] = A / det; o[1] = -(b * i - c * h) / det; o[2] = (b * f - c * e) / det;
  o[3] = B / det; o[4] = (a * i - c * g) / det; o[5] = -(a * f - c * d) / det;
  o[6] = C / det; o[7] = -(a * h - b * g) / det; o[8] = (a * e - b * d) / det;
}
static void mul3_d(const double* a, const double* b, double* o) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
}
static bool build_constants(TablesGlobal* t, ColorConstants* c) {
  for (int k = 0; k < 256; k++) {
    t->srgb_dec[k] = (float)eotf_srgb_d(k / 255.0);
    t->unorm8[k] = (float)k / 255.0f;
  }
  t->srgb_thr[0] = -INFINITY;
  for (int k = 1; k < 256; k++) {
    double th = eotf_srgb_d((k - 0.5) / 255.0);
    float f = (float)th;
    if ((double)f < th) f = nextafterf(f, INFINITY);
    t->srgb_thr[k] = f;
  }
  for (int k = 256; k < 260; k++) t->srgb_thr[k] = INFINITY;
  {  // bucket table of the exact encoder (texel.cuh); a bucket with two thresholds would break it: checked here
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    for (int b = 0; b < ZOS_ENC_N; b++) {
      const uint32_t top = (uint32_t)(ZOS_ENC_B0 + b), start = top << 16;
      uint32_t base = 0;
      while (base < 255 && bits(t->srgb_thr[base + 1]) <= start) base++;
      uint32_t t16 = 0x10000u;
      if (base < 255 && (bits(t->srgb_thr[base + 1]) >> 16) == top) {
        t16 = bits(t->srgb_thr[base + 1]) & 0xffffu;
        if (base + 2 <= 255 && (bits(t->srgb_thr[base + 2]) >> 16) == top) return false;  // never happens for the sRGB curve; zos_ctx_create reports it
      }
      t->srgb_enc[b] = (base << 16) + (0x10000u - t16) - (top << 16);
    }
  }
  {  // biased-key bucket table (texel.cuh): buckets are intervals of bit patterns of x found by bisection
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto fl = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    auto key = [&](uint32_t idx) { volatile float y = fl(idx) + ZOS_ENC2_BIAS; return bits(y) >> 16; };
    const uint32_t LOW = (uint32_t)ZOS_ENC2_LOW, HIGH = bits(1.0f) + 64u;  // blended values may exceed 1 by a rounding
    uint32_t lo[ZOS_ENC2_N + 1];
    for (int k = 0; k <= ZOS_ENC2_N; k++) {  // lo[k] = smallest pattern in [LOW, HIGH] whose key is >= K0 + k (HIGH + 1 if none)
      uint32_t a = LOW, b = HIGH + 1;
      while (a < b) { uint32_t m = a + (b - a) / 2; if (key(m) >= (uint32_t)(ZOS_ENC2_K0 + k)) b = m; else a = m + 1; }
      lo[k] = a;
    }
    if (lo[0] != LOW) return false;
    for (int k = 0; k < ZOS_ENC2_N; k++) {
      const uint32_t l = lo[k], h = lo[k + 1];  // bucket = [l, h)
      if (h - l >= (1u << 24)) return false;
      uint32_t base = 0;
      while (base < 255 && bits(t->srgb_thr[base + 1]) <= l) base++;
      uint32_t T = l + 0x00ffffffu;
      if (base < 255 && bits(t->srgb_thr[base + 1]) < h) {
        T = bits(t->srgb_thr[base + 1]);
        if (base + 2 <= 255 && bits(t->srgb_thr[base + 2]) < h) return false;  // two thresholds in one bucket
      }
      t->srgb_enc2[k] = (base << 24) + (1u << 24) - T;
    }
  }
  // Oklab M1, M2 (row-major; lib/std/src/oklab.frag:14-24 lists them column-major)
  static const float m1[9] = {0.8189330101f, 0.3618667424f, -0.1288597137f, 0.0329845436f, 0.9293118715f,
                              0.0361456387f, 0.0482003018f, 0.2643662691f, 0.6338517070f};
  static const float m2[9] = {0.2104542553f, 0.7936177850f, -0.0040720468f, 1.9779984951f, -2.4285922050f,
                              0.4505937099f, 0.0259040371f, 0.7827717662f, -0.8086757660f};
  // CAT02 / Hunt-Pointer-Estevez (lib/std/src/srlab2.frag:14-24)
  static const double cat[9] = {0.7328, 0.4296, -0.1624, -0.7036, 1.6975, 0.0061, 0.0030, 0.0136, 0.9834};
  static const double hpe[9] = {0.38971, 0.68898, -0.07868, -0.22981, 1.18340, 0.04641, 0.0, 0.0, 1.0};
  double m[9], inv[9], cati[9], hpei[9], tmp[9];
  for (int i = 0; i < 9; i++) { c->ok_m1[i] = m1[i]; c->ok_m2[i] = m2[i]; }
  for (int i = 0; i < 9; i++) m[i] = (double)m1[i];
  inv3_d(m, inv);
  for (int i = 0; i < 9; i++) c->ok_m1i[i] = (float)inv[i];
  for (int i = 0; i < 9; i++) m[i] = (double)m2[i];
  inv3_d(m, inv);
  for (int i = 0; i < 9; i++) c->ok_m2i[i] = (float)inv[i];
  inv3_d(cat, cati);
  inv3_d(hpe, hpei);
  for (int i = 0; i < 9; i++) {
    c->sr_cat[i] = (float)cat[i]; c->sr_cati[i] = (float)cati[i];
    c->sr_hpe[i] = (float)hpe[i]; c->sr_hpei[i] = (float)hpei[i];
  }
  mul3_d(hpe, cati, tmp);
  for (int i = 0; i < 9; i++) c->sr_hpe_cati[i] = (float)tmp[i];
  mul3_d(cat, hpei, tmp);
  for (int i = 0; i < 9; i++) c->sr_cat_hpei[i] = (float)tmp[i];
  return true;
}

static const float YUV_K[3][2] = {{0.299f, 0.114f}, {0.2126f, 0.0722f}, {0.2627f, 0.0593f}};

zos_status make_dev_image(zos_ctx* ctx, const zos_image* img, DevImage* out, const char* name) {
  if (!img || !img->data) return fail(ctx, ZOS_ERR_INVALID, "%s: null image", name);
  const zos_desc& d = img->desc;
  if (d.width == 0 || d.height == 0) return fail(ctx, ZOS_ERR_INVALID, "%s: empty image", name);
  memset(out, 0, sizeof *out);
  zos_texfmt f;
  zos_status st = zos_desc_texfmt(&d, &f);
  if (st != ZOS_OK) return fail(ctx, st, "%s: no texture representation for bits=%u parts=%u color=%u", name, d.bits, d.parts, d.color);
  out->fmt = f;
  out->p0 = (uint8_t*)img->data;
  out->p1 = (uint8_t*)img->plane1;
  out->p2 = (uint8_t*)img->plane2;
  out->pitch = d.row_stride;
  out->cpitch = img->chroma_stride;
  out->bstride = img->batch_stride;
  out->cbstride = img->chroma_batch_stride;
  out->w = (int32_t)d.width;
  out->h = (int32_t)d.height;
  out->block = d.block;
  if (d.block == ZOS_BLOCK_PIXEL) {
    out->bpp = (int32_t)zos_bits_bytes(d.bits);
    if (d.texel_stride != (uint32_t)out->bpp) return fail(ctx, ZOS_ERR_INVALID, "%s: texel_stride %u != %d (Descriptor::is_consistent)", name, d.texel_stride, out->bpp);
    if (d.row_stride < (uint64_t)d.width * out->bpp) return fail(ctx, ZOS_ERR_INVALID, "%s: row_stride too small", name);
  } else {
    if (d.yuv_matrix > 2) return fail(ctx, ZOS_ERR_INVALID, "%s: bad yuv_matrix", name);
    if (!img->plane1 || (d.block == ZOS_BLOCK_YUV420_PLANAR && !img->plane2)) return fail(ctx, ZOS_ERR_INVALID, "%s: missing chroma plane", name);
    out->bpp = 1;
    out->kr = YUV_K[d.yuv_matrix][0];
    out->kb = YUV_K[d.yuv_matrix][1];
    {  // the same four matrix coefficients the oracle uses: evaluated in double, rounded once
      const double kr = (double)out->kr, kb = (double)out->kb, kg = 1.0 - kr - kb;
      out->yoff = d.yuv_full_range ? 0.0f : 16.0f;
      out->ysc = d.yuv_full_range ? 1.0f / 255.0f : 1.0f / 219.0f;
      out->csc = d.yuv_full_range ? 1.0f / 255.0f : 1.0f / 224.0f;
      out->r_cr = (float)(2.0 * (1.0 - kr)); out->b_cb = (float)(2.0 * (1.0 - kb));
      out->g_cr = (float)(2.0 * kr * (1.0 - kr) / kg); out->g_cb = (float)(2.0 * kb * (1.0 - kb) / kg);
    }
    out->full_range = d.yuv_full_range;
    out->chroma_filter = d.chroma_filter;
    if (d.block == ZOS_BLOCK_YUV420_NV12) out->p2 = out->p1 + 1;
  }
  return ZOS_OK;
}

// The kernels' form of the Oklab steps (colorops.cuh): the caller's xyz_transform folded with the constant M1 / M1^-1, in double.
void fold_steps(zos_step* steps, uint32_t n) {
  static const double m1[9] = {0.8189330101f, 0.3618667424f, -0.1288597137f, 0.0329845436f, 0.9293118715f,
                               0.0361456387f, 0.0482003018f, 0.2643662691f, 0.6338517070f};  // the float constants of build_constants, widened
  static double m1i[9];
  static bool have = false;
  if (!have) { inv3_d(m1, m1i); have = true; }
  for (uint32_t i = 0; i < n; i++) {
    if (steps[i].kind != ZOS_STEP_OKLAB_ENC && steps[i].kind != ZOS_STEP_OKLAB_DEC) continue;
    double m[9], o[9];
    for (int k = 0; k < 9; k++) m[k] = (double)steps[i].m[k];
    if (steps[i].kind == ZOS_STEP_OKLAB_ENC) mul3_d(m1, m, o); else mul3_d(m, m1i, o);
    for (int k = 0; k < 9; k++) steps[i].m[k] = (float)o[k];
  }
}

zos_status validate_steps(zos_ctx* ctx, const zos_step* steps, uint32_t n) {
  if (n > ZOS_MAX_STEPS) return fail(ctx, ZOS_ERR_INVALID, "too many steps (%u > %d)", n, ZOS_MAX_STEPS);
  if (n && !steps) return fail(ctx, ZOS_ERR_INVALID, "null steps");
  for (uint32_t i = 0; i < n; i++) {
    uint32_t k = steps[i].kind;
    if (k < ZOS_STEP_MATRIX || k > ZOS_STEP_F16 || k == ZOS_STEP_INJECT) return fail(ctx, ZOS_ERR_UNSUPPORTED, "step %u: kind %u", i, k);
    if (k == ZOS_STEP_REQUANT && steps[i].fmt.storage > ZOS_STORAGE_FLOAT) return fail(ctx, ZOS_ERR_INVALID, "step %u: bad requant format", i);
  }
  return ZOS_OK;
}
}  // namespace zos

using namespace zos;

// Size classes of the arena: powers of two up to 2 MiB, multiples of 2 MiB above (the driver's own granularity), so that the
// registers of a relaunched program -- and of another program with images of the same shape -- find their blocks again.
static uint64_t arena_class(uint64_t bytes) {
  if (bytes <= 256) return 256;
  if (bytes <= (2ull << 20)) { uint64_t c = 256; while (c < bytes) c <<= 1; return c; }
  return (bytes + (2ull << 20) - 1) & ~((2ull << 20) - 1);
}
static void arena_release_parked(zos_ctx* ctx) {
  for (auto& cls : ctx->arena_free)
    for (void* p : cls.second) { cudaFree(p); ctx->arena.bytes_reserved -= cls.first; ctx->arena.bytes_parked -= cls.first; }
  ctx->arena_free.clear();
}

extern "C" {

uint32_t zos_abi_version(void) { return ZOS_ABI_VERSION; }

uint32_t zos_bits_bytes(uint32_t bits) {
  switch (bits) {
    case ZOS_BITS_UINT8: case ZOS_BITS_UINT332: case ZOS_BITS_UINT233: return 1;
    case ZOS_BITS_UINT16: case ZOS_BITS_UINT4X4: case ZOS_BITS_UINT_444: case ZOS_BITS_UINT444_: case ZOS_BITS_UINT565: case ZOS_BITS_UINT8X2: return 2;
    case ZOS_BITS_UINT8X3: return 3;
    case ZOS_BITS_UINT8X4: case ZOS_BITS_UINT16X2: case ZOS_BITS_UINT2101010: case ZOS_BITS_UINT1010102: case ZOS_BITS_UINT101010_: case ZOS_BITS_UINT_101010: return 4;
    case ZOS_BITS_UINT16X3: return 6;
    case ZOS_BITS_UINT16X4: case ZOS_BITS_FLOAT16X4: return 8;
    case ZOS_BITS_FLOAT32X4: return 16;
  }
  return 0;
}

uint64_t zos_aligned_row_stride(uint32_t width, uint32_t texel_stride) {
  uint64_t b = (uint64_t)width * texel_stride;
  return (b + 255) / 256 * 256;
}

zos_status zos_desc_texfmt(const zos_desc* d, zos_texfmt* out) {
  if (!d || !out) return ZOS_ERR_INVALID;
  if (d->block != ZOS_BLOCK_PIXEL) {
    if (d->block > ZOS_BLOCK_YUV420_NV12) return ZOS_ERR_INVALID;
    *out = zos_texfmt{d->transfer, ZOS_PARTS_YUV, ZOS_BITS_UINT8, ZOS_STORAGE_YUV420};
    return ZOS_OK;
  }
  uint32_t bytes = zos_bits_bytes(d->bits);
  if (bytes == 0) return ZOS_ERR_INVALID;
  const bool rgb = d->color == ZOS_COLOR_RGB, scalars = d->color == ZOS_COLOR_SCALARS;
  const bool flt = d->bits == ZOS_BITS_FLOAT16X4 || d->bits == ZOS_BITS_FLOAT32X4;
  if (rgb && d->bits == ZOS_BITS_UINT8X4 && (d->parts == ZOS_PARTS_RGBA || d->parts == ZOS_PARTS_BGRA) &&
      (d->transfer == ZOS_TRANSFER_SRGB || d->transfer == ZOS_TRANSFER_LINEAR)) {
    *out = zos_texfmt{d->transfer, d->parts, d->bits, d->transfer == ZOS_TRANSFER_SRGB ? (uint32_t)ZOS_STORAGE_SRGB8 : (uint32_t)ZOS_STORAGE_UNORM8};
    return ZOS_OK;
  }
  if (rgb || scalars) {
    if (flt) { *out = zos_texfmt{d->transfer, d->parts, d->bits, ZOS_STORAGE_FLOAT}; return ZOS_OK; }
    if (bytes != 1 && bytes != 2 && bytes != 4 && d->bits != ZOS_BITS_UINT16X4) return ZOS_ERR_UNSUPPORTED;  // stage.rs:63-72 (+ UInt16x4, ours)
    *out = zos_texfmt{d->transfer, d->parts, d->bits, ZOS_STORAGE_STAGED};
    return ZOS_OK;
  }
  if ((d->color == ZOS_COLOR_OKLAB || d->color == ZOS_COLOR_SRLAB2) && (d->parts == ZOS_PARTS_LCHA || d->parts == ZOS_PARTS_LABA)) {
    uint32_t tr = d->parts == ZOS_PARTS_LCHA ? (uint32_t)ZOS_TRANSFER_LABLCH : (uint32_t)ZOS_TRANSFER_LINEAR;
    if (flt) { *out = zos_texfmt{tr, ZOS_PARTS_LCHA, d->bits, ZOS_STORAGE_FLOAT}; return ZOS_OK; }
    if (bytes != 1 && bytes != 2 && bytes != 4 && d->bits != ZOS_BITS_UINT16X4) return ZOS_ERR_UNSUPPORTED;
    *out = zos_texfmt{tr, ZOS_PARTS_LCHA, d->bits, ZOS_STORAGE_STAGED};  // program.rs:882-890
    return ZOS_OK;
  }
  return ZOS_ERR_UNSUPPORTED;
}

uint64_t zos_desc_device_bytes(const zos_desc* d) {
  if (!d) return 0;
  uint64_t y = d->row_stride * d->height;
  if (d->block == ZOS_BLOCK_PIXEL) return y;
  uint64_t cw = (d->width + 1) / 2, ch = (d->height + 1) / 2;
  uint64_t cstride = zos_aligned_row_stride((uint32_t)(d->block == ZOS_BLOCK_YUV420_NV12 ? 2 * cw : cw), 1);
  return y + cstride * ch * (d->block == ZOS_BLOCK_YUV420_NV12 ? 1 : 2);
}

zos_status zos_ctx_create(int32_t device, zos_ctx** out) {
  if (!out) return ZOS_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(nullptr, ZOS_ERR_CUDA, "no CUDA device (%s); this backend has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, ZOS_ERR_INVALID, "device %d out of range (%d devices)", device, n);
  zos_ctx* ctx = new zos_ctx();
  ctx->device = device;
  if (const char* f = getenv("ZOS_CTX_FLAGS")) ctx->flags = (uint32_t)strtoul(f, nullptr, 0);  // A/B of kernel variants from outside (profiles/)
  zos_status st;
  if ((st = check_cuda(ctx, cudaSetDevice(device), "cudaSetDevice")) != ZOS_OK) goto bad;
  {
    cudaDeviceProp prop;
    if ((st = check_cuda(ctx, cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) != ZOS_OK) goto bad;
    if (prop.major != 10) { st = fail(ctx, ZOS_ERR_UNSUPPORTED, "device is sm_%d%d; this library carries sm_100a code only", prop.major, prop.minor); goto bad; }
    ctx->sm_count = prop.multiProcessorCount;
  }
  if ((st = check_cuda(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking), "cudaStreamCreate")) != ZOS_OK) goto bad;
  if (cudaHostAlloc((void**)&ctx->fault_host, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *ctx->fault_host = 0;
    if (cudaHostGetDevicePointer((void**)&ctx->fault_dev, ctx->fault_host, 0) != cudaSuccess) ctx->fault_dev = nullptr;
  }
  {
    void* wc = nullptr;
    if ((st = check_cuda(ctx, cudaMalloc(&wc, 256), "cudaMalloc(work counter)")) != ZOS_OK) goto bad;
    ctx->scratch.push_back(wc);
    ctx->work_counter = reinterpret_cast<float*>(wc);
  }
  {
    TablesGlobal* t = new TablesGlobal();
    ColorConstants* c = new ColorConstants();
    if (!build_constants(t, c)) { delete t; delete c; st = fail(ctx, ZOS_ERR_INVALID, "internal: sRGB bucket table has two thresholds in one bucket"); goto bad; }
    cudaError_t e1 = upload_constants_rowwise(t, c, ctx->stream);
    cudaError_t e2 = upload_constants_gather(t, c, ctx->stream);
    cudaError_t e3 = upload_constants_misc(t, c, ctx->stream);
    cudaError_t e5 = upload_constants_rowwise_u8(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_frame(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_affine(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_rowwise_lut(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_rowwise_lab(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_yuv_chain(t, c, ctx->stream);
    if (e5 == cudaSuccess) e5 = upload_constants_rowwise_rgb10(t, c, ctx->stream);
    cudaError_t e4 = cudaStreamSynchronize(ctx->stream);
    delete t;
    delete c;
    cudaError_t ee = e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3 != cudaSuccess ? e3 : e5 != cudaSuccess ? e5 : e4;
    if ((st = check_cuda(ctx, ee, "constant upload")) != ZOS_OK) goto bad;
  }
  *out = ctx;
  return ZOS_OK;
bad:
  g_create_error = ctx->err;
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return st;
}

void zos_ctx_destroy(zos_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  destroy_dynamic_cache(ctx);
  arena_release_parked(ctx);
  for (void* p : ctx->scratch) cudaFree(p);
  if (ctx->fault_host) cudaFreeHost(ctx->fault_host);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
const char* zos_last_error(const zos_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int32_t zos_ctx_device(const zos_ctx* ctx) { return ctx ? ctx->device : -1; }
void* zos_ctx_stream(const zos_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t zos_ctx_launch_count(const zos_ctx* ctx) { return ctx ? ctx->launches : 0; }
zos_status zos_ctx_set_flags(zos_ctx* ctx, uint32_t flags) {
  if (!ctx) return ZOS_ERR_INVALID;
  ctx->flags = flags;
  return ZOS_OK;
}
// Host-only view of the tables behind the exact sRGB8 encoders (texel.cuh), for verification without a GPU.
zos_status zos_srgb_encoder_tables(float* thresholds260, uint32_t* buckets, uint32_t* n_buckets, uint32_t* buckets2, uint32_t* n_buckets2) {
  TablesGlobal* t = new TablesGlobal();
  ColorConstants* c = new ColorConstants();
  const bool ok = build_constants(t, c);
  if (ok) {
    if (thresholds260) memcpy(thresholds260, t->srgb_thr, sizeof t->srgb_thr);
    if (buckets) memcpy(buckets, t->srgb_enc, sizeof t->srgb_enc);
    if (buckets2) memcpy(buckets2, t->srgb_enc2, sizeof t->srgb_enc2);
    if (n_buckets) *n_buckets = ZOS_ENC_N;
    if (n_buckets2) *n_buckets2 = ZOS_ENC2_N;
  }
  delete t;
  delete c;
  return ok ? ZOS_OK : ZOS_ERR_INVALID;
}

zos_status zos_sync(zos_ctx* ctx) {
  if (!ctx) return ZOS_ERR_INVALID;
  zos_status st = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  if (st == ZOS_OK && ctx->fault_host && *ctx->fault_host) {
    const int what = *ctx->fault_host;
    *ctx->fault_host = 0;
    if (what == 2) return zos::fail(ctx, ZOS_ERR_CUDA, "a table kernel found its shared memory at an unexpected address and did not run: results of this launch are missing");
    return zos::fail(ctx, ZOS_ERR_CUDA, "a TMA wait timed out inside a kernel: results of this launch are incomplete");
  }
  return st;
}

zos_status zos_poll(zos_ctx* ctx, int32_t* done) {
  if (!ctx || !done) return ZOS_ERR_INVALID;
  *done = 0;
  cudaSetDevice(ctx->device);
  const cudaError_t e = cudaStreamQuery(ctx->stream);
  if (e == cudaErrorNotReady) return ZOS_OK;
  if (e != cudaSuccess) return check_cuda(ctx, e, "cudaStreamQuery");
  *done = 1;
  return zos_sync(ctx);  // nothing left to wait for: picks up the kernels' fault word
}

zos_status zos_buf_alloc(zos_ctx* ctx, uint64_t bytes, zos_buf** out) {
  if (!ctx || !out) return ZOS_ERR_INVALID;
  *out = nullptr;
  const uint64_t cap = arena_class(bytes);
  void* p = nullptr;
  auto it = ctx->arena_free.find(cap);
  if (it != ctx->arena_free.end() && !it->second.empty()) {
    p = it->second.back();  // most recently parked first: a relaunch gets the very pointers it had (its CUDA graph stays valid)
    it->second.pop_back();
    ctx->arena.reuses++;
    ctx->arena.bytes_parked -= cap;
  } else {
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(&p, cap);
    if (e == cudaErrorMemoryAllocation && ctx->arena.bytes_parked) {  // parked blocks of other classes are in the way: give them back, once
      cudaGetLastError();
      cudaStreamSynchronize(ctx->stream);
      arena_release_parked(ctx);
      e = cudaMalloc(&p, cap);
    }
    zos_status st = check_cuda(ctx, e, "cudaMalloc");
    if (st != ZOS_OK) return st;
    ctx->arena.device_allocs++;
    ctx->arena.bytes_reserved += cap;
  }
  ctx->arena.bytes_in_use += cap;
  zos_buf* b = new zos_buf();
  b->ptr = p;
  b->size = bytes;
  b->cap = cap;
  *out = b;
  return ZOS_OK;
}
void zos_buf_free(zos_ctx* ctx, zos_buf* buf) {
  if (!buf) return;
  if (!ctx) { cudaFree(buf->ptr); delete buf; return; }  // the context is gone: nothing to park the block in
  ctx->arena_free[buf->cap].push_back(buf->ptr);
  ctx->arena.bytes_in_use -= buf->cap;
  ctx->arena.bytes_parked += buf->cap;
  delete buf;
}
zos_status zos_ctx_arena_stats(const zos_ctx* ctx, zos_arena_stats* out) {
  if (!ctx || !out) return ZOS_ERR_INVALID;
  *out = ctx->arena;
  return ZOS_OK;
}
zos_status zos_ctx_arena_trim(zos_ctx* ctx) {
  if (!ctx) return ZOS_ERR_INVALID;
  if (ctx->arena_free.empty()) return ZOS_OK;
  cudaSetDevice(ctx->device);
  zos_status st = check_cuda(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");  // kernels may still read parked blocks
  arena_release_parked(ctx);
  return st;
}
void* zos_buf_ptr(const zos_buf* buf) { return buf ? buf->ptr : nullptr; }
uint64_t zos_buf_size(const zos_buf* buf) { return buf ? buf->size : 0; }

zos_status zos_host_alloc(zos_ctx* ctx, uint64_t bytes, void** out) {
  if (!ctx || !out) return ZOS_ERR_INVALID;
  cudaSetDevice(ctx->device);
  return check_cuda(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc");
}
void zos_host_free(zos_ctx* ctx, void* ptr) {
  (void)ctx;
  if (ptr) cudaFreeHost(ptr);
}

zos_status zos_buf_upload(zos_ctx* ctx, zos_buf* dst, uint64_t off, uint64_t dpitch, const void* host, uint64_t hpitch,
                          uint64_t row_bytes, uint64_t rows) {
  if (!ctx || !dst || !host) return ZOS_ERR_INVALID;
  if (rows == 0 || row_bytes == 0) return ZOS_OK;
  if (row_bytes > dpitch || row_bytes > hpitch || off + (rows - 1) * dpitch + row_bytes > dst->size)
    return fail(ctx, ZOS_ERR_INVALID, "upload out of bounds");
  cudaSetDevice(ctx->device);
  if (dpitch == row_bytes && hpitch == row_bytes)  // tight on both sides (e.g. 3840 x 4 bytes is already 256-aligned): one linear copy
    return check_cuda(ctx, cudaMemcpyAsync((uint8_t*)dst->ptr + off, host, row_bytes * rows, cudaMemcpyHostToDevice, ctx->stream), "upload");
  return check_cuda(ctx, cudaMemcpy2DAsync((uint8_t*)dst->ptr + off, dpitch, host, hpitch, row_bytes, rows, cudaMemcpyHostToDevice, ctx->stream), "upload");
}
zos_status zos_buf_download(zos_ctx* ctx, const zos_buf* src, uint64_t off, uint64_t spitch, void* host, uint64_t hpitch,
                            uint64_t row_bytes, uint64_t rows) {
  if (!ctx || !src || !host) return ZOS_ERR_INVALID;
  if (rows == 0 || row_bytes == 0) return ZOS_OK;
  if (row_bytes > spitch || row_bytes > hpitch || off + (rows - 1) * spitch + row_bytes > src->size)
    return fail(ctx, ZOS_ERR_INVALID, "download out of bounds");
  cudaSetDevice(ctx->device);
  if (spitch == row_bytes && hpitch == row_bytes)
    return check_cuda(ctx, cudaMemcpyAsync(host, (const uint8_t*)src->ptr + off, row_bytes * rows, cudaMemcpyDeviceToHost, ctx->stream), "download");
  return check_cuda(ctx, cudaMemcpy2DAsync(host, hpitch, (const uint8_t*)src->ptr + off, spitch, row_bytes, rows, cudaMemcpyDeviceToHost, ctx->stream), "download");
}
zos_status zos_buf_copy(zos_ctx* ctx, zos_buf* dst, uint64_t doff, const zos_buf* src, uint64_t soff, uint64_t bytes) {
  if (!ctx || !dst || !src) return ZOS_ERR_INVALID;
  if (doff + bytes > dst->size || soff + bytes > src->size) return fail(ctx, ZOS_ERR_INVALID, "copy out of bounds");
  cudaSetDevice(ctx->device);
  return check_cuda(ctx, cudaMemcpyAsync((uint8_t*)dst->ptr + doff, (const uint8_t*)src->ptr + soff, bytes, cudaMemcpyDeviceToDevice, ctx->stream), "copy");
}
zos_status zos_buf_fill(zos_ctx* ctx, zos_buf* dst, uint64_t off, uint64_t bytes, uint8_t value) {
  if (!ctx || !dst) return ZOS_ERR_INVALID;
  if (off + bytes > dst->size) return fail(ctx, ZOS_ERR_INVALID, "fill out of bounds");
  cudaSetDevice(ctx->device);
  return check_cuda(ctx, cudaMemsetAsync((uint8_t*)dst->ptr + off, value, bytes, ctx->stream), "fill");
}

static zos_status image_xfer(zos_ctx* ctx, const zos_image* img, uint32_t frame, void* host, bool up) {
  if (!ctx || !img || !img->data || !host) return ZOS_ERR_INVALID;
  const zos_desc& d = img->desc;
  cudaSetDevice(ctx->device);
  auto copy = [&](uint8_t* dev, uint64_t dpitch, uint8_t* h, uint64_t row, uint64_t rows) {
    if (dpitch == row)  // the device pitch has no padding: one linear copy instead of a row loop in the copy engine
      return check_cuda(ctx, cudaMemcpyAsync(up ? (void*)dev : (void*)h, up ? (const void*)h : (const void*)dev, row * rows,
                                             up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream), up ? "image upload" : "image download");
    cudaError_t e = up ? cudaMemcpy2DAsync(dev, dpitch, h, row, row, rows, cudaMemcpyHostToDevice, ctx->stream)
                       : cudaMemcpy2DAsync(h, row, dev, dpitch, row, rows, cudaMemcpyDeviceToHost, ctx->stream);
    return check_cuda(ctx, e, up ? "image upload" : "image download");
  };
  uint8_t* h = (uint8_t*)host;
  uint64_t row = (uint64_t)d.width * (d.block == ZOS_BLOCK_PIXEL ? d.texel_stride : 1);
  zos_status st = copy((uint8_t*)img->data + (uint64_t)frame * img->batch_stride, d.row_stride, h, row, d.height);
  if (st != ZOS_OK || d.block == ZOS_BLOCK_PIXEL) return st;
  uint64_t cw = (d.width + 1) / 2, ch = (d.height + 1) / 2;
  h += row * d.height;
  if (d.block == ZOS_BLOCK_YUV420_NV12) return copy((uint8_t*)img->plane1 + (uint64_t)frame * img->chroma_batch_stride, img->chroma_stride, h, 2 * cw, ch);
  if ((st = copy((uint8_t*)img->plane1 + (uint64_t)frame * img->chroma_batch_stride, img->chroma_stride, h, cw, ch)) != ZOS_OK) return st;
  return copy((uint8_t*)img->plane2 + (uint64_t)frame * img->chroma_batch_stride, img->chroma_stride, h + cw * ch, cw, ch);
}
zos_status zos_image_upload(zos_ctx* ctx, const zos_image* dst, uint32_t frame, const void* host) { return image_xfer(ctx, dst, frame, (void*)host, true); }
zos_status zos_image_download(zos_ctx* ctx, const zos_image* src, uint32_t frame, void* host) { return image_xfer(ctx, src, frame, host, false); }

zos_status zos_pixel_chain(zos_ctx* ctx, const zos_image* src, const zos_image* dst, const zos_step* steps, uint32_t nsteps, uint32_t batch) {
  if (!ctx) return ZOS_ERR_INVALID;
  DevImage s, d;
  zos_status st;
  if ((st = make_dev_image(ctx, src, &s, "src")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if ((st = validate_steps(ctx, steps, nsteps)) != ZOS_OK) return st;
  zos_step folded[ZOS_MAX_STEPS];
  for (uint32_t i = 0; i < nsteps; i++) folded[i] = steps[i];
  fold_steps(folded, nsteps);
  steps = folded;
  if (s.w != d.w || s.h != d.h) return fail(ctx, ZOS_ERR_TYPE, "pixel_chain: size mismatch %dx%d vs %dx%d", s.w, s.h, d.w, d.h);
  if (batch == 0) return ZOS_OK;
  cudaSetDevice(ctx->device);
  if (d.block != ZOS_BLOCK_PIXEL) return launch_yuv_chain(ctx, s, d, steps, nsteps, batch);  // planar destination (yuv_chain.cu)
  if (s.block != ZOS_BLOCK_PIXEL) {
    bool handled = false;  // planar -> native 8-bit, same size: the streaming kernel of yuv_chain.cu
    st = launch_yuv_fast(ctx, s, d, steps, nsteps, batch, &handled);
    if (handled || st != ZOS_OK) return st;
    // other planar sources go through the gather kernel with the identity mapping
    zos_compose_params cp;
    memset(&cp, 0, sizeof cp);
    cp.map = ZOS_MAP_RECT; cp.sampling = ZOS_SAMPLE_NEAREST; cp.blend = ZOS_BLEND_OVERWRITE;
    cp.sel[2] = s.w; cp.sel[3] = s.h; cp.tgt[2] = d.w; cp.tgt[3] = d.h;
    cp.n_dst_steps = nsteps;
    for (uint32_t i = 0; i < nsteps; i++) cp.dst_steps[i] = steps[i];
    return launch_gather(ctx, nullptr, s, d, cp, batch);
  }
  return launch_rowwise(ctx, &s, nullptr, d, nullptr, steps, nsteps, batch);
}

zos_status zos_compose(zos_ctx* ctx, const zos_image* below, const zos_image* above, const zos_image* dst,
                       const zos_compose_params* cp, uint32_t batch) {
  if (!ctx || !cp) return ZOS_ERR_INVALID;
  DevImage b, a, d;
  zos_status st;
  if (below && (st = make_dev_image(ctx, below, &b, "below")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, above, &a, "above")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if ((st = validate_steps(ctx, cp->src_steps, cp->n_src_steps)) != ZOS_OK) return st;
  if ((st = validate_steps(ctx, cp->dst_steps, cp->n_dst_steps)) != ZOS_OK) return st;
  zos_compose_params folded = *cp;
  fold_steps(folded.src_steps, folded.n_src_steps);
  fold_steps(folded.dst_steps, folded.n_dst_steps);
  cp = &folded;
  if (below && (b.w != d.w || b.h != d.h)) return fail(ctx, ZOS_ERR_TYPE, "compose: `below` and dst differ in size");
  if (cp->map < ZOS_MAP_RECT || cp->map > ZOS_MAP_SCALE) return fail(ctx, ZOS_ERR_INVALID, "compose: bad map %d", cp->map);
  if (cp->sampling != ZOS_SAMPLE_NEAREST && cp->sampling != ZOS_SAMPLE_BILINEAR) return fail(ctx, ZOS_ERR_INVALID, "compose: bad sampling");
  if (cp->blend < ZOS_BLEND_OVERWRITE || cp->blend > ZOS_BLEND_INJECT) return fail(ctx, ZOS_ERR_INVALID, "compose: bad blend mode");
  if (cp->blend != ZOS_BLEND_OVERWRITE && !below) return fail(ctx, ZOS_ERR_INVALID, "compose: blending needs `below`");
  if (cp->map == ZOS_MAP_RECT && (cp->sel[2] <= 0 || cp->sel[3] <= 0 || cp->tgt[2] <= 0 || cp->tgt[3] <= 0))
    return fail(ctx, ZOS_ERR_INVALID, "compose: empty selection / target");
  if (batch == 0) return ZOS_OK;
  cudaSetDevice(ctx->device);
  if (below && rowwise_can_compose(b, a, d, *cp)) return launch_rowwise(ctx, &b, &a, d, cp, nullptr, 0, batch);
  return launch_gather(ctx, below ? &b : nullptr, a, d, *cp, batch);
}

zos_status zos_generate(zos_ctx* ctx, const zos_image* dst, uint32_t kind, const float* p, uint32_t batch) {
  if (!ctx || !p) return ZOS_ERR_INVALID;
  if (kind > ZOS_GEN_FRACTAL_NOISE) return fail(ctx, ZOS_ERR_INVALID, "generate: unknown generator %u", kind);
  DevImage d;
  zos_status st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if (d.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "generate: planar destination");
  if (kind == ZOS_GEN_FRACTAL_NOISE && !(p[4] >= 0.0f && p[4] <= 64.0f)) return fail(ctx, ZOS_ERR_INVALID, "fractal noise: 0..64 octaves");
  cudaSetDevice(ctx->device);
  return batch ? launch_generate(ctx, d, p, batch, kind) : ZOS_OK;
}
zos_status zos_generate_from_buffer(zos_ctx* ctx, const zos_image* dst, uint32_t kind, const zos_buf* params, uint64_t offset, uint32_t batch) {
  if (!ctx || !params) return ZOS_ERR_INVALID;
  if (kind > ZOS_GEN_FRACTAL_NOISE) return fail(ctx, ZOS_ERR_INVALID, "generate: unknown generator %u", kind);
  if ((offset & 3) || offset + 96 > params->size) return fail(ctx, ZOS_ERR_INVALID, "generate: the parameter block (96 bytes, 4-byte aligned) does not fit the buffer");
  DevImage d;
  zos_status st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if (d.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "generate: planar destination");
  cudaSetDevice(ctx->device);
  return batch ? launch_generate(ctx, d, nullptr, batch, kind, reinterpret_cast<const float*>((const uint8_t*)params->ptr + offset)) : ZOS_OK;
}
zos_status zos_generate_bilinear(zos_ctx* ctx, const zos_image* dst, const float* p, uint32_t batch) {
  return zos_generate(ctx, dst, ZOS_GEN_BILINEAR, p, batch);
}
zos_status zos_generate_solid(zos_ctx* ctx, const zos_image* dst, const float* color, uint32_t batch) {
  if (!color) return ZOS_ERR_INVALID;
  float p[24] = {0};
  memcpy(p, color, 16);
  return zos_generate(ctx, dst, ZOS_GEN_SOLID, p, batch);
}
zos_status zos_box3(zos_ctx* ctx, const zos_image* src, const zos_image* dst, const float* m, uint32_t batch) {
  if (!ctx || !m) return ZOS_ERR_INVALID;
  DevImage s, d;
  zos_status st;
  if ((st = make_dev_image(ctx, src, &s, "src")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if (s.w != d.w || s.h != d.h) return fail(ctx, ZOS_ERR_TYPE, "box3: size mismatch");
  if (s.block != ZOS_BLOCK_PIXEL || d.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "box3: planar image");
  cudaSetDevice(ctx->device);
  return batch ? launch_box3(ctx, s, d, m, batch) : ZOS_OK;
}
zos_status zos_palette(zos_ctx* ctx, const zos_image* pal, const zos_image* idx, const zos_image* dst, const float* xc,
                       const float* yc, uint32_t batch) {
  if (!ctx || !xc || !yc) return ZOS_ERR_INVALID;
  DevImage p, i, d;
  zos_status st;
  if ((st = make_dev_image(ctx, pal, &p, "palette")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, idx, &i, "indices")) != ZOS_OK) return st;
  if ((st = make_dev_image(ctx, dst, &d, "dst")) != ZOS_OK) return st;
  if (i.w != d.w || i.h != d.h) return fail(ctx, ZOS_ERR_TYPE, "palette: indices and dst differ in size");
  if (p.block != ZOS_BLOCK_PIXEL || i.block != ZOS_BLOCK_PIXEL || d.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "palette: planar image");
  cudaSetDevice(ctx->device);
  return batch ? launch_palette(ctx, p, i, d, xc, yc, batch) : ZOS_OK;
}

}  // extern "C"
