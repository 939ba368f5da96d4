// texel.cuh -- device-side texel unpack / pack and transfer functions.
//
// Semantics follow the reference's staging shader (lib/std/src/stage.frag:280-809, codes from
// lib/zosimos/src/shaders/stage.rs:52-132) and the native-vs-staged split of
// lib/zosimos/src/program.rs:781-946:
//   STAGED : demux bits -> reorder parts -> inverse transfer -> value held in an Rgba16Float
//            texture (f16 RNE); pack = f16 -> transfer -> reorder -> clamp -> TRUNCATING quantise.
//   SRGB8 / UNORM8 : what the texture unit does for Rgba8Unorm[Srgb] / Bgra8Unorm[Srgb]:
//            exact decode, round-to-nearest encode (sRGB: correctly rounded via thresholds).
//   FLOAT  : (ours) Float16x4 / Float32x4 texels, transfer per colour, RNE, no clamp.
// Arithmetic is written so that everything except the transcendental calls is bit-identical to
// the CPU oracle: no implicit contraction (--fmad=false), explicit fmaf, IEEE division.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/zosimos_cuda.h"

namespace zos {

struct Tables {           // shared-memory copy, filled once per CTA
  float srgb_dec[256];    // exact sRGB EOTF of k/255
  float unorm8[256];      // k/255 (IEEE)
  float srgb_thr[260];    // thr[k] = smallest f32 whose correctly rounded sRGB8 code is >= k; [0] = -inf, [256] = +inf
};
// Bucket table of the exact sRGB8 encoder (rowwise_lut.cu, frame_pipeline.cu): the f32 bit patterns of
// [2^-13, 1] cut into buckets of 2^16 patterns (sign / exponent / 7 mantissa bits).  No bucket holds more
// than one rounding threshold, so with  e = (base << 16) + (0x10000 - t16) - (top16 << 16)  per bucket
// (base = code at the bucket's start, t16 = low 16 bits of the threshold inside it, or 0x10000),
// (e + bits(x)) >> 16 is the correctly rounded code of x.  Values below 2^-13 encode to 0.
constexpr int ZOS_ENC_B0 = 0x3900;                    // top 16 bits of 2^-13
constexpr int ZOS_ENC_N = 0x3f80 - ZOS_ENC_B0 + 1;    // up to and including the bucket of 1.0
// Second form of the same idea with 2.6x fewer buckets, so that a table with one private copy per lane
// (no shared-memory bank conflicts at all) fits: the bucket key is taken from the bit pattern of
// y = x + 2^-5 instead of x.  Adding the bias compresses the low end, where thresholds are sparse in
// log space, so 7 mantissa bits of y separate all 255 thresholds over [0, 1] with 646 buckets.  A bucket
// is now just an interval of bit patterns of x (float addition is monotone), not an aligned range, and
//   e = (base << 24) + 2^24 - T      (T = bit pattern of the bucket's threshold, or one past its end)
// gives (e + bits(x)) >> 24 == base + (bits(x) >= T) because every bucket is narrower than 2^24 patterns.
constexpr float ZOS_ENC2_BIAS = 0.03125f;
constexpr int ZOS_ENC2_K0 = 0x3d00;                     // bits(bias) >> 16
constexpr int ZOS_ENC2_N = 0x3f85 - ZOS_ENC2_K0 + 1;    // through the bucket of 1 + bias, one spare
constexpr int ZOS_ENC2_LOW = 0x39000000;                // 2^-13: everything below encodes like it (code 0)
struct TablesGlobal {
  float srgb_dec[256];
  float unorm8[256];
  float srgb_thr[260];
  uint32_t srgb_enc[ZOS_ENC_N];    // not part of the per-CTA copy `Tables`
  uint32_t srgb_enc2[ZOS_ENC2_N];  // ditto
};
// One copy per translation unit (no relocatable device code): every kernel .cu exports an
// upload function built from ZOS_DEFINE_CONSTANT_UPLOAD (colorops.cuh) that runtime.cu calls at
// zos_ctx_create.
static __device__ TablesGlobal g_tables;

__device__ __forceinline__ void load_tables(Tables& t) {
  const float* g = reinterpret_cast<const float*>(&g_tables);
  float* s = reinterpret_cast<float*>(&t);
  for (int i = threadIdx.x; i < (int)(sizeof(Tables) / 4); i += blockDim.x) s[i] = g[i];
  __syncthreads();
}

__device__ __forceinline__ float f16r(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

// pow for the transfer curves: exp2(y*log2(x)) on the SFU, the same construction GLSL's pow has.
__device__ __forceinline__ float pow_fast(float x, float y) {
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * l));
  return r;
}

// ---- transfer functions, stage.frag:280-408 ----
__device__ __forceinline__ float oe_bt709(float v) { return v >= 0.018f ? 1.099f * pow_fast(v, 0.45f) - 0.099f : 4.5f * v; }
__device__ __forceinline__ float eo_bt709(float v) {
  const float thr = 0.0812428582f;  // oe_bt709(0.018)
  return v >= thr ? pow_fast((v + 0.099f) * (1.0f / 1.099f), 1.0f / 0.45f) : v * (1.0f / 4.5f);
}
__device__ __forceinline__ float oe_smpte240(float v) { return v < 0.0228f ? 4.0f * v : 1.1115f * pow_fast(v, 0.45f) - 0.1115f; }
__device__ __forceinline__ float eo_smpte240(float v) { return v < 0.0913f ? v * 0.25f : pow_fast((v - 0.1115f) * (1.0f / 1.1115f), 1.0f / 0.45f); }
__device__ __forceinline__ float oe_srgb(float v) {
  if (v < -0.0031308f) return -1.055f * pow_fast(-v, 1.0f / 2.4f) + 0.055f;
  if (v <= 0.0031308f) return v * 12.92f;
  return 1.055f * pow_fast(v, 1.0f / 2.4f) - 0.055f;
}
__device__ __forceinline__ float eo_srgb(float v) {
  if (v < -0.04045f) return -pow_fast((-v + 0.055f) * (1.0f / 1.055f), 2.4f);
  if (v <= 0.04045f) return v * (1.0f / 12.92f);
  return pow_fast((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}
#define ZOS_PQ_M1 (2610.0f / 16384.0f)
#define ZOS_PQ_M2 (2523.0f / 4096.0f)
#define ZOS_PQ_C1 (3424.0f / 4096.0f)
#define ZOS_PQ_C2 (2413.0f / 128.0f)
#define ZOS_PQ_C3 (2392.0f / 128.0f)
__device__ __forceinline__ float pq_eo(float v) {
  float n = pow_fast(v, 1.0f / ZOS_PQ_M2);
  float nom = fmaxf(n - ZOS_PQ_C1, 0.0f);
  float den = ZOS_PQ_C2 - ZOS_PQ_C3 * n;
  return pow_fast(nom / den, 1.0f / ZOS_PQ_M1);
}
__device__ __forceinline__ float pq_oe(float v) {  // transfer_oe_smpte2084, stage.frag:403-405
  float sd = pow_fast(oe_bt709(59.5208f * v), 2.4f) / 100.0f;
  float y = pow_fast(sd, ZOS_PQ_M1);
  return pow_fast((ZOS_PQ_C1 + ZOS_PQ_C2 * y) / (ZOS_PQ_C3 * y + 1.0f), ZOS_PQ_M2);
}
__device__ __forceinline__ float oe_scalar(uint32_t tr, float v) {
  switch (tr) {
    case ZOS_TRANSFER_SRGB: return oe_srgb(v);
    case ZOS_TRANSFER_BT709: case ZOS_TRANSFER_BT2020_10BIT: case ZOS_TRANSFER_BT2020_12BIT: return oe_bt709(v);
    case ZOS_TRANSFER_BT470M: return pow_fast(v, 1.0f / 2.2f);
    case ZOS_TRANSFER_BT601: return eo_bt709(v);  // swapped pair, stage.frag:308-315
    case ZOS_TRANSFER_SMPTE240: return oe_smpte240(v);
    case ZOS_TRANSFER_SMPTE2084: return pq_oe(v);
    default: return v;
  }
}
__device__ __forceinline__ float eo_scalar(uint32_t tr, float v) {
  switch (tr) {
    case ZOS_TRANSFER_SRGB: return eo_srgb(v);
    case ZOS_TRANSFER_BT709: case ZOS_TRANSFER_BT2020_10BIT: case ZOS_TRANSFER_BT2020_12BIT: return eo_bt709(v);
    case ZOS_TRANSFER_BT470M: return pow_fast(v, 2.2f);
    case ZOS_TRANSFER_BT601: return oe_bt709(v);
    case ZOS_TRANSFER_SMPTE240: return eo_smpte240(v);
    case ZOS_TRANSFER_SMPTE2084: return pq_eo(v);
    default: return v;
  }
}

#define ZOS_PI_F 3.14159265358979323846f
// atan2 and sqrt for the Lab <-> LCh form (stage.frag:415-425).  GLSL leaves the precision of atan / sqrt to the
// device; the library versions cost ~45 and ~10 instructions with slow-path branches.  These cost ~22 and 2:
// odd minimax polynomial of degree 15 on [0, 1] (max error 4e-8 rad before rounding, i.e. the accuracy class of
// atan2f) after the usual octant reduction, and the SFU's rsqrt-based approximation.
__device__ __forceinline__ float atan2_fast(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(mx));
  const float t = mx > 0.0f ? mn * r : 0.0f;
  const float s = t * t;
  float p = -0.004054565913975239f;
  p = fmaf(p, s, 0.021862953901290894f);
  p = fmaf(p, s, -0.0559123232960701f);
  p = fmaf(p, s, 0.0964219719171524f);
  p = fmaf(p, s, -0.1390862911939621f);
  p = fmaf(p, s, 0.19946566224098206f);
  p = fmaf(p, s, -0.33329859375953674f);
  p = fmaf(p, s, 0.9999993443489075f);
  p = p * t;
  if (ay > ax) p = 1.57079632679489662f - p;
  if (x < 0.0f) p = 3.14159265358979324f - p;
  return copysignf(p, y);
}
__device__ __forceinline__ float sqrt_fast(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// parts_transfer / parts_untransfer (stage.frag:750-809): rgb through the curve, alpha untouched
__device__ __forceinline__ void transfer_encode(uint32_t tr, float4& c) {
  if (tr == ZOS_TRANSFER_LINEAR) return;
  if (tr == ZOS_TRANSFER_LABLCH) {  // stage.frag:415-420
    float a = c.y, b = c.z;
    c.y = sqrt_fast(a * a + b * b);
    c.z = (atan2_fast(b, a) * (180.0f / ZOS_PI_F)) / 360.0f + 0.5f;
    return;
  }
  c.x = oe_scalar(tr, c.x); c.y = oe_scalar(tr, c.y); c.z = oe_scalar(tr, c.z);
}
__device__ __forceinline__ void transfer_decode(uint32_t tr, float4& c) {
  if (tr == ZOS_TRANSFER_LINEAR) return;
  if (tr == ZOS_TRANSFER_LABLCH) {  // stage.frag:422-425
    float ang = (360.0f * (c.z - 0.5f)) * (ZOS_PI_F / 180.0f);
    float C = c.y, sn, cs;
    __sincosf(ang, &sn, &cs);
    c.y = C * cs; c.z = C * sn;
    return;
  }
  c.x = eo_scalar(tr, c.x); c.y = eo_scalar(tr, c.y); c.z = eo_scalar(tr, c.z);
}

// ---- bit fields, stage.frag:533-641 ----
// v / d for an integer field value: reciprocal multiply plus one Newton step == the IEEE quotient for every
// field width the formats use (3, 7, 15, 31, 63, 255, 1023, 65535 checked exhaustively; tests/ cover it on device)
__device__ __forceinline__ float fld(uint32_t v, float d) {
  const float r = 1.0f / d;
  const float c = (float)v;
  const float q = c * r;
  return fmaf(fmaf(-q, d, c), r, q);
}
__device__ __forceinline__ float4 demux(uint32_t n, uint32_t kind, const Tables& T) {
  switch (kind) {
    case ZOS_BITS_UINT8X4: return make_float4(T.unorm8[n & 255], T.unorm8[(n >> 8) & 255], T.unorm8[(n >> 16) & 255], T.unorm8[n >> 24]);
    case ZOS_BITS_UINT1010102: return make_float4(fld(n & 1023, 1023.0f), fld((n >> 10) & 1023, 1023.0f), fld((n >> 20) & 1023, 1023.0f), fld(n >> 30, 3.0f));
    case ZOS_BITS_UINT2101010: return make_float4(fld(n & 3, 3.0f), fld((n >> 2) & 1023, 1023.0f), fld((n >> 12) & 1023, 1023.0f), fld(n >> 22, 1023.0f));
    case ZOS_BITS_UINT8: { float x = T.unorm8[n & 255]; return make_float4(x, x, x, x); }
    case ZOS_BITS_UINT16: { float x = fld(n & 65535, 65535.0f); return make_float4(x, x, x, x); }
    case ZOS_BITS_UINT8X2: return make_float4(T.unorm8[n & 255], 0.0f, 0.0f, T.unorm8[(n >> 8) & 255]);
    case ZOS_BITS_UINT8X3: return make_float4(T.unorm8[n & 255], T.unorm8[(n >> 8) & 255], T.unorm8[(n >> 16) & 255], 1.0f);
    case ZOS_BITS_UINT16X2: return make_float4(fld(n & 65535, 65535.0f), fld(n >> 16, 65535.0f), 0.0f, 1.0f);
    case ZOS_BITS_UINT332: return make_float4(fld(n & 3, 3.0f), fld((n >> 2) & 7, 7.0f), fld((n >> 5) & 7, 7.0f), 1.0f);
    case ZOS_BITS_UINT233: return make_float4(fld(n & 7, 7.0f), fld((n >> 3) & 7, 7.0f), fld((n >> 6) & 3, 3.0f), 1.0f);
    case ZOS_BITS_UINT4X4: return make_float4(fld(n & 15, 15.0f), fld((n >> 4) & 15, 15.0f), fld((n >> 8) & 15, 15.0f), fld((n >> 12) & 15, 15.0f));
    case ZOS_BITS_UINT_444: return make_float4(fld(n & 15, 15.0f), fld((n >> 4) & 15, 15.0f), fld((n >> 8) & 15, 15.0f), 1.0f);
    case ZOS_BITS_UINT444_: return make_float4(fld((n >> 4) & 15, 15.0f), fld((n >> 9) & 15, 15.0f), fld((n >> 12) & 15, 15.0f), 1.0f);
    case ZOS_BITS_UINT565: return make_float4(fld(n & 31, 31.0f), fld((n >> 5) & 63, 63.0f), fld((n >> 11) & 31, 31.0f), 1.0f);
    case ZOS_BITS_UINT101010_: return make_float4(fld((n >> 2) & 1023, 1023.0f), fld((n >> 12) & 1023, 1023.0f), fld(n >> 22, 1023.0f), 1.0f);
    case ZOS_BITS_UINT_101010: return make_float4(fld(n & 1023, 1023.0f), fld((n >> 10) & 1023, 1023.0f), fld((n >> 20) & 1023, 1023.0f), 1.0f);
  }
  return make_float4(1.0f, 0.0f, 0.0f, 1.0f);  // BIT_DECODE_FAIL
}
__device__ __forceinline__ uint32_t qz(float c, float s) { return (uint32_t)(c * s); }  // truncation
__device__ __forceinline__ uint32_t mux(const float4& c, uint32_t kind) {
  switch (kind) {
    case ZOS_BITS_UINT8X4: return qz(c.x, 255.0f) + (qz(c.y, 255.0f) << 8) + (qz(c.z, 255.0f) << 16) + (qz(c.w, 255.0f) << 24);
    case ZOS_BITS_UINT1010102: return qz(c.x, 1023.0f) + (qz(c.y, 1023.0f) << 10) + (qz(c.z, 1023.0f) << 20) + (qz(c.w, 3.0f) << 30);
    case ZOS_BITS_UINT2101010: return qz(c.x, 3.0f) + (qz(c.y, 1023.0f) << 2) + (qz(c.z, 1023.0f) << 12) + (qz(c.w, 1023.0f) << 22);
    case ZOS_BITS_UINT8: return qz(c.x, 255.0f);
    case ZOS_BITS_UINT16: return qz(c.x, 65535.0f);
    case ZOS_BITS_UINT8X2: return qz(c.x, 255.0f) + (qz(c.w, 255.0f) << 8);
    case ZOS_BITS_UINT8X3: return qz(c.x, 255.0f) + (qz(c.y, 255.0f) << 8) + (qz(c.z, 255.0f) << 16);
    case ZOS_BITS_UINT16X2: return qz(c.x, 65535.0f) + (qz(c.w, 65535.0f) << 16);  // .w on encode (stage.frag:592,619)
    case ZOS_BITS_UINT332: return qz(c.x, 3.0f) + (qz(c.y, 7.0f) << 2) + (qz(c.z, 7.0f) << 5);
    case ZOS_BITS_UINT233: return qz(c.x, 7.0f) + (qz(c.y, 7.0f) << 3) + (qz(c.z, 3.0f) << 6);
    case ZOS_BITS_UINT4X4: return qz(c.x, 15.0f) + (qz(c.y, 15.0f) << 4) + (qz(c.z, 15.0f) << 8) + (qz(c.w, 15.0f) << 12);
    case ZOS_BITS_UINT_444: return qz(c.x, 15.0f) + (qz(c.y, 15.0f) << 4) + (qz(c.z, 15.0f) << 8);
    case ZOS_BITS_UINT444_: return (qz(c.x, 15.0f) << 4) + (qz(c.y, 15.0f) << 8) + (qz(c.z, 15.0f) << 12);
    case ZOS_BITS_UINT565: return qz(c.x, 31.0f) + (qz(c.y, 63.0f) << 5) + (qz(c.z, 31.0f) << 11);
    case ZOS_BITS_UINT101010_: return (qz(c.x, 1023.0f) << 2) + (qz(c.y, 1023.0f) << 12) + (qz(c.z, 1023.0f) << 22);
    case ZOS_BITS_UINT_101010: return qz(c.x, 1023.0f) + (qz(c.y, 1023.0f) << 10) + (qz(c.z, 1023.0f) << 20);
  }
  return 0x55445544u;  // BIT_ENCODE_FAIL
}
// parts_normalize, stage.frag:654-699
__device__ __forceinline__ float4 parts_norm(const float4& c, uint32_t parts) {
  switch (parts) {
    case ZOS_PARTS_RGBA: case ZOS_PARTS_BGRA: case ZOS_PARTS_LABA: case ZOS_PARTS_LCHA: return c;  // Bgra not swizzled on decode
    case ZOS_PARTS_A: return make_float4(0.0f, 0.0f, 0.0f, c.x);
    case ZOS_PARTS_R: return make_float4(c.x, 0.0f, 0.0f, 1.0f);
    case ZOS_PARTS_G: return make_float4(0.0f, c.x, 0.0f, 1.0f);
    case ZOS_PARTS_B: return make_float4(0.0f, 0.0f, c.x, 1.0f);
    case ZOS_PARTS_LUMA: return make_float4(c.x, c.x, c.x, 1.0f);
    case ZOS_PARTS_LUMAA: return make_float4(c.x, c.x, c.x, c.w);
    case ZOS_PARTS_RGB: case ZOS_PARTS_RGBX: case ZOS_PARTS_LAB: case ZOS_PARTS_LCH: return make_float4(c.x, c.y, c.z, 1.0f);
    case ZOS_PARTS_BGR: case ZOS_PARTS_BGRX: return make_float4(c.z, c.y, c.x, 1.0f);
    case ZOS_PARTS_ARGB: return make_float4(c.y, c.z, c.w, c.x);
    case ZOS_PARTS_ABGR: return make_float4(c.w, c.z, c.y, c.x);
    case ZOS_PARTS_XRGB: return make_float4(c.y, c.z, c.w, 1.0f);
    case ZOS_PARTS_XBGR: return make_float4(c.w, c.z, c.y, 1.0f);
  }
  return make_float4(1.0f, 0.0f, 0.0f, 1.0f);
}
// parts_denormalize, stage.frag:703-747
__device__ __forceinline__ float4 parts_denorm(const float4& c, uint32_t parts) {
  switch (parts) {
    case ZOS_PARTS_A: return make_float4(c.w, 0.0f, 0.0f, 1.0f);
    case ZOS_PARTS_R: return make_float4(c.x, 0.0f, 0.0f, 1.0f);
    case ZOS_PARTS_G: return make_float4(c.y, 0.0f, 0.0f, 1.0f);
    case ZOS_PARTS_B: return make_float4(c.z, 0.0f, 0.0f, 1.0f);
    case ZOS_PARTS_LUMA: return make_float4(c.x, c.x, c.x, 1.0f);
    case ZOS_PARTS_LUMAA: return make_float4(c.x, c.x, c.x, c.w);
    case ZOS_PARTS_RGB: case ZOS_PARTS_RGBX: case ZOS_PARTS_LAB: case ZOS_PARTS_LCH: return make_float4(c.x, c.y, c.z, 1.0f);
    case ZOS_PARTS_BGR: case ZOS_PARTS_BGRX: return make_float4(c.z, c.y, c.x, 1.0f);
    case ZOS_PARTS_BGRA: return make_float4(c.z, c.y, c.x, c.w);
    case ZOS_PARTS_ARGB: return make_float4(c.w, c.x, c.y, c.z);
    case ZOS_PARTS_ABGR: return make_float4(c.w, c.z, c.y, c.x);
    case ZOS_PARTS_XRGB: return make_float4(1.0f, c.x, c.y, c.z);
    case ZOS_PARTS_XBGR: return make_float4(1.0f, c.z, c.y, c.x);
  }
  return c;
}

// Correctly rounded sRGB8 code of a linear value: an SFU estimate fixed up against the threshold table.
__device__ __forceinline__ uint32_t srgb8_encode(float x, const Tables& T) {
  x = clamp01(x);  // NaN -> 0
  float e = x <= 0.0031308f ? 12.92f * x : 1.055f * pow_fast(x, 1.0f / 2.4f) - 0.055f;
  int k = __float2int_rn(e * 255.0f);
  k = min(max(k, 0), 255);
  k += (x >= T.srgb_thr[k + 1]) ? 1 : 0;
  k -= (x < T.srgb_thr[k]) ? 1 : 0;
  return (uint32_t)k;
}
__device__ __forceinline__ uint32_t unorm8_rne(float x) { return (uint32_t)__float2int_rn(clamp01(x) * 255.0f); }

// ---- whole texels.  A texel travels as a uint4 (only .x for <= 4 bytes, .x/.y for 8 bytes) ----
#ifndef ZOS_SLOW_ATTR
#define ZOS_SLOW_ATTR __noinline__
#endif
static __device__ ZOS_SLOW_ATTR float4 unpack_slow(zos_texfmt f, uint4 w, const Tables* Tp) {
  const Tables& T = *Tp;
  switch (f.storage) {
    case ZOS_STORAGE_SRGB8: {
      uint32_t n = w.x;
      float a = T.srgb_dec[n & 255], b = T.srgb_dec[(n >> 8) & 255], c = T.srgb_dec[(n >> 16) & 255], al = T.unorm8[n >> 24];
      return f.parts == ZOS_PARTS_BGRA ? make_float4(c, b, a, al) : make_float4(a, b, c, al);
    }
    case ZOS_STORAGE_UNORM8: {
      uint32_t n = w.x;
      float a = T.unorm8[n & 255], b = T.unorm8[(n >> 8) & 255], c = T.unorm8[(n >> 16) & 255], al = T.unorm8[n >> 24];
      return f.parts == ZOS_PARTS_BGRA ? make_float4(c, b, a, al) : make_float4(a, b, c, al);
    }
    case ZOS_STORAGE_FLOAT: {
      float4 c;
      if (f.bits == ZOS_BITS_FLOAT16X4) {
        __half2 lo = *reinterpret_cast<const __half2*>(&w.x), hi = *reinterpret_cast<const __half2*>(&w.y);
        float2 a = __half22float2(lo), b = __half22float2(hi);
        c = make_float4(a.x, a.y, b.x, b.y);
      } else {
        c = make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
      }
      c = parts_norm(c, f.parts);
      transfer_decode(f.transfer, c);
      return c;
    }
    default: {
      // UInt16x4 (8 bytes): ours -- the reference declares decode_rgba16ui / encode_rgba16ui (shaders/stage.rs:148-158)
      // but stage.frag never defines them; same pipeline as every staged texel, fields of 16 bits
      float4 d = f.bits == ZOS_BITS_UINT16X4 ? make_float4(fld(w.x & 65535u, 65535.0f), fld(w.x >> 16, 65535.0f), fld(w.y & 65535u, 65535.0f), fld(w.y >> 16, 65535.0f))
                                            : demux(w.x, f.bits, T);
      float4 c = parts_norm(d, f.parts);
      transfer_decode(f.transfer, c);
      return make_float4(f16r(c.x), f16r(c.y), f16r(c.z), f16r(c.w));
    }
  }
}
static __device__ ZOS_SLOW_ATTR uint4 pack_slow(zos_texfmt f, float4 v, const Tables* Tp) {
  const Tables& T = *Tp;
  uint4 w = make_uint4(0, 0, 0, 0);
  switch (f.storage) {
    case ZOS_STORAGE_SRGB8: {
      uint32_t r = srgb8_encode(v.x, T), g = srgb8_encode(v.y, T), b = srgb8_encode(v.z, T), a = unorm8_rne(v.w);
      w.x = f.parts == ZOS_PARTS_BGRA ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24));
      return w;
    }
    case ZOS_STORAGE_UNORM8: {
      uint32_t r = unorm8_rne(v.x), g = unorm8_rne(v.y), b = unorm8_rne(v.z), a = unorm8_rne(v.w);
      w.x = f.parts == ZOS_PARTS_BGRA ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24));
      return w;
    }
    case ZOS_STORAGE_FLOAT: {
      transfer_encode(f.transfer, v);
      float4 c = parts_denorm(v, f.parts);
      if (f.bits == ZOS_BITS_FLOAT16X4) {
        __half2 lo = __floats2half2_rn(c.x, c.y), hi = __floats2half2_rn(c.z, c.w);
        w.x = *reinterpret_cast<uint32_t*>(&lo); w.y = *reinterpret_cast<uint32_t*>(&hi);
      } else {
        w = make_uint4(__float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), __float_as_uint(c.w));
      }
      return w;
    }
    default: {
      v = make_float4(f16r(v.x), f16r(v.y), f16r(v.z), f16r(v.w));
      transfer_encode(f.transfer, v);
      float4 c = parts_denorm(v, f.parts);
      c = make_float4(clamp01(c.x), clamp01(c.y), clamp01(c.z), clamp01(c.w));
      if (f.bits == ZOS_BITS_UINT16X4) {
        w.x = qz(c.x, 65535.0f) + (qz(c.y, 65535.0f) << 16);
        w.y = qz(c.z, 65535.0f) + (qz(c.w, 65535.0f) << 16);
        return w;
      }
      w.x = mux(c, f.bits);
      return w;
    }
  }
}


// Inline fast paths for the texels the headline workloads use (native 8-bit RGBA/BGRA, linear
// half/float RGBA); everything else goes through the out-of-line generic codec above.
__device__ __forceinline__ bool is_plain_float(const zos_texfmt& f) {
  return f.storage == ZOS_STORAGE_FLOAT && f.transfer == ZOS_TRANSFER_LINEAR &&
         (f.parts == ZOS_PARTS_RGBA || f.parts == ZOS_PARTS_LCHA || f.parts == ZOS_PARTS_LABA);
}
__device__ __forceinline__ float4 unpack_texel(const zos_texfmt& f, const uint4& w, const Tables& T) {
  if (f.storage == ZOS_STORAGE_SRGB8 || f.storage == ZOS_STORAGE_UNORM8) {
    const float* lut = f.storage == ZOS_STORAGE_SRGB8 ? T.srgb_dec : T.unorm8;
    uint32_t n = w.x;
    float a = lut[n & 255], b = lut[(n >> 8) & 255], c = lut[(n >> 16) & 255], al = T.unorm8[n >> 24];
    return f.parts == ZOS_PARTS_BGRA ? make_float4(c, b, a, al) : make_float4(a, b, c, al);
  }
  if (is_plain_float(f)) {
    if (f.bits == ZOS_BITS_FLOAT16X4) {
      float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
      return make_float4(a.x, a.y, b.x, b.y);
    }
    return make_float4(__uint_as_float(w.x), __uint_as_float(w.y), __uint_as_float(w.z), __uint_as_float(w.w));
  }
  return unpack_slow(f, w, &T);
}
__device__ __forceinline__ uint4 pack_texel(const zos_texfmt& f, const float4& v, const Tables& T) {
  if (f.storage == ZOS_STORAGE_SRGB8) {
    uint32_t r = srgb8_encode(v.x, T), g = srgb8_encode(v.y, T), b = srgb8_encode(v.z, T), a = unorm8_rne(v.w);
    return make_uint4(f.parts == ZOS_PARTS_BGRA ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24)), 0, 0, 0);
  }
  if (f.storage == ZOS_STORAGE_UNORM8) {
    uint32_t r = unorm8_rne(v.x), g = unorm8_rne(v.y), b = unorm8_rne(v.z), a = unorm8_rne(v.w);
    return make_uint4(f.parts == ZOS_PARTS_BGRA ? (b | (g << 8) | (r << 16) | (a << 24)) : (r | (g << 8) | (b << 16) | (a << 24)), 0, 0, 0);
  }
  if (is_plain_float(f)) {
    if (f.bits == ZOS_BITS_FLOAT16X4) {
      __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      return make_uint4(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi), 0, 0);
    }
    return make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
  }
  return pack_slow(f, v, &T);
}

}  // namespace zos
