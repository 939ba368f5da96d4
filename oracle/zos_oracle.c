/* zos_oracle.c -- CPU restatement of the zosimos compositing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (zosimos_b200/, the C-ABI
 * library) may include, link or call this file.  It is used by tests/, by
 * __graft_entry__.smoke() and by the cpu_baseline / --impl reference legs of
 * bench.py, and only as the checker / the timed CPU baseline.
 *
 * What it restates (reference = /root/reference, 197g/zosimos; file:line):
 *   - texel decode/encode ("staging"):   lib/std/src/stage.frag:439-809 and the
 *     native-vs-staged decision of lib/zosimos/src/program.rs:781-946
 *   - colour operators:                   lib/std/src/linear.frag:12-17,
 *     lib/std/src/oklab.frag:34-64, lib/std/src/srlab2.frag:36-120
 *   - painting (crop/inscribe/affine):    lib/std/src/box.vert:43-59,
 *     lib/std/src/copy.frag:8-10, lib/zosimos/src/program.rs:1897-1935
 *   - resize = bilinear grid + palette:   lib/std/src/bilinear.frag:14-20,
 *     lib/std/src/palette.frag:21-32
 *   - inject / box3 / solid:              lib/std/src/inject.frag:20-25,
 *     lib/std/src/box3.frag:16-52, lib/std/src/solid_rgb.frag:9-11
 *
 * The reference runs every operation as its own render pass over textures:
 * a register lives in memory in its declared texel format; an operand is
 * DECODED into a texture before use (program.rs:1475-1478) and the result of
 * each draw is ENCODED back (program.rs:1531-1532).  Staged formats use an
 * Rgba16Float working texture (shaders/stage.rs:129-131) and quantise by
 * truncation (stage.frag:591-594); native formats (RGBA8/BGRA8, sRGB or
 * linear) are converted by the texture unit (round to nearest).  The oracle
 * keeps that pass structure: every function here is one pass over a float
 * RGBA "texture" (4 floats per texel, tightly packed).
 *
 * Features the reference does not implement (blend,
 * bilinear sampling, RGBA16F/RGBA32F texels, planar YUV) follow the semantics
 * written down in DESIGN.md ("Semantics we define"); parity for those is
 * UNPINNED by construction.
 *
 * Arithmetic rules (shared with the CUDA kernels so results can be compared
 * bit for bit wherever no transcendental function is involved): IEEE f32,
 * no implicit contraction (-ffp-contract=off), explicit fmaf where a fused
 * multiply-add is intended, f16 rounding = IEEE RNE.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ZO_API __attribute__((visibility("default")))

/* ---- numeric codes (identical to stage.frag:107-170 / shaders/stage.rs:74-119) ---- */
enum {
  TR_BT709 = 0, TR_BT470M = 1, TR_BT601 = 2, TR_SMPTE240 = 3, TR_LINEAR = 4, TR_SRGB = 5,
  TR_BT2020_10 = 6, TR_BT2020_12 = 7, TR_SMPTE2084 = 8, TR_BT2100PQ = 9, TR_BT2100HLG = 10,
  TR_LINEAR_SCENE = 11, TR_LABLCH = 0x100
};
enum {
  P_A = 0, P_R = 1, P_G = 2, P_B = 3, P_LUMA = 4, P_LUMAA = 5, P_RGB = 6, P_BGR = 7, P_RGBA = 8,
  P_RGBX = 9, P_BGRA = 10, P_BGRX = 11, P_ARGB = 12, P_XRGB = 13, P_ABGR = 14, P_XBGR = 15,
  P_YUV = 16, P_LAB = 17, P_LABA = 18, P_LCH = 19, P_LCHA = 20
};
enum {
  B_INT8 = 0, B_INT332 = 1, B_INT233 = 2, B_INT16 = 3, B_INT4X4 = 4, B_INTI444 = 5, B_INT444I = 6,
  B_INT565 = 7, B_INT8X2 = 8, B_INT8X3 = 9, B_INT8X4 = 10, B_INT16X2 = 11, B_INT16X3 = 12,
  B_INT16X4 = 13, B_INT1010102 = 14, B_INT2101010 = 15, B_INT101010I = 16, B_INTI101010 = 17,
  B_FLOAT16X4 = 18, B_FLOAT32X4 = 19
};
/* storage classes: how a register's bytes become a texture and back */
enum {
  ST_STAGED = 0,  /* stage.frag decode/encode through an Rgba16Float texture    */
  ST_SRGB8 = 1,   /* native Rgba8UnormSrgb / Bgra8UnormSrgb (program.rs:794-816) */
  ST_UNORM8 = 2,  /* native Rgba8Unorm / Bgra8Unorm        (program.rs:817-838) */
  ST_FLOAT = 3    /* OURS: Float16x4 / Float32x4 texels, transfer per colour    */
};

typedef struct zo_fmt {
  uint32_t transfer; /* TR_*                                 */
  uint32_t parts;    /* P_*                                  */
  uint32_t bits;     /* B_*                                  */
  uint32_t storage;  /* ST_*                                 */
} zo_fmt;

static inline float f16r(float x) { return (float)(_Float16)x; }
static inline float clamp01(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }

/* ------------------------------------------------------------------ */
/* Transfer functions, stage.frag:280-425                              */
/* ------------------------------------------------------------------ */
static float oe_bt709(float v) { return v >= 0.018f ? 1.099f * powf(v, 0.45f) - 0.099f : 4.5f * v; }
static float eo_bt709(float v) {
  const float thr = 1.099f * powf(0.018f, 0.45f) - 0.099f; /* oe_bt709(0.018), stage.frag:291 */
  return v >= thr ? powf((v + 0.099f) / 1.099f, 1.0f / 0.45f) : v / 4.5f;
}
static float oe_bt470m(float v) { return powf(v, 1.0f / 2.2f); }
static float eo_bt470m(float v) { return powf(v, 2.2f); }
/* stage.frag:308-315: the 601 pair is the 709 pair with the roles swapped */
static float oe_bt601(float v) { return eo_bt709(v); }
static float eo_bt601(float v) { return oe_bt709(v); }
static float oe_smpte240(float v) { return v < 0.0228f ? 4.0f * v : 1.1115f * powf(v, 0.45f) - 0.1115f; }
static float eo_smpte240(float v) { return v < 0.0913f ? v / 4.0f : powf((v - 0.1115f) / 1.1115f, 1.0f / 0.45f); }
static float oe_srgb(float v) {
  if (v < -0.0031308f) return -1.055f * powf(-v, 1.0f / 2.4f) + 0.055f;
  if (v <= 0.0031308f) return v * 12.92f;
  return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
static float eo_srgb(float v) {
  if (v < -0.04045f) return -powf((-v + 0.055f) / 1.055f, 2.4f);
  if (v <= 0.04045f) return v / 12.92f;
  return powf((v + 0.055f) / 1.055f, 2.4f);
}
#define PQ_M1 (2610.0f / 16384.0f)
#define PQ_M2 (2523.0f / 4096.0f)
#define PQ_C1 (3424.0f / 4096.0f)
#define PQ_C2 (2413.0f / 128.0f)
#define PQ_C3 (2392.0f / 128.0f)
static float pq_eo(float v) {
  float n = powf(v, 1.0f / PQ_M2);
  float nom = fmaxf(n - PQ_C1, 0.0f);
  float den = PQ_C2 - PQ_C3 * n;
  return powf(nom / den, 1.0f / PQ_M1);
}
static float pq_eo_inv(float v) {
  float y = powf(v, PQ_M1);
  return powf((PQ_C1 + PQ_C2 * y) / (PQ_C3 * y + 1.0f), PQ_M2);
}
static float pq_scene_display(float v) { return powf(oe_bt709(59.5208f * v), 2.4f) / 100.0f; }
static float pq_display_scene(float v) { return eo_bt709(powf(v * 100.0f, 1.0f / 2.4f)) / 59.5208f; }
static float oe_smpte2084(float v) { return pq_eo_inv(pq_scene_display(v)); }

static float oe_scalar(uint32_t tr, float v) {
  switch (tr) {
    case TR_BT709: return oe_bt709(v);
    case TR_BT470M: return oe_bt470m(v);
    case TR_BT601: return oe_bt601(v);
    case TR_SMPTE240: return oe_smpte240(v);
    case TR_SRGB: return oe_srgb(v);
    case TR_BT2020_10: case TR_BT2020_12: return oe_bt709(v);
    case TR_SMPTE2084: return oe_smpte2084(v);
    default: return v; /* Linear, Bt2100Pq, Bt2100Hlg, LinearScene: identity (stage.frag:771-777) */
  }
}
static float eo_scalar(uint32_t tr, float v) {
  switch (tr) {
    case TR_BT709: return eo_bt709(v);
    case TR_BT470M: return eo_bt470m(v);
    case TR_BT601: return eo_bt601(v);
    case TR_SMPTE240: return eo_smpte240(v);
    case TR_SRGB: return eo_srgb(v);
    case TR_BT2020_10: case TR_BT2020_12: return eo_bt709(v);
    case TR_SMPTE2084: return pq_eo(v); /* stage.frag:799-800 uses the plain EOTF on decode */
    default: return v;
  }
}

#define ZO_PI_F 3.14159265358979323846f
/* stage.frag:415-425.  degrees()/radians() are a multiply by 180/pi resp. pi/180 */
static void lab_to_lch(float* c) {
  float L = c[0], a = c[1], b = c[2];
  float C = sqrtf(a * a + b * b);
  float h = (atan2f(b, a) * (180.0f / ZO_PI_F)) / 360.0f + 0.5f;
  c[0] = L; c[1] = C; c[2] = h;
}
static void lch_to_lab(float* c) {
  float ang = (360.0f * (c[2] - 0.5f)) * (ZO_PI_F / 180.0f);
  float C = c[1];
  c[1] = C * cosf(ang);
  c[2] = C * sinf(ang);
}
/* parts_transfer / parts_untransfer, stage.frag:750-809: rgb through the curve, alpha untouched */
static void transfer_apply(uint32_t tr, float* c, int encode) {
  if (tr == TR_LABLCH) { if (encode) lab_to_lch(c); else lch_to_lab(c); return; }
  if (tr == TR_LINEAR) return;
  for (int i = 0; i < 3; i++) c[i] = encode ? oe_scalar(tr, c[i]) : eo_scalar(tr, c[i]);
}

/* ------------------------------------------------------------------ */
/* Bit (de)multiplexing, stage.frag:533-641                            */
/* ------------------------------------------------------------------ */
static int bits_bytes(uint32_t bits) {
  switch (bits) {
    case B_INT8: case B_INT332: case B_INT233: return 1;
    case B_INT16: case B_INT4X4: case B_INTI444: case B_INT444I: case B_INT565: case B_INT8X2: return 2;
    case B_INT8X3: return 3;
    case B_INT8X4: case B_INT16X2: case B_INT1010102: case B_INT2101010: case B_INT101010I: case B_INTI101010: return 4;
    case B_INT16X3: return 6;
    case B_INT16X4: case B_FLOAT16X4: return 8;
    case B_FLOAT32X4: return 16;
  }
  return 0;
}
ZO_API int zo_bits_bytes(uint32_t bits) { return bits_bytes(bits); }

static const float FAIL_DEC[4] = {1.0f, 0.0f, 0.0f, 1.0f}; /* BIT_DECODE_FAIL */
#define FAIL_ENC 0x55445544u                                /* BIT_ENCODE_FAIL */

static void demux(uint32_t n, uint32_t kind, float* o) {
#define F(v, d) ((float)(v) / (d))
  switch (kind) {
    case B_INT8: o[0] = o[1] = o[2] = o[3] = F(n, 255.0f); return;
    case B_INT332: o[0] = F(n & 3, 3.0f); o[1] = F((n >> 2) & 7, 7.0f); o[2] = F(n >> 5, 7.0f); o[3] = 1.0f; return;
    case B_INT233: o[0] = F(n & 7, 7.0f); o[1] = F((n >> 3) & 7, 7.0f); o[2] = F(n >> 6, 3.0f); o[3] = 1.0f; return;
    case B_INT16: o[0] = o[1] = o[2] = o[3] = F(n, 65535.0f); return;
    case B_INT4X4: o[0] = F(n & 15, 15.0f); o[1] = F((n >> 4) & 15, 15.0f); o[2] = F((n >> 8) & 15, 15.0f); o[3] = F(n >> 12, 15.0f); return;
    case B_INTI444: o[0] = F(n & 15, 15.0f); o[1] = F((n >> 4) & 15, 15.0f); o[2] = F((n >> 8) & 15, 15.0f); o[3] = 1.0f; return;
    case B_INT444I: o[0] = F((n >> 4) & 15, 15.0f); o[1] = F((n >> 9) & 15, 15.0f); o[2] = F((n >> 12) & 15, 15.0f); o[3] = 1.0f; return;
    case B_INT565: o[0] = F(n & 31, 31.0f); o[1] = F((n >> 5) & 63, 63.0f); o[2] = F(n >> 11, 31.0f); o[3] = 1.0f; return;
    case B_INT8X2: o[0] = F(n & 255, 255.0f); o[1] = 0.0f; o[2] = 0.0f; o[3] = F((n >> 8) & 255, 255.0f); return;
    case B_INT8X3: o[0] = F(n & 255, 255.0f); o[1] = F((n >> 8) & 255, 255.0f); o[2] = F((n >> 16) & 255, 255.0f); o[3] = 1.0f; return;
    case B_INT8X4: o[0] = F(n & 255, 255.0f); o[1] = F((n >> 8) & 255, 255.0f); o[2] = F((n >> 16) & 255, 255.0f); o[3] = F(n >> 24, 255.0f); return;
    case B_INT16X2: o[0] = F(n & 65535, 65535.0f); o[1] = F((n >> 16) & 65535, 65535.0f); o[2] = 0.0f; o[3] = 1.0f; return;
    case B_INT1010102: o[0] = F(n & 3, 3.0f); o[1] = F((n >> 2) & 1023, 1023.0f); o[2] = F((n >> 12) & 1023, 1023.0f); o[3] = F(n >> 22, 1023.0f); return;
    case B_INT2101010: o[0] = F(n & 1023, 1023.0f); o[1] = F((n >> 10) & 1023, 1023.0f); o[2] = F((n >> 20) & 1023, 1023.0f); o[3] = F(n >> 30, 3.0f); return;
    case B_INT101010I: o[0] = F((n >> 2) & 1023, 1023.0f); o[1] = F((n >> 12) & 1023, 1023.0f); o[2] = F(n >> 22, 1023.0f); o[3] = 1.0f; return;
    case B_INTI101010: o[0] = F(n & 1023, 1023.0f); o[1] = F((n >> 10) & 1023, 1023.0f); o[2] = F((n >> 20) & 1023, 1023.0f); o[3] = 1.0f; return;
  }
#undef F
  memcpy(o, FAIL_DEC, sizeof FAIL_DEC);
}

/* uint(c * (2^n-1)): float->uint conversion truncates toward zero */
static inline uint32_t q(float c, float scale) { return (uint32_t)(c * scale); }

static uint32_t mux(const float* c, uint32_t kind) {
  switch (kind) {
    case B_INT8: return q(c[0], 255.0f);
    case B_INT332: return q(c[0], 3.0f) + (q(c[1], 7.0f) << 2) + (q(c[2], 7.0f) << 5);
    case B_INT233: return q(c[0], 7.0f) + (q(c[1], 7.0f) << 3) + (q(c[2], 3.0f) << 6);
    case B_INT16: return q(c[0], 65535.0f);
    case B_INT4X4: return q(c[0], 15.0f) + (q(c[1], 15.0f) << 4) + (q(c[2], 15.0f) << 8) + (q(c[3], 15.0f) << 12);
    case B_INTI444: return q(c[0], 15.0f) + (q(c[1], 15.0f) << 4) + (q(c[2], 15.0f) << 8);
    case B_INT444I: return (q(c[0], 15.0f) << 4) + (q(c[1], 15.0f) << 8) + (q(c[2], 15.0f) << 12);
    case B_INT565: return q(c[0], 31.0f) + (q(c[1], 63.0f) << 5) + (q(c[2], 31.0f) << 11);
    case B_INT8X2: return q(c[0], 255.0f) + (q(c[3], 255.0f) << 8);
    case B_INT8X3: return q(c[0], 255.0f) + (q(c[1], 255.0f) << 8) + (q(c[2], 255.0f) << 16);
    case B_INT8X4: return q(c[0], 255.0f) + (q(c[1], 255.0f) << 8) + (q(c[2], 255.0f) << 16) + (q(c[3], 255.0f) << 24);
    case B_INT16X2: return q(c[0], 65535.0f) + (q(c[3], 65535.0f) << 16); /* encode takes .w (stage.frag:592,619) */
    case B_INT1010102: return q(c[0], 3.0f) + (q(c[1], 1023.0f) << 2) + (q(c[2], 1023.0f) << 12) + (q(c[3], 1023.0f) << 22);
    case B_INT2101010: return q(c[0], 1023.0f) + (q(c[1], 1023.0f) << 10) + (q(c[2], 1023.0f) << 20) + (q(c[3], 3.0f) << 30);
    case B_INT101010I: return (q(c[0], 1023.0f) << 2) + (q(c[1], 1023.0f) << 12) + (q(c[2], 1023.0f) << 22);
    case B_INTI101010: return q(c[0], 1023.0f) + (q(c[1], 1023.0f) << 10) + (q(c[2], 1023.0f) << 20);
  }
  return FAIL_ENC;
}

/* parts_normalize, stage.frag:654-699 */
static void parts_norm(const float* c, uint32_t parts, float* o) {
  float x = c[0], y = c[1], z = c[2], w = c[3];
  switch (parts) {
    case P_A: o[0] = 0; o[1] = 0; o[2] = 0; o[3] = x; return;
    case P_R: o[0] = x; o[1] = 0; o[2] = 0; o[3] = 1; return;
    case P_G: o[0] = 0; o[1] = x; o[2] = 0; o[3] = 1; return;
    case P_B: o[0] = 0; o[1] = 0; o[2] = x; o[3] = 1; return;
    case P_LUMA: o[0] = x; o[1] = x; o[2] = x; o[3] = 1; return;
    case P_LUMAA: o[0] = x; o[1] = x; o[2] = x; o[3] = w; return;
    case P_RGB: case P_RGBX: case P_LAB: case P_LCH: o[0] = x; o[1] = y; o[2] = z; o[3] = 1; return;
    case P_BGR: case P_BGRX: o[0] = z; o[1] = y; o[2] = x; o[3] = 1; return;
    case P_RGBA: case P_BGRA: /* Bgra is NOT swizzled on decode (stage.frag:676-677) */
    case P_LABA: case P_LCHA: o[0] = x; o[1] = y; o[2] = z; o[3] = w; return;
    case P_ARGB: o[0] = y; o[1] = z; o[2] = w; o[3] = x; return;
    case P_ABGR: o[0] = w; o[1] = z; o[2] = y; o[3] = x; return;
    case P_XRGB: o[0] = y; o[1] = z; o[2] = w; o[3] = 1; return;
    case P_XBGR: o[0] = w; o[1] = z; o[2] = y; o[3] = 1; return;
  }
  memcpy(o, FAIL_DEC, sizeof FAIL_DEC);
}
/* parts_denormalize, stage.frag:703-747 */
static void parts_denorm(const float* c, uint32_t parts, float* o) {
  float r = c[0], g = c[1], b = c[2], a = c[3];
  switch (parts) {
    case P_A: o[0] = a; o[1] = 0; o[2] = 0; o[3] = 1; return;
    case P_R: o[0] = r; o[1] = 0; o[2] = 0; o[3] = 1; return;
    case P_G: o[0] = g; o[1] = 0; o[2] = 0; o[3] = 1; return;
    case P_B: o[0] = b; o[1] = 0; o[2] = 0; o[3] = 1; return;
    case P_LUMA: o[0] = r; o[1] = r; o[2] = r; o[3] = 1; return;
    case P_LUMAA: o[0] = r; o[1] = r; o[2] = r; o[3] = a; return;
    case P_RGB: case P_RGBX: case P_LAB: case P_LCH: o[0] = r; o[1] = g; o[2] = b; o[3] = 1; return;
    case P_BGR: case P_BGRX: o[0] = b; o[1] = g; o[2] = r; o[3] = 1; return;
    case P_BGRA: o[0] = b; o[1] = g; o[2] = r; o[3] = a; return;
    case P_ARGB: o[0] = a; o[1] = r; o[2] = g; o[3] = b; return;
    case P_ABGR: o[0] = a; o[1] = b; o[2] = g; o[3] = r; return;
    case P_XRGB: o[0] = 1; o[1] = r; o[2] = g; o[3] = b; return;
    case P_XBGR: o[0] = 1; o[1] = b; o[2] = g; o[3] = r; return;
  }
  o[0] = r; o[1] = g; o[2] = b; o[3] = a;
}

/* ------------------------------------------------------------------ */
/* Native sRGB8: what the texture unit does for Rgba8UnormSrgb.         */
/*   decode: exact EOTF of k/255 (table, evaluated in double).          */
/*   encode: the correctly rounded 8-bit code, i.e. the number of       */
/*   thresholds eotf((k-0.5)/255) that are <= x.                        */
/* ------------------------------------------------------------------ */
static float SRGB_DEC[256];
static float SRGB_THR[257]; /* THR[k], k=1..255; THR[0] = -inf, THR[256] = +inf */
static int tables_ready = 0;
static double eotf_d(double v) { return v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4); }
static void init_tables(void) {
  if (tables_ready) return;
  for (int k = 0; k < 256; k++) SRGB_DEC[k] = (float)eotf_d(k / 255.0);
  SRGB_THR[0] = -INFINITY;
  SRGB_THR[256] = INFINITY;
  for (int k = 1; k < 256; k++) {
    double t = eotf_d((k - 0.5) / 255.0);
    float f = (float)t;
    if ((double)f < t) f = nextafterf(f, INFINITY); /* smallest f32 >= t */
    SRGB_THR[k] = f;
  }
  tables_ready = 1;
}
ZO_API void zo_srgb_tables(float* dec256, float* thr257) {
  init_tables();
  memcpy(dec256, SRGB_DEC, sizeof SRGB_DEC);
  memcpy(thr257, SRGB_THR, sizeof SRGB_THR);
}
static uint32_t srgb8_encode(float x) {
  if (!(x > 0.0f)) return 0; /* also NaN */
  int lo = 0, hi = 255;      /* largest k with THR[k] <= x */
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (SRGB_THR[mid] <= x) lo = mid; else hi = mid - 1;
  }
  return (uint32_t)lo;
}
static inline uint32_t unorm8_rne(float x) { return (uint32_t)rintf(clamp01(x) * 255.0f); }

static inline float half_bits_to_float(uint16_t h) { _Float16 v; memcpy(&v, &h, 2); return (float)v; }
static inline uint16_t float_to_half_bits(float f) { _Float16 v = (_Float16)f; uint16_t h; memcpy(&h, &v, 2); return h; }

/* ------------------------------------------------------------------ */
/* decode: register bytes -> texture                                    */
/* ------------------------------------------------------------------ */
static void decode_texel(const zo_fmt* f, const uint8_t* p, float* o) {
  switch (f->storage) {
    case ST_SRGB8: case ST_UNORM8: {
      float c[4];
      for (int i = 0; i < 3; i++) c[i] = f->storage == ST_SRGB8 ? SRGB_DEC[p[i]] : (float)p[i] / 255.0f;
      c[3] = (float)p[3] / 255.0f;
      if (f->parts == P_BGRA) { o[0] = c[2]; o[1] = c[1]; o[2] = c[0]; } else { o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; }
      o[3] = c[3];
      return;
    }
    case ST_FLOAT: {
      float c[4], n[4];
      if (f->bits == B_FLOAT16X4) { for (int i = 0; i < 4; i++) { uint16_t h; memcpy(&h, p + 2 * i, 2); c[i] = half_bits_to_float(h); } }
      else memcpy(c, p, 16);
      parts_norm(c, f->parts, n);
      transfer_apply(f->transfer, n, 0);
      memcpy(o, n, 16);
      return;
    }
    default: {
      uint32_t n = 0;
      int nb = bits_bytes(f->bits);
      if (f->bits == B_INT16X4) { /* OURS: the reference declares decode_rgba16ui (shaders/stage.rs:158) but stage.frag never defines it */
        float c[4], e[4];
        for (int i = 0; i < 4; i++) { uint16_t v; memcpy(&v, p + 2 * i, 2); c[i] = (float)v / 65535.0f; }
        parts_norm(c, f->parts, e);
        transfer_apply(f->transfer, e, 0);
        for (int i = 0; i < 4; i++) o[i] = f16r(e[i]);
        return;
      }
      if (nb < 1 || nb > 4 || nb == 3) { memcpy(o, FAIL_DEC, 16); return; }
      memcpy(&n, p, nb); /* little endian sub-word of the R32Uint staging texel */
      float c[4], e[4];
      demux(n, f->bits, c);
      parts_norm(c, f->parts, e);
      transfer_apply(f->transfer, e, 0);
      for (int i = 0; i < 4; i++) o[i] = f16r(e[i]); /* Rgba16Float working texture */
      return;
    }
  }
}
static void encode_texel(const zo_fmt* f, const float* t, uint8_t* p) {
  switch (f->storage) {
    case ST_SRGB8: case ST_UNORM8: {
      uint32_t c[4];
      for (int i = 0; i < 3; i++) c[i] = f->storage == ST_SRGB8 ? srgb8_encode(t[i]) : unorm8_rne(t[i]);
      c[3] = unorm8_rne(t[3]);
      if (f->parts == P_BGRA) { p[0] = c[2]; p[1] = c[1]; p[2] = c[0]; } else { p[0] = c[0]; p[1] = c[1]; p[2] = c[2]; }
      p[3] = c[3];
      return;
    }
    case ST_FLOAT: {
      float e[4], c[4];
      memcpy(e, t, 16);
      transfer_apply(f->transfer, e, 1);
      parts_denorm(e, f->parts, c);
      if (f->bits == B_FLOAT16X4) { for (int i = 0; i < 4; i++) { uint16_t h = float_to_half_bits(c[i]); memcpy(p + 2 * i, &h, 2); } }
      else memcpy(p, c, 16);
      return;
    }
    default: {
      int nb = bits_bytes(f->bits);
      if (f->bits == B_INT16X4) { /* OURS, the mirror image of the decode above: f16 attachment, transfer, clamp, truncation */
        float e[4], c[4];
        for (int i = 0; i < 4; i++) e[i] = f16r(t[i]);
        transfer_apply(f->transfer, e, 1);
        parts_denorm(e, f->parts, c);
        for (int i = 0; i < 4; i++) { uint16_t v = (uint16_t)(uint32_t)(clamp01(c[i]) * 65535.0f); memcpy(p + 2 * i, &v, 2); }
        return;
      }
      if (nb < 1 || nb > 4 || nb == 3) return;
      float e[4], c[4];
      for (int i = 0; i < 4; i++) e[i] = f16r(t[i]); /* the draw wrote an Rgba16Float attachment */
      transfer_apply(f->transfer, e, 1);
      parts_denorm(e, f->parts, c);
      for (int i = 0; i < 4; i++) c[i] = clamp01(c[i]);
      uint32_t n = mux(c, f->bits);
      memcpy(p, &n, nb);
      return;
    }
  }
}

ZO_API void zo_decode(const zo_fmt* f, const uint8_t* src, size_t pitch, int w, int h, float* tex) {
  init_tables();
  int nb = bits_bytes(f->bits);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) decode_texel(f, src + (size_t)y * pitch + (size_t)x * nb, tex + ((size_t)y * w + x) * 4);
}
ZO_API void zo_encode(const zo_fmt* f, const float* tex, int w, int h, uint8_t* dst, size_t pitch) {
  init_tables();
  int nb = bits_bytes(f->bits);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) encode_texel(f, tex + ((size_t)y * w + x) * 4, dst + (size_t)y * pitch + (size_t)x * nb);
}

/* ------------------------------------------------------------------ */
/* Colour operators                                                     */
/* ------------------------------------------------------------------ */
/* row-major M * v with a fixed evaluation order (shared with the kernels) */
static inline void mat3_mul(const float* M, const float* v, float* o) {
  float x = v[0], y = v[1], z = v[2];
  o[0] = fmaf(M[2], z, fmaf(M[1], y, M[0] * x));
  o[1] = fmaf(M[5], z, fmaf(M[4], y, M[3] * x));
  o[2] = fmaf(M[8], z, fmaf(M[7], y, M[6] * x));
}
/* linear.frag:12-17 */
ZO_API void zo_linear(const float* M, const float* src, float* dst, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    float o[3];
    mat3_mul(M, src + 4 * i, o);
    dst[4 * i] = o[0]; dst[4 * i + 1] = o[1]; dst[4 * i + 2] = o[2]; dst[4 * i + 3] = src[4 * i + 3];
  }
}

/* Oklab constants (oklab.frag:14-24 holds them column-major; these are the row-major forms).
 * The shader evaluates inverse(M1), inverse(M2) in f32; we fix them as the f32 roundings of
 * the double-precision inverses. */
static const float OK_M1[9] = {0.8189330101f, 0.3618667424f, -0.1288597137f, 0.0329845436f, 0.9293118715f,
                               0.0361456387f, 0.0482003018f, 0.2643662691f, 0.6338517070f};
static const float OK_M2[9] = {0.2104542553f, 0.7936177850f, -0.0040720468f, 1.9779984951f, -2.4285922050f,
                               0.4505937099f, 0.0259040371f, 0.7827717662f, -0.8086757660f};
static float OK_M1I[9], OK_M2I[9];
/* SrLab2 (srlab2.frag:14-24), row-major */
static const double CAT02_D[9] = {0.7328, 0.4296, -0.1624, -0.7036, 1.6975, 0.0061, 0.0030, 0.0136, 0.9834};
static const double HPE_D[9] = {0.38971, 0.68898, -0.07868, -0.22981, 1.18340, 0.04641, 0.0, 0.0, 1.0};
static float SR_CAT[9], SR_CATI[9], SR_HPE[9], SR_HPEI[9], SR_HPE_CATI[9], SR_CAT_HPEI[9];
static int mats_ready = 0;

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
static void init_mats(void) {
  if (mats_ready) return;
  double m[9], inv[9], t[9];
  for (int i = 0; i < 9; i++) m[i] = (double)OK_M1[i];
  inv3_d(m, inv); for (int i = 0; i < 9; i++) OK_M1I[i] = (float)inv[i];
  for (int i = 0; i < 9; i++) m[i] = (double)OK_M2[i];
  inv3_d(m, inv); for (int i = 0; i < 9; i++) OK_M2I[i] = (float)inv[i];
  double cati[9], hpei[9];
  inv3_d(CAT02_D, cati); inv3_d(HPE_D, hpei);
  for (int i = 0; i < 9; i++) { SR_CAT[i] = (float)CAT02_D[i]; SR_CATI[i] = (float)cati[i]; SR_HPE[i] = (float)HPE_D[i]; SR_HPEI[i] = (float)hpei[i]; }
  mul3_d(HPE_D, cati, t); for (int i = 0; i < 9; i++) SR_HPE_CATI[i] = (float)t[i];
  mul3_d(CAT02_D, hpei, t); for (int i = 0; i < 9; i++) SR_CAT_HPEI[i] = (float)t[i];
  mats_ready = 1;
}
ZO_API void zo_constants(float* out /* 8 * 9 floats */) {
  init_mats();
  const float* src[8] = {OK_M1, OK_M2, OK_M1I, OK_M2I, SR_CAT, SR_CATI, SR_HPE_CATI, SR_CAT_HPEI};
  for (int k = 0; k < 8; k++) memcpy(out + 9 * k, src[k], 36);
}

static inline float scbrt(float v) { return v == 0.0f ? 0.0f : copysignf(cbrtf(fabsf(v)), v); }

/* oklab.frag:34-47 */
ZO_API void zo_oklab_encode(const float* T, const float* src, float* dst, size_t n) {
  init_mats();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    float xyz[3], lms[3], l3[3], lab[3];
    mat3_mul(T, src + 4 * i, xyz);
    mat3_mul(OK_M1, xyz, lms);
    for (int k = 0; k < 3; k++) l3[k] = scbrt(lms[k]);
    mat3_mul(OK_M2, l3, lab);
    dst[4 * i] = lab[0]; dst[4 * i + 1] = lab[1]; dst[4 * i + 2] = lab[2]; dst[4 * i + 3] = src[4 * i + 3];
  }
}
/* oklab.frag:50-64 */
ZO_API void zo_oklab_decode(const float* T, const float* src, float* dst, size_t n) {
  init_mats();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    float l3[3], lms[3], xyz[3], rgb[3];
    mat3_mul(OK_M2I, src + 4 * i, l3);
    for (int k = 0; k < 3; k++) lms[k] = l3[k] * l3[k] * l3[k];
    mat3_mul(OK_M1I, lms, xyz);
    mat3_mul(T, xyz, rgb);
    dst[4 * i] = clamp01(rgb[0]); dst[4 * i + 1] = clamp01(rgb[1]); dst[4 * i + 2] = clamp01(rgb[2]);
    dst[4 * i + 3] = src[4 * i + 3];
  }
}

/* srlab2.frag:87-120 */
static inline float sr_nl(float v) { return fabsf(v) < 216.0f / 24389.0f ? v * 24389.0f / 2700.0f : 1.16f * powf(v, 1.0f / 3.0f) - 0.16f; }
static inline float sr_nl_inv(float v) {
  if (fabsf(v) < 0.08f) return v * 2700.0f / 24389.0f;
  float vp = (v + 0.16f) / 1.16f;
  return vp * vp * vp;
}
/* srlab2.frag:36-57 (the whitepoint is NOT used by the encoder: wp_rgb = 1) */
ZO_API void zo_srlab2_encode(const float* T, const float* src, float* dst, size_t n) {
  init_mats();
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    float xyz[3], rgbw[3], lms[3], nl[3], e[3];
    mat3_mul(T, src + 4 * i, xyz);
    mat3_mul(SR_CAT, xyz, rgbw);
    mat3_mul(SR_HPE_CATI, rgbw, lms);
    for (int k = 0; k < 3; k++) nl[k] = sr_nl(lms[k]);
    mat3_mul(SR_HPEI, nl, e);
    dst[4 * i] = e[1];
    dst[4 * i + 1] = (e[0] - e[1]) * 5.0f / 1.16f;
    dst[4 * i + 2] = (e[2] - e[1]) * 2.0f / 1.16f;
    dst[4 * i + 3] = src[4 * i + 3];
  }
}
/* srlab2.frag:60-85 */
ZO_API void zo_srlab2_decode(const float* T, const float* wp_xyz, const float* src, float* dst, size_t n) {
  init_mats();
  float wp_rgb[3];
  mat3_mul(SR_CAT, wp_xyz, wp_rgb);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const float* Lab = src + 4 * i;
    float e[3] = {Lab[1] * 1.16f / 5.0f + Lab[0], Lab[0], Lab[2] * 1.16f / 2.0f + Lab[0]};
    float t[3], lms[3], rgbw[3], xyz[3], rgb[3];
    mat3_mul(SR_HPE, e, t);
    for (int k = 0; k < 3; k++) lms[k] = sr_nl_inv(t[k]);
    mat3_mul(SR_CAT_HPEI, lms, rgbw);
    for (int k = 0; k < 3; k++) rgbw[k] = rgbw[k] * wp_rgb[k];
    mat3_mul(SR_CATI, rgbw, xyz);
    mat3_mul(T, xyz, rgb);
    dst[4 * i] = clamp01(rgb[0]); dst[4 * i + 1] = clamp01(rgb[1]); dst[4 * i + 2] = clamp01(rgb[2]);
    dst[4 * i + 3] = Lab[3];
  }
}

/* ------------------------------------------------------------------ */
/* Painting.  box.vert + copy.frag with the Nearest / ClampToEdge       */
/* sampler (encoder.rs:1533-1537).                                      */
/* ------------------------------------------------------------------ */
/* PaintToSelection with an axis aligned target (crop / inscribe; QuadTarget::Rect):
 * source `sel` (x,y,w,h in source texels) is stretched over `tgt` (x,y,w,h in destination
 * pixels).  A destination pixel is written iff its centre lies inside the target; the source
 * texel is floor(((2*(i-tx)+1) * sw) / (2*tw)) -- the exact rational value of the interpolated
 * coordinate at the pixel centre, so the index is reproducible bit for bit. */
ZO_API void zo_paint_rect(const float* src, int sw_tex, int sh_tex, const int* sel, float* dst, int dw, int dh,
                          const int* tgt) {
  int sx = sel[0], sy = sel[1], sw = sel[2], sh = sel[3];
  int tx = tgt[0], ty = tgt[1], tw = tgt[2], th = tgt[3];
#pragma omp parallel for schedule(static)
  for (int j = ty; j < ty + th; j++) {
    if (j < 0 || j >= dh) continue;
    int64_t v = sy + ((int64_t)(2 * (j - ty) + 1) * sh) / (2 * (int64_t)th);
    if (v < 0) v = 0; if (v > sh_tex - 1) v = sh_tex - 1;
    for (int i = tx; i < tx + tw; i++) {
      if (i < 0 || i >= dw) continue;
      int64_t u = sx + ((int64_t)(2 * (i - tx) + 1) * sw) / (2 * (int64_t)tw);
      if (u < 0) u = 0; if (u > sw_tex - 1) u = sw_tex - 1;
      memcpy(dst + ((size_t)j * dw + i) * 4, src + ((size_t)v * sw_tex + u) * 4, 16);
    }
  }
}

/* PaintToSelection onto QuadTarget::Absolute (affine, command.rs:2642-2678): `inv` is the
 * row-major inverse of the affine matrix (destination pixel -> source pixel coordinates).
 * p = inv * (i+0.5, j+0.5, 1); written iff 0 <= p.x < sw and 0 <= p.y < sh.
 * sampling 0 = nearest (reference), 1 = bilinear (OURS: texel centres at k+0.5, clamp to edge). */
static inline void affine_point(const float* inv, float cx, float cy, float* px, float* py) {
  *px = fmaf(inv[1], cy, fmaf(inv[0], cx, inv[2]));
  *py = fmaf(inv[4], cy, fmaf(inv[3], cx, inv[5]));
}
/* `src` holds rows [yoff, ...) of an image of sw x sh texels (yoff = 0: the whole image) */
static inline void bilinear_tap_win(const float* src, int sw, int sh, int yoff, float px, float py, float* o) {
  float fx = px - 0.5f, fy = py - 0.5f;
  float x0f = floorf(fx), y0f = floorf(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  if (x0 < 0) x0 = 0; if (x1 < 0) x1 = 0; if (x0 > sw - 1) x0 = sw - 1; if (x1 > sw - 1) x1 = sw - 1;
  if (y0 < 0) y0 = 0; if (y1 < 0) y1 = 0; if (y0 > sh - 1) y0 = sh - 1; if (y1 > sh - 1) y1 = sh - 1;
  const float* p00 = src + ((size_t)(y0 - yoff) * sw + x0) * 4; const float* p10 = src + ((size_t)(y0 - yoff) * sw + x1) * 4;
  const float* p01 = src + ((size_t)(y1 - yoff) * sw + x0) * 4; const float* p11 = src + ((size_t)(y1 - yoff) * sw + x1) * 4;
  for (int k = 0; k < 4; k++) {
    float top = fmaf(ax, p10[k] - p00[k], p00[k]);
    float bot = fmaf(ax, p11[k] - p01[k], p01[k]);
    o[k] = fmaf(ay, bot - top, top);
  }
}
static inline void bilinear_tap(const float* src, int sw, int sh, float px, float py, float* o) {
  bilinear_tap_win(src, sw, sh, 0, px, py, o);
}
/* Windowed form (row-band sharding): `dst` holds rows [dst_y0, dst_y0 + dh) of the full
 * destination and `src` rows [src_y0, ...) of the full sw x sh source; coordinates are the FULL image's,
 * so a banded run yields exactly the bytes of the whole-image run. */
ZO_API void zo_paint_affine_window(const float* src, int sw, int sh, int src_y0, const float* inv, int sampling, float* dst,
                                   int dw, int dh, int dst_y0) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < dh; j++)
    for (int i = 0; i < dw; i++) {
      float px, py;
      affine_point(inv, (float)i + 0.5f, (float)(j + dst_y0) + 0.5f, &px, &py);
      if (!(px >= 0.0f && px < (float)sw && py >= 0.0f && py < (float)sh)) continue;
      float* o = dst + ((size_t)j * dw + i) * 4;
      if (sampling == 0) {
        int u = (int)floorf(px), v = (int)floorf(py);
        memcpy(o, src + ((size_t)(v - src_y0) * sw + u) * 4, 16);
      } else {
        bilinear_tap_win(src, sw, sh, src_y0, px, py, o);
      }
    }
}
ZO_API void zo_paint_affine(const float* src, int sw, int sh, const float* inv, int sampling, float* dst, int dw, int dh) {
  zo_paint_affine_window(src, sw, sh, 0, inv, sampling, dst, dw, dh, 0);
}

/* bilinear.frag:14-20; mix(a,b,t) = a*(1-t) + b*t; uv = pixel centre / size */
ZO_API void zo_gen_bilinear(const float* p /* u_min,u_max,v_min,v_max,uv_min,uv_max */, float* dst, int w, int h) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      float u = ((float)i + 0.5f) / (float)w, v = ((float)j + 0.5f) / (float)h, uv = u * v;
      for (int k = 0; k < 4; k++) {
        float a = p[k] * (1.0f - u) + p[4 + k] * u;
        float b = p[8 + k] * (1.0f - v) + p[12 + k] * v;
        float c = p[16 + k] * (1.0f - uv) + p[20 + k] * uv;
        dst[((size_t)j * w + i) * 4 + k] = a + b + c;
      }
    }
}

/* distribution_normal2d.frag:43-55: p = expectation[2], covariance_inverse (row major [a b; c d]), pseudo determinant */
ZO_API void zo_gen_normal2d(const float* p, float* dst, int w, int h) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      float u = ((float)i + 0.5f) / (float)w, v = ((float)j + 0.5f) / (float)h;
      float px = 2.0f * (u - 0.5f) - p[0], py = 2.0f * (v - 0.5f) - p[1];
      float tx = p[2] * px + p[3] * py, ty = p[4] * px + p[5] * py;
      float exponent = 0.5f * (px * tx + py * ty);
      float value = expf(-exponent) / sqrtf(p[6]);
      float* o = dst + ((size_t)j * w + i) * 4;
      o[0] = o[1] = o[2] = value; o[3] = 1.0f;
    }
}

/* fractal_noise.frag: pcg4d hash (jcgt.org/published/0009/03/02) of the cell corners, smoothstep interpolation,
 * `octaves` iterations with the point rotated by 0.5 rad and doubled in between */
static void zo_pcg4d(uint32_t v[4]) {
  for (int k = 0; k < 4; k++) v[k] = v[k] * 1664525u + 1013904223u;
  v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
  for (int k = 0; k < 4; k++) v[k] ^= v[k] >> 16;
  v[0] += v[1] * v[3]; v[1] += v[2] * v[0]; v[2] += v[0] * v[1]; v[3] += v[1] * v[2];
}
static void zo_hash2(uint32_t sx, uint32_t sy, float out[4]) {
  uint32_t v[4] = {sx, sy, 0u, 0u};
  zo_pcg4d(v);
  for (int k = 0; k < 4; k++) out[k] = (float)v[k] / 4294967296.0f; /* float(0xFFFFFFFFu) == 2^32 */
}
ZO_API void zo_gen_fractal_noise(const float* p /* scale.x, scale.y, amplitude, damping, octaves */, float* dst, int w, int h) {
  const int octaves = (int)p[4];
  const float c2 = 2.0f * 0.87758255f, s2 = 2.0f * 0.47942555f; /* 2 cos(0.5), 2 sin(0.5) */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      float x = ((float)i + 0.5f) / (float)w, y = ((float)j + 0.5f) / (float)h, z = 1.0f;
      float amp = p[2], acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int o = 0; o < octaves; o++) {
        float ptx = x * p[0], pty = y * p[1];
        float flx = floorf(ptx), fly = floorf(pty);
        float fx = ptx - flx, fy = pty - fly;
        uint32_t sx = (uint32_t)(int32_t)flx, sy = (uint32_t)(int32_t)fly;
        float a[4], b[4], c[4], d[4];
        zo_hash2(sx, sy, a); zo_hash2(sx + 1u, sy, b); zo_hash2(sx, sy + 1u, c); zo_hash2(sx + 1u, sy + 1u, d);
        float ux = fx * fx * (3.0f - 2.0f * fx), uy = fy * fy * (3.0f - 2.0f * fy);
        for (int k = 0; k < 4; k++) {
          float n = (a[k] * (1.0f - ux) + b[k] * ux) + (c[k] - a[k]) * uy * (1.0f - ux) + (d[k] - b[k]) * ux * uy;
          acc[k] = acc[k] + amp * n;
        }
        float nx = c2 * x - s2 * y, ny = s2 * x + c2 * y, nz = 2.0f * x + 2.0f * y + 2.0f * z;
        x = nx; y = ny; z = nz;
        amp = amp * p[3];
      }
      for (int k = 0; k < 4; k++) dst[((size_t)j * w + i) * 4 + k] = acc[k];
    }
}

/* palette.frag:21-32: coordinates come from the index image `rhs` (same size as dst) */
ZO_API void zo_palette(const float* lhs, int lw, int lh, const float* rhs, int w, int h, const float* xc, const float* yc,
                       float* dst) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      const float* b = rhs + ((size_t)j * w + i) * 4;
      float pu = xc[0] * b[0] + xc[1] * b[1] + xc[2] * b[2] + xc[3] * b[3] + 0.5f / (float)w;
      float pv = yc[0] * b[0] + yc[1] * b[1] + yc[2] * b[2] + yc[3] * b[3] + 0.5f / (float)h;
      int x = (int)floorf(pu * (float)lw), y = (int)floorf(pv * (float)lh);
      if (x < 0) x = 0; if (x > lw - 1) x = lw - 1;
      if (y < 0) y = 0; if (y > lh - 1) y = lh - 1;
      memcpy(dst + ((size_t)j * w + i) * 4, lhs + ((size_t)y * lw + x) * 4, 16);
    }
}

/* OURS: exact nearest / bilinear resize (the `ideal` resize; half-pixel centres, clamp to edge) */
ZO_API void zo_resize(const float* src, int sw, int sh, float* dst, int dw, int dh, int sampling) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < dh; j++)
    for (int i = 0; i < dw; i++) {
      float* o = dst + ((size_t)j * dw + i) * 4;
      if (sampling == 0) {
        int64_t u = ((int64_t)(2 * i + 1) * sw) / (2 * (int64_t)dw), v = ((int64_t)(2 * j + 1) * sh) / (2 * (int64_t)dh);
        memcpy(o, src + ((size_t)v * sw + u) * 4, 16);
      } else {
        float px = ((float)i + 0.5f) * ((float)sw / (float)dw), py = ((float)j + 0.5f) * ((float)sh / (float)dh);
        bilinear_tap(src, sw, sh, px, py, o);
      }
    }
}

/* inject.frag:20-25: mix(bg, vec4(dot(fg, color)), select) */
ZO_API void zo_inject(const float* bg, const float* fg, const float* mixv, const float* color, float* dst, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) {
    const float* f = fg + 4 * i;
    float d = f[0] * color[0] + f[1] * color[1] + f[2] * color[2] + f[3] * color[3];
    for (int k = 0; k < 4; k++) dst[4 * i + k] = bg[4 * i + k] * (1.0f - mixv[k]) + d * mixv[k];
  }
}

/* box3.frag:16-52: rgb = sum_{dy,dx} M[dy+1][dx+1] * src(x+dx, y+dy) (clamp to edge), alpha = 1 */
ZO_API void zo_box3(const float* M, const float* src, float* dst, int w, int h) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      float acc[3] = {0, 0, 0};
      for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
          int x = i + dx, y = j + dy;
          if (x < 0) x = 0; if (x > w - 1) x = w - 1; if (y < 0) y = 0; if (y > h - 1) y = h - 1;
          const float* s = src + ((size_t)y * w + x) * 4;
          float wgt = M[3 * (dy + 1) + (dx + 1)];
          for (int k = 0; k < 3; k++) acc[k] = fmaf(wgt, s[k], acc[k]);
        }
      float* o = dst + ((size_t)j * w + i) * 4;
      o[0] = acc[0]; o[1] = acc[1]; o[2] = acc[2]; o[3] = 1.0f;
    }
}

ZO_API void zo_fill(const float* color, float* dst, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) memcpy(dst + 4 * i, color, 16);
}

/* OURS: Porter-Duff on straight alpha in linear light.  mode: 0 clear 1 src 2 dst
 * 3 src-over 4 dst-over 5 src-in 6 dst-in 7 src-out 8 dst-out 9 src-atop 10 dst-atop 11 xor.
 * `dst` holds `below` on entry; pixels inside tgt (x,y,w,h) are blended with `src` (same size as
 * the target, no scaling). */
static inline void pd_factors(int mode, float as, float ad, float* fa, float* fb) {
  switch (mode) {
    case 0: *fa = 0; *fb = 0; return;
    case 1: *fa = 1; *fb = 0; return;
    case 2: *fa = 0; *fb = 1; return;
    case 3: *fa = 1; *fb = 1.0f - as; return;
    case 4: *fa = 1.0f - ad; *fb = 1; return;
    case 5: *fa = ad; *fb = 0; return;
    case 6: *fa = 0; *fb = as; return;
    case 7: *fa = 1.0f - ad; *fb = 0; return;
    case 8: *fa = 0; *fb = 1.0f - as; return;
    case 9: *fa = ad; *fb = 1.0f - as; return;
    case 10: *fa = 1.0f - ad; *fb = as; return;
    default: *fa = 1.0f - ad; *fb = 1.0f - as; return;
  }
}
static inline void pd_blend(int mode, const float* s, const float* d, float* o) {
  float as = s[3], ad = d[3], fa, fb;
  pd_factors(mode, as, ad, &fa, &fb);
  float wa = as * fa, wb = ad * fb;
  float ao = wa + wb;
  float rcp = ao > 0.0f ? 1.0f / ao : 0.0f; /* one correctly rounded reciprocal per pixel */
  for (int k = 0; k < 3; k++) {
    float pm = fmaf(wb, d[k], wa * s[k]);
    o[k] = pm * rcp;
  }
  o[3] = ao;
}
ZO_API void zo_blend(const float* src, int sw, int sh, float* dst, int dw, int dh, int tx, int ty, int mode) {
#pragma omp parallel for schedule(static)
  for (int j = 0; j < sh; j++) {
    int y = ty + j;
    if (y < 0 || y >= dh) continue;
    for (int i = 0; i < sw; i++) {
      int x = tx + i;
      if (x < 0 || x >= dw) continue;
      float* d = dst + ((size_t)y * dw + x) * 4;
      float o[4];
      pd_blend(mode, src + ((size_t)j * sw + i) * 4, d, o);
      memcpy(d, o, 16);
    }
  }
}

/* OURS: planar YUV 4:2:0 (I420: Y plane, then U, then V; or NV12: Y then interleaved UV) ->
 * texture of non-linear R'G'B' passed through the EOTF `transfer`.  kr,kb = luma coefficients;
 * range 0 = limited (16..235 / 16..240), 1 = full.  Chroma is sited at the centre of each 2x2
 * luma block and upsampled nearest (chroma_filter 0) or bilinearly (1). */
typedef struct zo_yuv {
  float kr, kb;
  uint32_t full_range, nv12, chroma_filter, transfer;
} zo_yuv;
static inline void yuv_to_rgb(const zo_yuv* p, float Y, float U, float V, float* o) {
  /* range scaling by f32 reciprocals, then the standard matrix form; the four coefficients are
   * evaluated in double and rounded once (the kernels get the same four floats from the host) */
  const float yoff = p->full_range ? 0.0f : 16.0f;
  const float ysc = p->full_range ? 1.0f / 255.0f : 1.0f / 219.0f, csc = p->full_range ? 1.0f / 255.0f : 1.0f / 224.0f;
  const double kr = (double)p->kr, kb = (double)p->kb, kg = 1.0 - kr - kb;
  const float r_cr = (float)(2.0 * (1.0 - kr)), b_cb = (float)(2.0 * (1.0 - kb));
  const float g_cr = (float)(2.0 * kr * (1.0 - kr) / kg), g_cb = (float)(2.0 * kb * (1.0 - kb) / kg);
  float y = (Y - yoff) * ysc, cb = (U - 128.0f) * csc, cr = (V - 128.0f) * csc;
  o[0] = fmaf(r_cr, cr, y);
  o[1] = fmaf(-g_cb, cb, fmaf(-g_cr, cr, y));
  o[2] = fmaf(b_cb, cb, y);
}
static inline float chroma_sample(const uint8_t* plane, size_t pitch, int step, int cw, int ch, int x, int y, int filter) {
  if (!filter) return (float)plane[(size_t)(y >> 1) * pitch + (size_t)(x >> 1) * step];
  /* luma centre (x+.5, y+.5) in chroma-sample units is ((x+.5)/2, (y+.5)/2); chroma centres at k+.5 */
  float fx = ((float)x + 0.5f) * 0.5f - 0.5f, fy = ((float)y + 0.5f) * 0.5f - 0.5f;
  float x0f = floorf(fx), y0f = floorf(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
  if (x0 < 0) x0 = 0; if (x1 > cw - 1) x1 = cw - 1; if (x0 > cw - 1) x0 = cw - 1; if (x1 < 0) x1 = 0;
  if (y0 < 0) y0 = 0; if (y1 > ch - 1) y1 = ch - 1; if (y0 > ch - 1) y0 = ch - 1; if (y1 < 0) y1 = 0;
  float p00 = plane[(size_t)y0 * pitch + (size_t)x0 * step], p10 = plane[(size_t)y0 * pitch + (size_t)x1 * step];
  float p01 = plane[(size_t)y1 * pitch + (size_t)x0 * step], p11 = plane[(size_t)y1 * pitch + (size_t)x1 * step];
  float top = fmaf(ax, p10 - p00, p00), bot = fmaf(ax, p11 - p01, p01);
  return fmaf(ay, bot - top, top);
}
ZO_API void zo_decode_yuv420(const zo_yuv* p, const uint8_t* yp, size_t ypitch, const uint8_t* up, const uint8_t* vp,
                             size_t cpitch, int w, int h, float* tex) {
  init_tables();
  int cw = (w + 1) / 2, ch = (h + 1) / 2, step = p->nv12 ? 2 : 1;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      float Y = (float)yp[(size_t)j * ypitch + i];
      float U = chroma_sample(up, cpitch, step, cw, ch, i, j, p->chroma_filter);
      float V = chroma_sample(vp, cpitch, step, cw, ch, i, j, p->chroma_filter);
      float* o = tex + ((size_t)j * w + i) * 4;
      yuv_to_rgb(p, Y, U, V, o);
      for (int k = 0; k < 3; k++) o[k] = eo_scalar(p->transfer, o[k]);
      o[3] = 1.0f;
    }
}
/* OURS: texture (linear) -> OETF -> Y'CbCr -> 4:2:0 planes; chroma = mean of the 2x2 block, RNE */
ZO_API void zo_encode_yuv420(const zo_yuv* p, const float* tex, int w, int h, uint8_t* yp, size_t ypitch, uint8_t* up,
                             uint8_t* vp, size_t cpitch) {
  int cw = (w + 1) / 2, ch = (h + 1) / 2, step = p->nv12 ? 2 : 1;
  float kg = 1.0f - p->kr - p->kb;
  /* chroma scaling by ONE rounded reciprocal each (like the decode side's host-prepared constants) */
  const float rcb = 1.0f / (2.0f * (1.0f - p->kb)), rcr = 1.0f / (2.0f * (1.0f - p->kr));
#pragma omp parallel for schedule(static)
  for (int cj = 0; cj < ch; cj++)
    for (int ci = 0; ci < cw; ci++) {
      float cbs = 0.0f, crs = 0.0f; int cnt = 0;
      for (int dy = 0; dy < 2; dy++)
        for (int dx = 0; dx < 2; dx++) {
          int x = 2 * ci + dx, y = 2 * cj + dy;
          if (x >= w || y >= h) continue;
          const float* t = tex + ((size_t)y * w + x) * 4;
          float r = oe_scalar(p->transfer, t[0]), g = oe_scalar(p->transfer, t[1]), b = oe_scalar(p->transfer, t[2]);
          float yy = fmaf(p->kb, b, fmaf(kg, g, p->kr * r));
          float cb = (b - yy) * rcb, cr = (r - yy) * rcr;
          float Yq = p->full_range ? yy * 255.0f : fmaf(yy, 219.0f, 16.0f);
          yp[(size_t)y * ypitch + x] = (uint8_t)rintf(fminf(fmaxf(Yq, 0.0f), 255.0f));
          cbs += cb; crs += cr; cnt++;
        }
      float cbm = cbs / (float)cnt, crm = crs / (float)cnt;
      float Uq = p->full_range ? fmaf(cbm, 255.0f, 128.0f) : fmaf(cbm, 224.0f, 128.0f);
      float Vq = p->full_range ? fmaf(crm, 255.0f, 128.0f) : fmaf(crm, 224.0f, 128.0f);
      up[(size_t)cj * cpitch + (size_t)ci * step] = (uint8_t)rintf(fminf(fmaxf(Uq, 0.0f), 255.0f));
      vp[(size_t)cj * cpitch + (size_t)ci * step] = (uint8_t)rintf(fminf(fmaxf(Vq, 0.0f), 255.0f));
    }
}

ZO_API void zo_round_f16(float* tex, size_t nfloats) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < nfloats; i++) tex[i] = f16r(tex[i]);
}
