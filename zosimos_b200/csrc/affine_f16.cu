// affine_f16.cu -- the resampling kernel of BASELINE config 3: `affine(below, M, above)` with nearest or
// bilinear sampling where all three images are plain linear RGBA16F (command.rs:1636-1673; the
// quad rasterisation it replaces is program.rs:1897-1935 + box.vert + copy.frag, bilinear is ours).
// Same results, bit for bit, as the general gather kernel (gather.cu) and the oracle
// (zo_paint_affine_window): identical coordinate arithmetic (fmaf order of map_point), identical
// tap clamping and lerp order.  What is different is the instruction count per pixel:
//
//   * a CTA produces 32x32 destination tiles (persistent, grid-stride, tiles walked in compact 2-D bands);
//     a dedicated PRODUCER warp works out each tile's geometry (4 corner mappings, box origin, alignment),
//     publishes it in shared memory and issues ONE TMA bulk tensor copy for the tile's source bounding box,
//     up to 3 tiles ahead; full / empty mbarriers per stage, no CTA-wide barrier in the loop;
//   * the x half of the inverse mapping is hoisted out of the 4-row loop of a thread, coverage of a
//     covered pixel bounds its taps so each clamp is one instruction, taps are LDS.64 at 32-bit
//     shared addresses, rows of a thread are fully unrolled so their 16 loads are in flight together.
#include <cuda_fp16.h>

#include "colorops.cuh"
#include "tma.cuh"
#include "zos_internal.h"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_affine)

namespace {
constexpr int TILE = 32;
constexpr int THREADS = 256;
constexpr int ROWS = TILE / (THREADS / 32);  // rows of a tile per thread

struct AffParams {
  const uint8_t* above;
  const uint8_t* below;
  uint8_t* dst;
  uint64_t above_pitch, above_bstride, below_pitch, below_bstride, dst_pitch, dst_bstride;
  int32_t aw, ah;  // size of the window of the source that `above` holds
  int32_t dw, dh;
  int32_t has_below, blend;
  float inv[6];
  int32_t dox, doy, sox, soy, sfw, sfh;
  uint32_t tiles_x, tiles_y, total_tiles;
  FastDiv div_frame, div_band;  // tiles per frame, tiles per band of 8 tile rows
  int32_t box_w, box_h;
  int32_t margin;    // texels a tap may lie outside of [floor(min), floor(max)] of the mapped corners: 1 bilinear, 0 nearest
  float sfwf, sfhf;  // (float)sfw, (float)sfh
  int* fault;  // mapped host word set when an mbarrier wait runs away
  float* counter;  // tile dispenser (zero at launch): see the producer
  int32_t below_map;  // TensorMaps::m1 describes `below`: tiles without covered pixels are staged through it too
};

struct __align__(16) Geo {
  // what the compute warps of the staged path read (the producer works all of it out once per tile)
  uint32_t base;                // shared address of texel (0, 0) of the FULL source, relative to this stage's box
  int32_t staged;               // any && fits: the box is in shared memory
  float cx0, cy0;               // centre of the tile's first pixel in full-destination coordinates
  uint64_t dst_off, below_off;  // byte offsets of the tile's first pixel
  int32_t nx, ny;               // live columns / rows of the tile
  // the rest is for the producer and the direct path
  int32_t frame, x0, y0;  // destination tile
  int32_t bx, by;         // box origin in the window `above` holds
  int32_t fits, any;
  int32_t last;  // this is the CTA's final tile (set by the producer one iteration later, before the entry is published)
  int32_t interior;  // a full 32 x 32 tile whose every bilinear tap lies inside the source without clamping (so every pixel is covered)
  int32_t pad_[3];
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}
__device__ __forceinline__ void map_point(const AffParams& P, float cx, float cy, float& px, float& py) {
  px = fmaf(P.inv[1], cy, fmaf(P.inv[0], cx, P.inv[2]));
  py = fmaf(P.inv[4], cy, fmaf(P.inv[3], cx, P.inv[5]));
}

// the bounding box of the tile's taps (the same construction as gather.cu's tile_info)
__device__ void tile_geometry(const AffParams& P, uint32_t t, Geo& g) {
  // tiles are walked in bands of 8 tile rows, column by column inside a band: the ~700 tiles in flight at any
  // time then form a compact 2-D block whose source footprints overlap in L2 instead of three full-width strips
  const uint32_t per_frame = P.tiles_x * P.tiles_y;
  const uint32_t fr = fastdiv(t, P.div_frame), tf = t - fr * per_frame;
  const uint32_t band_tiles = 8u * P.tiles_x;
  const uint32_t band = fastdiv(tf, P.div_band), in_band = tf - band * band_tiles;
  const uint32_t rows = min(8u, P.tiles_y - band * 8u);
  const uint32_t txi = rows == 8u ? in_band >> 3 : in_band / rows, tyi = band * 8u + (in_band - txi * rows);
  g.frame = (int)fr; g.x0 = (int)txi * TILE; g.y0 = (int)tyi * TILE;
  g.bx = g.by = 0; g.fits = 0; g.any = 0; g.interior = 0;
  const int x1 = min(g.x0 + TILE, P.dw) - 1, y1 = min(g.y0 + TILE, P.dh) - 1;
  float px[4], py[4];
  const float gx0 = (float)(g.x0 + P.dox) + 0.5f, gx1 = (float)(x1 + P.dox) + 0.5f;
  const float gy0 = (float)(g.y0 + P.doy) + 0.5f, gy1 = (float)(y1 + P.doy) + 0.5f;
  map_point(P, gx0, gy0, px[0], py[0]);
  map_point(P, gx1, gy0, px[1], py[1]);
  map_point(P, gx0, gy1, px[2], py[2]);
  map_point(P, gx1, gy1, px[3], py[3]);
  float minx = fminf(fminf(px[0], px[1]), fminf(px[2], px[3])), maxx = fmaxf(fmaxf(px[0], px[1]), fmaxf(px[2], px[3]));
  float miny = fminf(fminf(py[0], py[1]), fminf(py[2], py[3])), maxy = fmaxf(fmaxf(py[0], py[1]), fmaxf(py[2], py[3]));
  // taps floor(p - 0.5), + 1 need no clamp when 0.5 <= p < size - 0.5 for every pixel; the corner values bound all of them
  g.interior = x1 - g.x0 == TILE - 1 && y1 - g.y0 == TILE - 1 && minx >= 0.5f && miny >= 0.5f && maxx < (float)P.sfw - 0.5f && maxy < (float)P.sfh - 0.5f;
  minx = fmaxf(minx, 0.0f); miny = fmaxf(miny, 0.0f);
  maxx = fminf(maxx, (float)P.sfw); maxy = fminf(maxy, (float)P.sfh);
  if (minx > maxx || miny > maxy) return;
  // The mapping is monotone in x and in y separately, in floating point too (fmaf is monotone in each
  // argument), so the corner values bound every pixel of the tile exactly; bilinear taps of a pixel at p
  // are floor(p - 0.5) and that + 1, i.e. within [floor(min) - 1, floor(max) + 1].
  // (a nearest tap of a pixel at p is floor(p): no margin)
  int lox = max((int)floorf(minx) - P.margin - P.sox, 0), hix = min((int)floorf(maxx) + P.margin - P.sox, P.aw - 1);
  int loy = max((int)floorf(miny) - P.margin - P.soy, 0), hiy = min((int)floorf(maxy) + P.margin - P.soy, P.ah - 1);
  if (lox > hix || loy > hiy) return;
  lox &= ~1;  // 8-byte texels: a 16-byte aligned TMA source address
  g.any = 1;
  g.bx = lox; g.by = loy;
  g.fits = (hix - lox + 1 <= P.box_w) && (hiy - loy + 1 <= P.box_h);
}

__device__ __forceinline__ float4 half4_to_float4(const uint2& w) {
  float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 w;
  asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "r"(a));
  return w;
}

// The taps of one covered pixel.  SMEM: `base` is the shared address of texel (0, 0) of the FULL source
// (i.e. the stage address minus the box / window origin), `rowb` the bytes of a box row.  Otherwise
// `gbase` points at texel (0, 0) of the full source in global memory and `rowb` is the pitch.
template <bool BILINEAR, bool SMEM>
__device__ __forceinline__ float4 sample(const AffParams& P, float px, float py, uint32_t base, const uint8_t* gbase, uint32_t rowb) {
  if (!BILINEAR) {
    const int u = (int)floorf(px), w = (int)floorf(py);
    if (SMEM) return half4_to_float4(lds64(base + (uint32_t)w * rowb + (uint32_t)u * 8u));
    return half4_to_float4(*reinterpret_cast<const uint2*>(gbase + (int64_t)w * rowb + (int64_t)u * 8));
  }
  const float fx = px - 0.5f, fy = py - 0.5f;
  const float x0f = floorf(fx), y0f = floorf(fy);
  const float ax = fx - x0f, ay = fy - y0f;
  // covered: px in [0, sfw) so x0 in [-1, sfw-1]: clamp(x0) = max(x0, 0), clamp(x0 + 1) = min(x0 + 1, sfw - 1)
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int xa = max(x0, 0), xb = min(x0 + 1, P.sfw - 1), ya = max(y0, 0), yb = min(y0 + 1, P.sfh - 1);
  uint2 w00, w10, w01, w11;
  if (SMEM) {
    const uint32_t ra = base + (uint32_t)ya * rowb, rb = base + (uint32_t)yb * rowb;
    w00 = lds64(ra + (uint32_t)xa * 8u); w10 = lds64(ra + (uint32_t)xb * 8u);
    w01 = lds64(rb + (uint32_t)xa * 8u); w11 = lds64(rb + (uint32_t)xb * 8u);
  } else {
    const uint8_t* ra = gbase + (int64_t)ya * rowb;
    const uint8_t* rb = gbase + (int64_t)yb * rowb;
    w00 = *reinterpret_cast<const uint2*>(ra + (int64_t)xa * 8); w10 = *reinterpret_cast<const uint2*>(ra + (int64_t)xb * 8);
    w01 = *reinterpret_cast<const uint2*>(rb + (int64_t)xa * 8); w11 = *reinterpret_cast<const uint2*>(rb + (int64_t)xb * 8);
  }
  const float4 p00 = half4_to_float4(w00), p10 = half4_to_float4(w10), p01 = half4_to_float4(w01), p11 = half4_to_float4(w11);
  float4 o;
#define ZOS_LERP2(c) { float top = fmaf(ax, p10.c - p00.c, p00.c), bot = fmaf(ax, p11.c - p01.c, p01.c); o.c = fmaf(ay, bot - top, top); }
  ZOS_LERP2(x) ZOS_LERP2(y) ZOS_LERP2(z) ZOS_LERP2(w)
#undef ZOS_LERP2
  return o;
}

// predicated loads: no branch, no valid address needed when the predicate is off
__device__ __forceinline__ uint2 lds64_if(uint32_t a, bool p) {
  uint2 w = make_uint2(0u, 0u);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.shared.v2.u32 {%0, %1}, [%2];\n\t}" : "+r"(w.x), "+r"(w.y) : "r"(a), "r"((uint32_t)p));
  return w;
}
__device__ __forceinline__ uint2 ldg64_cs_if(const uint8_t* ptr, bool p) {
  uint2 w = make_uint2(0u, 0u);
  asm("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %3, 0;\n\t@q ld.global.cs.v2.u32 {%0, %1}, [%2];\n\t}" : "+r"(w.x), "+r"(w.y) : "l"(ptr), "r"((uint32_t)p));
  return w;
}

// sm_100 mixed-precision add (FHADD): (float)h - c with ONE rounding, i.e. exactly the f32 subtraction of the
// converted half (the conversion is exact) -- it saves the separate f16 -> f32 conversion of the minuend.
__device__ __forceinline__ float hsub_lo(uint32_t w, float c) {
  float d;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tsub.rn.f32.f16 %0, l, %2;\n\t}" : "=f"(d) : "r"(w), "f"(c));
  return d;
}
__device__ __forceinline__ float hsub_hi(uint32_t w, float c) {
  float d;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tsub.rn.f32.f16 %0, h, %2;\n\t}" : "=f"(d) : "r"(w), "f"(c));
  return d;
}
__device__ __forceinline__ float h2f_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h2f_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

// fma(ay, bot - top, top) of top = fma(ax, p10 - p00, p00), bot = fma(ax, p11 - p01, p01): the lerp order of the
// oracle (zo_paint_affine_window) and gather.cu, 8 instructions per channel
#define ZOS_BILERP(EXT, SUB, W00, W10, W01, W11, AX, AY, OUT) \
  {                                                           \
    const float a_ = EXT(W00), b_ = EXT(W01);                 \
    const float top_ = fmaf(AX, SUB(W10, a_), a_);            \
    const float bot_ = fmaf(AX, SUB(W11, b_), b_);            \
    OUT = fmaf(AY, bot_ - top_, top_);                        \
  }

struct Lane {  // per-thread constants of a compute thread
  uint32_t lx, ly;
  float lxf, lyf;
  uint64_t dst_thr, below_thr;  // ly * pitch + lx * 8
};

// The rows of one thread from the staged box, branch-free: GROUP rows are set up together (coordinates,
// coverage, tap addresses), their loads are issued back to back, then the arithmetic runs for all of them.
// ALLCOV: every pixel of the warp's rows is live and covered (the common case: interior tiles) -- plain loads,
// no `below`, no select.  Otherwise loads are predicated and uncovered pixels take `below` (or the clear colour)
// through a select.
template <bool BILINEAR, bool ALLCOV, int GROUP, bool NOCLAMP = false>
__device__ __forceinline__ void compute_rows(const AffParams& P, const Geo& g, const Lane& L, const float (&px)[ROWS], const float (&py)[ROWS],
                                             const bool (&live)[ROWS], const bool (&cov)[ROWS]) {
  const uint32_t rowb = (uint32_t)P.box_w * 8u;
  const uint32_t base = g.base;
  const int xmax = P.sfw - 1, ymax = P.sfh - 1;
  uint8_t* dp = P.dst + (g.dst_off + L.dst_thr);
  const uint8_t* bp = P.below + (g.below_off + L.below_thr);
  const uint64_t dstep = (uint64_t)(THREADS / 32) * P.dst_pitch, bstep = (uint64_t)(THREADS / 32) * P.below_pitch;
  const bool has_below = P.has_below != 0;
#pragma unroll
  for (int k0 = 0; k0 < ROWS; k0 += GROUP) {
    float ax[GROUP], ay[GROUP];
    uint2 w00[GROUP], w10[GROUP], w01[GROUP], w11[GROUP], wb[GROUP];
#pragma unroll
    for (int q = 0; q < GROUP; q++) {
      const int k = k0 + q;
      const bool c = ALLCOV || cov[k];
      if (BILINEAR) {
        const float fx = px[k] - 0.5f, fy = py[k] - 0.5f;
        // floor as ONE conversion; back to float on the integer pipe (exact: |fx| < 2^23 for covered pixels)
        const int x0 = __float2int_rd(fx), y0 = __float2int_rd(fy);
        const float x0f = (float)x0, y0f = (float)y0;
        ax[q] = fx - x0f; ay[q] = fy - y0f;
        // covered: px in [0, sfw) so x0 in [-1, sfw-1]: clamp(x0) = max(x0, 0), clamp(x0 + 1) = min(x0 + 1, sfw - 1)
        if (NOCLAMP) {
          // interior tile: the four taps are x0, x0 + 1 on rows y0, y0 + 1 -- one address, constant offsets
          const uint32_t a00 = base + (uint32_t)y0 * rowb + (uint32_t)x0 * 8u, a01 = a00 + rowb;
          w00[q] = lds64(a00); w10[q] = lds64(a00 + 8u);
          w01[q] = lds64(a01); w11[q] = lds64(a01 + 8u);
          continue;
        }
        const int xa = max(x0, 0), xb = min(x0 + 1, xmax), ya = max(y0, 0), yb = min(y0 + 1, ymax);
        const uint32_t ra = base + (uint32_t)ya * rowb, rb = base + (uint32_t)yb * rowb;
        if (ALLCOV) {
          w00[q] = lds64(ra + (uint32_t)xa * 8u); w10[q] = lds64(ra + (uint32_t)xb * 8u);
          w01[q] = lds64(rb + (uint32_t)xa * 8u); w11[q] = lds64(rb + (uint32_t)xb * 8u);
        } else {
          w00[q] = lds64_if(ra + (uint32_t)xa * 8u, c); w10[q] = lds64_if(ra + (uint32_t)xb * 8u, c);
          w01[q] = lds64_if(rb + (uint32_t)xa * 8u, c); w11[q] = lds64_if(rb + (uint32_t)xb * 8u, c);
        }
      } else {
        const int u = __float2int_rd(px[k]), w = __float2int_rd(py[k]);
        const uint32_t a = base + (uint32_t)w * rowb + (uint32_t)u * 8u;
        w00[q] = ALLCOV ? lds64(a) : lds64_if(a, c);
      }
      if (!ALLCOV) wb[q] = ldg64_cs_if(bp + (uint64_t)k * bstep, has_below && live[k] && !c);
    }
#pragma unroll
    for (int q = 0; q < GROUP; q++) {
      const int k = k0 + q;
      uint2 o;
      if (BILINEAR) {
        float4 v;
        ZOS_BILERP(h2f_lo, hsub_lo, w00[q].x, w10[q].x, w01[q].x, w11[q].x, ax[q], ay[q], v.x)
        ZOS_BILERP(h2f_hi, hsub_hi, w00[q].x, w10[q].x, w01[q].x, w11[q].x, ax[q], ay[q], v.y)
        ZOS_BILERP(h2f_lo, hsub_lo, w00[q].y, w10[q].y, w01[q].y, w11[q].y, ax[q], ay[q], v.z)
        ZOS_BILERP(h2f_hi, hsub_hi, w00[q].y, w10[q].y, w01[q].y, w11[q].y, ax[q], ay[q], v.w)
        __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
      } else {
        // the working value of a nearest tap is the texel itself, and f16 -> f32 -> f16 is the identity
        o = w00[q];
      }
      if (ALLCOV) {
        __stcs(reinterpret_cast<uint2*>(dp + (uint64_t)k * dstep), o);
      } else {
        // uncovered: `below` as it is, else the clear colour (0, 0, 1, 1)
        if (!cov[k]) o = has_below ? wb[q] : make_uint2(0u, 0x3c003c00u);
        if (live[k]) __stcs(reinterpret_cast<uint2*>(dp + (uint64_t)k * dstep), o);
      }
    }
  }
}

// The width of the staged box decides which shared-memory banks the taps of a warp meet: lane i of a row reads texel
// (x + i * inv[0], y + i * inv[3]), an LDS.64 serves a half-warp per wavefront when its 16 lanes hit 16 different bank
// pairs, and bank pair = (column + row * box_w) mod 16.  A fixed rule (round 1: box_w = 2 mod 4) is good for some rotations
// and very bad for others (30 degrees: 2.0 wavefronts per half-warp load, -30 degrees: 6.3, because the row term cancels the
// column term; ncu had 45 % of this kernel's shared wavefronts as conflict replays, and the L1 data pipe is its busiest
// unit: profiles/r02_kernel_facts.md).  So the launcher simulates the four taps of a row of 32 lanes on a lattice of
// sub-texel phases for the 8 even widths from the minimum up, and takes the cheapest (ties: the narrowest, every extra
// column is fetched by the TMA unit).
int pick_box_width(int min_w, float dxl, float dyl) {
  int best_w = min_w;
  long best = -1;
  for (int w = min_w; w < min_w + 16; w += 2) {
    long cost = 0;
    for (int ph = 0; ph < 16; ph++) {
      const float x0 = 64.0f + 0.25f * (float)(ph & 3) + 0.125f, y0 = 64.0f + 0.25f * (float)(ph >> 2) + 0.125f;
      for (int tap = 0; tap < 4; tap++)
        for (int half = 0; half < 2; half++) {
          int addr[16], worst = 0;
          for (int l = 0; l < 16; l++) {
            const int i = half * 16 + l;
            const int X = (int)floorf(x0 + (float)i * dxl - 0.5f) + (tap & 1), Y = (int)floorf(y0 + (float)i * dyl - 0.5f) + (tap >> 1);
            addr[l] = Y * w + X;
          }
          for (int b = 0; b < 16; b++) {
            int distinct = 0;
            for (int l = 0; l < 16; l++) {
              if ((addr[l] & 15) != b) continue;
              bool seen = false;
              for (int k = 0; k < l; k++) seen = seen || addr[k] == addr[l];
              distinct += !seen;
            }
            worst = distinct > worst ? distinct : worst;
          }
          cost += worst;
        }
    }
    if (best < 0 || cost < best) { best = cost; best_w = w; }
  }
  return best_w;
}

template <bool BILINEAR, int GROUP>
__device__ __forceinline__ void compute_tile_smem(const AffParams& P, const Geo& g, const Lane& L) {
  // == (float)(i + dox) + 0.5f: every term is an exact float below 2^22 (checked by the launcher)
  const float cx = g.cx0 + L.lxf;
  const float tx = fmaf(P.inv[0], cx, P.inv[2]), ty = fmaf(P.inv[3], cx, P.inv[5]);
  if (g.interior) {
    // most tiles: full, every pixel covered, no tap needs a clamp -- no coverage tests, no vote
    const float cyb = g.cy0 + L.lyf;
    float px[ROWS], py[ROWS];
    bool live[ROWS], cov[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; k++) {
      const float cy = cyb + (float)((THREADS / 32) * k);
      px[k] = fmaf(P.inv[1], cy, tx); py[k] = fmaf(P.inv[4], cy, ty);
      live[k] = cov[k] = true;
    }
    compute_rows<BILINEAR, true, GROUP, true>(P, g, L, px, py, live, cov);
    return;
  }
  // covered <=> 0 <= p < size.  No coordinate is -0 (the launcher turns -0 coefficients into +0, sums that
  // cancel give +0), so the bit patterns of non-negative floats order like unsigned integers and every negative
  // or NaN pattern is above bits(size): one unsigned compare per axis.
  const uint32_t wbits = __float_as_uint(P.sfwf), hbits = __float_as_uint(P.sfhf);
  const float cyb = g.cy0 + L.lyf;
  const bool col_live = (int)L.lx < g.nx;
  const int rows_left = g.ny - (int)L.ly;  // row k of this thread is live when (THREADS / 32) * k < rows_left
  float px[ROWS], py[ROWS];
  bool live[ROWS], cov[ROWS];
  bool all = true;
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    const float cy = cyb + (float)((THREADS / 32) * k);  // == (float)(j + doy) + 0.5f, exact
    px[k] = fmaf(P.inv[1], cy, tx); py[k] = fmaf(P.inv[4], cy, ty);
    live[k] = col_live && (THREADS / 32) * k < rows_left;
    cov[k] = live[k] && __float_as_uint(px[k]) < wbits && __float_as_uint(py[k]) < hbits;
    all = all && cov[k];
  }
  if (__all_sync(0xffffffffu, all)) compute_rows<BILINEAR, true, GROUP>(P, g, L, px, py, live, cov);
  else compute_rows<BILINEAR, false, GROUP>(P, g, L, px, py, live, cov);
}

// Tiles whose footprint does not fit the box (strong minification) or that no covered pixel touches:
// taps straight from global memory.
template <bool BILINEAR>
__device__ __forceinline__ void compute_tile_global(const AffParams& P, const Geo& g) {
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int i = g.x0 + lx;
  if (i >= P.dw) return;
  const float cx = (float)(i + P.dox) + 0.5f;
  const float tx = fmaf(P.inv[0], cx, P.inv[2]), ty = fmaf(P.inv[3], cx, P.inv[5]);
  const float sfw = (float)P.sfw, sfh = (float)P.sfh;
  const uint32_t rowb = (uint32_t)P.above_pitch;
  const uint8_t* gbase = P.above + (uint64_t)g.frame * P.above_bstride - (int64_t)P.soy * (int64_t)P.above_pitch - (int64_t)P.sox * 8;
  const uint64_t off0 = (uint64_t)g.frame * P.dst_bstride + (uint64_t)(g.y0 + ly) * P.dst_pitch + (uint64_t)i * 8u;
  const uint64_t boff0 = (uint64_t)g.frame * P.below_bstride + (uint64_t)(g.y0 + ly) * P.below_pitch + (uint64_t)i * 8u;
#pragma unroll 1
  for (int k = 0; k < ROWS; k++) {
    const int j = g.y0 + ly + (THREADS / 32) * k;
    if (j >= P.dh) break;
    const float cy = (float)(j + P.doy) + 0.5f;
    const float px = fmaf(P.inv[1], cy, tx), py = fmaf(P.inv[4], cy, ty);
    const bool covered = px >= 0.0f && px < sfw && py >= 0.0f && py < sfh;
    uint2 o = make_uint2(0u, 0x3c003c00u);  // Target::Discard clear colour (0, 0, 1, 1)
    if (covered) {
      const float4 v = sample<BILINEAR, false>(P, px, py, 0u, gbase, rowb);
      __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      o = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else if (P.has_below) {
      o = __ldcs(reinterpret_cast<const uint2*>(P.below + boff0 + (uint64_t)((THREADS / 32) * k) * P.below_pitch));
    }
    __stcs(reinterpret_cast<uint2*>(P.dst + off0 + (uint64_t)((THREADS / 32) * k) * P.dst_pitch), o);
  }
}

// A tile no covered pixel can touch: `below` (or the clear colour) as it is.  All rows of a thread are loaded
// before the first store so that a warp pays one memory round trip per tile, not one per row.
__device__ __forceinline__ void copy_tile(const AffParams& P, const Geo& g, const Lane& L) {
  const bool col_live = (int)L.lx < g.nx;
  const int rows_left = g.ny - (int)L.ly;
  uint8_t* dp = P.dst + (g.dst_off + L.dst_thr);
  const uint8_t* bp = P.below + (g.below_off + L.below_thr);
  const uint64_t dstep = (uint64_t)(THREADS / 32) * P.dst_pitch, bstep = (uint64_t)(THREADS / 32) * P.below_pitch;
  uint2 w[ROWS];
  bool live[ROWS];
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    live[k] = col_live && (THREADS / 32) * k < rows_left;
    w[k] = make_uint2(0u, 0x3c003c00u);  // Target::Discard clear colour (0, 0, 1, 1)
    if (P.has_below && live[k]) w[k] = __ldcs(reinterpret_cast<const uint2*>(bp + (uint64_t)k * bstep));
  }
#pragma unroll
  for (int k = 0; k < ROWS; k++)
    if (live[k]) __stcs(reinterpret_cast<uint2*>(dp + (uint64_t)k * dstep), w[k]);
}

// The same tile when the producer has staged `below`'s 32 x 32 texels in the stage buffer (one TMA box, like a source
// box): the round trip to DRAM is hidden by the pipeline instead of being paid by all eight warps at once.
__device__ __forceinline__ void copy_tile_staged(const AffParams& P, const Geo& g, const Lane& L) {
  const bool col_live = (int)L.lx < g.nx;
  const int rows_left = g.ny - (int)L.ly;
  uint8_t* dp = P.dst + (g.dst_off + L.dst_thr);
  const uint64_t dstep = (uint64_t)(THREADS / 32) * P.dst_pitch;
  const uint32_t a = g.base + L.ly * (TILE * 8u) + L.lx * 8u;
#pragma unroll
  for (int k = 0; k < ROWS; k++) {
    const uint2 w = lds64(a + (uint32_t)k * ((THREADS / 32) * TILE * 8u));
    if (col_live && (THREADS / 32) * k < rows_left) __stcs(reinterpret_cast<uint2*>(dp + (uint64_t)k * dstep), w);
  }
}

// Warp-specialised pipeline: warps 0..7 compute tiles, warp 8 (one lane) is the producer.  The producer works out
// the geometry of a tile AHEAD iterations before the tile's turn and stores it in a ring in shared memory, so that
// the moment the tile's stage is free the TMA load goes out (the ~250 serial instructions of the geometry are
// not between "stage released" and "load issued").  `full[s]` completes when the box has landed (or at once for
// tiles that need none).  (Measured and dropped: cp.async.bulk.prefetch.tensor of the box into L2 at geometry
// time -- 0.72 -> 0.61 of the HBM peak at 2 frames, 0.49 at 16: the prefetched lines displace the overlap between
// neighbouring boxes that L2 serves today.  Also measured and dropped: the producer warp staging the box with
// 16-byte cp.async copies instead of the TMA box load -- 0.72 -> 0.32-0.38; prefetch.global.L2 of the box's lines
// from the producer warp's idle lanes 1, 2 or 4 tiles ahead -- 0.73 -> 0.60 (nearest), 0.50 (bilinear).  Every
// extra request for a line costs L2 slice throughput, and that is the bound: SM-side traffic (2.3x the source
// texels used + `below` + stores, 1.6 GB per launch) plus the DRAM-side fills and write-backs (1.06 GB) pass
// through the slices at ~12 TB/s, their measured cap; a 0.5 degree rotation (box 1.35x) runs at the same speed.)  A consumer warp that is done with stage s arrives on
// `empty[s]` (8 arrivals free the stage).  There is no CTA-wide barrier in the loop: warps drift apart by up
// to STAGES tiles, and the ~250 serial instructions of the geometry are off the compute warps' critical path.
constexpr int STAGES = 3;
constexpr int AHEAD = 4;
constexpr int RING = 8;  // > AHEAD + STAGES: an entry is rewritten only after its tile has been consumed
constexpr int CONSUMER_WARPS = THREADS / 32;
constexpr int THREADS_ALL = THREADS + 32;

template <bool BILINEAR, int GROUP, int CTAS>
__global__ void __launch_bounds__(THREADS_ALL, CTAS) k_affine_f16(const __grid_constant__ AffParams P, const __grid_constant__ TensorMaps M) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  __shared__ Geo geo[RING];
  const uint32_t box_bytes = (uint32_t)P.box_w * P.box_h * 8u;
  const uint32_t stage_bytes = (box_bytes + 127u) & ~127u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the barriers exist before anybody arms or polls them
  const CUtensorMap* const m0 = &M.m0;  // stays in param space (see gather.cu)
  const CUtensorMap* const m1 = &M.m1;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == CONSUMER_WARPS) {
    if ((threadIdx.x & 31) != 0) return;
    // Tiles are handed out dynamically: the producer draws tile numbers from a counter in global memory instead of
    // walking blockIdx.x + k * gridDim.x.  Tiles do not cost the same (one without covered pixels is a copy), so
    // equal static shares are unequal work; numbers drawn at about the same time are neighbours in the banded
    // order, so the tiles in flight still form one compact block.  The draw for the next tile is in flight while this
    // tile's geometry is worked out, AHEAD tiles before its load.  The counter is a FLOAT (tile numbers are below
    // 2^24, every value exact): ptxas wraps an integer atom.add into its warp-aggregation sequence, whose shuffle
    // waits for the atomic's round trip on the spot.
    auto draw = [&]() { float t; asm volatile("atom.global.add.f32 %0, [%1], 0f3F800000;" : "=f"(t) : "l"(P.counter) : "memory"); return t; };
    const uint32_t rowb = (uint32_t)P.box_w * 8u;
    // (TILES_PER_DRAW tiles per draw: one draw per tile ran into the throughput of atomics on a single address,
    //  ~0.6 per ns were needed for 130 k tiles in 0.22 ms -- nearest fell from 0.75 to 0.65)
    constexpr uint32_t TILES_PER_DRAW = 4;
    uint32_t tile = (uint32_t)draw() * TILES_PER_DRAW, left = TILES_PER_DRAW;
    float next_draw = draw();
    bool done = false;
    uint32_t s = 0, round = 0;  // stage and use count of the stage for the tile being issued
    for (uint32_t it = 0;; it++) {
      if (!done) {  // geometry of this CTA's tile number `it`, AHEAD iterations ahead of its load
        Geo& g_ = geo[it % RING];
        if (tile < P.total_tiles) {
          tile_geometry(P, tile, g_);
          g_.staged = (g_.any && g_.fits) ? 1 : (!g_.any && P.below_map) ? 2 : 0;  // 1: source box staged, 2: `below` tile staged
          g_.base = smem_u32(dyn + (size_t)(it % STAGES) * stage_bytes);
          if (g_.staged == 1) g_.base -= (uint32_t)(g_.by + P.soy) * rowb + (uint32_t)(g_.bx + P.sox) * 8u;
          g_.cx0 = (float)(g_.x0 + P.dox) + 0.5f; g_.cy0 = (float)(g_.y0 + P.doy) + 0.5f;
          g_.dst_off = (uint64_t)g_.frame * P.dst_bstride + (uint64_t)g_.y0 * P.dst_pitch + (uint64_t)g_.x0 * 8u;
          g_.below_off = (uint64_t)g_.frame * P.below_bstride + (uint64_t)g_.y0 * P.below_pitch + (uint64_t)g_.x0 * 8u;
          g_.nx = min(TILE, P.dw - g_.x0); g_.ny = min(TILE, P.dh - g_.y0);
          g_.last = 0;
          tile++;
          if (--left == 0) { tile = (uint32_t)next_draw * TILES_PER_DRAW; left = TILES_PER_DRAW; next_draw = draw(); }
        } else {
          // no tile left.  The previous entry (not yet published: entries go out AHEAD iterations after their geometry)
          // becomes the CTA's last one; a CTA that drew nothing publishes one empty entry.
          if (it == 0) g_.staged = -1; else geo[(it - 1) % RING].last = 1;
          done = true;
        }
      }
      if (it < AHEAD) continue;
      const Geo& g = geo[(it - AHEAD) % RING];
      if (round > 0) {  // wait until the consumers have released the stage's previous tile
        uint32_t spins = 0;
        while (!mbar_try_wait_sleep(&empty[s], (round - 1) & 1, 20000u)) {
          if (++spins > (1u << 20)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; break; }
        }
      }
      if (g.staged == 1) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[s], box_bytes);
        tma_load_3d(dyn + (size_t)s * stage_bytes, m0, g.bx * 2, g.by, g.frame, &full[s]);
      } else if (g.staged == 2) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&full[s], TILE * TILE * 8u);
        tma_load_3d(dyn + (size_t)s * stage_bytes, m1, g.x0 * 2, g.y0, g.frame, &full[s]);
      } else {
        mbar_arrive(&full[s]);  // nothing to load: the geometry alone is the payload (release: the ring entry is visible)
      }
      if (g.staged < 0 || g.last) break;
      if (++s == STAGES) { s = 0; round++; }
    }
    return;
  }
  Lane L;
  L.lx = threadIdx.x & 31; L.ly = threadIdx.x >> 5;
  L.lxf = (float)L.lx; L.lyf = (float)L.ly;
  L.dst_thr = (uint64_t)L.ly * P.dst_pitch + (uint64_t)L.lx * 8u;
  L.below_thr = (uint64_t)L.ly * P.below_pitch + (uint64_t)L.lx * 8u;
  uint32_t s = 0, parity = 0, r = 0;
  for (;;) {
    uint32_t spins = 0;
    bool lost = false;
    while (!mbar_try_wait_sleep(&full[s], parity, 20000u)) {
      if (++spins > (1u << 20)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; lost = true; break; }
    }
    const Geo& g = geo[r];
    if (g.staged < 0 || lost) break;  // this CTA drew no tile at all (or a barrier ran away: reported through zos_sync)
    const bool last = g.last != 0;
    if (g.staged == 1) compute_tile_smem<BILINEAR, GROUP>(P, g, L);
    else if (g.staged == 2) copy_tile_staged(P, g, L);  // no covered pixel, `below` staged
    else if (!g.any) copy_tile(P, g, L);                // no covered pixel
    else compute_tile_global<BILINEAR>(P, g);       // a footprint larger than the box: taps straight from global memory
    if (last) break;                                // the producer has left: nobody waits for this stage any more
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);  // this warp is done with stage s (its shared-memory reads have completed)
    if (++s == STAGES) { s = 0; parity ^= 1u; }
    r = (r + 1) & (RING - 1);
  }
}

bool plain_f16(const DevImage& im) {
  return im.block == ZOS_BLOCK_PIXEL && im.fmt.storage == ZOS_STORAGE_FLOAT && im.fmt.bits == ZOS_BITS_FLOAT16X4 &&
         im.fmt.transfer == ZOS_TRANSFER_LINEAR && (im.fmt.parts == ZOS_PARTS_RGBA || im.fmt.parts == ZOS_PARTS_LCHA || im.fmt.parts == ZOS_PARTS_LABA) &&
         ((uintptr_t)im.p0 % 16) == 0 && (im.pitch % 16) == 0 && (im.bstride % 16) == 0;
}
}  // namespace

// Serves: ZOS_MAP_AFFINE, nearest / bilinear, all images plain linear RGBA16F, no steps, any
// Porter-Duff mode or overwrite.  *handled stays false when the launch is left to gather.cu.
zos_status launch_affine_f16(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                             const zos_compose_params& cp, uint32_t batch, bool* handled) {
  *handled = false;
  if ((ctx->flags & ZOS_CTX_NO_FAST_PATHS) || !cp.use_tma || cp.map != ZOS_MAP_AFFINE) return ZOS_OK;
  if (cp.blend != ZOS_BLEND_OVERWRITE || cp.n_src_steps || cp.n_dst_steps) return ZOS_OK;  // the reference's affine: no blending (encoder.rs:1493)
  if (cp.sampling != ZOS_SAMPLE_NEAREST && cp.sampling != ZOS_SAMPLE_BILINEAR) return ZOS_OK;
  if (!plain_f16(above) || !plain_f16(dst) || (below && !plain_f16(*below))) return ZOS_OK;
  AffParams P;
  memset(&P, 0, sizeof P);
  P.fault = ctx->fault_dev;
  P.above = above.p0; P.above_pitch = above.pitch; P.above_bstride = above.bstride; P.aw = above.w; P.ah = above.h;
  if (below) { P.below = below->p0; P.below_pitch = below->pitch; P.below_bstride = below->bstride; }
  P.has_below = below != nullptr;
  P.dst = dst.p0; P.dst_pitch = dst.pitch; P.dst_bstride = dst.bstride; P.dw = dst.w; P.dh = dst.h;
  P.blend = cp.blend;
  for (int k = 0; k < 6; k++) P.inv[k] = cp.inv[k] == 0.0f ? 0.0f : cp.inv[k];  // -0 -> +0: same results (see compute_tile_smem), no -0 coordinate
  P.dox = cp.dst_origin[0]; P.doy = cp.dst_origin[1]; P.sox = cp.src_origin[0]; P.soy = cp.src_origin[1];
  P.sfw = cp.src_full[0] > 0 ? cp.src_full[0] : above.w; P.sfh = cp.src_full[1] > 0 ? cp.src_full[1] : above.h;
  if (P.dox < 0 || P.doy < 0 || P.sox < 0 || P.soy < 0 || P.sox + above.w > P.sfw || P.soy + above.h > P.sfh) return ZOS_OK;  // gather.cu reports it
  if (above.pitch >= (1ull << 31)) return ZOS_OK;
  if ((int64_t)P.dox + dst.w >= (1 << 22) || (int64_t)P.doy + dst.h >= (1 << 22) || P.sfw >= (1 << 22) || P.sfh >= (1 << 22)) return ZOS_OK;
  P.sfwf = (float)P.sfw; P.sfhf = (float)P.sfh;
  P.tiles_x = (dst.w + TILE - 1) / TILE; P.tiles_y = (dst.h + TILE - 1) / TILE;
  const uint64_t total = (uint64_t)P.tiles_x * P.tiles_y * batch;
  if (total == 0 || total >= (1ull << 24)) return ZOS_OK;  // (tile numbers are drawn from a float counter)
  P.total_tiles = (uint32_t)total;
  P.div_frame = make_fastdiv(P.tiles_x * P.tiles_y); P.div_band = make_fastdiv(8u * P.tiles_x);
  const float ex = (TILE - 1) * (fabsf(P.inv[0]) + fabsf(P.inv[1])), ey = (TILE - 1) * (fabsf(P.inv[3]) + fabsf(P.inv[4]));
  if (!(ex < 200.0f) || !(ey < 200.0f)) return ZOS_OK;
  // taps span ceil(extent) + 2 texels + the sampling margin on both sides, + 1 for the even box origin; a row of 8 * (4k + 2) bytes puts
  // vertically adjacent taps 4 banks apart (a multiple of 128 bytes would put them on the same bank)
  P.margin = cp.sampling == ZOS_SAMPLE_BILINEAR ? 1 : 0;
  P.box_w = (int)ceilf(ex) + 3 + 2 * P.margin;
  P.box_w += P.box_w & 1;  // (the TMA box's inner extent is a multiple of 16 bytes)
  if (ctx->box_cache.w == 0 || ctx->box_cache.min_w != P.box_w || ctx->box_cache.dxl != P.inv[0] || ctx->box_cache.dyl != P.inv[3]) {
    ctx->box_cache.w = pick_box_width(P.box_w, P.inv[0], P.inv[3]);  // ~1 ms of host work: once per mapping, not per launch
    ctx->box_cache.min_w = P.box_w; ctx->box_cache.dxl = P.inv[0]; ctx->box_cache.dyl = P.inv[3];
  }
  P.box_w = ctx->box_cache.w;
  P.box_h = (int)ceilf(ey) + 2 + 2 * P.margin;
  const size_t stage = ((size_t)P.box_w * P.box_h * 8 + 127) & ~(size_t)127;
  const size_t smem = STAGES * stage;
  if (P.box_w * 2 > 256 || P.box_h > 256 || smem > 144 * 1024) return ZOS_OK;
  TensorMaps M;
  memset(&M, 0, sizeof M);
  if (!make_map(ctx, &M.m0, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, above.p0, (uint64_t)above.w * 2, above.h, above.pitch, batch, above.bstride,
                (uint32_t)P.box_w * 2, (uint32_t)P.box_h))
    return ZOS_OK;
  // `below` tiles (32 x 32 texels) go through the same stages; without the map they are read directly
  P.below_map = below && stage >= (size_t)TILE * TILE * 8 && (P.dox | P.doy) == 0 &&
                make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, below->p0, (uint64_t)below->w * 2, below->h, below->pitch, batch, below->bstride,
                         TILE * 2, TILE);
  // 3 stages of ~19 KB (30 degree rotation): 4 CTAs of 9 warps per SM, 56 registers per thread
  int per_sm = (int)((228 * 1024) / (smem + 1024 + 256));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
  const int grid = (int)(total < cap ? total : cap);
  P.counter = ctx->work_counter;
  if (cudaMemsetAsync(ctx->work_counter, 0, sizeof(float), ctx->stream) != cudaSuccess) return check_cuda(ctx, cudaGetLastError(), "affine_f16 counter reset");
#define ZOS_AFF_LAUNCH(B, G, C)                                                                          \
  do {                                                                                                    \
    ensure_dyn_smem(ctx, k_affine_f16<B, G, C>, 144 * 1024);                                                \
    k_affine_f16<B, G, C><<<grid, THREADS_ALL, smem, ctx->stream>>>(P, M);                                    \
  } while (0)
  if (cp.sampling == ZOS_SAMPLE_BILINEAR) ZOS_AFF_LAUNCH(true, 2, 4);   // GROUP = rows of a thread set up together
  else ZOS_AFF_LAUNCH(false, 2, 4);
#undef ZOS_AFF_LAUNCH
  ctx->launches++;
  *handled = true;
  return check_cuda(ctx, cudaGetLastError(), "k_affine_f16 launch");
}

}  // namespace zos

// Host-only view of the staged-box width rule (no device needed): tests compare it with an independent model of the banks.
extern "C" int32_t zos_affine_box_width(int32_t min_width, float step_x_per_lane, float step_y_per_lane) {
  if (min_width < 2 || min_width > 256) return -1;
  return zos::pick_box_width(min_width + (min_width & 1), step_x_per_lane, step_y_per_lane);
}
