// gather.cu -- destination-centric composition / resampling kernel.
//
//   dst(i,j) = pack( dst_steps( inside(i,j) ? blend(sample(above, map(i,j)), below(i,j)) : below(i,j) ) )
//
// Covers the reference's PaintToSelection draws (box.vert:43-59 + copy.frag:8-10 with the Nearest /
// ClampToEdge sampler of encoder.rs:1533-1537) for crop / inscribe / affine
// (command.rs:2507-2526, 2642-2740), resize (command.rs:1675-1702: bilinear.frag grid + palette.frag
// lookup, coordinates truncated to 8 bits), and this backend's additions: bilinear sampling,
// Porter-Duff blending, planar YUV 4:2:0 sources.
//
// A CTA of 256 threads produces a 32x32 destination tile, 4 consecutive pixels per thread.  The
// footprint of the tile in the source (the bounding box of the inverse-mapped tile) is staged
// into shared memory by ONE TMA bulk tensor copy per plane (cp.async.bulk.tensor.3d, completion on
// an mbarrier), double buffered so the next tile's box lands while the current tile is computed;
// out-of-image parts of the box are zero-filled by the TMA unit and never read (taps are clamped
// to the image first).  Tiles whose footprint does not fit the box fall back to direct loads.
#include "colorops.cuh"
#include "tma.cuh"
#include "zos_internal.h"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_gather)

constexpr int TILE = 32;
constexpr int THREADS = 256;

struct GatherParams {
  DevImage below, above, dst;
  int32_t has_below, below_vec, dst_vec;
  int32_t map, sampling, blend;
  int32_t sel[4], tgt[4];
  float inv[6];
  float rx, ry;  // sel.w / tgt.w, sel.h / tgt.h (bilinear rect mapping)
  float inj[8];  // ZOS_BLEND_INJECT: mix, color
  // row-band / tile sharding: dst holds the window of the full destination that starts
  // at (dox, doy); `above` holds the window of the full sfw x sfh source that starts at (sox, soy).
  // All mapping arithmetic uses FULL-image coordinates, so a windowed run gives the whole run's bytes.
  int32_t dox, doy, sox, soy, sfw, sfh;
  uint32_t tiles_x, tiles_y, total_tiles;
  FastDiv div_tx, div_ty;
  int32_t use_tma;        // 0 direct, 1 TMA (pixel source), 2 TMA (planar YUV source)
  int32_t box_w, box_h;   // texels, plane 0
  int32_t cbox_w, cbox_h; // chroma box
  int32_t ept;            // 4-byte elements per texel in the tensor map (bpp >= 4), else 0
  int32_t conv_w, conv_h; // FAST=2: size of the converted-footprint buffer (texels)
  int* fault;             // mapped host word set when an mbarrier wait runs away (zos_sync reports it)
  int32_t align_x;        // box origin x is rounded down to this many texels: the TMA source address must be 16-byte aligned
  StepList src_steps, dst_steps;
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}

// ---------------- tile geometry ----------------
struct TileInfo {
  uint32_t frame;
  int x0, y0;    // destination tile origin
  int bx, by;    // source box origin (plane 0 texels)
  int cbx, cby;  // chroma box origin
  int fx0, fy0;  // FAST=2: origin (source texels) of the tile's tap footprint
  bool fits;     // footprint fits the TMA box
  bool any;      // tile intersects the covered area at all
};

__device__ __forceinline__ void map_point(const GatherParams& P, float cx, float cy, float& px, float& py) {
  px = fmaf(P.inv[1], cy, fmaf(P.inv[0], cx, P.inv[2]));
  py = fmaf(P.inv[4], cy, fmaf(P.inv[3], cx, P.inv[5]));
}

__device__ TileInfo tile_info(const GatherParams& P, uint32_t t) {
  TileInfo ti;
  uint32_t r = fastdiv(t, P.div_tx);
  uint32_t txi = t - r * P.tiles_x;
  ti.frame = fastdiv(r, P.div_ty);
  uint32_t tyi = r - ti.frame * P.tiles_y;
  ti.x0 = (int)txi * TILE;
  ti.y0 = (int)tyi * TILE;
  ti.fits = false; ti.any = true;
  ti.bx = ti.by = ti.cbx = ti.cby = 0;
  ti.fx0 = ti.fy0 = 0;
  if (!P.use_tma) return ti;
  int x1 = min(ti.x0 + TILE, P.dst.w) - 1, y1 = min(ti.y0 + TILE, P.dst.h) - 1;  // inclusive
  float minx, maxx, miny, maxy;
  if (P.map == ZOS_MAP_AFFINE) {
    float px[4], py[4];
    const float gx0 = (float)(ti.x0 + P.dox) + 0.5f, gx1 = (float)(x1 + P.dox) + 0.5f;
    const float gy0 = (float)(ti.y0 + P.doy) + 0.5f, gy1 = (float)(y1 + P.doy) + 0.5f;
    map_point(P, gx0, gy0, px[0], py[0]);
    map_point(P, gx1, gy0, px[1], py[1]);
    map_point(P, gx0, gy1, px[2], py[2]);
    map_point(P, gx1, gy1, px[3], py[3]);
    minx = fminf(fminf(px[0], px[1]), fminf(px[2], px[3])); maxx = fmaxf(fmaxf(px[0], px[1]), fmaxf(px[2], px[3]));
    miny = fminf(fminf(py[0], py[1]), fminf(py[2], py[3])); maxy = fmaxf(fmaxf(py[0], py[1]), fmaxf(py[2], py[3]));
  } else {  // ZOS_MAP_RECT (also the exact resize): separable, monotone
    int ix0 = max(ti.x0 + P.dox, P.tgt[0]), ix1 = min(x1 + P.dox, P.tgt[0] + P.tgt[2] - 1);
    int iy0 = max(ti.y0 + P.doy, P.tgt[1]), iy1 = min(y1 + P.doy, P.tgt[1] + P.tgt[3] - 1);
    if (ix0 > ix1 || iy0 > iy1) { ti.any = false; return ti; }
    minx = P.sel[0] + ((float)(ix0 - P.tgt[0]) + 0.5f) * P.rx; maxx = P.sel[0] + ((float)(ix1 - P.tgt[0]) + 0.5f) * P.rx;
    miny = P.sel[1] + ((float)(iy0 - P.tgt[1]) + 0.5f) * P.ry; maxy = P.sel[1] + ((float)(iy1 - P.tgt[1]) + 0.5f) * P.ry;
  }
  ti.fx0 = min(max((int)floorf(minx - 0.5f), 0), P.sfw - 1);
  ti.fy0 = min(max((int)floorf(miny - 0.5f), 0), P.sfh - 1);
  // clip the footprint to the image: taps are clamped before they are fetched
  minx = fmaxf(minx, 0.0f); miny = fmaxf(miny, 0.0f);
  maxx = fminf(maxx, (float)P.sfw); maxy = fminf(maxy, (float)P.sfh);
  if (minx > maxx || miny > maxy) { ti.any = false; return ti; }
  // from here on in the coordinates of the window of the source that `above` holds
  int lox = max((int)floorf(minx) - 2 - P.sox, 0), hix = min((int)floorf(maxx) + 2 - P.sox, P.above.w - 1);
  int loy = max((int)floorf(miny) - 2 - P.soy, 0), hiy = min((int)floorf(maxy) + 2 - P.soy, P.above.h - 1);
  if (lox > hix || loy > hiy) { ti.any = false; return ti; }
  if (P.use_tma == 2) {  // chroma needs an even origin; bilinear chroma reaches one chroma texel further
    lox = max(lox - 2, 0) & ~(P.align_x - 1); loy = max((loy & ~1) - 2, 0);
    hix = min(hix + 2, P.above.w - 1); hiy = min(hiy + 2, P.above.h - 1);
    ti.cbx = lox >> 1; ti.cby = loy >> 1;
    ti.fits = ((hix >> 1) - ti.cbx + 1 <= P.cbox_w) && ((hiy >> 1) - ti.cby + 1 <= P.cbox_h);
  } else {
    lox &= ~(P.align_x - 1);
    ti.fits = true;
  }
  ti.bx = lox; ti.by = loy;
  ti.fits = ti.fits && (hix - lox + 1 <= P.box_w) && (hiy - loy + 1 <= P.box_h);
  return ti;
}

// ---------------- source fetch ----------------
struct Stage {          // one pipeline stage in dynamic shared memory
  const uint8_t* p0;    // plane 0 box
  const uint8_t* p1;    // U box
  const uint8_t* p2;    // V box
  const float4* conv;   // FAST=2: the tile's footprint converted to working values, [conv_h][conv_w]
};

template <bool SMEM>
__device__ __forceinline__ uint4 fetch_word(const GatherParams& P, const TileInfo& ti, const Stage& st, int u, int v) {
  const int bpp = P.above.bpp;
  const uint8_t* p;
  u -= P.sox; v -= P.soy;  // full-image coordinates -> the window `above` holds
  if (SMEM) p = st.p0 + ((size_t)(v - ti.by) * P.box_w + (u - ti.bx)) * bpp;
  else p = P.above.p0 + ti.frame * P.above.bstride + (uint64_t)v * P.above.pitch + (uint64_t)u * bpp;
  uint4 w = make_uint4(0, 0, 0, 0);
  switch (bpp) {
    case 1: w.x = *p; break;
    case 2: w.x = *reinterpret_cast<const uint16_t*>(p); break;
    case 4: w.x = *reinterpret_cast<const uint32_t*>(p); break;
    case 8: { uint2 t = *reinterpret_cast<const uint2*>(p); w.x = t.x; w.y = t.y; break; }
    default: w = *reinterpret_cast<const uint4*>(p); break;
  }
  return w;
}

template <bool SMEM>
__device__ __forceinline__ float chroma_at(const GatherParams& P, const TileInfo& ti, const uint8_t* box, const uint8_t* plane,
                                           int cx, int cy) {
  if (SMEM) return (float)box[(size_t)(cy - ti.cby) * P.cbox_w * (P.above.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1) +
                              (size_t)(cx - ti.cbx) * (P.above.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1)];
  const int step = P.above.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1;
  return (float)plane[ti.frame * P.above.cbstride + (uint64_t)cy * P.above.cpitch + (uint64_t)cx * step];
}

// one source texel, unpacked to working values (linear light), with the src steps applied
// EOTF applied to video samples: the BT.709 family (by far the common case) inline on the SFU with
// reciprocal multiplies, everything else through the generic curves of texel.cuh.
__device__ __forceinline__ float yuv_eotf(uint32_t tr, float v) {
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

__device__ __forceinline__ float4 half4_to_float4(const uint2& w) {
  float2 a = __half22float2(*reinterpret_cast<const __half2*>(&w.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

// FAST = 1: `above`, `below` and dst are all plain linear RGBA16F and there are no steps (BASELINE
// config 3): texels are 8-byte loads and two conversions, none of the generic codec is instantiated.
template <bool SMEM, int FAST>
__device__ __forceinline__ float4 fetch_texel(const GatherParams& P, const TileInfo& ti, const Stage& st, int u, int v, const Tables& T) {
  if (FAST == 1) {
    u -= P.sox; v -= P.soy;
    const uint8_t* p = SMEM ? st.p0 + ((size_t)(v - ti.by) * P.box_w + (u - ti.bx)) * 8
                            : P.above.p0 + ti.frame * P.above.bstride + (uint64_t)v * P.above.pitch + (uint64_t)u * 8;
    return half4_to_float4(*reinterpret_cast<const uint2*>(p));
  }
  if (FAST == 2 && SMEM) {  // already unpacked, converted and run through the src steps (phase 1 of the tile)
    return st.conv[(v - ti.fy0) * P.conv_w + (u - ti.fx0)];
  }
  float4 c;
  if (P.above.block == ZOS_BLOCK_PIXEL) {
    c = unpack_texel(P.above.fmt, fetch_word<SMEM>(P, ti, st, u, v), T);
  } else {
    float Y;  // (planar sources are never windowed: sox = soy = 0)
    if (SMEM) Y = (float)st.p0[(size_t)(v - ti.by) * P.box_w + (u - ti.bx)];
    else Y = (float)P.above.p0[ti.frame * P.above.bstride + (uint64_t)v * P.above.pitch + u];
    const int cw = (P.above.w + 1) >> 1, ch = (P.above.h + 1) >> 1;
    float U, V;
    if (!P.above.chroma_filter) {
      U = chroma_at<SMEM>(P, ti, st.p1, P.above.p1, u >> 1, v >> 1);
      V = chroma_at<SMEM>(P, ti, st.p2, P.above.p2, u >> 1, v >> 1);
    } else {
      float fx = ((float)u + 0.5f) * 0.5f - 0.5f, fy = ((float)v + 0.5f) * 0.5f - 0.5f;
      float x0f = floorf(fx), y0f = floorf(fy);
      float ax = fx - x0f, ay = fy - y0f;
      int x0 = (int)x0f, y0 = (int)y0f;
      int x1 = min(max(x0 + 1, 0), cw - 1), y1 = min(max(y0 + 1, 0), ch - 1);
      x0 = min(max(x0, 0), cw - 1); y0 = min(max(y0, 0), ch - 1);
      float u00 = chroma_at<SMEM>(P, ti, st.p1, P.above.p1, x0, y0), u10 = chroma_at<SMEM>(P, ti, st.p1, P.above.p1, x1, y0);
      float u01 = chroma_at<SMEM>(P, ti, st.p1, P.above.p1, x0, y1), u11 = chroma_at<SMEM>(P, ti, st.p1, P.above.p1, x1, y1);
      float v00 = chroma_at<SMEM>(P, ti, st.p2, P.above.p2, x0, y0), v10 = chroma_at<SMEM>(P, ti, st.p2, P.above.p2, x1, y0);
      float v01 = chroma_at<SMEM>(P, ti, st.p2, P.above.p2, x0, y1), v11 = chroma_at<SMEM>(P, ti, st.p2, P.above.p2, x1, y1);
      float ut = fmaf(ax, u10 - u00, u00), ub = fmaf(ax, u11 - u01, u01);
      float vt = fmaf(ax, v10 - v00, v00), vb = fmaf(ax, v11 - v01, v01);
      U = fmaf(ay, ub - ut, ut); V = fmaf(ay, vb - vt, vt);
    }
    // R'G'B' = matrix * (range-scaled Y'CbCr), then the colour's EOTF (semantics: DESIGN.md section 3)
    const DevImage& A = P.above;
    const float y = (Y - A.yoff) * A.ysc, cb = (U - 128.0f) * A.csc, cr = (V - 128.0f) * A.csc;
    const float r = fmaf(A.r_cr, cr, y), g = fmaf(-A.g_cb, cb, fmaf(-A.g_cr, cr, y)), b = fmaf(A.b_cb, cb, y);
    c = make_float4(yuv_eotf(A.fmt.transfer, r), yuv_eotf(A.fmt.transfer, g), yuv_eotf(A.fmt.transfer, b), 1.0f);
  }
  apply_steps(P.src_steps, c, T);
  return c;
}

template <bool SMEM, int FAST>
__device__ __forceinline__ float4 sample_bilinear(const GatherParams& P, const TileInfo& ti, const Stage& st, float px, float py,
                                                  const Tables& T) {
  float fx = px - 0.5f, fy = py - 0.5f;
  float x0f = floorf(fx), y0f = floorf(fy);
  float ax = fx - x0f, ay = fy - y0f;
  int x0 = (int)x0f, y0 = (int)y0f;
  const int sw = P.sfw, sh = P.sfh;
  int x1 = min(max(x0 + 1, 0), sw - 1), y1 = min(max(y0 + 1, 0), sh - 1);
  x0 = min(max(x0, 0), sw - 1); y0 = min(max(y0, 0), sh - 1);
  float4 p00, p10, p01, p11;
  if (FAST == 1) {
    // one base address, the three neighbours are at +8 bytes / +one row (or 0 where a tap was clamped)
    const int dx = (x1 - x0) * 8;
    if (SMEM) {  // 32-bit shared addresses, LDS.64
      const uint32_t a = smem_u32(st.p0) + (uint32_t)(((y0 - P.soy - ti.by) * P.box_w + (x0 - P.sox - ti.bx)) * 8);
      const uint32_t dy = (uint32_t)((y1 - y0) * P.box_w * 8);
      uint2 w00, w10, w01, w11;
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w00.x), "=r"(w00.y) : "r"(a));
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w10.x), "=r"(w10.y) : "r"(a + dx));
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w01.x), "=r"(w01.y) : "r"(a + dy));
      asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w11.x), "=r"(w11.y) : "r"(a + dy + dx));
      p00 = half4_to_float4(w00); p10 = half4_to_float4(w10); p01 = half4_to_float4(w01); p11 = half4_to_float4(w11);
    } else {
      const uint8_t* p = P.above.p0 + ti.frame * P.above.bstride + (uint64_t)(y0 - P.soy) * P.above.pitch + (uint64_t)(x0 - P.sox) * 8;
      const int64_t dy = (int64_t)(y1 - y0) * (int64_t)P.above.pitch;
      p00 = half4_to_float4(*reinterpret_cast<const uint2*>(p));
      p10 = half4_to_float4(*reinterpret_cast<const uint2*>(p + dx));
      p01 = half4_to_float4(*reinterpret_cast<const uint2*>(p + dy));
      p11 = half4_to_float4(*reinterpret_cast<const uint2*>(p + dy + dx));
    }
  } else {
    p00 = fetch_texel<SMEM, FAST>(P, ti, st, x0, y0, T); p10 = fetch_texel<SMEM, FAST>(P, ti, st, x1, y0, T);
    p01 = fetch_texel<SMEM, FAST>(P, ti, st, x0, y1, T); p11 = fetch_texel<SMEM, FAST>(P, ti, st, x1, y1, T);
  }
  float4 o;
#define ZOS_LERP2(c) { float top = fmaf(ax, p10.c - p00.c, p00.c), bot = fmaf(ax, p11.c - p01.c, p01.c); o.c = fmaf(ay, bot - top, top); }
  ZOS_LERP2(x) ZOS_LERP2(y) ZOS_LERP2(z) ZOS_LERP2(w)
#undef ZOS_LERP2
  return o;
}

__device__ __forceinline__ int rect_index(int k, int s, int t) {  // floor(((2k+1)*s) / (2t)), exact
  if (s == t) return k;
  return (int)(((uint64_t)(2 * (uint32_t)k + 1) * (uint32_t)s) / (2ull * (uint32_t)t));
}

// CommandBuffer::resize the way the reference does it (command.rs:1675-1702): the coordinate is
// written to an RGBA8 linear "Scalars" register (f16 texture, truncating pack), read back, biased
// by half a GRID texel and looked up with the nearest / clamp-to-edge sampler (palette.frag:21-32).
__device__ __forceinline__ int grid8_index(int i, int n_dst, int n_src, const Tables& T) {
  float u = ((float)i + 0.5f) / (float)n_dst;
  uint32_t q = (uint32_t)(clamp01(f16r(u)) * 255.0f);
  float c = f16r(T.unorm8[q]);
  float pu = c + 0.5f / (float)n_dst;
  int x = (int)floorf(pu * (float)n_src);
  return min(max(x, 0), n_src - 1);
}

__device__ __forceinline__ uint4 load_px(const DevImage& im, uint32_t frame, int x, int y) {
  const uint8_t* p = im.p0 + frame * im.bstride + (uint64_t)y * im.pitch + (uint64_t)x * im.bpp;
  uint4 w = make_uint4(0, 0, 0, 0);
  switch (im.bpp) {
    case 1: w.x = *p; break;
    case 2: w.x = *reinterpret_cast<const uint16_t*>(p); break;
    case 4: w.x = __ldcs(reinterpret_cast<const uint32_t*>(p)); break;
    case 8: { uint2 t = __ldcs(reinterpret_cast<const uint2*>(p)); w.x = t.x; w.y = t.y; break; }
    default: w = __ldcs(reinterpret_cast<const uint4*>(p)); break;
  }
  return w;
}
__device__ __forceinline__ void store_px(const DevImage& im, uint32_t frame, int x, int y, const uint4& w) {
  uint8_t* p = im.p0 + frame * im.bstride + (uint64_t)y * im.pitch + (uint64_t)x * im.bpp;
  switch (im.bpp) {
    case 1: *p = (uint8_t)w.x; break;
    case 2: *reinterpret_cast<uint16_t*>(p) = (uint16_t)w.x; break;
    case 4: __stcs(reinterpret_cast<uint32_t*>(p), w.x); break;
    case 8: __stcs(reinterpret_cast<uint2*>(p), make_uint2(w.x, w.y)); break;
    default: __stcs(reinterpret_cast<uint4*>(p), w); break;
  }
}

// A warp owns one 32-pixel row segment of the tile per iteration: all global accesses to `below`
// and dst are fully coalesced without per-thread vectors, for every texel size.
template <bool SMEM, int FAST>
__device__ __forceinline__ void compute_tile(const GatherParams& P, const TileInfo& ti, const Stage& st, const Tables& T) {
  const int i = ti.x0 + (threadIdx.x & 31);
  if (i >= P.dst.w) return;
#pragma unroll 1
  for (int k = 0; k < TILE / 8; k++) {
    const int j = ti.y0 + (threadIdx.x >> 5) + 8 * k;
    if (j >= P.dst.h) break;
    bool covered = false;
    float4 v = make_float4(0.0f, 0.0f, 1.0f, 1.0f);  // Target::Discard clear colour
    if (P.map == ZOS_MAP_AFFINE) {
      float px, py;
      map_point(P, (float)(i + P.dox) + 0.5f, (float)(j + P.doy) + 0.5f, px, py);
      if (px >= 0.0f && px < (float)P.sfw && py >= 0.0f && py < (float)P.sfh) {
        covered = true;
        v = P.sampling == ZOS_SAMPLE_NEAREST ? fetch_texel<SMEM, FAST>(P, ti, st, (int)floorf(px), (int)floorf(py), T)
                                             : sample_bilinear<SMEM, FAST>(P, ti, st, px, py, T);
      }
    } else if (P.map == ZOS_MAP_GRID8) {
      covered = true;
      v = fetch_texel<SMEM, FAST>(P, ti, st, grid8_index(i, P.dst.w, P.above.w, T), grid8_index(j, P.dst.h, P.above.h, T), T);
    } else {
      const int kx = i + P.dox - P.tgt[0], ky = j + P.doy - P.tgt[1];
      if (kx >= 0 && kx < P.tgt[2] && ky >= 0 && ky < P.tgt[3]) {
        covered = true;
        if (P.sampling == ZOS_SAMPLE_NEAREST) {
          int u = min(max(P.sel[0] + rect_index(kx, P.sel[2], P.tgt[2]), 0), P.sfw - 1);
          int w = min(max(P.sel[1] + rect_index(ky, P.sel[3], P.tgt[3]), 0), P.sfh - 1);
          v = fetch_texel<SMEM, FAST>(P, ti, st, u, w, T);
        } else {
          float px = (float)P.sel[0] + ((float)kx + 0.5f) * P.rx, py = (float)P.sel[1] + ((float)ky + 0.5f) * P.ry;
          v = sample_bilinear<SMEM, FAST>(P, ti, st, px, py, T);
        }
      }
    }
    if (FAST == 1) {
      if (P.has_below && !(covered && P.blend == ZOS_BLEND_OVERWRITE)) {
        const uint2 w = __ldcs(reinterpret_cast<const uint2*>(P.below.p0 + ti.frame * P.below.bstride + (uint64_t)j * P.below.pitch + (uint64_t)i * 8));
        float4 b = half4_to_float4(w);
        v = covered ? porter_duff(P.blend, v, b) : b;
      }
      __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      __stcs(reinterpret_cast<uint2*>(P.dst.p0 + ti.frame * P.dst.bstride + (uint64_t)j * P.dst.pitch + (uint64_t)i * 8),
             make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi)));
      continue;
    }
    if (FAST == 2) {
      // `below` and dst are native 8-bit texels, blend is overwrite or source-over, no dst steps
      if (P.has_below && !(covered && P.blend == ZOS_BLEND_OVERWRITE)) {
        const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(P.below.p0 + ti.frame * P.below.bstride + (uint64_t)j * P.below.pitch + (uint64_t)i * 4));
        const float4 b = unpack_texel(P.below.fmt, make_uint4(w, 0, 0, 0), T);
        if (covered) {  // porter_duff(ZOS_BLEND_SRC_OVER, v, b) written out
          const float wb = b.w * (1.0f - v.w), ao = v.w + wb;
          const float rcp = ao > 0.0f ? __frcp_rn(ao) : 0.0f;
          v = make_float4(fmaf(wb, b.x, v.w * v.x) * rcp, fmaf(wb, b.y, v.w * v.y) * rcp, fmaf(wb, b.z, v.w * v.z) * rcp, ao);
        } else {
          v = b;
        }
      }
      __stcs(reinterpret_cast<uint32_t*>(P.dst.p0 + ti.frame * P.dst.bstride + (uint64_t)j * P.dst.pitch + (uint64_t)i * 4), pack_texel(P.dst.fmt, v, T).x);
      continue;
    }
    if (P.has_below && !(covered && P.blend == ZOS_BLEND_OVERWRITE)) {
      float4 b = unpack_texel(P.below.fmt, load_px(P.below, ti.frame, i, j), T);
      v = !covered ? b : P.blend == ZOS_BLEND_INJECT ? inject_blend(P.inj, v, b) : porter_duff(P.blend, v, b);
    }
    apply_steps(P.dst_steps, v, T);
    store_px(P.dst, ti.frame, i, j, pack_texel(P.dst.fmt, v, T));
  }
}

template <int FAST>
__global__ void __launch_bounds__(THREADS) k_gather_direct(const __grid_constant__ GatherParams P) {
  __shared__ Tables T;
  load_tables(T);
  Stage st{nullptr, nullptr, nullptr, nullptr};
  for (uint32_t t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
    TileInfo ti = tile_info(P, t);
    compute_tile<false, FAST>(P, ti, st, T);
  }
}

template <int FAST>
__global__ void __launch_bounds__(THREADS) k_gather_tma(const __grid_constant__ GatherParams P, const __grid_constant__ TensorMaps M) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ Tables T;
  __shared__ __align__(8) uint64_t bar[2];
  load_tables(T);
  const int bpp = P.above.bpp;
  const uint32_t box0 = (uint32_t)P.box_w * P.box_h * bpp;
  const uint32_t cstep = P.above.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1;
  const uint32_t cbox = P.use_tma == 2 ? (uint32_t)P.cbox_w * P.cbox_h * cstep : 0;
  const uint32_t box0_al = (box0 + 127) & ~127u, cbox_al = (cbox + 127) & ~127u;
  const uint32_t nchroma = P.use_tma == 2 ? (cstep == 2 ? 1 : 2) : 0;
  const uint32_t stage_bytes = box0_al + nchroma * cbox_al;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // NB: the tensor maps are read by the TMA unit through their param-space address; never let them
  // be copied (no by-value captures / locals), a map in local memory is an illegal instruction.
  const CUtensorMap* const m0 = &M.m0;
  const CUtensorMap* const m1 = &M.m1;
  const CUtensorMap* const m2 = &M.m2;
#define ZOS_ISSUE(TILE_INDEX, STAGE)                                                                           \
  do {                                                                                                          \
    TileInfo ti_ = tile_info(P, (TILE_INDEX));                                                                  \
    if (ti_.fits && ti_.any) {                                                                                  \
      uint8_t* base_ = dyn + (size_t)(STAGE) * stage_bytes;                                                     \
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                            \
      mbar_expect_tx(&bar[(STAGE)], box0 + nchroma * cbox);                                                     \
      tma_load_3d(base_, m0, ti_.bx * (P.ept ? P.ept : 1), ti_.by, (int)ti_.frame, &bar[(STAGE)]);              \
      if (nchroma >= 1) tma_load_3d(base_ + box0_al, m1, ti_.cbx, ti_.cby, (int)ti_.frame, &bar[(STAGE)]);      \
      if (nchroma == 2) tma_load_3d(base_ + box0_al + cbox_al, m2, ti_.cbx, ti_.cby, (int)ti_.frame, &bar[(STAGE)]); \
    }                                                                                                           \
  } while (0)

  uint32_t phase[2] = {0, 0};
  int s = 0;
  if (threadIdx.x == 0 && blockIdx.x < P.total_tiles) ZOS_ISSUE(blockIdx.x, 0);
  for (uint32_t t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
    const uint32_t next = t + gridDim.x;
    if (threadIdx.x == 0 && next < P.total_tiles) ZOS_ISSUE(next, s ^ 1);
    TileInfo ti = tile_info(P, t);
    if (ti.any) {
      if (ti.fits) {
        uint32_t spins = 0;
        while (!mbar_try_wait(&bar[s], phase[s])) {
          if (++spins > (1u << 24)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; break; }
        }
        phase[s] ^= 1;
        uint8_t* base = dyn + (size_t)s * stage_bytes;
        float4* conv = reinterpret_cast<float4*>(dyn + 2 * (size_t)stage_bytes);
        Stage st{base, base + box0_al, cstep == 2 ? base + box0_al + 1 : base + box0_al + cbox_al, conv};
        if (FAST == 2) {
          // phase 1: every source texel of the tile's tap footprint is unpacked / converted / stepped ONCE
          // (a bilinear tap pattern would otherwise convert each of them ~4 / scale^2 times)
          const int n = P.conv_w * P.conv_h;
          for (int idx = threadIdx.x; idx < n; idx += THREADS) {
            const int ry = idx / P.conv_w, rx = idx - ry * P.conv_w;
            const int u = min(ti.fx0 + rx, P.sfw - 1), v = min(ti.fy0 + ry, P.sfh - 1);
            conv[idx] = fetch_texel<true, 0>(P, ti, st, u, v, T);
          }
          __syncthreads();
        }
        compute_tile<true, FAST>(P, ti, st, T);
      } else {
        Stage st{nullptr, nullptr, nullptr, nullptr};
        compute_tile<false, FAST>(P, ti, st, T);
      }
    } else {
      Stage st{nullptr, nullptr, nullptr, nullptr};
      compute_tile<false, FAST>(P, ti, st, T);  // nothing of `above` lands here: copies `below`
    }
    __syncthreads();  // everyone is done with stage s before it is refilled two tiles later
    s ^= 1;
  }
}

// ---------------- host side ----------------
static bool img_vec_ok(const DevImage& im) {
  return ((uintptr_t)im.p0 % 16) == 0 && (im.pitch % 16) == 0 && (im.bstride % 16) == 0 &&
         im.pitch >= (uint64_t)((im.w + 3) / 4) * 4 * im.bpp && (im.bpp == 4 || im.bpp == 8);
}

zos_status launch_gather(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                         const zos_compose_params& cp, uint32_t batch) {
  if (dst.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "planar YUV destinations are not implemented yet");
  if (frame_pipeline_eligible(ctx, below, above, dst, cp)) {  // the dedicated video-frame kernel (frame_pipeline.cu)
    bool handled = false;
    zos_status st = launch_frame_pipeline(ctx, below, above, dst, cp, batch, &handled);
    if (handled || st != ZOS_OK) return st;
  }
  {  // the dedicated RGBA16F resampler (affine_f16.cu)
    bool handled = false;
    zos_status st = launch_affine_f16(ctx, below, above, dst, cp, batch, &handled);
    if (handled || st != ZOS_OK) return st;
  }
  if (below && below->block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "planar `below`");
  GatherParams P;
  memset(&P, 0, sizeof P);
  P.above = above; P.dst = dst;
  P.fault = ctx->fault_dev;
  P.has_below = below != nullptr;
  if (below) { P.below = *below; P.below_vec = img_vec_ok(*below); }
  P.dst_vec = img_vec_ok(dst);
  P.map = cp.map; P.sampling = cp.sampling; P.blend = cp.blend;
  for (int k = 0; k < 4; k++) { P.sel[k] = cp.sel[k]; P.tgt[k] = cp.tgt[k]; }
  if (cp.map == ZOS_MAP_SCALE || cp.map == ZOS_MAP_GRID8) {
    P.sel[0] = P.sel[1] = 0; P.sel[2] = above.w; P.sel[3] = above.h;
    P.tgt[0] = P.tgt[1] = 0; P.tgt[2] = dst.w; P.tgt[3] = dst.h;
    if (cp.map == ZOS_MAP_SCALE) P.map = ZOS_MAP_RECT;
  }
  for (int k = 0; k < 6; k++) P.inv[k] = cp.inv[k];
  memcpy(P.inj, cp.inject_mix, 16); memcpy(P.inj + 4, cp.inject_color, 16);
  P.dox = cp.dst_origin[0]; P.doy = cp.dst_origin[1]; P.sox = cp.src_origin[0]; P.soy = cp.src_origin[1];
  P.sfw = cp.src_full[0] > 0 ? cp.src_full[0] : above.w; P.sfh = cp.src_full[1] > 0 ? cp.src_full[1] : above.h;
  const bool windowed = P.dox || P.doy || P.sox || P.soy || P.sfw != above.w || P.sfh != above.h;
  if (windowed && (cp.map == ZOS_MAP_GRID8 || cp.map == ZOS_MAP_SCALE || above.block != ZOS_BLOCK_PIXEL))
    return fail(ctx, ZOS_ERR_UNSUPPORTED, "windowed (sharded) launches support ZOS_MAP_RECT / ZOS_MAP_AFFINE on pixel sources");
  if (P.dox < 0 || P.doy < 0 || P.sox < 0 || P.soy < 0 || P.sox + above.w > P.sfw || P.soy + above.h > P.sfh)
    return fail(ctx, ZOS_ERR_INVALID, "window outside of the full image");
  P.rx = (float)P.sel[2] / (float)(P.tgt[2] > 0 ? P.tgt[2] : 1);
  P.ry = (float)P.sel[3] / (float)(P.tgt[3] > 0 ? P.tgt[3] : 1);
  P.src_steps.n = cp.n_src_steps;
  for (uint32_t i = 0; i < cp.n_src_steps; i++) P.src_steps.s[i] = cp.src_steps[i];
  P.dst_steps.n = cp.n_dst_steps;
  for (uint32_t i = 0; i < cp.n_dst_steps; i++) P.dst_steps.s[i] = cp.dst_steps[i];
  P.tiles_x = (dst.w + TILE - 1) / TILE;
  P.tiles_y = (dst.h + TILE - 1) / TILE;
  uint64_t total = (uint64_t)P.tiles_x * P.tiles_y * batch;
  if (total >= (1ull << 31)) return fail(ctx, ZOS_ERR_UNSUPPORTED, "gather: too many tiles in one launch");
  P.total_tiles = (uint32_t)total;
  P.div_tx = make_fastdiv(P.tiles_x);
  P.div_ty = make_fastdiv(P.tiles_y);

  // ---- TMA plan: box = bounding box of a 32x32 destination tile in the source, plus margins
  bool tma = cp.use_tma && P.map != ZOS_MAP_GRID8;
  bool two_phase = false;
  TensorMaps M;
  memset(&M, 0, sizeof M);
  size_t smem = 0;
  if (tma) {
    float ex, ey;
    if (P.map == ZOS_MAP_AFFINE) {
      ex = (TILE - 1) * (fabsf(P.inv[0]) + fabsf(P.inv[1]));
      ey = (TILE - 1) * (fabsf(P.inv[3]) + fabsf(P.inv[4]));
    } else {
      ex = (TILE - 1) * P.rx; ey = (TILE - 1) * P.ry;
    }
    const bool yuv = above.block != ZOS_BLOCK_PIXEL;
    int bpp = above.bpp;
    int gran = bpp >= 16 ? 1 : 16 / bpp;  // box rows must be a multiple of 16 bytes, and so must the box origin
    if (yuv) gran = 32;                   // so that the chroma box (half width) is too
    P.align_x = gran;
    int need_w = (int)ceilf(ex) + (yuv ? 12 : 6) + (gran - 1), need_h = (int)ceilf(ey) + (yuv ? 12 : 6);
    int bw = (need_w + gran - 1) / gran * gran, bh = need_h;
    P.ept = bpp >= 4 ? bpp / 4 : 0;
    uint32_t bw_elems = P.ept ? bw * P.ept : bw;
    size_t stage = ((size_t)bw * bh * bpp + 127) & ~(size_t)127;
    int cstep = above.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1;
    if (yuv) {
      P.cbox_w = bw / 2; P.cbox_h = (bh + 1) / 2 + 1;
      size_t cb = ((size_t)P.cbox_w * P.cbox_h * cstep + 127) & ~(size_t)127;
      stage += cb * (cstep == 2 ? 1 : 2);
    }
    smem = 2 * stage;
    bool ok = bw_elems <= 256 && bh <= 256 && smem <= 96 * 1024;
    if (ok) {
      CUtensorMapDataType dt = bpp == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : bpp == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
      int eb = bpp >= 4 ? 4 : bpp;
      uint64_t w_elems = P.ept ? (uint64_t)above.w * P.ept : (uint64_t)above.w;
      ok = make_map(ctx, &M.m0, dt, eb, above.p0, w_elems, above.h, above.pitch, batch, above.bstride, bw_elems, bh);
      if (ok && yuv) {
        uint64_t cw = (above.w + 1) / 2, ch = (above.h + 1) / 2;
        if (cstep == 2) {
          ok = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
        } else {
          ok = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h) &&
               make_map(ctx, &M.m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p2, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
        }
      }
    }
    // FAST=2 (two-phase tiles): planar source, separable mapping; the converted footprint lives behind the stages
    auto native8 = [](const DevImage& im) {
      return im.block == ZOS_BLOCK_PIXEL && im.bpp == 4 && (im.fmt.storage == ZOS_STORAGE_SRGB8 || im.fmt.storage == ZOS_STORAGE_UNORM8) &&
             ((uintptr_t)im.p0 % 4) == 0 && (im.pitch % 4) == 0 && (im.bstride % 4) == 0;
    };
    two_phase = ok && yuv && P.map == ZOS_MAP_RECT && !(ctx->flags & ZOS_CTX_NO_FAST_PATHS) && cp.n_dst_steps == 0 &&
                (cp.blend == ZOS_BLEND_OVERWRITE || cp.blend == ZOS_BLEND_SRC_OVER) && native8(dst) && (!below || native8(*below));
    if (two_phase) {
      P.conv_w = (int)ceilf(ex) + 4; P.conv_h = (int)ceilf(ey) + 4;
      size_t conv_bytes = (size_t)P.conv_w * P.conv_h * sizeof(float4);
      if (smem + conv_bytes <= 150 * 1024) smem += conv_bytes; else two_phase = false;
    }
    if (ok) { P.use_tma = yuv ? 2 : 1; P.box_w = bw; P.box_h = bh; } else { tma = false; P.use_tma = 0; smem = 0; two_phase = false; }
  }

  auto plain_f16 = [](const DevImage& im) {
    return im.block == ZOS_BLOCK_PIXEL && im.fmt.storage == ZOS_STORAGE_FLOAT && im.fmt.bits == ZOS_BITS_FLOAT16X4 &&
           im.fmt.transfer == ZOS_TRANSFER_LINEAR && (im.fmt.parts == ZOS_PARTS_RGBA || im.fmt.parts == ZOS_PARTS_LCHA || im.fmt.parts == ZOS_PARTS_LABA) &&
           ((uintptr_t)im.p0 % 8) == 0 && (im.pitch % 8) == 0 && (im.bstride % 8) == 0;
  };
  const bool fast = !(ctx->flags & ZOS_CTX_NO_FAST_PATHS) && P.map != ZOS_MAP_GRID8 && cp.blend != ZOS_BLEND_INJECT && cp.n_src_steps == 0 && cp.n_dst_steps == 0 &&
                    plain_f16(above) && plain_f16(dst) && (!below || plain_f16(*below));
  cudaError_t e;
  if (tma) {
    ensure_dyn_smem(ctx, k_gather_tma<0>, 96 * 1024);
    ensure_dyn_smem(ctx, k_gather_tma<1>, 96 * 1024);
    ensure_dyn_smem(ctx, k_gather_tma<2>, 150 * 1024);
    int per_sm = (int)((200 * 1024) / (smem + sizeof(Tables) + 1024));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
    int grid = (int)(total < cap ? total : cap);
    if (two_phase) k_gather_tma<2><<<grid, THREADS, smem, ctx->stream>>>(P, M);
    else if (fast) k_gather_tma<1><<<grid, THREADS, smem, ctx->stream>>>(P, M);
    else k_gather_tma<0><<<grid, THREADS, smem, ctx->stream>>>(P, M);
    e = cudaGetLastError();
  } else {
    uint64_t cap = (uint64_t)ctx->sm_count * 8;
    int grid = (int)(total < cap ? total : cap);
    if (fast) k_gather_direct<1><<<grid, THREADS, 0, ctx->stream>>>(P);
    else k_gather_direct<0><<<grid, THREADS, 0, ctx->stream>>>(P);
    e = cudaGetLastError();
  }
  ctx->launches++;
  return check_cuda(ctx, e, "k_gather launch");
}

}  // namespace zos
