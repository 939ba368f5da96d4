// frame_pipeline.cu -- the fused video-frame pipeline of BASELINE config 4 as ONE kernel:
//
//   planar YUV 4:2:0 (I420 / NV12) unpack -> Y'CbCr -> R'G'B' -> EOTF -> [3x3 primaries matrix]
//   -> nearest / bilinear resize (axis-aligned placement) -> [source-over onto an 8-bit background]
//   -> sRGB8 / unorm8 pack,
//
// for a whole batch of frames per launch.  Same arithmetic, bit for bit, as the general gather
// kernel (gather.cu) that also serves this case; this file is the lean version of it:
//
//   * a CTA produces 32x32 destination tiles (persistent, grid-stride over tiles x frames);
//   * the Y and chroma footprints of the NEXT tile are fetched by TMA bulk tensor copies
//     (cp.async.bulk.tensor.3d + mbarrier) while the current tile is computed (double buffered);
//   * phase 1 converts the staged footprint to linear-light RGB exactly once per source texel, one
//     thread per 2x2 luma block (the block shares its chroma sample and all index arithmetic), into
//     a float4 tile in shared memory;
//   * phase 2 samples that tile (4 LDS.128 + 3 lerps per channel for bilinear), composites over the
//     background and packs; global accesses of a warp are 32 consecutive pixels of one row;
//   * tile geometry is computed by one thread for the next tile and broadcast through shared memory.
//
// None of this exists in the reference (no planar texels, no bilinear, no blend: SURVEY.md 0.2);
// semantics are DESIGN.md section 3, the oracle is oracle/zos_oracle.c (zo_decode_yuv420, zo_linear,
// zo_resize, zo_blend, zo_encode).
#include <stdlib.h>

#include "colorops.cuh"
#include "f32x2.cuh"
#include "tma.cuh"
#include "zos_internal.h"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_frame)

namespace {
constexpr int TILE = 32;
constexpr int THREADS = 256;

struct FrameParams {
  int32_t sw, sh, nv12;
  float yoff, ysc, csc, r_cr, b_cb, g_cr, g_cb;
  uint32_t transfer;
  int32_t nmat;
  float m[9];
  int32_t sel[4], tgt[4];
  float rx, ry;
  int32_t sampling, blend;
  const uint8_t* below;
  uint64_t below_pitch, below_bstride;
  int32_t has_below;
  zos_texfmt below_fmt, dst_fmt;
  uint8_t* dst;
  uint64_t dst_pitch, dst_bstride;
  int32_t dw, dh;
  uint32_t tiles_x, tiles_y, total_tiles;
  FastDiv div_tx, div_ty;
  int32_t box_w, box_h, cbox_w, cbox_h, conv_w, conv_h;
  int* fault;  // mapped host word set when an mbarrier wait runs away
  // k_frame_fast
  int32_t tile_h;       // rows of a destination tile (TILE, or SPEC_TH for k_frame_spec)
  uint32_t clear_word;  // the encoded Target::Discard clear colour (0, 0, 1, 1)
  uint32_t spack;       // byte-permute selector of the destination word (RGBA / BGRA)
  uint32_t plane_bytes; // bytes of one plane of the converted footprint
  FastDiv div_cbw;      // division by conv_w / 2
};

struct TileGeo {
  int32_t frame, x0, y0;  // destination tile
  int32_t bx, by;         // luma box origin (multiple of 32 / even), chroma box origin is half of it
  int32_t fx0, fy0;       // origin of the converted footprint (even)
  int32_t any;            // does the placement of the frame touch this tile at all
  int32_t full;           // every pixel of the tile is a destination pixel covered by the frame (k_frame_fast's straight path)
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}

__device__ void tile_geometry(const FrameParams& P, uint32_t t, TileGeo& g) {
  uint32_t r = fastdiv(t, P.div_tx);
  uint32_t txi = t - r * P.tiles_x;
  uint32_t fr = fastdiv(r, P.div_ty);
  uint32_t tyi = r - fr * P.tiles_y;
  g.frame = (int)fr; g.x0 = (int)txi * TILE; g.y0 = (int)tyi * P.tile_h;
  g.bx = g.by = g.fx0 = g.fy0 = 0; g.any = 0; g.full = 0;
  const int x1 = min(g.x0 + TILE, P.dw) - 1, y1 = min(g.y0 + P.tile_h, P.dh) - 1;
  const int ix0 = max(g.x0, P.tgt[0]), ix1 = min(x1, P.tgt[0] + P.tgt[2] - 1);
  const int iy0 = max(g.y0, P.tgt[1]), iy1 = min(y1, P.tgt[1] + P.tgt[3] - 1);
  if (ix0 > ix1 || iy0 > iy1) return;
  const float minx = (float)P.sel[0] + ((float)(ix0 - P.tgt[0]) + 0.5f) * P.rx;
  const float miny = (float)P.sel[1] + ((float)(iy0 - P.tgt[1]) + 0.5f) * P.ry;
  g.fx0 = min(max((int)floorf(minx - 0.5f), 0), P.sw - 1) & ~1;
  g.fy0 = min(max((int)floorf(miny - 0.5f), 0), P.sh - 1) & ~1;
  g.bx = g.fx0 & ~31;  // 16-byte aligned TMA source address for the luma AND the (half width) chroma planes
  g.by = g.fy0;
  g.any = 1;
  g.full = ix0 == g.x0 && iy0 == g.y0 && ix1 == g.x0 + TILE - 1 && iy1 == g.y0 + P.tile_h - 1;
}

// EOTF of video samples (same code as gather.cu's yuv_eotf)
__device__ __forceinline__ float eotf(uint32_t tr, float v) {
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

__device__ __forceinline__ float4 convert(const FrameParams& P, float Y, float cb, float cr) {
  const float y = (Y - P.yoff) * P.ysc;
  float r = fmaf(P.r_cr, cr, y), g = fmaf(-P.g_cb, cb, fmaf(-P.g_cr, cr, y)), b = fmaf(P.b_cb, cb, y);
  r = eotf(P.transfer, r); g = eotf(P.transfer, g); b = eotf(P.transfer, b);
  if (P.nmat) {
    float3 t = mat3_mul(P.m, r, g, b);
    r = t.x; g = t.y; b = t.z;
  }
  return make_float4(r, g, b, 1.0f);
}

__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

__device__ __forceinline__ int rect_index(int k, int s, int t) {  // floor(((2k+1)*s) / (2t)), exact
  if (s == t) return k;
  return (int)(((uint64_t)(2 * (uint32_t)k + 1) * (uint32_t)s) / (2ull * (uint32_t)t));
}

__global__ void __launch_bounds__(THREADS, 3) k_frame_pipeline(const __grid_constant__ FrameParams P, const __grid_constant__ TensorMaps M) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ Tables T;
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ TileGeo geo[2];
  load_tables(T);
  const uint32_t cstep = P.nv12 ? 2 : 1;
  const uint32_t ybox = (uint32_t)P.box_w * P.box_h, cbox = (uint32_t)P.cbox_w * P.cbox_h * cstep;
  const uint32_t ybox_al = (ybox + 127) & ~127u, cbox_al = (cbox + 127) & ~127u;
  const uint32_t nchroma = P.nv12 ? 1 : 2;
  const uint32_t stage_bytes = ybox_al + nchroma * cbox_al;
  const uint32_t conv_base = smem_u32(dyn + 2 * (size_t)stage_bytes);
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the barriers exist before anybody (thread 0 included) arms or polls them
  const CUtensorMap* const m0 = &M.m0;
  const CUtensorMap* const m1 = &M.m1;
  const CUtensorMap* const m2 = &M.m2;
#define ZOS_FRAME_ISSUE(TILE_INDEX, STAGE)                                                                 \
  do {                                                                                                      \
    TileGeo g_;                                                                                             \
    tile_geometry(P, (TILE_INDEX), g_);                                                                     \
    geo[(STAGE)] = g_;                                                                                      \
    if (g_.any) {                                                                                           \
      uint8_t* base_ = dyn + (size_t)(STAGE) * stage_bytes;                                                 \
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                        \
      mbar_expect_tx(&bar[(STAGE)], ybox + nchroma * cbox);                                                 \
      tma_load_3d(base_, m0, g_.bx, g_.by, g_.frame, &bar[(STAGE)]);                                        \
      tma_load_3d(base_ + ybox_al, m1, g_.bx >> 1, g_.by >> 1, g_.frame, &bar[(STAGE)]);                    \
      if (nchroma == 2) tma_load_3d(base_ + ybox_al + cbox_al, m2, g_.bx >> 1, g_.by >> 1, g_.frame, &bar[(STAGE)]); \
    }                                                                                                       \
  } while (0)

  if (threadIdx.x == 0 && blockIdx.x < P.total_tiles) ZOS_FRAME_ISSUE(blockIdx.x, 0);
  __syncthreads();
  uint32_t phase[2] = {0, 0};
  int s = 0;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (uint32_t t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
    const uint32_t next = t + gridDim.x;
    if (threadIdx.x == 0 && next < P.total_tiles) ZOS_FRAME_ISSUE(next, s ^ 1);
    const TileGeo g = geo[s];
    if (g.any) {
      uint32_t spins = 0;
      while (!mbar_try_wait(&bar[s], phase[s])) {
        if (++spins > (1u << 24)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; break; }
      }
      phase[s] ^= 1;
      // ---- phase 1: the footprint, one 2x2 luma block per thread
      const uint32_t ybase = smem_u32(dyn + (size_t)s * stage_bytes);
      const uint32_t ubase = ybase + ybox_al, vbase = P.nv12 ? ubase + 1 : ubase + cbox_al;
      const int offx = g.fx0 - g.bx, offy = g.fy0 - g.by;  // both even
      const int cbw = P.conv_w >> 1, cbh = P.conv_h >> 1;
      if (lx < cbw) {
        for (int byi = ly; byi < cbh; byi += THREADS / 32) {
          const uint32_t ya = ybase + (uint32_t)((offy + 2 * byi) * P.box_w + offx + 2 * lx);
          const uint32_t ca = (uint32_t)(((offy >> 1) + byi) * P.cbox_w + (offx >> 1) + lx) * cstep;
          uint32_t y01, y23, u8, v8;
          asm("ld.shared.u16 %0, [%1];" : "=r"(y01) : "r"(ya));
          asm("ld.shared.u16 %0, [%1];" : "=r"(y23) : "r"(ya + (uint32_t)P.box_w));
          asm("ld.shared.u8 %0, [%1];" : "=r"(u8) : "r"(ubase + ca));
          asm("ld.shared.u8 %0, [%1];" : "=r"(v8) : "r"(vbase + ca));
          const float cb = ((float)u8 - 128.0f) * P.csc, cr = ((float)v8 - 128.0f) * P.csc;
          const uint32_t o = conv_base + (uint32_t)((2 * byi) * P.conv_w + 2 * lx) * 16u;
          const float4 c00 = convert(P, (float)(y01 & 255u), cb, cr), c10 = convert(P, (float)(y01 >> 8), cb, cr);
          const float4 c01 = convert(P, (float)(y23 & 255u), cb, cr), c11 = convert(P, (float)(y23 >> 8), cb, cr);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(o), "f"(c00.x), "f"(c00.y), "f"(c00.z), "f"(c00.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(o + 16u), "f"(c10.x), "f"(c10.y), "f"(c10.z), "f"(c10.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(o + (uint32_t)P.conv_w * 16u), "f"(c01.x), "f"(c01.y), "f"(c01.z), "f"(c01.w) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(o + (uint32_t)P.conv_w * 16u + 16u), "f"(c11.x), "f"(c11.y), "f"(c11.z), "f"(c11.w) : "memory");
        }
      }
      __syncthreads();
    }
    // ---- phase 2: sample, composite, pack
    const int i = g.x0 + lx;
    if (i < P.dw) {
      const int kx = i - P.tgt[0];
      const bool col_in = g.any && kx >= 0 && kx < P.tgt[2];
      // horizontal taps are the same for the 4 rows this thread produces
      int xa = 0, xb = 0;
      float ax = 0.0f;
      if (col_in) {
        if (P.sampling == ZOS_SAMPLE_NEAREST) {
          xa = xb = min(max(P.sel[0] + rect_index(kx, P.sel[2], P.tgt[2]), 0), P.sw - 1);
        } else {
          const float px = (float)P.sel[0] + ((float)kx + 0.5f) * P.rx;
          const float fx = px - 0.5f, x0f = floorf(fx);
          ax = fx - x0f;
          const int x0 = (int)x0f;
          xb = min(max(x0 + 1, 0), P.sw - 1);
          xa = min(max(x0, 0), P.sw - 1);
        }
        xa -= g.fx0; xb -= g.fx0;
      }
#pragma unroll 1
      for (int k = 0; k < TILE / 8; k++) {
        const int j = g.y0 + ly + 8 * k;
        if (j >= P.dh) break;
        const int ky = j - P.tgt[1];
        const bool covered = col_in && ky >= 0 && ky < P.tgt[3];
        float4 v = make_float4(0.0f, 0.0f, 1.0f, 1.0f);  // Target::Discard clear colour
        if (covered) {
          if (P.sampling == ZOS_SAMPLE_NEAREST) {
            const int w = min(max(P.sel[1] + rect_index(ky, P.sel[3], P.tgt[3]), 0), P.sh - 1) - g.fy0;
            v = lds128(conv_base + (uint32_t)(w * P.conv_w + xa) * 16u);
          } else {
            const float py = (float)P.sel[1] + ((float)ky + 0.5f) * P.ry;
            const float fy = py - 0.5f, y0f = floorf(fy);
            const float ay = fy - y0f;
            const int y0 = (int)y0f;
            const int yb = min(max(y0 + 1, 0), P.sh - 1) - g.fy0, ya = min(max(y0, 0), P.sh - 1) - g.fy0;
            const uint32_t ra = conv_base + (uint32_t)(ya * P.conv_w) * 16u, rb = conv_base + (uint32_t)(yb * P.conv_w) * 16u;
            const float4 p00 = lds128(ra + xa * 16u), p10 = lds128(ra + xb * 16u), p01 = lds128(rb + xa * 16u), p11 = lds128(rb + xb * 16u);
#define ZOS_LERP2(c) { float top = fmaf(ax, p10.c - p00.c, p00.c), bot = fmaf(ax, p11.c - p01.c, p01.c); v.c = fmaf(ay, bot - top, top); }
            ZOS_LERP2(x) ZOS_LERP2(y) ZOS_LERP2(z) ZOS_LERP2(w)
#undef ZOS_LERP2
          }
        }
        if (P.has_below && !(covered && P.blend == ZOS_BLEND_OVERWRITE)) {
          const uint32_t w = __ldcs(reinterpret_cast<const uint32_t*>(P.below + (uint64_t)g.frame * P.below_bstride + (uint64_t)j * P.below_pitch + (uint64_t)i * 4));
          const float4 b = unpack_texel(P.below_fmt, make_uint4(w, 0, 0, 0), T);
          if (covered) {  // porter_duff(ZOS_BLEND_SRC_OVER, v, b) written out
            const float wb = b.w * (1.0f - v.w), ao = v.w + wb;
            const float rcp = ao > 0.0f ? __frcp_rn(ao) : 0.0f;
            v = make_float4(fmaf(wb, b.x, v.w * v.x) * rcp, fmaf(wb, b.y, v.w * v.y) * rcp, fmaf(wb, b.z, v.w * v.z) * rcp, ao);
          } else {
            v = b;
          }
        }
        __stcs(reinterpret_cast<uint32_t*>(P.dst + (uint64_t)g.frame * P.dst_bstride + (uint64_t)j * P.dst_pitch + (uint64_t)i * 4),
               pack_texel(P.dst_fmt, v, T).x);
      }
    }
    __syncthreads();  // conv and stage s are free again; geo[s ^ 1] (written by thread 0 above) is visible
    s ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// k_frame_fast: the same pipeline specialised at compile time for what BASELINE config 4 runs:
// BT.709-family or linear EOTF, native 8-bit destination, background (if any) in the destination's
// texel.  Same arithmetic as k_frame_pipeline, far fewer instructions:
//   * a planar-YUV frame is opaque (alpha = 1 exactly), so source-over of a covered pixel IS the frame's
//     colour (fmaf(0, b, 1 * v) * 1 == v): neither decode nor blend, and `below` is only read - as raw
//     words, decode -> encode being the identity - where the frame does not cover the destination;
//   * the converted footprint is kept as three f32 planes (no alpha plane), sized by the exact tap span;
//   * sRGB8 encode through the bucket table (texel.cuh): one look-up per channel, no transcendental;
//   * no run-time format / transfer / sampling switches.
constexpr int FER = 2;  // copies of the (biased-key, texel.cuh) encoder bucket table in shared memory: 5 KB, so that 4 CTAs fit
                        // an SM (8 copies and 3 CTAs: 0.170 of the HBM peak on C4; 2 copies and 4 CTAs: 0.178 -- the three encoder
                        // look-ups of a pixel conflict more, the two CTA barriers per tile hurt less)

template <int TRK>
__device__ __forceinline__ float eotf_k(uint32_t tr, float v) {
  if (TRK == 0) {  // BT.709 / BT.2020 10- and 12-bit: one curve
    float lin = v * (1.0f / 4.5f);
    float l2, pw;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"((v + 0.099f) * (1.0f / 1.099f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(l2 * (1.0f / 0.45f)));
    return v >= 0.0812428582f ? pw : lin;
  }
  if (TRK == 1) return v;
  return eo_scalar(tr, v);
}

template <int TRK>
__device__ __forceinline__ float3 convert_rgb(const FrameParams& P, float Y, float cb, float cr) {
  const float y = (Y - P.yoff) * P.ysc;
  float r = fmaf(P.r_cr, cr, y), g = fmaf(-P.g_cb, cb, fmaf(-P.g_cr, cr, y)), b = fmaf(P.b_cb, cb, y);
  r = eotf_k<TRK>(P.transfer, r); g = eotf_k<TRK>(P.transfer, g); b = eotf_k<TRK>(P.transfer, b);
  if (P.nmat) return mat3_mul(P.m, r, g, b);
  return make_float3(r, g, b);
}
// two horizontally adjacent texels of one plane: 8-byte stores, a warp writes 256 contiguous bytes
__device__ __forceinline__ void sts64(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}

// ---- round 2: two texels at once on the packed f32x2 instructions (f32x2.cuh); same operations, same order, same bits.
__device__ __forceinline__ void sts64(uint32_t addr, F2 v) { sts64(addr, f2_lo(v), f2_hi(v)); }
template <int TRK>
__device__ __forceinline__ F2 eotf_k2(uint32_t tr, F2 v) {
  if (TRK == 0) {
    const F2 lin = f2_mul(v, f2(1.0f / 4.5f));
    const F2 arg = f2_mul(f2_add(v, f2(0.099f)), f2(1.0f / 1.099f));
    float l0, l1, p0, p1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l0) : "f"(f2_lo(arg)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l1) : "f"(f2_hi(arg)));
    const F2 e = f2_mul(f2(l0, l1), f2(1.0f / 0.45f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(f2_lo(e)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(f2_hi(e)));
    return f2(f2_lo(v) >= 0.0812428582f ? p0 : f2_lo(lin), f2_hi(v) >= 0.0812428582f ? p1 : f2_hi(lin));
  }
  if (TRK == 1) return v;
  return f2(eo_scalar(tr, f2_lo(v)), eo_scalar(tr, f2_hi(v)));
}
// convert_rgb on the two luma samples in bytes 0 and 1 of `ypair` (they share the block's chroma)
template <int TRK>
__device__ __forceinline__ void convert_rgb2(const FrameParams& P, uint32_t ypair, float cb, float cr, F2& r, F2& g, F2& b) {
  // (float)code as 0x4b0000cc - 2^23, exact, on the FMA pipe instead of the conversion unit
  const F2 Y = f2_sub(f2(__uint_as_float(__byte_perm(ypair, 0x4b000000u, 0x7540)), __uint_as_float(__byte_perm(ypair, 0x4b000000u, 0x7541))), f2(8388608.0f));
  const F2 y = f2_mul(f2_sub(Y, f2(P.yoff)), f2(P.ysc));
  r = f2_fma(f2(P.r_cr), f2(cr), y); g = f2_fma(f2(-P.g_cb), f2(cb), f2_fma(f2(-P.g_cr), f2(cr), y)); b = f2_fma(f2(P.b_cb), f2(cb), y);
  r = eotf_k2<TRK>(P.transfer, r); g = eotf_k2<TRK>(P.transfer, g); b = eotf_k2<TRK>(P.transfer, b);
  if (P.nmat) f2_mat3(P.m, r, g, b);
}

__device__ __forceinline__ float lds32(uint32_t a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}

// correctly rounded sRGB8 code of x in [0, 1], in byte 3 of the result (biased-key table, texel.cuh)
template <int R = FER>
__device__ __forceinline__ uint32_t srgb_code_b3(float x, uint32_t enc_lane) {
  const float y = x + ZOS_ENC2_BIAS;
  const int idx = max(__float_as_int(x), ZOS_ENC2_LOW);
  const uint32_t a = (((__float_as_uint(y) >> 16) - (uint32_t)ZOS_ENC2_K0) * (R * 4u)) + enc_lane;
  uint32_t e;
  asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(a));
  return e + (uint32_t)idx;
}

template <bool BILINEAR, bool SRGB_DST, int TRK>
__global__ void __launch_bounds__(THREADS, 4) k_frame_fast(const __grid_constant__ FrameParams P, const __grid_constant__ TensorMaps M) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ TileGeo geo[2];
  // taps of a fully covered tile are separable: one entry per destination column and per destination row, worked
  // out once per tile by 64 threads (the coordinates of a row are the same for all 32 lanes of a warp)
  __shared__ __align__(16) uint4 ctab[TILE], rtab[TILE];  // {offset of tap a, offset of tap b, weight, -}
  const uint32_t cstep = P.nv12 ? 2 : 1;
  const uint32_t ybox = (uint32_t)P.box_w * P.box_h, cbox = (uint32_t)P.cbox_w * P.cbox_h * cstep;
  const uint32_t ybox_al = (ybox + 127) & ~127u, cbox_al = (cbox + 127) & ~127u;
  const uint32_t nchroma = P.nv12 ? 1 : 2;
  const uint32_t stage_bytes = ybox_al + nchroma * cbox_al;
  const uint32_t conv_base = smem_u32(dyn + 2 * (size_t)stage_bytes);
  uint32_t* enc = reinterpret_cast<uint32_t*>(dyn + 2 * (size_t)stage_bytes + 3 * (size_t)P.plane_bytes);
  if (SRGB_DST) {
    for (int i = threadIdx.x; i < ZOS_ENC2_N * FER; i += THREADS) enc[i] = g_tables.srgb_enc2[i / FER];
  }
  const uint32_t enc_lane = smem_u32(enc) + (threadIdx.x & (FER - 1)) * 4u;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();  // the barriers exist before anybody (thread 0 included) arms or polls them
  const CUtensorMap* const m0 = &M.m0;
  const CUtensorMap* const m1 = &M.m1;
  const CUtensorMap* const m2 = &M.m2;
  if (threadIdx.x == 0 && blockIdx.x < P.total_tiles) ZOS_FRAME_ISSUE(blockIdx.x, 0);
  __syncthreads();
  uint32_t phase[2] = {0, 0};
  int s = 0;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const uint32_t cw4 = (uint32_t)P.conv_w * 4u;
  for (uint32_t t = blockIdx.x; t < P.total_tiles; t += gridDim.x) {
    const uint32_t next = t + gridDim.x;
    if (threadIdx.x == 0 && next < P.total_tiles) ZOS_FRAME_ISSUE(next, s ^ 1);
    const TileGeo g = geo[s];
    if (g.any) {
      uint32_t spins = 0;
      while (!mbar_try_wait(&bar[s], phase[s])) {
        if (++spins > (1u << 24)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; break; }
      }
      phase[s] ^= 1;
      if (BILINEAR && g.full && threadIdx.x < 2 * TILE) {
        // the arithmetic of the per-pixel path below, once per column (threads 0..31) and per row (32..63)
        const bool col = threadIdx.x < TILE;
        const int q = col ? (int)threadIdx.x : (int)threadIdx.x - TILE;
        const int kq = (col ? g.x0 - P.tgt[0] : g.y0 - P.tgt[1]) + q;
        const float p = (float)(col ? P.sel[0] : P.sel[1]) + ((float)kq + 0.5f) * (col ? P.rx : P.ry);
        const float f = p - 0.5f, f0 = floorf(f);
        const int i0 = (int)f0, lim = (col ? P.sw : P.sh) - 1, org = col ? g.fx0 : g.fy0;
        const uint32_t unit = col ? 4u : cw4;
        const uint4 e = make_uint4((uint32_t)(min(max(i0, 0), lim) - org) * unit, (uint32_t)(min(max(i0 + 1, 0), lim) - org) * unit,
                                   __float_as_uint(f - f0), 0u);
        if (col) ctab[q] = e; else rtab[q] = e;
      }
      // ---- phase 1: the footprint, one 2x2 luma block per thread
      const uint32_t ybase = smem_u32(dyn + (size_t)s * stage_bytes);
      const uint32_t ubase = ybase + ybox_al, vbase = P.nv12 ? ubase + 1 : ubase + cbox_al;
      const int offx = g.fx0 - g.bx, offy = g.fy0 - g.by;  // both even
      const int cbw = P.conv_w >> 1, cbh = P.conv_h >> 1;
      // (2x2 blocks are numbered row-major and dealt to the threads round robin: all lanes busy)
      for (uint32_t blk = threadIdx.x; blk < (uint32_t)(cbw * cbh); blk += THREADS) {
        const int byi = (int)fastdiv(blk, P.div_cbw), bxi = (int)blk - byi * cbw;
        const uint32_t ya = ybase + (uint32_t)((offy + 2 * byi) * P.box_w + offx + 2 * bxi);
        const uint32_t ca = (uint32_t)(((offy >> 1) + byi) * P.cbox_w + (offx >> 1) + bxi) * cstep;
        uint32_t y01, y23, u8, v8;
        asm("ld.shared.u16 %0, [%1];" : "=r"(y01) : "r"(ya));
        asm("ld.shared.u16 %0, [%1];" : "=r"(y23) : "r"(ya + (uint32_t)P.box_w));
        asm("ld.shared.u8 %0, [%1];" : "=r"(u8) : "r"(ubase + ca));
        asm("ld.shared.u8 %0, [%1];" : "=r"(v8) : "r"(vbase + ca));
        const float cb = ((float)u8 - 128.0f) * P.csc, cr = ((float)v8 - 128.0f) * P.csc;
        const uint32_t o = conv_base + (uint32_t)((2 * byi) * P.conv_w + 2 * bxi) * 4u;
        F2 r0, g0, b0, r1, g1, b1;  // the block's upper and lower pair of texels
        convert_rgb2<TRK>(P, y01, cb, cr, r0, g0, b0);
        convert_rgb2<TRK>(P, y23, cb, cr, r1, g1, b1);
        sts64(o, r0); sts64(o + P.plane_bytes, g0); sts64(o + 2u * P.plane_bytes, b0);
        sts64(o + cw4, r1); sts64(o + cw4 + P.plane_bytes, g1); sts64(o + cw4 + 2u * P.plane_bytes, b1);
      }
      __syncthreads();
    }
    // ---- phase 2: sample, pack
    const int i = g.x0 + lx;
    if (BILINEAR && g.full) {
      // straight path: no coverage tests, no `below`; a pixel is 12 loads at column + row offsets, 3 lerps, 3 encodes
      const uint4 ct = ctab[lx];
      const float ax = __uint_as_float(ct.z);
      const uint32_t ca = conv_base + ct.x, cb = conv_base + ct.y;
      uint8_t* dp = P.dst + ((uint64_t)g.frame * P.dst_bstride + (uint64_t)(g.y0 + ly) * P.dst_pitch + (uint64_t)i * 4u);
      const uint64_t dstep = 8u * P.dst_pitch;
#pragma unroll
      for (int k = 0; k < TILE / 8; k += 2) {  // two of the thread's four rows at a time: the lerps are packed
        const uint4 rt0 = rtab[ly + 8 * k], rt1 = rtab[ly + 8 * k + 8];
        const F2 ay = f2(__uint_as_float(rt0.z), __uint_as_float(rt1.z)), axx = f2(ax);
        const uint32_t a00 = ca + rt0.x, a10 = cb + rt0.x, a01 = ca + rt0.y, a11 = cb + rt0.y;
        const uint32_t c00 = ca + rt1.x, c10 = cb + rt1.x, c01 = ca + rt1.y, c11 = cb + rt1.y;
        F2 r, gg, b;
#define ZOS_TAP(dst_, off_) { const F2 p00 = f2(lds32(a00 + (off_)), lds32(c00 + (off_))), p10 = f2(lds32(a10 + (off_)), lds32(c10 + (off_))); \
                              const F2 p01 = f2(lds32(a01 + (off_)), lds32(c01 + (off_))), p11 = f2(lds32(a11 + (off_)), lds32(c11 + (off_))); \
                              const F2 top = f2_fma(axx, f2_sub(p10, p00), p00), bot = f2_fma(axx, f2_sub(p11, p01), p01); dst_ = f2_fma(ay, f2_sub(bot, top), top); }
        ZOS_TAP(r, 0u) ZOS_TAP(gg, P.plane_bytes) ZOS_TAP(b, 2u * P.plane_bytes)
#undef ZOS_TAP
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float rr = fminf(fmaxf(h ? f2_hi(r) : f2_lo(r), 0.0f), 1.0f), g1 = fminf(fmaxf(h ? f2_hi(gg) : f2_lo(gg), 0.0f), 1.0f),
                      bb = fminf(fmaxf(h ? f2_hi(b) : f2_lo(b), 0.0f), 1.0f);
          uint32_t t1, t2;
          if (SRGB_DST) {
            t1 = __byte_perm(srgb_code_b3(rr, enc_lane), srgb_code_b3(g1, enc_lane), 0x0073);
            t2 = __byte_perm(srgb_code_b3(bb, enc_lane), 0xffu, 0x0043);
          } else {
            t1 = __byte_perm(__float_as_uint(rr * 255.0f + 8388608.0f), __float_as_uint(g1 * 255.0f + 8388608.0f), 0x0040);
            t2 = __byte_perm(__float_as_uint(bb * 255.0f + 8388608.0f), 0xffu, 0x0040);
          }
          __stcs(reinterpret_cast<uint32_t*>(dp + (uint64_t)(k + h) * dstep), __byte_perm(t1, t2, P.spack));
        }
      }
    } else if (i < P.dw) {
      const int kx = i - P.tgt[0];
      const bool col_in = g.any && kx >= 0 && kx < P.tgt[2];
      uint32_t xa4 = 0, xb4 = 0;  // byte offsets of the horizontal taps inside a footprint row
      float ax = 0.0f;
      if (col_in) {
        int xa, xb;
        if (!BILINEAR) {
          xa = xb = min(max(P.sel[0] + rect_index(kx, P.sel[2], P.tgt[2]), 0), P.sw - 1);
        } else {
          const float px = (float)P.sel[0] + ((float)kx + 0.5f) * P.rx;
          const float fx = px - 0.5f, x0f = floorf(fx);
          ax = fx - x0f;
          const int x0 = (int)x0f;
          xb = min(max(x0 + 1, 0), P.sw - 1);
          xa = min(max(x0, 0), P.sw - 1);
        }
        xa4 = (uint32_t)(xa - g.fx0) * 4u; xb4 = (uint32_t)(xb - g.fx0) * 4u;
      }
      const uint64_t off0 = (uint64_t)g.frame * P.dst_bstride + (uint64_t)(g.y0 + ly) * P.dst_pitch + (uint64_t)i * 4u;
      const uint64_t boff0 = (uint64_t)g.frame * P.below_bstride + (uint64_t)(g.y0 + ly) * P.below_pitch + (uint64_t)i * 4u;
#pragma unroll
      for (int k = 0; k < TILE / 8; k++) {
        const int j = g.y0 + ly + 8 * k;
        if (j < P.dh) {
          const int ky = j - P.tgt[1];
          const bool covered = col_in && ky >= 0 && ky < P.tgt[3];
          uint32_t word = P.clear_word;
          if (covered) {
            float r, gg, b;
            if (!BILINEAR) {
              const int w = min(max(P.sel[1] + rect_index(ky, P.sel[3], P.tgt[3]), 0), P.sh - 1) - g.fy0;
              const uint32_t a = conv_base + (uint32_t)w * cw4 + xa4;
              r = lds32(a); gg = lds32(a + P.plane_bytes); b = lds32(a + 2u * P.plane_bytes);
            } else {
              const float py = (float)P.sel[1] + ((float)ky + 0.5f) * P.ry;
              const float fy = py - 0.5f, y0f = floorf(fy);
              const float ay = fy - y0f;
              const int y0 = (int)y0f;
              const int yb = min(max(y0 + 1, 0), P.sh - 1) - g.fy0, ya = min(max(y0, 0), P.sh - 1) - g.fy0;
              const uint32_t ra = conv_base + (uint32_t)ya * cw4, rb = conv_base + (uint32_t)yb * cw4;
              const uint32_t a00 = ra + xa4, a10 = ra + xb4, a01 = rb + xa4, a11 = rb + xb4;
#define ZOS_TAP(dst_, off_) { const float p00 = lds32(a00 + (off_)), p10 = lds32(a10 + (off_)), p01 = lds32(a01 + (off_)), p11 = lds32(a11 + (off_)); \
                              const float top = fmaf(ax, p10 - p00, p00), bot = fmaf(ax, p11 - p01, p01); dst_ = fmaf(ay, bot - top, top); }
              ZOS_TAP(r, 0u) ZOS_TAP(gg, P.plane_bytes) ZOS_TAP(b, 2u * P.plane_bytes)
#undef ZOS_TAP
            }
            // alpha of the frame is exactly 1: source-over == the frame's colour, alpha code 255
            r = fminf(fmaxf(r, 0.0f), 1.0f); gg = fminf(fmaxf(gg, 0.0f), 1.0f); b = fminf(fmaxf(b, 0.0f), 1.0f);
            uint32_t t1, t2;
            if (SRGB_DST) {
              t1 = __byte_perm(srgb_code_b3(r, enc_lane), srgb_code_b3(gg, enc_lane), 0x0073);
              t2 = __byte_perm(srgb_code_b3(b, enc_lane), 0xffu, 0x0043);
            } else {
              t1 = __byte_perm(__float_as_uint(r * 255.0f + 8388608.0f), __float_as_uint(gg * 255.0f + 8388608.0f), 0x0040);
              t2 = __byte_perm(__float_as_uint(b * 255.0f + 8388608.0f), 0xffu, 0x0040);
            }
            word = __byte_perm(t1, t2, P.spack);
          } else if (P.has_below) {
            word = __ldcs(reinterpret_cast<const uint32_t*>(P.below + boff0 + (uint64_t)(8 * k) * P.below_pitch));
          }
          __stcs(reinterpret_cast<uint32_t*>(P.dst + off0 + (uint64_t)(8 * k) * P.dst_pitch), word);
        }
      }
    }
    __syncthreads();  // conv and stage s are free again; geo[s ^ 1] (written by thread 0 above) is visible
    s ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// k_frame_spec (round 2): k_frame_fast with the two phases on DIFFERENT warps.  In k_frame_fast every warp converts, waits at a
// CTA barrier, samples, waits again; 676 conversion blocks over 256 threads are 2.64 passes, so two of eight warps idle through
// the third, and ncu has 1.7 warps per issue stalled at the barrier with 70 % of the issue slots used
// (profiles/r02_c4_fused_kernel_packed.txt).  Here 8 conversion warps and 4 sampling warps (the phases are 2 : 1 by
// instruction count) pass footprint tiles through TWO buffers with full / empty mbarriers: conversion of tile j + 1 runs while
// tile j is sampled, nobody waits for the slowest warp of the other phase, and a warp that finished its share of a tile goes
// on to the next one.  The first lane of the first sampling warp also is the TMA producer: when the footprint of tile j is
// complete its YUV stage is free, so it works out the geometry of tile j + 2 and issues the loads (a thirteenth warp for that role
// lowers the register cap from 80 to 72: measured 5 % slower).  Same arithmetic, same
// bytes as k_frame_fast (the phase bodies are the same code).
#ifndef ZOS_SPEC_CW
#define ZOS_SPEC_CW 8
#define ZOS_SPEC_SW 4
#define ZOS_SPEC_CTAS 2
#endif
constexpr int SPEC_CW = ZOS_SPEC_CW, SPEC_SW = ZOS_SPEC_SW, SPEC_CTAS = ZOS_SPEC_CTAS;  // conversion / sampling warps of a CTA, CTAs per SM
constexpr int SPEC_C = SPEC_CW * 32, SPEC_THREADS = (SPEC_CW + SPEC_SW) * 32;
#ifndef ZOS_SPEC_FER
#define ZOS_SPEC_FER 2
#endif
constexpr int SFER = ZOS_SPEC_FER;  // copies of the encoder bucket table (8 copies measured the same as 2: the look-ups of the sampling warps are not what binds)
// Tile height of this kernel.  Two footprints of a 32 x 32 tile are 85 KB of shared memory per CTA: two CTAs = 24 warps per SM.  Lower
// tiles fit three CTAs but measured slower (config 4: 32 rows x 2 CTAs 0.131 of the HBM figure, 24 rows x 3 CTAs 0.117, 24 x 2 0.120,
// 16 x 3 0.106): the per-tile work that does not shrink with the tile (geometry and loads by one thread, tap tables, hand-shakes)
// and the conversions of the footprint's margin weigh more than the occupancy buys.
#ifndef ZOS_SPEC_TH
#define ZOS_SPEC_TH 32
#endif
constexpr int SPEC_TH = ZOS_SPEC_TH;
static_assert(SPEC_TH % (2 * SPEC_SW) == 0 && SPEC_TH <= TILE, "a sampling thread owns an even number of rows");

#define ZOS_SPEC_WAIT(BAR, PARITY)                                                                  \
  {                                                                                                 \
    uint32_t spins_ = 0;                                                                            \
    while (!mbar_try_wait_sleep((BAR), (PARITY), 2000u)) {                                          \
      if (++spins_ > (1u << 22)) { if (P.fault) *reinterpret_cast<volatile int*>(P.fault) = 1; break; } \
    }                                                                                               \
  }

template <bool BILINEAR, bool SRGB_DST, int TRK>
__global__ void __launch_bounds__(SPEC_THREADS, SPEC_CTAS) k_frame_spec(const __grid_constant__ FrameParams P, const __grid_constant__ TensorMaps M) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ __align__(8) uint64_t bar_yuv[2], bar_full[2], bar_empty[2];
  __shared__ TileGeo geo[4];
  __shared__ __align__(16) uint4 ctab[2][TILE], rtab[2][TILE];  // per footprint buffer: {offset of tap a, offset of tap b, weight, -}
  const uint32_t cstep = P.nv12 ? 2 : 1;
  const uint32_t ybox = (uint32_t)P.box_w * P.box_h, cbox = (uint32_t)P.cbox_w * P.cbox_h * cstep;
  const uint32_t ybox_al = (ybox + 127) & ~127u, cbox_al = (cbox + 127) & ~127u;
  const uint32_t nchroma = P.nv12 ? 1 : 2;
  const uint32_t stage_bytes = ybox_al + nchroma * cbox_al;
  const uint32_t conv_bytes = 3u * P.plane_bytes;
  uint32_t* enc = reinterpret_cast<uint32_t*>(dyn + 2 * (size_t)stage_bytes + 2 * (size_t)conv_bytes);
  if (SRGB_DST) {
    for (int i = threadIdx.x; i < ZOS_ENC2_N * SFER; i += SPEC_THREADS) enc[i] = g_tables.srgb_enc2[i / SFER];
  }
  const uint32_t enc_lane = smem_u32(enc) + (threadIdx.x & (SFER - 1)) * 4u;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 2; k++) { mbar_init(&bar_yuv[k], 1); mbar_init(&bar_full[k], SPEC_CW); mbar_init(&bar_empty[k], SPEC_SW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const CUtensorMap* const m0 = &M.m0;
  const CUtensorMap* const m1 = &M.m1;
  const CUtensorMap* const m2 = &M.m2;
  const uint32_t ntiles = blockIdx.x < P.total_tiles ? (P.total_tiles - 1u - blockIdx.x) / gridDim.x + 1u : 0u;  // tiles of this CTA: blockIdx.x + j * gridDim.x
  // geometry + loads of tile J into YUV stage J & 1; a tile the frame does not touch completes its barrier without a load
#define ZOS_SPEC_ISSUE(J)                                                                                   \
  do {                                                                                                      \
    TileGeo g_;                                                                                             \
    tile_geometry(P, blockIdx.x + (J) * gridDim.x, g_);                                                     \
    geo[(J) & 3u] = g_;                                                                                     \
    uint64_t* bar_ = &bar_yuv[(J) & 1u];                                                                    \
    if (g_.any) {                                                                                           \
      uint8_t* base_ = dyn + (size_t)((J) & 1u) * stage_bytes;                                              \
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                                        \
      mbar_expect_tx(bar_, ybox + nchroma * cbox);                                                          \
      tma_load_3d(base_, m0, g_.bx, g_.by, g_.frame, bar_);                                                 \
      tma_load_3d(base_ + ybox_al, m1, g_.bx >> 1, g_.by >> 1, g_.frame, bar_);                             \
      if (nchroma == 2) tma_load_3d(base_ + ybox_al + cbox_al, m2, g_.bx >> 1, g_.by >> 1, g_.frame, bar_); \
    } else {                                                                                                \
      mbar_arrive(bar_);                                                                                    \
    }                                                                                                       \
  } while (0)
  const uint32_t cw4 = (uint32_t)P.conv_w * 4u;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < SPEC_C) {
    // ================= conversion warps: YUV stage j & 1 -> footprint buffer j & 1
    const int tC = (int)threadIdx.x;
    for (uint32_t j = 0; j < ntiles; j++) {
      const uint32_t b = j & 1u, s = j & 1u;
      ZOS_SPEC_WAIT(&bar_yuv[s], (j >> 1) & 1u)                       // geometry published, planes landed
      if (j >= 2) ZOS_SPEC_WAIT(&bar_empty[b], ((j - 2u) >> 1) & 1u)  // the sampling warps are done with this buffer (tile j - 2)
      const TileGeo g = geo[j & 3u];
      const uint32_t conv_base = smem_u32(dyn + 2 * (size_t)stage_bytes + (size_t)b * conv_bytes);
      if (g.any) {
      if (BILINEAR && g.full && tC < TILE + SPEC_TH) {
        // the arithmetic of the per-pixel path below, once per column (threads 0..31) and per row (32..63)
        const bool col = tC < TILE;
        const int q = col ? tC : tC - TILE;
        const int kq = (col ? g.x0 - P.tgt[0] : g.y0 - P.tgt[1]) + q;
        const float p = (float)(col ? P.sel[0] : P.sel[1]) + ((float)kq + 0.5f) * (col ? P.rx : P.ry);
        const float f = p - 0.5f, f0 = floorf(f);
        const int i0 = (int)f0, lim = (col ? P.sw : P.sh) - 1, org = col ? g.fx0 : g.fy0;
        const uint32_t unit = col ? 4u : cw4;
        const uint4 e = make_uint4((uint32_t)(min(max(i0, 0), lim) - org) * unit, (uint32_t)(min(max(i0 + 1, 0), lim) - org) * unit,
                                   __float_as_uint(f - f0), 0u);
        if (col) ctab[b][q] = e; else rtab[b][q] = e;
      }
      const uint32_t ybase = smem_u32(dyn + (size_t)s * stage_bytes);
      const uint32_t ubase = ybase + ybox_al, vbase = P.nv12 ? ubase + 1 : ubase + cbox_al;
      const int offx = g.fx0 - g.bx, offy = g.fy0 - g.by;  // both even
      const int cbw = P.conv_w >> 1, cbh = P.conv_h >> 1;
      // (2x2 blocks are numbered row-major and dealt to the threads round robin: all lanes busy)
      for (uint32_t blk = (uint32_t)tC; blk < (uint32_t)(cbw * cbh); blk += SPEC_C) {
        const int byi = (int)fastdiv(blk, P.div_cbw), bxi = (int)blk - byi * cbw;
        const uint32_t ya = ybase + (uint32_t)((offy + 2 * byi) * P.box_w + offx + 2 * bxi);
        const uint32_t ca = (uint32_t)(((offy >> 1) + byi) * P.cbox_w + (offx >> 1) + bxi) * cstep;
        uint32_t y01, y23, u8, v8;
        asm("ld.shared.u16 %0, [%1];" : "=r"(y01) : "r"(ya));
        asm("ld.shared.u16 %0, [%1];" : "=r"(y23) : "r"(ya + (uint32_t)P.box_w));
        asm("ld.shared.u8 %0, [%1];" : "=r"(u8) : "r"(ubase + ca));
        asm("ld.shared.u8 %0, [%1];" : "=r"(v8) : "r"(vbase + ca));
        const float cb = ((float)u8 - 128.0f) * P.csc, cr = ((float)v8 - 128.0f) * P.csc;
        const uint32_t o = conv_base + (uint32_t)((2 * byi) * P.conv_w + 2 * bxi) * 4u;
        F2 r0, g0, b0, r1, g1, b1;  // the block's upper and lower pair of texels
        convert_rgb2<TRK>(P, y01, cb, cr, r0, g0, b0);
        convert_rgb2<TRK>(P, y23, cb, cr, r1, g1, b1);
        sts64(o, r0); sts64(o + P.plane_bytes, g0); sts64(o + 2u * P.plane_bytes, b0);
        sts64(o + cw4, r1); sts64(o + cw4 + P.plane_bytes, g1); sts64(o + cw4 + 2u * P.plane_bytes, b1);
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[b]);
    }
  } else {
    // ================= sampling warps (+ the TMA producer): footprint buffer j & 1 -> destination tile
    const int tS = (int)threadIdx.x - SPEC_C;
    const int lx = tS & 31, ly = tS >> 5;
    if (tS == 0) {
      if (ntiles > 0) ZOS_SPEC_ISSUE(0u);
      if (ntiles > 1) ZOS_SPEC_ISSUE(1u);
    }
    for (uint32_t j = 0; j < ntiles; j++) {
      const uint32_t b = j & 1u;
      ZOS_SPEC_WAIT(&bar_full[b], (j >> 1) & 1u)  // every conversion warp finished tile j: its footprint is complete, its YUV stage free
      if (tS == 0 && j + 2u < ntiles) ZOS_SPEC_ISSUE(j + 2u);
      const TileGeo g = geo[j & 3u];
      const uint32_t conv_base = smem_u32(dyn + 2 * (size_t)stage_bytes + (size_t)b * conv_bytes);
    const int i = g.x0 + lx;
    if (BILINEAR && g.full) {
      // straight path: no coverage tests, no `below`; a pixel is 12 loads at column + row offsets, 3 lerps, 3 encodes
      const uint4 ct = ctab[b][lx];
      const float ax = __uint_as_float(ct.z);
      const uint32_t ca = conv_base + ct.x, cb = conv_base + ct.y;
      uint8_t* dp = P.dst + ((uint64_t)g.frame * P.dst_bstride + (uint64_t)(g.y0 + ly) * P.dst_pitch + (uint64_t)i * 4u);
      const uint64_t dstep = (uint64_t)SPEC_SW * P.dst_pitch;
#pragma unroll
      for (int k = 0; k < SPEC_TH / SPEC_SW; k += 2) {  // two of the thread's rows at a time: the lerps are packed
        const uint4 rt0 = rtab[b][ly + SPEC_SW * k], rt1 = rtab[b][ly + SPEC_SW * (k + 1)];
        const F2 ay = f2(__uint_as_float(rt0.z), __uint_as_float(rt1.z)), axx = f2(ax);
        const uint32_t a00 = ca + rt0.x, a10 = cb + rt0.x, a01 = ca + rt0.y, a11 = cb + rt0.y;
        const uint32_t c00 = ca + rt1.x, c10 = cb + rt1.x, c01 = ca + rt1.y, c11 = cb + rt1.y;
        F2 r, gg, b;
#define ZOS_TAP(dst_, off_) { const F2 p00 = f2(lds32(a00 + (off_)), lds32(c00 + (off_))), p10 = f2(lds32(a10 + (off_)), lds32(c10 + (off_))); \
                              const F2 p01 = f2(lds32(a01 + (off_)), lds32(c01 + (off_))), p11 = f2(lds32(a11 + (off_)), lds32(c11 + (off_))); \
                              const F2 top = f2_fma(axx, f2_sub(p10, p00), p00), bot = f2_fma(axx, f2_sub(p11, p01), p01); dst_ = f2_fma(ay, f2_sub(bot, top), top); }
        ZOS_TAP(r, 0u) ZOS_TAP(gg, P.plane_bytes) ZOS_TAP(b, 2u * P.plane_bytes)
#undef ZOS_TAP
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const float rr = fminf(fmaxf(h ? f2_hi(r) : f2_lo(r), 0.0f), 1.0f), g1 = fminf(fmaxf(h ? f2_hi(gg) : f2_lo(gg), 0.0f), 1.0f),
                      bb = fminf(fmaxf(h ? f2_hi(b) : f2_lo(b), 0.0f), 1.0f);
          uint32_t t1, t2;
          if (SRGB_DST) {
            t1 = __byte_perm(srgb_code_b3<SFER>(rr, enc_lane), srgb_code_b3<SFER>(g1, enc_lane), 0x0073);
            t2 = __byte_perm(srgb_code_b3<SFER>(bb, enc_lane), 0xffu, 0x0043);
          } else {
            t1 = __byte_perm(__float_as_uint(rr * 255.0f + 8388608.0f), __float_as_uint(g1 * 255.0f + 8388608.0f), 0x0040);
            t2 = __byte_perm(__float_as_uint(bb * 255.0f + 8388608.0f), 0xffu, 0x0040);
          }
          __stcs(reinterpret_cast<uint32_t*>(dp + (uint64_t)(k + h) * dstep), __byte_perm(t1, t2, P.spack));
        }
      }
    } else if (i < P.dw) {
      const int kx = i - P.tgt[0];
      const bool col_in = g.any && kx >= 0 && kx < P.tgt[2];
      uint32_t xa4 = 0, xb4 = 0;  // byte offsets of the horizontal taps inside a footprint row
      float ax = 0.0f;
      if (col_in) {
        int xa, xb;
        if (!BILINEAR) {
          xa = xb = min(max(P.sel[0] + rect_index(kx, P.sel[2], P.tgt[2]), 0), P.sw - 1);
        } else {
          const float px = (float)P.sel[0] + ((float)kx + 0.5f) * P.rx;
          const float fx = px - 0.5f, x0f = floorf(fx);
          ax = fx - x0f;
          const int x0 = (int)x0f;
          xb = min(max(x0 + 1, 0), P.sw - 1);
          xa = min(max(x0, 0), P.sw - 1);
        }
        xa4 = (uint32_t)(xa - g.fx0) * 4u; xb4 = (uint32_t)(xb - g.fx0) * 4u;
      }
      const uint64_t off0 = (uint64_t)g.frame * P.dst_bstride + (uint64_t)(g.y0 + ly) * P.dst_pitch + (uint64_t)i * 4u;
      const uint64_t boff0 = (uint64_t)g.frame * P.below_bstride + (uint64_t)(g.y0 + ly) * P.below_pitch + (uint64_t)i * 4u;
#pragma unroll
      for (int k = 0; k < SPEC_TH / SPEC_SW; k++) {
        const int j = g.y0 + ly + SPEC_SW * k;
        if (j < P.dh) {
          const int ky = j - P.tgt[1];
          const bool covered = col_in && ky >= 0 && ky < P.tgt[3];
          uint32_t word = P.clear_word;
          if (covered) {
            float r, gg, b;
            if (!BILINEAR) {
              const int w = min(max(P.sel[1] + rect_index(ky, P.sel[3], P.tgt[3]), 0), P.sh - 1) - g.fy0;
              const uint32_t a = conv_base + (uint32_t)w * cw4 + xa4;
              r = lds32(a); gg = lds32(a + P.plane_bytes); b = lds32(a + 2u * P.plane_bytes);
            } else {
              const float py = (float)P.sel[1] + ((float)ky + 0.5f) * P.ry;
              const float fy = py - 0.5f, y0f = floorf(fy);
              const float ay = fy - y0f;
              const int y0 = (int)y0f;
              const int yb = min(max(y0 + 1, 0), P.sh - 1) - g.fy0, ya = min(max(y0, 0), P.sh - 1) - g.fy0;
              const uint32_t ra = conv_base + (uint32_t)ya * cw4, rb = conv_base + (uint32_t)yb * cw4;
              const uint32_t a00 = ra + xa4, a10 = ra + xb4, a01 = rb + xa4, a11 = rb + xb4;
#define ZOS_TAP(dst_, off_) { const float p00 = lds32(a00 + (off_)), p10 = lds32(a10 + (off_)), p01 = lds32(a01 + (off_)), p11 = lds32(a11 + (off_)); \
                              const float top = fmaf(ax, p10 - p00, p00), bot = fmaf(ax, p11 - p01, p01); dst_ = fmaf(ay, bot - top, top); }
              ZOS_TAP(r, 0u) ZOS_TAP(gg, P.plane_bytes) ZOS_TAP(b, 2u * P.plane_bytes)
#undef ZOS_TAP
            }
            // alpha of the frame is exactly 1: source-over == the frame's colour, alpha code 255
            r = fminf(fmaxf(r, 0.0f), 1.0f); gg = fminf(fmaxf(gg, 0.0f), 1.0f); b = fminf(fmaxf(b, 0.0f), 1.0f);
            uint32_t t1, t2;
            if (SRGB_DST) {
              t1 = __byte_perm(srgb_code_b3<SFER>(r, enc_lane), srgb_code_b3<SFER>(gg, enc_lane), 0x0073);
              t2 = __byte_perm(srgb_code_b3<SFER>(b, enc_lane), 0xffu, 0x0043);
            } else {
              t1 = __byte_perm(__float_as_uint(r * 255.0f + 8388608.0f), __float_as_uint(gg * 255.0f + 8388608.0f), 0x0040);
              t2 = __byte_perm(__float_as_uint(b * 255.0f + 8388608.0f), 0xffu, 0x0040);
            }
            word = __byte_perm(t1, t2, P.spack);
          } else if (P.has_below) {
            word = __ldcs(reinterpret_cast<const uint32_t*>(P.below + boff0 + (uint64_t)(SPEC_SW * k) * P.below_pitch));
          }
          __stcs(reinterpret_cast<uint32_t*>(P.dst + off0 + (uint64_t)(SPEC_SW * k) * P.dst_pitch), word);
        }
      }
    }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_empty[b]);
    }
  }
#undef ZOS_SPEC_ISSUE
}
#undef ZOS_SPEC_WAIT

bool native8(const DevImage& im) {
  return im.block == ZOS_BLOCK_PIXEL && im.bpp == 4 && (im.fmt.storage == ZOS_STORAGE_SRGB8 || im.fmt.storage == ZOS_STORAGE_UNORM8) &&
         ((uintptr_t)im.p0 % 4) == 0 && (im.pitch % 4) == 0 && (im.bstride % 4) == 0;
}
}  // namespace

// The dedicated kernel serves: planar 4:2:0 source with nearest chroma, axis-aligned placement (RECT / SCALE),
// at most one matrix step on the source side, overwrite or source-over onto native 8-bit images.
bool frame_pipeline_eligible(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst, const zos_compose_params& cp) {
  if (ctx->flags & ZOS_CTX_NO_FAST_PATHS) return false;
  if (!cp.use_tma) return false;
  if (above.block == ZOS_BLOCK_PIXEL || above.chroma_filter != 0) return false;
  if (cp.map != ZOS_MAP_RECT && cp.map != ZOS_MAP_SCALE) return false;
  if (cp.blend != ZOS_BLEND_OVERWRITE && cp.blend != ZOS_BLEND_SRC_OVER) return false;
  if (cp.n_dst_steps != 0 || cp.n_src_steps > 1 || (cp.n_src_steps == 1 && cp.src_steps[0].kind != ZOS_STEP_MATRIX)) return false;
  if (cp.dst_origin[0] || cp.dst_origin[1] || cp.src_origin[0] || cp.src_origin[1] || cp.src_full[0] || cp.src_full[1]) return false;
  if (!native8(dst) || (below && !native8(*below))) return false;
  if ((above.w & 1) || (above.h & 1)) return false;  // 2x2 blocks: even frame sizes (all video formats)
  return true;
}

zos_status launch_frame_pipeline(zos_ctx* ctx, const DevImage* below, const DevImage& above, const DevImage& dst,
                                 const zos_compose_params& cp, uint32_t batch, bool* handled) {
  *handled = false;
  FrameParams P;
  memset(&P, 0, sizeof P);
  P.fault = ctx->fault_dev;
  P.sw = above.w; P.sh = above.h; P.nv12 = above.block == ZOS_BLOCK_YUV420_NV12;
  P.yoff = above.yoff; P.ysc = above.ysc; P.csc = above.csc; P.r_cr = above.r_cr; P.b_cb = above.b_cb; P.g_cr = above.g_cr; P.g_cb = above.g_cb;
  P.transfer = above.fmt.transfer;
  P.nmat = (int32_t)cp.n_src_steps;
  if (P.nmat) memcpy(P.m, cp.src_steps[0].m, sizeof P.m);
  for (int k = 0; k < 4; k++) { P.sel[k] = cp.sel[k]; P.tgt[k] = cp.tgt[k]; }
  if (cp.map == ZOS_MAP_SCALE) {
    P.sel[0] = P.sel[1] = 0; P.sel[2] = above.w; P.sel[3] = above.h;
    P.tgt[0] = P.tgt[1] = 0; P.tgt[2] = dst.w; P.tgt[3] = dst.h;
  }
  P.rx = (float)P.sel[2] / (float)(P.tgt[2] > 0 ? P.tgt[2] : 1);
  P.ry = (float)P.sel[3] / (float)(P.tgt[3] > 0 ? P.tgt[3] : 1);
  P.sampling = cp.sampling; P.blend = cp.blend;
  P.has_below = below != nullptr;
  if (below) { P.below = below->p0; P.below_pitch = below->pitch; P.below_bstride = below->bstride; P.below_fmt = below->fmt; }
  P.dst = dst.p0; P.dst_pitch = dst.pitch; P.dst_bstride = dst.bstride; P.dst_fmt = dst.fmt; P.dw = dst.w; P.dh = dst.h;
  P.tile_h = TILE;
  P.tiles_x = (dst.w + TILE - 1) / TILE; P.tiles_y = (dst.h + TILE - 1) / TILE;
  uint64_t total = (uint64_t)P.tiles_x * P.tiles_y * batch;
  if (total == 0 || total >= (1ull << 31)) return ZOS_OK;
  P.total_tiles = (uint32_t)total;
  P.div_tx = make_fastdiv(P.tiles_x); P.div_ty = make_fastdiv(P.tiles_y);
  // footprint of a 32-pixel run: taps from floor(min - 0.5) (rounded down to even) to floor(max - 0.5) + 1
  P.conv_w = ((int)ceilf((TILE - 1) * P.rx) + 6) & ~1;
  P.conv_h = ((int)ceilf((TILE - 1) * P.ry) + 6) & ~1;
  if (P.conv_w > 64 || P.conv_h > 96) return ZOS_OK;  // strong minification: the general kernel
  P.box_w = (P.conv_w + 31 + 31) & ~31;
  P.box_h = P.conv_h;
  P.cbox_w = P.box_w / 2; P.cbox_h = P.box_h / 2;
  const int cstep = P.nv12 ? 2 : 1;
  size_t ybox_al = ((size_t)P.box_w * P.box_h + 127) & ~(size_t)127, cbox_al = ((size_t)P.cbox_w * P.cbox_h * cstep + 127) & ~(size_t)127;
  size_t stage = ybox_al + (P.nv12 ? 1 : 2) * cbox_al;
  size_t smem = 2 * stage + (size_t)P.conv_w * P.conv_h * 16;
  if (smem > 160 * 1024) return ZOS_OK;
  TensorMaps M;
  memset(&M, 0, sizeof M);
  const uint64_t cw = (above.w + 1) / 2, ch = (above.h + 1) / 2;
  bool ok = make_map(ctx, &M.m0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p0, above.w, above.h, above.pitch, batch, above.bstride, P.box_w, P.box_h);
  if (ok && P.nv12) ok = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
  if (ok && !P.nv12)
    ok = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h) &&
         make_map(ctx, &M.m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p2, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
  if (!ok) return ZOS_OK;
  // the compile-time specialised kernel: known EOTF class, background (if any) in the destination's texel
  const uint32_t tr = above.fmt.transfer;
  const int trk = (tr == ZOS_TRANSFER_BT709 || tr == ZOS_TRANSFER_BT2020_10BIT || tr == ZOS_TRANSFER_BT2020_12BIT) ? 0 : tr == ZOS_TRANSFER_LINEAR ? 1 : 2;
  const bool same_texel = !below || (below->fmt.storage == dst.fmt.storage && below->fmt.parts == dst.fmt.parts);
  const bool rgba_like = dst.fmt.parts == ZOS_PARTS_RGBA || dst.fmt.parts == ZOS_PARTS_BGRA;
  if (trk != 2 && same_texel && rgba_like) {
    const bool srgb = dst.fmt.storage == ZOS_STORAGE_SRGB8;
    const bool bgra = dst.fmt.parts == ZOS_PARTS_BGRA;
    P.clear_word = bgra ? 0xff0000ffu : 0xffff0000u;  // (0, 0, 1, 1): both codecs map 0 -> 0 and 1 -> 255
    P.spack = bgra ? 0x5014u : 0x5410u;
    // exact tap span of 32 destination pixels: floor(min - 0.5) rounded down to even ... floor(max - 0.5) + 1
    P.conv_w = ((int)ceilf((TILE - 1) * P.rx) + 5) & ~1;
    P.conv_h = ((int)ceilf((TILE - 1) * P.ry) + 5) & ~1;
    P.box_w = (P.conv_w + 31 + 31) & ~31;
    P.box_h = P.conv_h;
    P.cbox_w = P.box_w / 2; P.cbox_h = P.box_h / 2;
    ybox_al = ((size_t)P.box_w * P.box_h + 127) & ~(size_t)127; cbox_al = ((size_t)P.cbox_w * P.cbox_h * cstep + 127) & ~(size_t)127;
    stage = ybox_al + (P.nv12 ? 1 : 2) * cbox_al;
    P.plane_bytes = (uint32_t)(P.conv_w * P.conv_h * 4);
    P.div_cbw = make_fastdiv((uint32_t)(P.conv_w / 2));
    const size_t fsmem = 2 * stage + 3 * (size_t)P.plane_bytes + (srgb ? (size_t)ZOS_ENC2_N * FER * 4 : 0);
    bool ok2 = make_map(ctx, &M.m0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p0, above.w, above.h, above.pitch, batch, above.bstride, P.box_w, P.box_h);
    if (ok2 && P.nv12) ok2 = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
    if (ok2 && !P.nv12)
      ok2 = make_map(ctx, &M.m1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p1, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h) &&
            make_map(ctx, &M.m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p2, cw, ch, above.cpitch, batch, above.cbstride, P.cbox_w, P.cbox_h);
    // ---- k_frame_spec: its own tile height, hence its own footprint, boxes and tensor maps
    if (!(ctx->flags & ZOS_CTX_FRAME_FAST_ONLY)) {
      FrameParams S = P;
      TensorMaps MS;
      memset(&MS, 0, sizeof MS);
      S.tile_h = SPEC_TH;
      S.tiles_y = (dst.h + SPEC_TH - 1) / SPEC_TH;
      const uint64_t stotal = (uint64_t)S.tiles_x * S.tiles_y * batch;
      S.div_ty = make_fastdiv(S.tiles_y);
      S.conv_h = ((int)ceilf((SPEC_TH - 1) * P.ry) + 5) & ~1;
      S.box_h = S.conv_h;
      S.cbox_h = S.box_h / 2;
      const size_t sy = ((size_t)S.box_w * S.box_h + 127) & ~(size_t)127, sc = ((size_t)S.cbox_w * S.cbox_h * cstep + 127) & ~(size_t)127;
      const size_t sstage = sy + (P.nv12 ? 1 : 2) * sc;
      S.plane_bytes = (uint32_t)(S.conv_w * S.conv_h * 4);
      const size_t ssmem = 2 * sstage + 6 * (size_t)S.plane_bytes + (srgb ? (size_t)ZOS_ENC2_N * SFER * 4 : 0);  // two footprint buffers
      bool oks = stotal > 0 && stotal < (1ull << 32) && ssmem <= 110 * 1024 &&
                 make_map(ctx, &MS.m0, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p0, above.w, above.h, above.pitch, batch, above.bstride, S.box_w, S.box_h);
      if (oks && P.nv12) oks = make_map(ctx, &MS.m1, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, above.p1, cw, ch, above.cpitch, batch, above.cbstride, S.cbox_w, S.cbox_h);
      if (oks && !P.nv12)
        oks = make_map(ctx, &MS.m1, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p1, cw, ch, above.cpitch, batch, above.cbstride, S.cbox_w, S.cbox_h) &&
              make_map(ctx, &MS.m2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, above.p2, cw, ch, above.cpitch, batch, above.cbstride, S.cbox_w, S.cbox_h);
      if (oks) {
        S.total_tiles = (uint32_t)stotal;
        int per_sm = (int)((227 * 1024) / (ssmem + 1024 + 2560));
        per_sm = per_sm < 1 ? 1 : (per_sm > SPEC_CTAS ? SPEC_CTAS : per_sm);
        const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
        const int grid = (int)(stotal < cap ? stotal : cap);
        const bool bil = cp.sampling != ZOS_SAMPLE_NEAREST;
#define ZOS_FS(B, S_, T)                                                                                         \
  do {                                                                                                            \
    ensure_dyn_smem(ctx, k_frame_spec<B, S_, T>, 110 * 1024);                                                       \
    k_frame_spec<B, S_, T><<<grid, SPEC_THREADS, ssmem, ctx->stream>>>(S, MS);                                      \
  } while (0)
        if (trk == 0) { if (bil) { if (srgb) ZOS_FS(true, true, 0); else ZOS_FS(true, false, 0); } else { if (srgb) ZOS_FS(false, true, 0); else ZOS_FS(false, false, 0); } }
        else { if (bil) { if (srgb) ZOS_FS(true, true, 1); else ZOS_FS(true, false, 1); } else { if (srgb) ZOS_FS(false, true, 1); else ZOS_FS(false, false, 1); } }
#undef ZOS_FS
        ctx->launches++;
        *handled = true;
        return check_cuda(ctx, cudaGetLastError(), "k_frame_spec launch");
      }
    }
    if (ok2 && fsmem <= 160 * 1024) {
      int per_sm = (int)((226 * 1024) / (fsmem + 1024 + 256));
      per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
      const uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
      const int grid = (int)(total < cap ? total : cap);
      const bool bil = cp.sampling != ZOS_SAMPLE_NEAREST;
#define ZOS_FF(B, S, T)                                                                                          \
  do {                                                                                                            \
    ensure_dyn_smem(ctx, k_frame_fast<B, S, T>, 160 * 1024);                                                        \
    k_frame_fast<B, S, T><<<grid, THREADS, fsmem, ctx->stream>>>(P, M);                                            \
  } while (0)
      if (trk == 0) { if (bil) { if (srgb) ZOS_FF(true, true, 0); else ZOS_FF(true, false, 0); } else { if (srgb) ZOS_FF(false, true, 0); else ZOS_FF(false, false, 0); } }
      else { if (bil) { if (srgb) ZOS_FF(true, true, 1); else ZOS_FF(true, false, 1); } else { if (srgb) ZOS_FF(false, true, 1); else ZOS_FF(false, false, 1); } }
#undef ZOS_FF
      ctx->launches++;
      *handled = true;
      return check_cuda(ctx, cudaGetLastError(), "k_frame_fast launch");
    }
    return ZOS_OK;  // (maps were rebuilt for the smaller box: leave this launch to the general kernel)
  }
  ensure_dyn_smem(ctx, k_frame_pipeline, 160 * 1024);
  int per_sm = (int)((220 * 1024) / (smem + sizeof(Tables) + 2048));
  per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);
  uint64_t cap = (uint64_t)ctx->sm_count * per_sm;
  int grid = (int)(total < cap ? total : cap);
  k_frame_pipeline<<<grid, THREADS, smem, ctx->stream>>>(P, M);
  ctx->launches++;
  *handled = true;
  return check_cuda(ctx, cudaGetLastError(), "k_frame_pipeline launch");
}

}  // namespace zos
