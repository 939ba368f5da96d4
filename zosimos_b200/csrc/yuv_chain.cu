// yuv_chain.cu -- per-pixel chains that END in a planar YUV 4:2:0 image (I420 / NV12), the
// YUV420 -> YUV420 and RGBA -> YUV420 rows of BASELINE config 5.  The reference has no planar texels
// at all (program.rs:794-938 lowers Block::Pixel only); semantics are ours and
// the oracle is zo_encode_yuv420 (oracle/zos_oracle.c): OETF of the colour, Y' = Kr R' + Kg G' + Kb B',
// Cb = (B' - Y') * (1 / (2 (1 - Kb))) with one rounded reciprocal, Cr likewise, range scaling, round to nearest even; one chroma sample
// per 2x2 block = the mean of the block's (up to 4) chroma values.
//
// One thread owns one 2x2 block: 4 source texels in, 4 Y' bytes and one Cb / Cr pair out; a warp
// covers 64 x 2 pixels.  Sources: any pixel texel (generic codec) or planar YUV with nearest chroma.
#include "colorops.cuh"
#include "f32x2.cuh"
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
  const float rcb = 1.0f / (2.0f * (1.0f - D.kb)), rcr = 1.0f / (2.0f * (1.0f - D.kr));
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
        const float cb = (b - yy) * rcb, cr = (r - yy) * rcr;
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

// ---- the same chains with everything known at compile time: planar source with nearest chroma and a
// BT.709-family EOTF, 0..2 matrix steps, and one of three destinations.  Same arithmetic as k_yuv_chain
// (and, for the pixel destinations, as gather.cu / rowwise.cu), a fraction of the instructions.
enum { D_YUV709 = 0, D_SRGB8 = 1, D_UNORM8 = 2 };
constexpr int YER = 8;  // copies of the biased-key sRGB encoder table (texel.cuh)

struct YuvFastParams {
  DevImage src, dst;
  int32_t nmat;
  float m[2][9];
  uint32_t bw, bh, total;
  FastDiv div_bw, div_bh;
  uint32_t spack;
};

__device__ __forceinline__ float eotf709(float v) {  // == yuv_eotf for the BT.709 family
  float lin = v * (1.0f / 4.5f);
  float l2, pw;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"((v + 0.099f) * (1.0f / 1.099f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pw) : "f"(l2 * (1.0f / 0.45f)));
  return v >= 0.0812428582f ? pw : lin;
}
__device__ __forceinline__ uint32_t srgb_code_b3(float x, uint32_t enc_lane) {  // x in [0, 1]; code in byte 3
  const float y = x + ZOS_ENC2_BIAS;
  const int idx = max(__float_as_int(x), ZOS_ENC2_LOW);
  const uint32_t a = (((__float_as_uint(y) >> 16) - (uint32_t)ZOS_ENC2_K0) * (YER * 4u)) + enc_lane;
  uint32_t e;
  asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(a));
  return e + (uint32_t)idx;
}

// Round 2: the block's two pixels of a row go through the arithmetic TOGETHER, on the packed f32x2 instructions of
// sm_100a (f32x2.cuh): the kernel is bound by instruction issue and by the SFU (12 MUFU per pixel for YUV -> YUV), the FMA
// pipe is 39 % busy (profiles/r01_c5_yuv_yuv_kernel.txt), so every packed instruction frees an issue slot.  Same IEEE
// operations in the same order as k_yuv_chain: where that code rounds a product and a sum separately and the product has
// no other user the sum stays scalar (ptxas would contract the packed pair into one FFMA2, see f32x2.cuh).  Integer ->
// float conversions (SFU pipe) are replaced by the exact 2^23 construction on the FMA pipe, 8-bit results leave through
// the same construction, and the two Y' bytes of a row are one 16-bit store.
__device__ __forceinline__ float lg2a(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2a(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ F2 eotf709x2(F2 v) {  // == eotf709 on both halves
  const F2 lin = f2_mul(v, f2(1.0f / 4.5f));
  const F2 arg = f2_mul(f2_add(v, f2(0.099f)), f2(1.0f / 1.099f));
  const F2 e = f2_mul(f2(lg2a(f2_lo(arg)), lg2a(f2_hi(arg))), f2(1.0f / 0.45f));
  const float p0 = ex2a(f2_lo(e)), p1 = ex2a(f2_hi(e));
  return f2(f2_lo(v) >= 0.0812428582f ? p0 : f2_lo(lin), f2_hi(v) >= 0.0812428582f ? p1 : f2_hi(lin));
}
__device__ __forceinline__ F2 oe709x2(F2 v) {  // == oe_bt709 (texel.cuh) on both halves
  const F2 e = f2_mul(f2(lg2a(f2_lo(v)), lg2a(f2_hi(v))), f2(0.45f));   // pow_fast(v, 0.45f) = ex2(0.45 * lg2 v)
  const F2 m = f2_mul(f2(1.099f), f2(ex2a(f2_lo(e)), ex2a(f2_hi(e))));
  const F2 lin = f2_mul(f2(4.5f), v);
  return f2(f2_lo(v) >= 0.018f ? f2_lo(m) - 0.099f : f2_lo(lin), f2_hi(v) >= 0.018f ? f2_hi(m) - 0.099f : f2_hi(lin));  // (scalar subtractions)
}
// the float of byte k of `w`: 0x4b0000cc - 2^23, exact
__device__ __forceinline__ F2 bytes_to_f2(uint32_t w0, uint32_t sel0, uint32_t w1, uint32_t sel1) {
  return f2_sub(f2(__uint_as_float(__byte_perm(w0, 0x4b000000u, sel0)), __uint_as_float(__byte_perm(w1, 0x4b000000u, sel1))), f2(8388608.0f));
}
__device__ __forceinline__ F2 clamp255x2(F2 v) {
  return f2(fminf(fmaxf(f2_lo(v), 0.0f), 255.0f), fminf(fmaxf(f2_hi(v), 0.0f), 255.0f));
}

template <int DST>
__global__ void __launch_bounds__(256) k_yuv_fast(const __grid_constant__ YuvFastParams P) {
  __shared__ uint32_t enc[DST == D_SRGB8 ? ZOS_ENC2_N * YER : 1];
  if (DST == D_SRGB8) {
    for (int i = threadIdx.x; i < ZOS_ENC2_N * YER; i += blockDim.x) enc[i] = g_tables.srgb_enc2[i / YER];
    __syncthreads();
  }
  const uint32_t enc_lane = (uint32_t)__cvta_generic_to_shared(enc) + (threadIdx.x & (YER - 1)) * 4u;
  const DevImage& A = P.src;
  const DevImage& D = P.dst;
  const int cstep_s = A.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1, cstep_d = D.block == ZOS_BLOCK_YUV420_NV12 ? 2 : 1;
  const float kg = 1.0f - D.kr - D.kb;
  const float rcb = 1.0f / (2.0f * (1.0f - D.kb)), rcr = 1.0f / (2.0f * (1.0f - D.kr));
  const float cscale = D.full_range ? 255.0f : 224.0f;
  const float yk = D.full_range ? 255.0f : 219.0f, y0k = D.full_range ? 0.0f : 16.0f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < P.total; idx += stride) {
    const uint32_t rowid = fastdiv(idx, P.div_bw);
    const int ci = (int)(idx - rowid * P.bw);
    const uint32_t frame = fastdiv(rowid, P.div_bh);
    const int cj = (int)(rowid - frame * P.bh);
    const uint64_t sco = frame * A.cbstride + (uint64_t)cj * A.cpitch + (uint64_t)ci * cstep_s;
    // chroma: ((float)code - 128) * csc, both samples at once
    const F2 cc = f2_mul(f2_sub(bytes_to_f2((uint32_t)A.p1[sco], 0x7540, (uint32_t)A.p2[sco], 0x7540), f2(128.0f)), f2(A.csc));
    const float cb_in = f2_lo(cc), cr_in = f2_hi(cc);
    // both pixels of a block row come from one 16-bit load (the frame width is even for video formats; odd tails take bytes)
    const int x0 = 2 * ci, y0 = 2 * cj;
    const bool two_x = x0 + 1 < D.w, two_y = y0 + 1 < D.h;
    float cbs = 0.0f, crs = 0.0f;  // summed in k_yuv_chain's order (0,0), (1,0), (0,1), (1,1): scalar adds
#pragma unroll
    for (int dy = 0; dy < 2; dy++) {
      if (dy == 1 && !two_y) break;
      const uint8_t* yrow = A.p0 + frame * A.bstride + (uint64_t)(y0 + dy) * A.pitch + x0;
      const uint32_t ypair = two_x ? (uint32_t)*reinterpret_cast<const uint16_t*>(yrow) : (uint32_t)*yrow;
      const F2 yy = f2_mul(f2_sub(bytes_to_f2(ypair, 0x7540, ypair, 0x7541), f2(A.yoff)), f2(A.ysc));
      F2 r = f2_fma(f2(A.r_cr), f2(cr_in), yy), g = f2_fma(f2(-A.g_cb), f2(cb_in), f2_fma(f2(-A.g_cr), f2(cr_in), yy)), b = f2_fma(f2(A.b_cb), f2(cb_in), yy);
      r = eotf709x2(r); g = eotf709x2(g); b = eotf709x2(b);
#pragma unroll
      for (int k = 0; k < 2; k++)
        if (k < P.nmat) f2_mat3(P.m[k], r, g, b);
      if (DST == D_YUV709) {
        const F2 er = oe709x2(r), eg = oe709x2(g), eb = oe709x2(b);
        const F2 yo = f2_fma(f2(D.kb), eb, f2_fma(f2(kg), eg, f2_mul(f2(D.kr), er)));
        const F2 cb = f2_mul(f2_sub(eb, yo), f2(rcb)), cr = f2_mul(f2_sub(er, yo), f2(rcr));
        // full range: yo * 255 (a product); limited: fmaf(yo, 219, 16) -- fma(yo, 255, 0) is that product exactly
        const F2 yq = f2_add(clamp255x2(f2_fma(yo, f2(yk), f2(y0k))), f2(8388608.0f));  // low byte = round to nearest even
        uint8_t* drow = D.p0 + frame * D.bstride + (uint64_t)(y0 + dy) * D.pitch + x0;
        if (two_x) *reinterpret_cast<uint16_t*>(drow) = (uint16_t)__byte_perm(__float_as_uint(f2_lo(yq)), __float_as_uint(f2_hi(yq)), 0x0040);
        else *drow = (uint8_t)__float_as_uint(f2_lo(yq));
        cbs += f2_lo(cb); crs += f2_lo(cr);
        if (two_x) { cbs += f2_hi(cb); crs += f2_hi(cr); }
      } else {
        const float v0[3] = {fminf(fmaxf(f2_lo(r), 0.0f), 1.0f), fminf(fmaxf(f2_lo(g), 0.0f), 1.0f), fminf(fmaxf(f2_lo(b), 0.0f), 1.0f)};
        const float v1[3] = {fminf(fmaxf(f2_hi(r), 0.0f), 1.0f), fminf(fmaxf(f2_hi(g), 0.0f), 1.0f), fminf(fmaxf(f2_hi(b), 0.0f), 1.0f)};
        uint32_t words[2];
        if (DST == D_SRGB8) {
          words[0] = __byte_perm(__byte_perm(srgb_code_b3(v0[0], enc_lane), srgb_code_b3(v0[1], enc_lane), 0x0073), __byte_perm(srgb_code_b3(v0[2], enc_lane), 0xffu, 0x0043), P.spack);
          words[1] = __byte_perm(__byte_perm(srgb_code_b3(v1[0], enc_lane), srgb_code_b3(v1[1], enc_lane), 0x0073), __byte_perm(srgb_code_b3(v1[2], enc_lane), 0xffu, 0x0043), P.spack);
        } else {
          const F2 k = f2(255.0f);
          const F2 qr = f2_mul(f2(v0[0], v1[0]), k), qg = f2_mul(f2(v0[1], v1[1]), k), qb = f2_mul(f2(v0[2], v1[2]), k);
          const float m = 8388608.0f;  // (scalar adds: v * 255 and + 2^23 are two roundings)
          words[0] = __byte_perm(__byte_perm(__float_as_uint(f2_lo(qr) + m), __float_as_uint(f2_lo(qg) + m), 0x0040), __byte_perm(__float_as_uint(f2_lo(qb) + m), 0xffu, 0x0040), P.spack);
          words[1] = __byte_perm(__byte_perm(__float_as_uint(f2_hi(qr) + m), __float_as_uint(f2_hi(qg) + m), 0x0040), __byte_perm(__float_as_uint(f2_hi(qb) + m), 0xffu, 0x0040), P.spack);
        }
        uint8_t* drow = D.p0 + frame * D.bstride + (uint64_t)(y0 + dy) * D.pitch + (uint64_t)x0 * 4u;
        if (two_x) __stcs(reinterpret_cast<uint2*>(drow), make_uint2(words[0], words[1]));
        else *reinterpret_cast<uint32_t*>(drow) = words[0];
      }
    }
    if (DST == D_YUV709) {
      const int cnt = (two_x ? 2 : 1) * (two_y ? 2 : 1);
      const float inv = cnt == 4 ? 0.25f : cnt == 2 ? 0.5f : 1.0f;  // == dividing by 1, 2 or 4
      const uint64_t co = frame * D.cbstride + (uint64_t)cj * D.cpitch + (uint64_t)ci * cstep_d;
      const F2 q = f2_add(clamp255x2(f2_fma(f2_mul(f2(cbs, crs), f2(inv)), f2(cscale), f2(128.0f))), f2(8388608.0f));
      D.p1[co] = (uint8_t)__float_as_uint(f2_lo(q));
      D.p2[co] = (uint8_t)__float_as_uint(f2_hi(q));
    }
  }
}

bool bt709_family(uint32_t tr) { return tr == ZOS_TRANSFER_BT709 || tr == ZOS_TRANSFER_BT2020_10BIT || tr == ZOS_TRANSFER_BT2020_12BIT; }
}  // namespace

// Planar source (nearest chroma, BT.709-family EOTF), matrix-only steps, same size -> BT.709-family planar or native
// 8-bit RGBA / BGRA.  *handled stays false when the chain is not of that shape.
zos_status launch_yuv_fast(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch, bool* handled) {
  *handled = false;
  if (ctx->flags & ZOS_CTX_NO_FAST_PATHS) return ZOS_OK;
  if (src.block == ZOS_BLOCK_PIXEL || src.chroma_filter != 0 || !bt709_family(src.fmt.transfer)) return ZOS_OK;
  if (src.w != dst.w || src.h != dst.h || nsteps > 2) return ZOS_OK;
  for (uint32_t i = 0; i < nsteps; i++)
    if (steps[i].kind != ZOS_STEP_MATRIX) return ZOS_OK;
  int kind;
  if (dst.block != ZOS_BLOCK_PIXEL) {
    if (!bt709_family(dst.fmt.transfer)) return ZOS_OK;
    kind = D_YUV709;
  } else {
    if (dst.bpp != 4 || (dst.fmt.parts != ZOS_PARTS_RGBA && dst.fmt.parts != ZOS_PARTS_BGRA)) return ZOS_OK;
    if (dst.fmt.storage == ZOS_STORAGE_SRGB8) kind = D_SRGB8;
    else if (dst.fmt.storage == ZOS_STORAGE_UNORM8) kind = D_UNORM8;
    else return ZOS_OK;
    if (((uintptr_t)dst.p0 % 8) || (dst.pitch % 8) || (dst.bstride % 8)) return ZOS_OK;
  }
  if ((src.pitch % 2) || ((uintptr_t)src.p0 % 2) || (src.bstride % 2)) return ZOS_OK;
  if (kind == D_YUV709 && ((dst.pitch % 2) || ((uintptr_t)dst.p0 % 2) || (dst.bstride % 2))) return ZOS_OK;  // 16-bit Y' stores
  YuvFastParams P;
  memset(&P, 0, sizeof P);
  P.src = src; P.dst = dst;
  P.nmat = (int32_t)nsteps;
  for (uint32_t i = 0; i < nsteps; i++) memcpy(P.m[i], steps[i].m, sizeof(float) * 9);
  P.spack = dst.fmt.parts == ZOS_PARTS_BGRA ? 0x5014u : 0x5410u;
  P.bw = (uint32_t)(dst.w + 1) / 2; P.bh = (uint32_t)(dst.h + 1) / 2;
  const uint64_t total = (uint64_t)P.bw * P.bh * batch;
  if (total == 0 || total >= (1ull << 32)) return ZOS_OK;
  P.total = (uint32_t)total;
  P.div_bw = make_fastdiv(P.bw); P.div_bh = make_fastdiv(P.bh);
  const int grid = grid_for(ctx, total, 256, 32);  // (more CTAs than resident, +10 %: DESIGN.md "grid size of the streaming kernels"; each CTA fills a 5-20 KB table)
  if (kind == D_YUV709) k_yuv_fast<D_YUV709><<<grid, 256, 0, ctx->stream>>>(P);
  else if (kind == D_SRGB8) k_yuv_fast<D_SRGB8><<<grid, 256, 0, ctx->stream>>>(P);
  else k_yuv_fast<D_UNORM8><<<grid, 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  *handled = true;
  return check_cuda(ctx, cudaGetLastError(), "k_yuv_fast launch");
}

zos_status launch_yuv_chain(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch) {
  {
    bool handled = false;
    zos_status st = launch_yuv_fast(ctx, src, dst, steps, nsteps, batch, &handled);
    if (handled || st != ZOS_OK) return st;
  }
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
  const int grid = grid_for(ctx, total, 256, 32);
  k_yuv_chain<<<grid, 256, 0, ctx->stream>>>(P);
  ctx->launches++;
  return check_cuda(ctx, cudaGetLastError(), "k_yuv_chain launch");
}

}  // namespace zos
