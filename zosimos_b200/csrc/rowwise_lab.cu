// rowwise_lab.cu -- BASELINE config 1 as one lean kernel: a native 8-bit image converted to a Lab-type
// colour held in an 8-bit staged register and back,
//
//   input(sRGB RGBA8) -> color_convert(Oklab | SrLab2, Texel{UInt8x4, LchA | LabA}) -> color_convert(sRGB RGBA8)
//
// i.e. the fused step chain [LAB_ENC, REQUANT(UInt8x4 staged), LAB_DEC] between two native 8-bit texels
// (command.rs:986-1107, oklab.frag:34-64, srlab2.frag:36-120; the register's quantisation is
// stage.frag's encode + decode, program.rs:1475-1478,1531-1532).  The generic kernel interprets
// texel formats and steps per pixel (~650 instructions); here everything is compile time and every
// 8-bit -> float map is a lane-private shared-memory table:
//
//   * source colour decode: one byte permute + one LDS per channel (rowwise_lut.cu's table);
//   * alpha never meets the colour maths: alpha_out = A[alpha_in], a 256-entry map built at kernel
//     start by running the generic arithmetic (decode, f16 store, truncating pack, unpack, f16, encode);
//   * the register: f16 rounding, [Lab -> LCh: sqrt, atan2], clamp, truncation to codes, then tables:
//     f16(k / 255), k / 255 and cos / sin of the hue a code stands for (built by the generic decode code);
//   * sRGB8 encode through the bucket table (texel.cuh), no transcendental.
//
// The colour steps themselves are the shared functions of colorops.cuh: results equal the generic
// kernel's bit for bit (tests compare both, and both with the oracle).
#include "colorops.cuh"
#include "f32x2.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_rowwise_lab)

namespace {
constexpr int LAB_THREADS = 1024;
static_assert(ZOS_ENC2_N <= LAB_THREADS, "one thread per encoder bucket in the table fill");
constexpr uint32_t DEC_BYTES = 256u * 256u;              // [code][0..31] colour decode, [code][32..63] alpha map
constexpr uint32_t Q_BYTES = 256u * 256u;                // [code][lane & 15] float4 {f16(code/255), code/255, cos(hue(code)), sin(hue(code))}
constexpr uint32_t ENC_BYTES = (uint32_t)ZOS_ENC2_N * 128u;  // biased-key bucket table (texel.cuh), one private copy per lane
constexpr uint32_t ENC_SHIFT = 16 - 7;
constexpr uint32_t ENC_MASK = 0x3ffu * 128u;
constexpr uint32_t ENC_VOFF = (ZOS_ENC2_K0 & 0x3ff) * 128u;

struct LabParams {
  FastParams F;  // below = the source
  zos_step enc, dec;
  int32_t src_srgb;
};

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

struct Ctx {
  uint32_t dec, q, enc;     // shared addresses (enc minus ENC_VOFF)
  uint32_t lane4;           // (lane & 31) * 4 with zero upper bytes
  uint32_t lane_q;          // (lane & 15) * 16
  uint32_t sr, sg, sb, sa, spack;
};

__device__ __forceinline__ uint32_t srgb_code_b3(float x, const Ctx& c) {  // x in [0, 1]; code in byte 3
  const float y = x + ZOS_ENC2_BIAS;
  const int idx = max(__float_as_int(x), ZOS_ENC2_LOW);
  uint32_t a;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a) : "r"(__float_as_uint(y) >> ENC_SHIFT), "r"(ENC_MASK), "r"(c.lane4));
  return lds_u32(a + c.enc) + (uint32_t)idx;
}

// the staged UInt8x4 register (stage.frag encode then decode): f16 attachment, [Lab -> LCh], clamp,
// truncating pack; unorm decode, [LCh -> Lab], f16 texture.  Everything after the truncation is a table.
__device__ __forceinline__ uint32_t code8(float x) { return (uint32_t)(clamp01(x) * 255.0f); }
template <bool LCH>
__device__ __forceinline__ void requant8(float4& v, const Ctx& c) {
  float4 t = make_float4(f16r(v.x), f16r(v.y), f16r(v.z), 1.0f);
  if (LCH) transfer_encode(ZOS_TRANSFER_LABLCH, t);  // C = sqrt(a^2 + b^2), h = atan2(b, a) / 2pi + 0.5
  const uint32_t qx = c.q + code8(t.x) * 256u + c.lane_q, qy = c.q + code8(t.y) * 256u + c.lane_q, qz_ = c.q + code8(t.z) * 256u + c.lane_q;
  v.x = lds_f32(qx);
  if (LCH) {
    const float C = lds_f32(qy + 4u), cs = lds_f32(qz_ + 8u), sn = lds_f32(qz_ + 12u);
    v.y = f16r(C * cs); v.z = f16r(C * sn);
  } else {
    v.y = lds_f32(qy); v.z = lds_f32(qz_);
  }
}

template <int LAB, bool SRGB_DST, bool LCH>
__device__ __forceinline__ uint32_t pixel(const LabParams& P, uint32_t w, const Ctx& c) {
  float4 v;
  v.x = lds_f32(__byte_perm(w, c.lane4, c.sr) + c.dec);
  v.y = lds_f32(__byte_perm(w, c.lane4, c.sg) + c.dec);
  v.z = lds_f32(__byte_perm(w, c.lane4, c.sb) + c.dec);
  v.w = 1.0f;
  const uint32_t acode = lds_u32(__byte_perm(w, c.lane4, c.sa) + c.dec + 128u);
  if (LAB == 0) oklab_enc(P.enc, v); else srlab2_enc(P.enc, v);
  requant8<LCH>(v, c);
  if (LAB == 0) oklab_dec(P.dec, v); else srlab2_dec(P.dec, v);  // ends with clamp01 on the colour
  uint32_t t1, t2;
  if (SRGB_DST) {
    t1 = __byte_perm(srgb_code_b3(v.x, c), srgb_code_b3(v.y, c), 0x0073);
    t2 = __byte_perm(srgb_code_b3(v.z, c), acode, 0x0043);
  } else {
    t1 = __byte_perm(__float_as_uint(v.x * 255.0f + 8388608.0f), __float_as_uint(v.y * 255.0f + 8388608.0f), 0x0040);
    t2 = __byte_perm(__float_as_uint(v.z * 255.0f + 8388608.0f), acode, 0x0040);
  }
  return __byte_perm(t1, t2, c.spack);
}

// ---- two pixels at once for the Oklab chain (round 2).  The kernel is bound by instruction issue (90 % busy, 186
// instructions per pixel, FMA pipe 1/3 busy: profiles/r01_c1_lab_kernel.txt) and about half of those instructions are
// IEEE multiplies, adds and fmas: on the packed f32x2 instructions of sm_100a (f32x2.cuh) a pair of pixels pays one issue
// slot for each.  Operation for operation the code of oklab_enc / requant8 / oklab_dec above and in colorops.cuh /
// texel.cuh (results stay bit-identical to the generic kernel: tests); what is not a plain multiply-add -- SFU calls,
// comparisons, sign transfers, the division by 360, float -> int -- stays scalar, and so does a sum whose only producer is
// a product the scalar code rounds separately (ptxas would contract the packed pair, f32x2.cuh).
__device__ __forceinline__ F2 f16r2(F2 v) {  // f16r on both halves: one packed conversion down, two up
  const float2 f = __half22float2(__floats2half2_rn(f2_lo(v), f2_hi(v)));
  return f2(f.x, f.y);
}
__device__ __forceinline__ float ex2_approx(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2_approx(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ F2 cbrt_signed2(F2 v) {  // cbrt_signed (colorops.cuh): pow_fast(|v|, 1/3) with the sign of v, 0 for 0
  const float v0 = f2_lo(v), v1 = f2_hi(v);
  const F2 e = f2_mul(f2(1.0f / 3.0f), f2(lg2_approx(fabsf(v0)), lg2_approx(fabsf(v1))));
  const float r0 = ex2_approx(f2_lo(e)), r1 = ex2_approx(f2_hi(e));
  return f2(v0 == 0.0f ? 0.0f : copysignf(r0, v0), v1 == 0.0f ? 0.0f : copysignf(r1, v1));
}
// atan2_fast (texel.cuh) followed by the hue scaling of transfer_encode(LABLCH)
__device__ __forceinline__ F2 hue2(F2 y, F2 x) {
  const float x0 = f2_lo(x), x1 = f2_hi(x), y0 = f2_lo(y), y1 = f2_hi(y);
  const float ax0 = fabsf(x0), ay0 = fabsf(y0), ax1 = fabsf(x1), ay1 = fabsf(y1);
  const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0), mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
  float q0, q1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(mx0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(mx1));
  const F2 tq = f2_mul(f2(mn0, mn1), f2(q0, q1));
  const F2 t = f2(mx0 > 0.0f ? f2_lo(tq) : 0.0f, mx1 > 0.0f ? f2_hi(tq) : 0.0f);
  const F2 s = f2_mul(t, t);
  F2 p = f2(-0.004054565913975239f);
  p = f2_fma(p, s, f2(0.021862953901290894f));
  p = f2_fma(p, s, f2(-0.0559123232960701f));
  p = f2_fma(p, s, f2(0.0964219719171524f));
  p = f2_fma(p, s, f2(-0.1390862911939621f));
  p = f2_fma(p, s, f2(0.19946566224098206f));
  p = f2_fma(p, s, f2(-0.33329859375953674f));
  p = f2_fma(p, s, f2(0.9999993443489075f));
  p = f2_mul(p, t);
  float p0 = f2_lo(p), p1 = f2_hi(p);
  if (ay0 > ax0) p0 = 1.57079632679489662f - p0;
  if (ay1 > ax1) p1 = 1.57079632679489662f - p1;
  if (x0 < 0.0f) p0 = 3.14159265358979324f - p0;
  if (x1 < 0.0f) p1 = 3.14159265358979324f - p1;
  p0 = copysignf(p0, y0); p1 = copysignf(p1, y1);
  const F2 d = f2_mul(f2(p0, p1), f2(180.0f / ZOS_PI_F));
  return f2(f2_lo(d) / 360.0f + 0.5f, f2_hi(d) / 360.0f + 0.5f);
}

template <bool SRGB_DST, bool LCH>
__device__ __forceinline__ void pixel2_oklab(const LabParams& P, uint32_t w0, uint32_t w1, const Ctx& c, uint32_t& o0, uint32_t& o1) {
  F2 x = f2(lds_f32(__byte_perm(w0, c.lane4, c.sr) + c.dec), lds_f32(__byte_perm(w1, c.lane4, c.sr) + c.dec));
  F2 y = f2(lds_f32(__byte_perm(w0, c.lane4, c.sg) + c.dec), lds_f32(__byte_perm(w1, c.lane4, c.sg) + c.dec));
  F2 z = f2(lds_f32(__byte_perm(w0, c.lane4, c.sb) + c.dec), lds_f32(__byte_perm(w1, c.lane4, c.sb) + c.dec));
  const uint32_t acode0 = lds_u32(__byte_perm(w0, c.lane4, c.sa) + c.dec + 128u), acode1 = lds_u32(__byte_perm(w1, c.lane4, c.sa) + c.dec + 128u);
  // oklab_enc
  f2_mat3(P.enc.m, x, y, z);  // (M1 * to_xyz, folded by the entry point: colorops.cuh)
  x = cbrt_signed2(x); y = cbrt_signed2(y); z = cbrt_signed2(z);
  f2_mat3(c_color.ok_m2, x, y, z);
  // requant8: the staged UInt8x4 register
  F2 tx = f16r2(x), ty = f16r2(y), tz = f16r2(z);
  if (LCH) {
    const F2 aa = f2_mul(ty, ty), bb = f2_mul(tz, tz);
    const F2 hue = hue2(tz, ty);
    ty = f2(sqrt_fast(f2_lo(aa) + f2_lo(bb)), sqrt_fast(f2_hi(aa) + f2_hi(bb)));  // (scalar sums of the rounded products)
    tz = hue;
  }
  const uint32_t qx0 = c.q + code8(f2_lo(tx)) * 256u + c.lane_q, qy0 = c.q + code8(f2_lo(ty)) * 256u + c.lane_q, qz0 = c.q + code8(f2_lo(tz)) * 256u + c.lane_q;
  const uint32_t qx1 = c.q + code8(f2_hi(tx)) * 256u + c.lane_q, qy1 = c.q + code8(f2_hi(ty)) * 256u + c.lane_q, qz1 = c.q + code8(f2_hi(tz)) * 256u + c.lane_q;
  x = f2(lds_f32(qx0), lds_f32(qx1));
  if (LCH) {
    const F2 C = f2(lds_f32(qy0 + 4u), lds_f32(qy1 + 4u));
    y = f16r2(f2_mul(C, f2(lds_f32(qz0 + 8u), lds_f32(qz1 + 8u))));
    z = f16r2(f2_mul(C, f2(lds_f32(qz0 + 12u), lds_f32(qz1 + 12u))));
  } else {
    y = f2(lds_f32(qy0), lds_f32(qy1)); z = f2(lds_f32(qz0), lds_f32(qz1));
  }
  // oklab_dec
  f2_mat3(c_color.ok_m2i, x, y, z);
  x = f2_mul(f2_mul(x, x), x); y = f2_mul(f2_mul(y, y), y); z = f2_mul(f2_mul(z, z), z);
  f2_mat3(P.dec.m, x, y, z);  // (from_xyz * M1^-1)
  const float v0[3] = {clamp01(f2_lo(x)), clamp01(f2_lo(y)), clamp01(f2_lo(z))}, v1[3] = {clamp01(f2_hi(x)), clamp01(f2_hi(y)), clamp01(f2_hi(z))};
  uint32_t t1, t2, u1, u2;
  if (SRGB_DST) {
    t1 = __byte_perm(srgb_code_b3(v0[0], c), srgb_code_b3(v0[1], c), 0x0073);
    t2 = __byte_perm(srgb_code_b3(v0[2], c), acode0, 0x0043);
    u1 = __byte_perm(srgb_code_b3(v1[0], c), srgb_code_b3(v1[1], c), 0x0073);
    u2 = __byte_perm(srgb_code_b3(v1[2], c), acode1, 0x0043);
  } else {
    const F2 k = f2(255.0f);
    const F2 qr = f2_mul(f2(v0[0], v1[0]), k), qg = f2_mul(f2(v0[1], v1[1]), k), qb = f2_mul(f2(v0[2], v1[2]), k);
    const float m = 8388608.0f;  // (scalar adds: two roundings)
    t1 = __byte_perm(__float_as_uint(f2_lo(qr) + m), __float_as_uint(f2_lo(qg) + m), 0x0040);
    t2 = __byte_perm(__float_as_uint(f2_lo(qb) + m), acode0, 0x0040);
    u1 = __byte_perm(__float_as_uint(f2_hi(qr) + m), __float_as_uint(f2_hi(qg) + m), 0x0040);
    u2 = __byte_perm(__float_as_uint(f2_hi(qb) + m), acode1, 0x0040);
  }
  o0 = __byte_perm(t1, t2, c.spack);
  o1 = __byte_perm(u1, u2, c.spack);
}

template <int LAB, bool SRGB_DST, bool LCH>
__global__ void __launch_bounds__(LAB_THREADS, 1) k_rowwise_lab(const __grid_constant__ LabParams P) {
  extern __shared__ __align__(256) uint8_t smem[];
  float* dec = reinterpret_cast<float*>(smem);
  float* q = reinterpret_cast<float*>(smem + DEC_BYTES);
  uint32_t* enc = reinterpret_cast<uint32_t*>(smem + DEC_BYTES + Q_BYTES);
  // table fill (see rowwise_lut.cu): 16-byte stores, a warp writes 512 contiguous bytes per instruction
#pragma unroll
  for (int k = 0; k < 256 * 64 / 4 / LAB_THREADS; k++) {
    const int f = k * LAB_THREADS + threadIdx.x, code = f >> 4;
    const float u = g_tables.unorm8[code];
    float v;
    if (f & 8) {
      // alpha through the chain: decode, register (f16, clamp, truncate, unorm, f16), encode (round to nearest)
      const uint32_t kq = (uint32_t)(clamp01(f16r(u)) * 255.0f);
      const float a2 = f16r(g_tables.unorm8[kq]);
      v = __uint_as_float((uint32_t)__float2int_rn(clamp01(a2) * 255.0f));
    } else {
      v = P.src_srgb ? g_tables.srgb_dec[code] : u;
    }
    reinterpret_cast<float4*>(dec)[f] = make_float4(v, v, v, v);
  }
#pragma unroll
  for (int k = 0; k < 256 * 16 / LAB_THREADS; k++) {
    const int f = k * LAB_THREADS + threadIdx.x;
    const float u = g_tables.unorm8[f >> 4];
    float4 c4 = make_float4(0.0f, 1.0f, u, 1.0f);  // the generic decode of hue code f / 16 at chroma 1
    transfer_decode(ZOS_TRANSFER_LABLCH, c4);       // c4.y = cos(hue), c4.z = sin(hue): the very code the generic path runs
    reinterpret_cast<float4*>(q)[f] = make_float4(f16r(u), u, c4.y, c4.z);
  }
  if (SRGB_DST) {
#pragma unroll
    for (int k = 0; k < (ZOS_ENC2_N * 8 + LAB_THREADS - 1) / LAB_THREADS; k++) {
      const int f = k * LAB_THREADS + threadIdx.x;
      if (f < ZOS_ENC2_N * 8) {
        const uint32_t e = g_tables.srgb_enc2[f >> 3];
        reinterpret_cast<uint4*>(enc)[f] = make_uint4(e, e, e, e);
      }
    }
  }
  __syncthreads();

  Ctx c;
  c.dec = (uint32_t)__cvta_generic_to_shared(dec);
  c.q = (uint32_t)__cvta_generic_to_shared(q);
  c.enc = (uint32_t)__cvta_generic_to_shared(enc) - ENC_VOFF;
  c.lane4 = (threadIdx.x & 31u) * 4u;
  c.lane_q = (threadIdx.x & 15u) * 16u;
  const uint32_t kr = P.F.src_bgra ? 2u : 0u, kb = P.F.src_bgra ? 0u : 2u;
  c.sr = 0x7604u | (kr << 4); c.sg = 0x7614u; c.sb = 0x7604u | (kb << 4); c.sa = 0x7634u;
  c.spack = P.F.dst_bgra ? 0x5014u : 0x5410u;

  const uint32_t stride = gridDim.x * LAB_THREADS;
  for (uint32_t idx = blockIdx.x * LAB_THREADS + threadIdx.x; idx < P.F.total_groups; idx += stride) {
    const Loc L = locate<0>(P.F, idx);
    const uint4 rb = __ldcs(reinterpret_cast<const uint4*>(P.F.below + L.ob));
    uint32_t o[4];
    if (LAB == 0) {
      pixel2_oklab<SRGB_DST, LCH>(P, rb.x, rb.y, c, o[0], o[1]);
      pixel2_oklab<SRGB_DST, LCH>(P, rb.z, rb.w, c, o[2], o[3]);
    } else {
      o[0] = pixel<LAB, SRGB_DST, LCH>(P, rb.x, c); o[1] = pixel<LAB, SRGB_DST, LCH>(P, rb.y, c);
      o[2] = pixel<LAB, SRGB_DST, LCH>(P, rb.z, c); o[3] = pixel<LAB, SRGB_DST, LCH>(P, rb.w, c);
    }
    uint8_t* dp = P.F.dst + L.od;
    if (L.npx == 4) {
      __stcs(reinterpret_cast<uint4*>(dp), make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i < L.npx) reinterpret_cast<uint32_t*>(dp)[i] = o[i];
    }
    if (idx + stride < idx) break;  // 32-bit wrap
  }
}

bool native8(const DevImage& im, bool* srgb, bool* bgra) {
  if (im.block != ZOS_BLOCK_PIXEL || im.bpp != 4) return false;
  if (im.fmt.storage != ZOS_STORAGE_SRGB8 && im.fmt.storage != ZOS_STORAGE_UNORM8) return false;
  *srgb = im.fmt.storage == ZOS_STORAGE_SRGB8;
  *bgra = im.fmt.parts == ZOS_PARTS_BGRA;
  return true;
}

template <int LAB, bool SRGB_DST, bool LCH>
cudaError_t launch_one(zos_ctx* ctx, const LabParams& P) {
  auto kern = k_rowwise_lab<LAB, SRGB_DST, LCH>;
  const uint32_t bytes = DEC_BYTES + Q_BYTES + (SRGB_DST ? ENC_BYTES : 0u);
  {
    cudaError_t e = ensure_dyn_smem(ctx, kern, (int)(DEC_BYTES + Q_BYTES + ENC_BYTES));
    if (e != cudaSuccess) return e;
  }
  const uint64_t ctas = ((uint64_t)P.F.total_groups + LAB_THREADS - 1) / LAB_THREADS;
  const int grid = (int)(ctas < (uint64_t)ctx->sm_count ? ctas : (uint64_t)ctx->sm_count);
  kern<<<grid, LAB_THREADS, bytes, ctx->stream>>>(P);
  return cudaGetLastError();
}
}  // namespace

cudaError_t launch_rowwise_lab(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch, bool* handled) {
  *handled = false;
  if (nsteps != 3) return cudaSuccess;
  const bool ok_chain = steps[0].kind == ZOS_STEP_OKLAB_ENC && steps[2].kind == ZOS_STEP_OKLAB_DEC;
  const bool sr_chain = steps[0].kind == ZOS_STEP_SRLAB2_ENC && steps[2].kind == ZOS_STEP_SRLAB2_DEC;
  if (!(ok_chain || sr_chain) || steps[1].kind != ZOS_STEP_REQUANT) return cudaSuccess;
  const zos_texfmt& rf = steps[1].fmt;
  if (rf.storage != ZOS_STORAGE_STAGED || rf.bits != ZOS_BITS_UINT8X4) return cudaSuccess;
  if (rf.transfer != ZOS_TRANSFER_LINEAR && rf.transfer != ZOS_TRANSFER_LABLCH) return cudaSuccess;  // LabA / LchA registers
  if (rf.parts != ZOS_PARTS_LABA && rf.parts != ZOS_PARTS_LCHA && rf.parts != ZOS_PARTS_RGBA) return cudaSuccess;  // pass-through swizzles only
  const bool lch = rf.transfer == ZOS_TRANSFER_LABLCH;
  bool ssrgb, sbgra, dsrgb, dbgra;
  if (!native8(src, &ssrgb, &sbgra) || !native8(dst, &dsrgb, &dbgra)) return cudaSuccess;
  if (src.w != dst.w || src.h != dst.h) return cudaSuccess;
  LabParams P;
  memset(&P, 0, sizeof P);
  P.F.below = src.p0; P.F.below_pitch = src.pitch; P.F.below_bstride = src.bstride;
  P.F.dst = dst.p0; P.F.dst_pitch = dst.pitch; P.F.dst_bstride = dst.bstride;
  P.F.w = dst.w; P.F.h = dst.h; P.F.has_below = 1;
  P.F.src_bgra = sbgra; P.F.dst_bgra = dbgra;
  P.src_srgb = ssrgb;
  P.enc = steps[0]; P.dec = steps[2];
  const uint64_t gpr = (uint64_t)(dst.w + 3) / 4, total = gpr * (uint64_t)dst.h * batch;
  if (total == 0 || total >= (1ull << 32)) return cudaSuccess;
  P.F.groups_per_row = (uint32_t)gpr; P.F.total_groups = (uint32_t)total;
  P.F.div_gpr = make_fastdiv((uint32_t)gpr); P.F.div_h = make_fastdiv((uint32_t)dst.h);
  *handled = true;
  if (ok_chain) {
    if (lch) return dsrgb ? launch_one<0, true, true>(ctx, P) : launch_one<0, false, true>(ctx, P);
    return dsrgb ? launch_one<0, true, false>(ctx, P) : launch_one<0, false, false>(ctx, P);
  }
  if (lch) return dsrgb ? launch_one<1, true, true>(ctx, P) : launch_one<1, false, true>(ctx, P);
  return dsrgb ? launch_one<1, true, false>(ctx, P) : launch_one<1, false, false>(ctx, P);
}

}  // namespace zos
