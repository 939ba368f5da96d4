// rowwise_lut.cu -- the streaming kernel for native 8-bit texels on both sides (Rgba8Unorm[Srgb],
// Bgra8Unorm[Srgb]; lib/zosimos/src/program.rs:794-838): BASELINE config 2 (inscribe / blend of two
// RGBA8 sRGB layers) and the RGBA8 row of config 5.  Bit for bit the results of the generic kernel
// (rowwise.cu) and of the oracle; the work per pixel is cut to ~55 instructions by turning both
// codecs into ONE shared-memory look-up per channel:
//
//   * decode: a 64 KB table with one 256-byte row per code: 32 lane-private copies of the exact
//     sRGB EOTF value followed by 32 copies of code/255.  The address of a look-up is ONE byte
//     permute, (code << 8) | (lane * 4), and no two lanes of a warp ever share a bank;
//   * sRGB encode, correctly rounded, without transcendental: the f32 bit pattern of x in [2^-13, 1]
//     is cut into 1665 buckets (sign/exponent/7 mantissa bits).  No bucket holds more than one of the
//     255 rounding thresholds, so   code = base[bucket] + (low16(x) >= low16(threshold[bucket])).
//     Both facts are folded into one 32-bit entry e = (base << 16) + (0x10000 - t16) - (top16 << 16):
//     (e + bits(x)) >> 16 is the code: one integer max (values below 2^-13 encode to 0), a shift,
//     a mask-or, the look-up and an add per channel.  The table is replicated 16x;
//   * 4 pixels per thread and iteration with 16-byte accesses, the next group's loads issued before
//     the current group is processed; a fully linear addressing mode when all layers share one
//     geometry (the blend workload); one 1024-thread CTA per SM owning ~171 KB of tables.
//
// The arithmetic between decode and encode is exactly the generic kernel's (source-over with one
// reciprocal, mat3_mul's fmaf order), so this is an optimisation of instruction count only.
#include "colorops.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

constexpr int LUT_THREADS = 1024;
static_assert(ZOS_ENC2_N <= LUT_THREADS, "one thread per encoder bucket in the table fill");
constexpr uint32_t DEC_BYTES = 256u * 256u;                    // [code][0..31] sRGB EOTF, [code][32..63] code/255
constexpr uint32_t ENC_BYTES = (uint32_t)ZOS_ENC2_N * 128u;    // [bucket][0..31]: one private copy per lane
constexpr uint32_t ENC_SHIFT = 16 - 7;                         // bits(y) >> 16 is the key, rows are 128 bytes apart
constexpr uint32_t ENC_MASK = 0x3ffu * 128u;                   // the low 10 bits of the key are unique over the table
constexpr uint32_t ENC_VOFF = (ZOS_ENC2_K0 & 0x3ff) * 128u;    // masked offset of the first row (32 KB: the decode table sits below)

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_rowwise_lut)

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

struct LutCtx {
  uint32_t dec;      // shared address of the decode table
  uint32_t enc;      // shared address of the encode table minus ENC_VOFF
  uint32_t lane4;    // (lane & 31) * 4, upper bytes zero: byte 0 of every decode address
  uint32_t sr, sg, sb, sa;  // byte-permute selectors building (code << 8) | lane4 for R, G, B, A of a source word
  uint32_t spack;           // final selector of the destination word (RGBA / BGRA)
};

struct Px { float r, g, b, a; };

template <int SK, bool WITH_ALPHA = true>
__device__ __forceinline__ Px decode8(uint32_t w, const LutCtx& c) {
  constexpr uint32_t col = SK == K_SRGB8 ? 0u : 128u;
  Px p;
  p.r = lds_f32(__byte_perm(w, c.lane4, c.sr) + c.dec + col);
  p.g = lds_f32(__byte_perm(w, c.lane4, c.sg) + c.dec + col);
  p.b = lds_f32(__byte_perm(w, c.lane4, c.sb) + c.dec + col);
  p.a = WITH_ALPHA ? lds_f32(__byte_perm(w, c.lane4, c.sa) + c.dec + 128u) : 1.0f;
  return p;
}

// correctly rounded sRGB8 code of x (0 <= x <= 1 up to one rounding), in byte 3 of the result: the
// biased-key bucket table of texel.cuh, one conflict-free look-up
__device__ __forceinline__ uint32_t srgb_code_b3(float x, const LutCtx& c) {
  const float y = x + ZOS_ENC2_BIAS;
  const int idx = max(__float_as_int(x), ZOS_ENC2_LOW);
  uint32_t a;  // ((bits(y) >> SHIFT) & MASK) | lane column, as ONE logic op
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(a) : "r"(__float_as_uint(y) >> ENC_SHIFT), "r"(ENC_MASK), "r"(c.lane4));
  return lds_u32(a + c.enc) + (uint32_t)idx;
}

// RAW_ALPHA: alpha was not touched between decode and encode (no blend; matrix steps act on colour
// only) and code -> code/255 -> code is the identity, so byte 3 of the source word `w` is the result.
template <int DK, bool CLAMP, bool RAW_ALPHA>
__device__ __forceinline__ uint32_t encode8(const Px& p, const LutCtx& c, uint32_t w) {
  float v[4] = {p.r, p.g, p.b, p.a};
  if (CLAMP) {
#pragma unroll
    for (int i = 0; i < 3; i++) v[i] = fminf(fmaxf(v[i], 0.0f), 1.0f);
  }
  uint32_t ca = w;  // code in byte 3
  if (!RAW_ALPHA) ca = __float_as_uint(v[3] * 255.0f + 8388608.0f);  // code in byte 0 (blended alpha is in [0, 1])
  uint32_t t1, t2;
  if constexpr (DK == K_SRGB8) {
    t1 = __byte_perm(srgb_code_b3(v[0], c), srgb_code_b3(v[1], c), 0x0073);
    t2 = __byte_perm(srgb_code_b3(v[2], c), ca, RAW_ALPHA ? 0x0073 : 0x0043);
  } else {
    t1 = __byte_perm(__float_as_uint(v[0] * 255.0f + 8388608.0f), __float_as_uint(v[1] * 255.0f + 8388608.0f), 0x0040);
    t2 = __byte_perm(__float_as_uint(v[2] * 255.0f + 8388608.0f), ca, RAW_ALPHA ? 0x0070 : 0x0040);
  }
  return __byte_perm(t1, t2, c.spack);
}

// One pixel.  MODE 0: `b` only; 2: `a` over `b` (source-over on straight alpha in linear light, the
// oracle's pd_blend mode 3: identical operation order).
template <int SK, int DK, int MODE, int NMAT>
__device__ __forceinline__ uint32_t pixel8(const FastParams& P, uint32_t b, uint32_t a, const LutCtx& c) {
  Px v = decode8<SK, MODE != 0>(b, c);
  if (MODE == 2) {
    Px s = decode8<SK>(a, c);
    float wbk = v.a * (1.0f - s.a);
    float ao = s.a + wbk;
    // ao is 0 or in [1/255, 1]: SFU reciprocal + one Newton step is the correctly rounded 1/ao there
    // (tested for all alpha pairs).  ao == 0 has a zero numerator: any finite reciprocal gives the
    // oracle's 0, so the guard is a max with a tiny normal number instead of a select.
    float aos = fmaxf(ao, 1e-30f);
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(aos));
    float rcp = fmaf(r0, -fmaf(aos, r0, -1.0f), r0);
    v.r = fmaf(wbk, v.r, s.a * s.r) * rcp;
    v.g = fmaf(wbk, v.g, s.a * s.g) * rcp;
    v.b = fmaf(wbk, v.b, s.a * s.b) * rcp;
    v.a = ao;
  }
#pragma unroll
  for (int k = 0; k < NMAT; k++) {
    float3 t = mat3_mul(P.m[k], v.r, v.g, v.b);
    v.r = t.x; v.g = t.y; v.b = t.z;
  }
  return encode8<DK, (NMAT > 0), MODE == 0>(v, c, b);
}

// LINEAR: every layer has the destination's geometry with rows and frames back to back, so the images
// are plain streams of 16-byte groups (the blend / convert workloads).  Two groups are in flight per
// thread, ping-pong, so that no register rotation is needed.
template <int SK, int DK, int MODE, int NMAT, bool LINEAR>
__global__ void __launch_bounds__(LUT_THREADS, 1) k_rowwise_lut(const __grid_constant__ FastParams P) {
  extern __shared__ __align__(256) uint8_t smem[];
  float* dec = reinterpret_cast<float*>(smem);
  uint32_t* enc = reinterpret_cast<uint32_t*>(smem + DEC_BYTES);
  // Table fill, the fixed cost of a launch (it dominates for small images): 16-byte stores, a warp writes 512
  // contiguous bytes per instruction, all (L2-resident) source loads of a thread are independent.
#pragma unroll
  for (int k = 0; k < 256 * 64 / 4 / LUT_THREADS; k++) {
    const int f = k * LUT_THREADS + threadIdx.x;                 // float4 index: row f / 16, floats [4 (f % 16), +4)
    const float v = (f & 8) ? g_tables.unorm8[f >> 4] : g_tables.srgb_dec[f >> 4];
    reinterpret_cast<float4*>(dec)[f] = make_float4(v, v, v, v);
  }
  if (DK == K_SRGB8) {
#pragma unroll
    for (int k = 0; k < (ZOS_ENC2_N * 8 + LUT_THREADS - 1) / LUT_THREADS; k++) {
      const int f = k * LUT_THREADS + threadIdx.x;               // uint4 index: bucket f / 8
      if (f < ZOS_ENC2_N * 8) {
        const uint32_t e = g_tables.srgb_enc2[f >> 3];
        reinterpret_cast<uint4*>(enc)[f] = make_uint4(e, e, e, e);
      }
    }
  }
  __syncthreads();

  LutCtx c;
  c.dec = (uint32_t)__cvta_generic_to_shared(dec);
  c.enc = (uint32_t)__cvta_generic_to_shared(enc) - ENC_VOFF;
  c.lane4 = (threadIdx.x & 31u) * 4u;
  // selector nibbles: byte 0 <- lane4.byte0 (4), byte 1 <- word byte k, bytes 2, 3 <- lane4's zero bytes (6, 7)
  const uint32_t kr = P.src_bgra ? 2u : 0u, kb = P.src_bgra ? 0u : 2u;
  c.sr = 0x7604u | (kr << 4); c.sg = 0x7614u; c.sb = 0x7604u | (kb << 4); c.sa = 0x7634u;
  c.spack = P.dst_bgra ? 0x5014u : 0x5410u;

  const uint32_t stride = gridDim.x * LUT_THREADS;
  uint32_t idx = blockIdx.x * LUT_THREADS + threadIdx.x;
  if (idx >= P.total_groups) return;

  if constexpr (LINEAR) {
    const uint4* pb = reinterpret_cast<const uint4*>(P.below) + idx;
    const uint4* pa = reinterpret_cast<const uint4*>(MODE ? P.above : P.below) + idx;
    uint4* pd = reinterpret_cast<uint4*>(P.dst) + idx;
    // Software pipeline: the loads of the next D - 1 groups of a thread are in flight while one group is computed.
    // With one group ahead (the first version) 57 % of the stall samples sat on the first use of a loaded word
    // (profiles/r01_c5_rgba8_kernel_v2.txt): one group of arithmetic on 8 warps per scheduler is ~0.7 us, less
    // than a DRAM round trip under load.
    constexpr int D = MODE ? 3 : 4;
    const uint32_t mine = (P.total_groups - 1u - idx) / stride + 1u;  // groups of this thread: idx + k * stride, k < mine
    uint4 b[D], a[D];
#pragma unroll
    for (int j = 0; j < D; j++) b[j] = a[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < D - 1; j++)
      if ((uint32_t)j < mine) { b[j] = __ldcs(pb + (size_t)j * stride); if (MODE) a[j] = __ldcs(pa + (size_t)j * stride); }
#define ZOS_LUT_GROUP(B, A, OUT)                                                        \
    __stcs(OUT, make_uint4(pixel8<SK, DK, MODE, NMAT>(P, B.x, A.x, c), pixel8<SK, DK, MODE, NMAT>(P, B.y, A.y, c), \
                           pixel8<SK, DK, MODE, NMAT>(P, B.z, A.z, c), pixel8<SK, DK, MODE, NMAT>(P, B.w, A.w, c)))
    for (uint32_t k = 0;; k += D) {
#pragma unroll
      for (int j = 0; j < D; j++) {
        if (k + j >= mine) return;
        if (k + j + (D - 1) < mine) {
          b[(j + D - 1) % D] = __ldcs(pb + (size_t)(j + D - 1) * stride);
          if (MODE) a[(j + D - 1) % D] = __ldcs(pa + (size_t)(j + D - 1) * stride);
        }
        ZOS_LUT_GROUP(b[j], a[j], pd + (size_t)j * stride);
      }
      pb += (size_t)D * stride; pa += (size_t)D * stride; pd += (size_t)D * stride;
    }
#undef ZOS_LUT_GROUP
    return;
  } else {
  Loc L = locate<MODE>(P, idx);
  uint4 rb = __ldcs(reinterpret_cast<const uint4*>(P.below + L.ob));
  uint4 ra = make_uint4(0, 0, 0, 0);
  if (MODE != 0 && L.ncov > 0) ra = __ldcs(reinterpret_cast<const uint4*>(P.above + L.oa));
  for (;;) {
    const uint32_t nidx = idx + stride;
    const bool more = nidx < P.total_groups && nidx >= stride;
    Loc NL = L;
    uint4 nb = rb, na = ra;
    if (more) {  // the next group's loads are in flight while this one is computed
      NL = locate<MODE>(P, nidx);
      nb = __ldcs(reinterpret_cast<const uint4*>(P.below + NL.ob));
      na = make_uint4(0, 0, 0, 0);
      if (MODE != 0 && NL.ncov > 0) na = __ldcs(reinterpret_cast<const uint4*>(P.above + NL.oa));
    }
    const uint32_t wb[4] = {rb.x, rb.y, rb.z, rb.w}, wa[4] = {ra.x, ra.y, ra.z, ra.w};
    uint32_t o[4];
    if (MODE == 0 || L.ncov == 4) {
#pragma unroll
      for (int i = 0; i < 4; i++) o[i] = pixel8<SK, DK, MODE, NMAT>(P, wb[i], wa[i], c);
    } else {
      // a group outside of / straddling the edge of `above`: covered pixels first, then the rest
#pragma unroll
      for (int i = 0; i < 4; i++)
        o[i] = i < L.ncov ? pixel8<SK, DK, MODE, NMAT>(P, wb[i], wa[i], c) : pixel8<SK, DK, 0, NMAT>(P, wb[i], wa[i], c);
    }
    uint8_t* dp = P.dst + L.od;
    if (L.npx == 4) {
      __stcs(reinterpret_cast<uint4*>(dp), make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i < L.npx) reinterpret_cast<uint32_t*>(dp)[i] = o[i];
    }
    if (!more) break;
    L = NL; rb = nb; ra = na; idx = nidx;
  }
  }
}

template <int SK, int DK, int MODE, int NMAT>
static cudaError_t launch_one(zos_ctx* ctx, const FastParams& P) {
  const uint32_t bytes = DEC_BYTES + (DK == K_SRGB8 ? ENC_BYTES : 0u);
  auto kern = k_rowwise_lut<SK, DK, MODE, NMAT, false>;
  auto kern_lin = k_rowwise_lut<SK, DK, MODE, NMAT, true>;
  {
    cudaError_t e = ensure_dyn_smem(ctx, kern, (int)(DEC_BYTES + ENC_BYTES));
    if (e == cudaSuccess) e = ensure_dyn_smem(ctx, kern_lin, (int)(DEC_BYTES + ENC_BYTES));
    if (e != cudaSuccess) return e;
  }
  const uint64_t ctas = (P.total_groups + LUT_THREADS - 1) / LUT_THREADS;
  const int grid = (int)(ctas < (uint64_t)ctx->sm_count ? ctas : (uint64_t)ctx->sm_count);
  if (P.linear) kern_lin<<<grid, LUT_THREADS, bytes, ctx->stream>>>(P);
  else kern<<<grid, LUT_THREADS, bytes, ctx->stream>>>(P);
  return cudaGetLastError();
}

cudaError_t launch_rowwise_lut(zos_ctx* ctx, FastParams& P, int sk, int dk, int mode, int nmat) {
  // linear addressing: every layer has the destination's geometry, rows and frames back to back
  const uint64_t row = (uint64_t)P.w * 4u, frame = row * P.h;
  const bool one = (uint64_t)P.total_groups == (uint64_t)P.groups_per_row * (uint64_t)P.h;  // a single frame: frame strides are not used
  const bool full = mode == 0 || (P.tx == 0 && P.ty == 0 && P.aw == P.w && P.ah == P.h && P.above_pitch == row && (one || P.above_bstride == frame));
  P.linear = (P.w % 4 == 0) && full && P.below_pitch == row && P.dst_pitch == row && (one || (P.below_bstride == frame && P.dst_bstride == frame));
#define ZOS_LUT(SK_, DK_)                                                 \
  if (sk == SK_ && dk == DK_) {                                           \
    if (mode == 0 && nmat == 0) return launch_one<SK_, DK_, 0, 0>(ctx, P); \
    if (mode == 0 && nmat == 1) return launch_one<SK_, DK_, 0, 1>(ctx, P); \
    if (mode == 0) return launch_one<SK_, DK_, 0, 2>(ctx, P);              \
    if (nmat == 0) return launch_one<SK_, DK_, 2, 0>(ctx, P);              \
    if (nmat == 1) return launch_one<SK_, DK_, 2, 1>(ctx, P);              \
    return launch_one<SK_, DK_, 2, 2>(ctx, P);                             \
  }
  ZOS_LUT(K_SRGB8, K_SRGB8) ZOS_LUT(K_SRGB8, K_UNORM8) ZOS_LUT(K_UNORM8, K_SRGB8) ZOS_LUT(K_UNORM8, K_UNORM8)
#undef ZOS_LUT
  return cudaErrorInvalidValue;
}

}  // namespace zos
