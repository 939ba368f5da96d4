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
//     geometry (the blend workload); one 1024-thread CTA per SM owning ~162 KB of tables;
//   * round 2 (the kernel is bound by instruction issue and, over a second of back-to-back launches, by the 1 kW power
//     cap: profiles/r02_bench_all_v1.json): the floating-point part runs on PAIRS of pixels with the packed
//     f32x2 instructions of sm_100a (FFMA2 / FMUL2 / FADD2, f32x2.cuh: same IEEE results, half the issue slots), and
//     the encoder rows are 256 bytes apart so that the address of a look-up is ONE byte permute of the biased value's
//     upper half with the lane offset (was shift + logic op); the decode table lives in the unused upper halves of
//     those rows.
//
// The arithmetic between decode and encode is exactly the generic kernel's (source-over with one
// reciprocal, mat3_mul's fmaf order), so this is an optimisation of instruction count only.
#include "colorops.cuh"
#include "f32x2.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

constexpr int LUT_THREADS = 1024;
static_assert(ZOS_ENC2_N <= LUT_THREADS, "one thread per encoder bucket in the table fill");
// Shared memory: ZOS_ENC2_N rows of 256 bytes.  Row k, bytes [0, 128): the encoder entry of bucket k, one private copy per
// lane.  Bytes [128, 256) of row c (c < 256): the exact sRGB EOTF of code c, one copy per lane; of row 256 + c: c / 255.
constexpr uint32_t ROW = 256u;
static_assert(ZOS_ENC2_N >= 512, "the decode columns need 512 rows");
constexpr uint32_t SMEM_BYTES = (uint32_t)ZOS_ENC2_N * ROW;
constexpr uint32_t DEC_SRGB = 128u, DEC_UNORM = 256u * ROW + 128u;   // byte offsets of the two decode columns
constexpr uint32_t ENC_KEY0 = (uint32_t)ZOS_ENC2_K0 * ROW;           // (bits(y) >> 16) * 256 of the first bucket
// The table's shared address is a COMPILE-TIME constant, so that it costs no instruction: it is the immediate offset of
// every look-up (the compiler kept the base in a uniform register for the blend variant but spent one add per look-up, 6 per
// pixel, in the convert variants).  Dynamic shared memory starts at 0x400 on sm_100 (1 KB is reserved in front of it) when a
// kernel has no static shared memory; the kernel checks that and, if the layout is ever different, writes nothing and raises
// the context's fault word (zos_sync then fails with ZOS_ERR_CUDA) -- it does not trap, the context stays usable.
// (Measured dead end: placing the table at the next multiple of 64 KB instead, so that the base folds into the byte permute,
// needs 63 KB of padding -- the larger carve-out leaves the L1 too small to keep the streaming loads in flight and the
// kernel lost 15 %.)
constexpr uint32_t TABLE_AT = 0x400u;

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

// Streaming 16-byte accesses of the linear path.  ZOS_LUT_LDST selects the cache operators; measured side by side on one box
// (gpurun_out/ab8_*, c2_blend burst / sustained): 0 = .cs both ways 0.918 / 0.828, 1 = ld.global.nc.L1::no_allocate 0.921 / 0.829,
// 2 = ld.global.L1::no_allocate 0.922 / 0.829, 3 = .cg both ways (L2 only) 0.949 / 0.831 -- the kernel's busiest unit is the L1 data
// pipe (look-ups AND global accesses), and .cg takes the global share of it down.  (On a second box: burst 0.903 -> 0.908 for the blend,
// 0.884 -> 0.908 for the RGBA8 conversion, sustained unchanged at 0.78: that chip is power-bound at 1.51 GHz.  The same change made
// k_affine_f16 slower, 0.908 -> 0.878 sustained for the nearest variant, and did nothing for k_rowwise_rgb10: both keep .cs.)
#ifndef ZOS_LUT_LDST
#define ZOS_LUT_LDST 3
#endif
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
#if ZOS_LUT_LDST == 0
  return __ldcs(p);
#elif ZOS_LUT_LDST == 1
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#elif ZOS_LUT_LDST == 2
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) {
#if ZOS_LUT_LDST == 3
  __stcg(p, v);
#else
  __stcs(p, v);
#endif
}

struct LutCtx {
  uint32_t lane4;    // (lane & 31) * 4, upper bytes zero: byte 0 of every look-up address
  uint32_t sr, sg, sb, sa;  // byte-permute selectors building (code << 8) | lane4 for R, G, B, A of a source word
  uint32_t spack;           // final selector of the destination word (RGBA / BGRA)
};

struct Px { float r, g, b, a; };

template <int SK, bool WITH_ALPHA = true>
__device__ __forceinline__ Px decode8(uint32_t w, const LutCtx& c) {
  constexpr uint32_t col = SK == K_SRGB8 ? DEC_SRGB : DEC_UNORM;
  Px p;
  p.r = lds_f32(__byte_perm(w, c.lane4, c.sr) + (TABLE_AT + col));
  p.g = lds_f32(__byte_perm(w, c.lane4, c.sg) + (TABLE_AT + col));
  p.b = lds_f32(__byte_perm(w, c.lane4, c.sb) + (TABLE_AT + col));
  p.a = WITH_ALPHA ? lds_f32(__byte_perm(w, c.lane4, c.sa) + (TABLE_AT + DEC_UNORM)) : 1.0f;
  return p;
}

// correctly rounded sRGB8 code of x (0 <= x <= 1 up to one rounding), in byte 3 of the result: the biased-key bucket table
// of texel.cuh, one conflict-free look-up.  `y` = x + ZOS_ENC2_BIAS (computed for two pixels at once by the caller).
__device__ __forceinline__ uint32_t srgb_code_b3(float x, float y, const LutCtx& c) {
  const int idx = max(__float_as_int(x), ZOS_ENC2_LOW);
  // (bits(y) >> 16) * 256 + lane * 4 as ONE byte permute: byte 0 <- lane4.0, bytes 1, 2 <- y.2, y.3, byte 3 <- lane4.1 (zero)
  return lds_u32(__byte_perm(__float_as_uint(y), c.lane4, 0x5324) + (TABLE_AT - ENC_KEY0)) + (uint32_t)idx;  // (wraps below zero and back)
}

struct Px2 { F2 r, g, b, a; };  // two pixels, channel by channel

// code / 255 of byte 3 of two words, exactly (IEEE division; rowwise_fast.cu's unorm8_exact, verified for all 256 codes),
// without a look-up.  Measured dead end, kept behind ZOS_LUT_ALPHA_ARITHMETIC: two look-ups per pixel less and two
// instructions more lose 2 % both over 10 ms and over 1 s (profiles/r02_lut_kernel_ab.md).
__device__ __forceinline__ F2 alpha_pair(uint32_t w0, uint32_t w1) {
  const F2 c = f2_sub(f2(__uint_as_float(__byte_perm(w0, 0x4b000000u, 0x7543)), __uint_as_float(__byte_perm(w1, 0x4b000000u, 0x7543))), f2(8388608.0f));
  const F2 r = f2(0.003921568859368563f);
  const F2 q = f2_mul(c, r);
  return f2_fma(f2_fma(q, f2(-255.0f), c), r, q);
}

// RAW_ALPHA: alpha was not touched between decode and encode (no blend; matrix steps act on colour
// only) and code -> code/255 -> code is the identity, so byte 3 of the source words is the result.
template <int DK, bool CLAMP, bool RAW_ALPHA>
__device__ __forceinline__ void encode8x2(const Px2& p, const LutCtx& c, uint32_t w0, uint32_t w1, uint32_t& o0, uint32_t& o1) {
  float v0[3] = {f2_lo(p.r), f2_lo(p.g), f2_lo(p.b)}, v1[3] = {f2_hi(p.r), f2_hi(p.g), f2_hi(p.b)};
  if (CLAMP) {
#pragma unroll
    for (int i = 0; i < 3; i++) { v0[i] = fminf(fmaxf(v0[i], 0.0f), 1.0f); v1[i] = fminf(fmaxf(v1[i], 0.0f), 1.0f); }
  }
  uint32_t ca0 = w0, ca1 = w1;  // code in byte 3
  if (!RAW_ALPHA) {             // code in byte 0 (blended alpha is in [0, 1])
    // v * 255.0f + 8388608.0f is TWO roundings: the product packed, the add scalar -- ptxas contracts a packed mul.rn into a
    // packed add.rn that is its only user (FFMA2), explicit rounding modifiers and --fmad=false notwithstanding
    const F2 q = f2_mul(p.a, f2(255.0f));
    ca0 = __float_as_uint(f2_lo(q) + 8388608.0f); ca1 = __float_as_uint(f2_hi(q) + 8388608.0f);
  }
  uint32_t t1, t2, u1, u2;
  if constexpr (DK == K_SRGB8) {
    const F2 bias = f2(ZOS_ENC2_BIAS);
    const F2 yr = f2_add(f2(v0[0], v1[0]), bias), yg = f2_add(f2(v0[1], v1[1]), bias), yb = f2_add(f2(v0[2], v1[2]), bias);
    t1 = __byte_perm(srgb_code_b3(v0[0], f2_lo(yr), c), srgb_code_b3(v0[1], f2_lo(yg), c), 0x0073);
    t2 = __byte_perm(srgb_code_b3(v0[2], f2_lo(yb), c), ca0, RAW_ALPHA ? 0x0073 : 0x0043);
    u1 = __byte_perm(srgb_code_b3(v1[0], f2_hi(yr), c), srgb_code_b3(v1[1], f2_hi(yg), c), 0x0073);
    u2 = __byte_perm(srgb_code_b3(v1[2], f2_hi(yb), c), ca1, RAW_ALPHA ? 0x0073 : 0x0043);
  } else {
    const F2 k = f2(255.0f);
    const float m = 8388608.0f;  // (scalar adds: see the alpha above)
    const F2 qr = f2_mul(f2(v0[0], v1[0]), k), qg = f2_mul(f2(v0[1], v1[1]), k), qb = f2_mul(f2(v0[2], v1[2]), k);
    t1 = __byte_perm(__float_as_uint(f2_lo(qr) + m), __float_as_uint(f2_lo(qg) + m), 0x0040);
    t2 = __byte_perm(__float_as_uint(f2_lo(qb) + m), ca0, RAW_ALPHA ? 0x0070 : 0x0040);
    u1 = __byte_perm(__float_as_uint(f2_hi(qr) + m), __float_as_uint(f2_hi(qg) + m), 0x0040);
    u2 = __byte_perm(__float_as_uint(f2_hi(qb) + m), ca1, RAW_ALPHA ? 0x0070 : 0x0040);
  }
  o0 = __byte_perm(t1, t2, c.spack);
  o1 = __byte_perm(u1, u2, c.spack);
}

// Two pixels.  MODE 0: `b` only; 2: `a` over `b` (source-over on straight alpha in linear light, the
// oracle's pd_blend mode 3: identical operation order, each step on both pixels with one packed instruction).
template <int SK, int DK, int MODE, int NMAT>
__device__ __forceinline__ void pixel8x2(const FastParams& P, uint32_t b0, uint32_t b1, uint32_t a0, uint32_t a1, const LutCtx& c, uint32_t& o0, uint32_t& o1) {
  const Px v0 = decode8<SK, false>(b0, c), v1 = decode8<SK, false>(b1, c);
  Px2 v = {f2(v0.r, v1.r), f2(v0.g, v1.g), f2(v0.b, v1.b), f2(1.0f)};
  if (MODE == 2) {
    const Px s0 = decode8<SK, false>(a0, c), s1 = decode8<SK, false>(a1, c);
#ifdef ZOS_LUT_ALPHA_ARITHMETIC
    v.a = alpha_pair(b0, b1);
    const F2 sa = alpha_pair(a0, a1);
#else
    v.a = f2(lds_f32(__byte_perm(b0, c.lane4, c.sa) + (TABLE_AT + DEC_UNORM)), lds_f32(__byte_perm(b1, c.lane4, c.sa) + (TABLE_AT + DEC_UNORM)));
    const F2 sa = f2(lds_f32(__byte_perm(a0, c.lane4, c.sa) + (TABLE_AT + DEC_UNORM)), lds_f32(__byte_perm(a1, c.lane4, c.sa) + (TABLE_AT + DEC_UNORM)));
#endif
    const F2 wbk = f2_mul(v.a, f2_sub(f2(1.0f), sa));
    const F2 ao = f2_add(sa, wbk);
    // ao is 0 or in [1/255, 1]: SFU reciprocal + one Newton step is the correctly rounded 1/ao there
    // (tested for all alpha pairs).  ao == 0 has a zero numerator: any finite reciprocal gives the
    // oracle's 0, so the guard is a max with a tiny normal number instead of a select.
    const float aos0 = fmaxf(f2_lo(ao), 1e-30f), aos1 = fmaxf(f2_hi(ao), 1e-30f);
    float q0, q1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(aos0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(aos1));
    const F2 r0 = f2(q0, q1);
    // fmaf(r0, -fmaf(aos, r0, -1), r0), the negation moved onto r0 (exact)
    const F2 rcp = f2_fma(f2_sub(f2(0.0f), r0), f2_fma(f2(aos0, aos1), r0, f2(-1.0f)), r0);
    v.r = f2_mul(f2_fma(wbk, v.r, f2_mul(sa, f2(s0.r, s1.r))), rcp);
    v.g = f2_mul(f2_fma(wbk, v.g, f2_mul(sa, f2(s0.g, s1.g))), rcp);
    v.b = f2_mul(f2_fma(wbk, v.b, f2_mul(sa, f2(s0.b, s1.b))), rcp);
    v.a = ao;
  }
#pragma unroll
  for (int k = 0; k < NMAT; k++) f2_mat3(P.m[k], v.r, v.g, v.b);
  encode8x2<DK, (NMAT > 0), MODE == 0>(v, c, b0, b1, o0, o1);
}

template <int SK, int DK, int MODE, int NMAT>
__device__ __forceinline__ uint32_t pixel8(const FastParams& P, uint32_t b, uint32_t a, const LutCtx& c) {
  uint32_t o0, o1;
  pixel8x2<SK, DK, MODE, NMAT>(P, b, b, a, a, c, o0, o1);  // (edge groups only: the duplicate half folds away or is cheap)
  return o0;
}

// LINEAR: every layer has the destination's geometry with rows and frames back to back, so the images
// are plain streams of 16-byte groups (the blend / convert workloads).  Two groups are in flight per
// thread, ping-pong, so that no register rotation is needed.
template <int SK, int DK, int MODE, int NMAT, bool LINEAR>
__global__ void __launch_bounds__(LUT_THREADS, 1) k_rowwise_lut(const __grid_constant__ FastParams P) {
  extern __shared__ __align__(256) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((uint32_t)__cvta_generic_to_shared(smem_raw) != TABLE_AT) {  // not the layout the look-up immediates were built for: report, do not compute
    if (threadIdx.x == 0 && P.fault) *reinterpret_cast<volatile int*>(P.fault) = 2;
    return;
  }
  // Table fill, the fixed cost of a launch (it dominates for small images): 16-byte stores, a warp writes 512
  // contiguous bytes per instruction, all (L2-resident) source loads of a thread are independent.
#pragma unroll
  for (int k = 0; k < 512 * 8 / LUT_THREADS; k++) {
    const int f = k * LUT_THREADS + threadIdx.x;                 // uint4 index within the decode columns: row f / 8
    const int row = f >> 3;
    const float v = row < 256 ? g_tables.srgb_dec[row] : g_tables.unorm8[row - 256];
    reinterpret_cast<float4*>(smem + (uint32_t)row * ROW + 128u)[f & 7] = make_float4(v, v, v, v);
  }
  if (DK == K_SRGB8) {
#pragma unroll
    for (int k = 0; k < (ZOS_ENC2_N * 8 + LUT_THREADS - 1) / LUT_THREADS; k++) {
      const int f = k * LUT_THREADS + threadIdx.x;               // uint4 index: bucket f / 8
      if (f < ZOS_ENC2_N * 8) {
        const uint32_t e = g_tables.srgb_enc2[f >> 3];
        reinterpret_cast<uint4*>(smem + (uint32_t)(f >> 3) * ROW)[f & 7] = make_uint4(e, e, e, e);
      }
    }
  }
  __syncthreads();

  LutCtx c;
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
      if ((uint32_t)j < mine) { b[j] = ld_stream(pb + (size_t)j * stride); if (MODE) a[j] = ld_stream(pa + (size_t)j * stride); }
#define ZOS_LUT_GROUP(B, A, OUT)                                                        \
    {                                                                                   \
      uint4 o_;                                                                         \
      pixel8x2<SK, DK, MODE, NMAT>(P, B.x, B.y, A.x, A.y, c, o_.x, o_.y);               \
      pixel8x2<SK, DK, MODE, NMAT>(P, B.z, B.w, A.z, A.w, c, o_.z, o_.w);               \
      st_stream(OUT, o_);                                                               \
    }
    for (uint32_t k = 0;; k += D) {
#pragma unroll
      for (int j = 0; j < D; j++) {
        if (k + j >= mine) return;
        if (k + j + (D - 1) < mine) {
          b[(j + D - 1) % D] = ld_stream(pb + (size_t)(j + D - 1) * stride);
          if (MODE) a[(j + D - 1) % D] = ld_stream(pa + (size_t)(j + D - 1) * stride);
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
      pixel8x2<SK, DK, MODE, NMAT>(P, wb[0], wb[1], wa[0], wa[1], c, o[0], o[1]);
      pixel8x2<SK, DK, MODE, NMAT>(P, wb[2], wb[3], wa[2], wa[3], c, o[2], o[3]);
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
  const uint32_t bytes = DK == K_SRGB8 ? SMEM_BYTES : 512u * ROW;
  auto kern = k_rowwise_lut<SK, DK, MODE, NMAT, false>;
  auto kern_lin = k_rowwise_lut<SK, DK, MODE, NMAT, true>;
  {
    cudaError_t e = ensure_dyn_smem(ctx, kern, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = ensure_dyn_smem(ctx, kern_lin, (int)SMEM_BYTES);
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
