// rowwise_rgb10.cu -- staged UInt1010102 RgbA texels on both sides (the RGB10A2 row of BASELINE config 5)
// with 0..2 matrix steps in between.  The reference handles this texel in stage.frag (decode
// :471-482,533-582, encode :484-501,590-641: demux, inverse transfer, Rgba16Float working texture; f16
// attachment, transfer, clamp, TRUNCATING quantisation).  Every stage of that codec is a function of few
// bits, so both directions become tables built at kernel start BY RUNNING THE GENERIC CODEC'S OWN CODE:
//
//   decode: D[k]  = f16(eotf(k / 1023)), 1024 entries (8 copies);   alpha: 4 entries, via A[] below
//   encode: E[h]  = uint(clamp01(oetf(half(h))) * 1023) for the f16 bit patterns h of [0, 1] (the value stored
//           in the f16 attachment is all the encoder ever sees; the clamp is applied BEFORE the f16 rounding,
//           which commutes with it because 0 and 1 are f16 values and rounding is monotone), as u16:
//           15361 entries + one for everything above 1 (oetf(1) may evaluate a hair below 1) = 30 KB, 4 copies
//           interleaved entry by entry ([h][lane & 3]: the address is one multiply-add, and the 16 lanes of copies
//           0-1 / 2-3 spread over the 16 even / odd banks: 2.65 wavefronts per look-up instead of 3.2 for one
//           128 KB table; word-interleaved copies were measured twice and are no faster, their address costs
//           three more integer instructions per channel);
//   alpha:  A[a2] = encode(decode(a2)) (not the identity: f16(1/3) * 3 truncates to 0).
//
// Results equal k_rowwise_fast<K_RGB10, K_RGB10> and the generic kernel bit for bit (tests); instead of
// six pow evaluations a pixel costs three decode and three encode look-ups.
#include <cuda_fp16.h>

#include "colorops.cuh"
#include "zos_internal.h"
#include "rowwise_params.cuh"

namespace zos {

ZOS_DEFINE_CONSTANT_UPLOAD(upload_constants_rowwise_rgb10)

namespace {
constexpr int THREADS = 1024;
constexpr int DR = 16;                                 // copies of the decode table
constexpr uint32_t DEC_BYTES = 1024u * DR * 4u;        // [code][lane & 15]
constexpr int ER = 4;                                  // copies of the encode table
constexpr uint32_t ENC_N = 0x3c00u + 2u;               // f16 patterns of [0, 1], padded to whole words
constexpr uint32_t ENC_BYTES = ENC_N * ER * 2u;        // [h][lane & 3] u16 entries
constexpr uint32_t SMEM_BYTES = DEC_BYTES + ENC_BYTES;

__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t half_bits(float v) {
  return (uint32_t)__half_as_ushort(__float2half_rn(v));
}

// byte address of entry h = f16(clamp01(v)) in this lane's copy
__device__ __forceinline__ uint32_t enc_addr(uint32_t enc_lane, float v) {
  const uint32_t h = min(half_bits(fmaxf(v, 0.0f)), ENC_N - 1u);  // fmaxf(NaN, 0) = 0 like clamp01; the last entry stands for everything above 1
  return enc_lane + h * (ER * 2u);
}

template <int NMAT>
__device__ __forceinline__ uint32_t pixel(const FastParams& P, uint32_t w, uint32_t dec_lane, uint32_t enc, uint32_t amap) {
  float r = lds_f32(dec_lane + (w & 1023u) * (DR * 4u));
  float g = lds_f32(dec_lane + ((w >> 10) & 1023u) * (DR * 4u));
  float b = lds_f32(dec_lane + ((w >> 20) & 1023u) * (DR * 4u));
#pragma unroll
  for (int k = 0; k < NMAT; k++) {
    float3 t = mat3_mul(P.m[k], r, g, b);
    r = t.x; g = t.y; b = t.z;
  }
  const uint32_t cr = lds_u16(enc_addr(enc, r)), cg = lds_u16(enc_addr(enc, g)), cb = lds_u16(enc_addr(enc, b));
  const uint32_t a2 = (amap >> ((w >> 30) * 2u)) & 3u;
  return cr + (cg << 10) + (cb << 20) + (a2 << 30);
}

template <int NMAT, bool LINEAR>
__global__ void __launch_bounds__(THREADS, 1) k_rowwise_rgb10(const __grid_constant__ FastParams P) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* dec = reinterpret_cast<float*>(smem);
  uint16_t* enc = reinterpret_cast<uint16_t*>(smem + DEC_BYTES);
  const bool src_srgb = P.src_tr == ZOS_TRANSFER_SRGB, dst_srgb = P.dst_tr == ZOS_TRANSFER_SRGB;
  // the codec of rowwise_fast.cu's K_RGB10 (== stage.frag), evaluated once per table entry
  for (int i = threadIdx.x; i < 1024 * DR; i += THREADS) {
    float p = fld((uint32_t)(i / DR), 1023.0f);
    if (src_srgb) p = eo_srgb(p);
    dec[i] = f16r(p);
  }
  for (int i = threadIdx.x; i < (int)(ENC_N * ER); i += THREADS) {
    const uint32_t h = (uint32_t)i / ER, c = (uint32_t)i % ER;
    float r = __half2float(__ushort_as_half((unsigned short)h));  // what the f16 attachment holds (0x3c01: any value above 1)
    if (dst_srgb) r = oe_srgb(r);
    enc[h * ER + c] = (uint16_t)(uint32_t)(clamp01(r) * 1023.0f);
  }
  uint32_t amap = 0;
#pragma unroll
  for (uint32_t a = 0; a < 4; a++) {
    const float v = f16r(fld(a, 3.0f));                   // decode
    amap |= ((uint32_t)(clamp01(f16r(v)) * 3.0f)) << (2u * a);  // encode
  }
  __syncthreads();
  const uint32_t dec_lane = (uint32_t)__cvta_generic_to_shared(dec) + (threadIdx.x & (DR - 1)) * 4u;
  const uint32_t enc_base = (uint32_t)__cvta_generic_to_shared(enc) + (threadIdx.x & (ER - 1)) * 2u;
  const uint32_t stride = gridDim.x * THREADS;
  uint32_t idx = blockIdx.x * THREADS + threadIdx.x;
  if (idx >= P.total_groups) return;
  if constexpr (LINEAR) {
    // source and destination are plain streams of 16-byte groups (rows and frames back to back, width a multiple
    // of 4): no index arithmetic (the two divisions + 64-bit offsets of locate() were 14 of 60 instructions per
    // pixel on a kernel bound by its integer work), the loads of the next two groups in flight
    constexpr int D = 3;
    const uint32_t mine = (P.total_groups - 1u - idx) / stride + 1u;  // groups of this thread: idx + k * stride, k < mine
    const uint4* pb = reinterpret_cast<const uint4*>(P.below) + idx;
    uint4* pd = reinterpret_cast<uint4*>(P.dst) + idx;
    uint4 b[D];
#pragma unroll
    for (int j = 0; j < D; j++) b[j] = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < D - 1; j++)
      if ((uint32_t)j < mine) b[j] = __ldcs(pb + (size_t)j * stride);
    for (uint32_t k = 0;; k += D) {
#pragma unroll
      for (int j = 0; j < D; j++) {
        if (k + j >= mine) return;
        if (k + j + (D - 1) < mine) b[(j + D - 1) % D] = __ldcs(pb + (size_t)(j + D - 1) * stride);
        __stcs(pd + (size_t)j * stride, make_uint4(pixel<NMAT>(P, b[j].x, dec_lane, enc_base, amap), pixel<NMAT>(P, b[j].y, dec_lane, enc_base, amap),
                                                    pixel<NMAT>(P, b[j].z, dec_lane, enc_base, amap), pixel<NMAT>(P, b[j].w, dec_lane, enc_base, amap)));
      }
      pb += (size_t)D * stride; pd += (size_t)D * stride;
    }
  } else {
  // software pipeline: the next group's 16 bytes are in flight while the current group goes through its
  // ~200 instructions (one CTA of 32 warps per SM: without it every warp idles for a full DRAM round trip
  // per group -- 41 % of the stall samples, profiles/r01_c5_rgb10_kernel.txt)
  Loc L = locate<0>(P, idx);
  uint4 rb = __ldcs(reinterpret_cast<const uint4*>(P.below + L.ob));
  for (;;) {
    const uint32_t nidx = idx + stride;
    const bool more = nidx > idx && nidx < P.total_groups;  // (32-bit wrap ends the walk)
    const Loc Ln = locate<0>(P, more ? nidx : idx);
    uint4 rn = rb;
    // (volatile + predicated: the compiler otherwise sinks the conditional load below the arithmetic)
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %5, 0;\n\t@q ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];\n\t}"
                 : "+r"(rn.x), "+r"(rn.y), "+r"(rn.z), "+r"(rn.w) : "l"(P.below + Ln.ob), "r"((uint32_t)more) : "memory");
    uint32_t o[4];
    o[0] = pixel<NMAT>(P, rb.x, dec_lane, enc_base, amap); o[1] = pixel<NMAT>(P, rb.y, dec_lane, enc_base, amap);
    o[2] = pixel<NMAT>(P, rb.z, dec_lane, enc_base, amap); o[3] = pixel<NMAT>(P, rb.w, dec_lane, enc_base, amap);
    uint8_t* dp = P.dst + L.od;
    if (L.npx == 4) {
      __stcs(reinterpret_cast<uint4*>(dp), make_uint4(o[0], o[1], o[2], o[3]));
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++)
        if (i < L.npx) reinterpret_cast<uint32_t*>(dp)[i] = o[i];
    }
    if (!more) break;
    idx = nidx; L = Ln; rb = rn;
  }
  }
}

template <int NMAT>
cudaError_t launch_one(zos_ctx* ctx, const FastParams& P) {
  {
    cudaError_t e = ensure_dyn_smem(ctx, k_rowwise_rgb10<NMAT, false>, (int)SMEM_BYTES);
    if (e == cudaSuccess) e = ensure_dyn_smem(ctx, k_rowwise_rgb10<NMAT, true>, (int)SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const uint64_t ctas = ((uint64_t)P.total_groups + THREADS - 1) / THREADS;
  const int grid = (int)(ctas < (uint64_t)ctx->sm_count ? ctas : (uint64_t)ctx->sm_count);
  // linear addressing: rows and frames of source and destination back to back, whole groups only
  const uint64_t row = (uint64_t)P.w * 4u, frame = row * P.h;
  const bool one = (uint64_t)P.total_groups == (uint64_t)P.groups_per_row * (uint64_t)P.h;  // a single frame: frame strides are not used
  const bool linear = (P.w % 4 == 0) && P.below_pitch == row && P.dst_pitch == row && (one || (P.below_bstride == frame && P.dst_bstride == frame)) &&
                      ((uintptr_t)P.below % 16 == 0) && ((uintptr_t)P.dst % 16 == 0);
  if (linear) k_rowwise_rgb10<NMAT, true><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(P);
  else k_rowwise_rgb10<NMAT, false><<<grid, THREADS, SMEM_BYTES, ctx->stream>>>(P);
  return cudaGetLastError();
}
}  // namespace

// staged RGB10A2 -> staged RGB10A2 with `nmat` matrix steps (P as prepared by launch_rowwise_u8)
cudaError_t launch_rowwise_rgb10(zos_ctx* ctx, const FastParams& P, int nmat) {
  if (nmat == 0) return launch_one<0>(ctx, P);
  if (nmat == 1) return launch_one<1>(ctx, P);
  return launch_one<2>(ctx, P);
}

}  // namespace zos
