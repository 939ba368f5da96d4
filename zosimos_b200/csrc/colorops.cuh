// colorops.cuh -- per-pixel colour steps fused between unpack and pack.
// linear.frag:12-17, oklab.frag:34-64, srlab2.frag:36-120, inject.frag:20-25 of the reference
// (lib/std/src/).  Matrices are row-major; evaluation order matches the CPU oracle.
#pragma once
#include "texel.cuh"

namespace zos {

struct ColorConstants {  // filled on the host in double precision, see runtime.cu
  float ok_m1[9], ok_m2[9], ok_m1i[9], ok_m2i[9];
  float sr_cat[9], sr_cati[9], sr_hpe[9], sr_hpei[9], sr_hpe_cati[9], sr_cat_hpei[9];
};
static __constant__ ColorConstants c_color;

#define ZOS_DEFINE_CONSTANT_UPLOAD(NAME)                                                              \
  cudaError_t NAME(const zos::TablesGlobal* t, const zos::ColorConstants* c, cudaStream_t stream) { \
    cudaError_t e = cudaMemcpyToSymbolAsync(zos::g_tables, t, sizeof(*t), 0, cudaMemcpyHostToDevice, stream); \
    if (e != cudaSuccess) return e;                                                                   \
    return cudaMemcpyToSymbolAsync(zos::c_color, c, sizeof(*c), 0, cudaMemcpyHostToDevice, stream);   \
  }

struct StepList {
  uint32_t n;
  zos_step s[ZOS_MAX_STEPS];
};

__device__ __forceinline__ float3 mat3_mul(const float* M, float x, float y, float z) {
  return make_float3(fmaf(M[2], z, fmaf(M[1], y, M[0] * x)), fmaf(M[5], z, fmaf(M[4], y, M[3] * x)),
                     fmaf(M[8], z, fmaf(M[7], y, M[6] * x)));
}
__device__ __forceinline__ float cbrt_signed(float v) {
  // pow(abs(v), 1/3) * sign(v), oklab.frag:43
  float r = pow_fast(fabsf(v), 1.0f / 3.0f);  // lg2.approx / ex2.approx on the SFU, like every other pow of the path
  return v == 0.0f ? 0.0f : copysignf(r, v);
}
__device__ __forceinline__ float sr_nl(float v) {
  return fabsf(v) < 216.0f / 24389.0f ? v * 24389.0f / 2700.0f : 1.16f * pow_fast(v, 1.0f / 3.0f) - 0.16f;
}
__device__ __forceinline__ float sr_nl_inv(float v) {
  if (fabsf(v) < 0.08f) return v * 2700.0f / 24389.0f;
  float vp = (v + 0.16f) / 1.16f;
  return vp * vp * vp;
}

// The four non-linear colour steps (oklab.frag:34-64, srlab2.frag:36-120), shared by the step interpreter
// below and by the specialised Lab kernel (rowwise_lab.cu): one definition, identical results.
// s.m of the two Oklab steps is the FOLDED matrix here: M1 * to_xyz for the encoder, from_xyz * M1^-1 for the decoder, multiplied
// in double precision by the C-ABI entry points (zos::fold_steps, runtime.cu) -- one 3x3 product per pixel and direction less
// than oklab.frag:34-64 spells out (two matrices applied in turn); the chain carries the <= 1 LSB tolerance of its SFU cube
// roots anyway.
__device__ __forceinline__ void oklab_enc(const zos_step& s, float4& v) {
  float3 lms = mat3_mul(s.m, v.x, v.y, v.z);
  float3 lab = mat3_mul(c_color.ok_m2, cbrt_signed(lms.x), cbrt_signed(lms.y), cbrt_signed(lms.z));
  v.x = lab.x; v.y = lab.y; v.z = lab.z;
}
__device__ __forceinline__ void oklab_dec(const zos_step& s, float4& v) {
  float3 l = mat3_mul(c_color.ok_m2i, v.x, v.y, v.z);
  float3 rgb = mat3_mul(s.m, l.x * l.x * l.x, l.y * l.y * l.y, l.z * l.z * l.z);
  v.x = clamp01(rgb.x); v.y = clamp01(rgb.y); v.z = clamp01(rgb.z);
}
__device__ __forceinline__ void srlab2_enc(const zos_step& s, float4& v) {
  float3 xyz = mat3_mul(s.m, v.x, v.y, v.z);
  float3 rw = mat3_mul(c_color.sr_cat, xyz.x, xyz.y, xyz.z);
  float3 lms = mat3_mul(c_color.sr_hpe_cati, rw.x, rw.y, rw.z);
  float3 e = mat3_mul(c_color.sr_hpei, sr_nl(lms.x), sr_nl(lms.y), sr_nl(lms.z));
  v.x = e.y; v.y = (e.x - e.y) * 5.0f / 1.16f; v.z = (e.z - e.y) * 2.0f / 1.16f;
}
__device__ __forceinline__ void srlab2_dec(const zos_step& s, float4& v) {
  float3 wp = mat3_mul(c_color.sr_cat, s.v[0], s.v[1], s.v[2]);
  float3 t = mat3_mul(c_color.sr_hpe, v.y * 1.16f / 5.0f + v.x, v.x, v.z * 1.16f / 2.0f + v.x);
  float3 rw = mat3_mul(c_color.sr_cat_hpei, sr_nl_inv(t.x), sr_nl_inv(t.y), sr_nl_inv(t.z));
  float3 xyz = mat3_mul(c_color.sr_cati, rw.x * wp.x, rw.y * wp.y, rw.z * wp.z);
  float3 rgb = mat3_mul(s.m, xyz.x, xyz.y, xyz.z);
  v.x = clamp01(rgb.x); v.y = clamp01(rgb.y); v.z = clamp01(rgb.z);
}

static __device__ ZOS_SLOW_ATTR float4 apply_step_slow(const zos_step* sp, float4 v, const Tables* Tp) {
  const zos_step& s = *sp;
  const Tables& T = *Tp;
  switch (s.kind) {
    case ZOS_STEP_MATRIX: {
      float3 o = mat3_mul(s.m, v.x, v.y, v.z);
      v.x = o.x; v.y = o.y; v.z = o.z;
      break;
    }
    case ZOS_STEP_OKLAB_ENC: oklab_enc(s, v); break;
    case ZOS_STEP_OKLAB_DEC: oklab_dec(s, v); break;
    case ZOS_STEP_SRLAB2_ENC: srlab2_enc(s, v); break;
    case ZOS_STEP_SRLAB2_DEC: srlab2_dec(s, v); break;
    case ZOS_STEP_REQUANT: {
      uint4 w = pack_slow(s.fmt, v, Tp);
      v = unpack_slow(s.fmt, w, Tp);
      break;
    }
    case ZOS_STEP_F16:
      v = make_float4(f16r(v.x), f16r(v.y), f16r(v.z), f16r(v.w));
      break;
    default: break;
  }
  return v;
}
// The 3x3 matrix (by far the most common step) is applied inline with its coefficients read
// straight from the constant bank; the other steps are out-of-line calls.
__device__ __forceinline__ void apply_steps(const StepList& sl, float4& v, const Tables& T) {
  for (uint32_t i = 0; i < sl.n; i++) {
    if (sl.s[i].kind == ZOS_STEP_MATRIX) {
      float3 o = mat3_mul(sl.s[i].m, v.x, v.y, v.z);
      v.x = o.x; v.y = o.y; v.z = o.z;
    } else {
      v = apply_step_slow(&sl.s[i], v, &T);
    }
  }
}

// Porter-Duff on straight alpha in linear light (ours; the reference's blend is unimplemented,
// lib/zosimos/src/command.rs:1510-1519).  s = above, d = below.
__device__ __forceinline__ float4 porter_duff(int mode, const float4& s, const float4& d) {
  float as = s.w, ad = d.w, fa, fb;
  switch (mode) {
    case ZOS_BLEND_CLEAR: fa = 0.0f; fb = 0.0f; break;
    case ZOS_BLEND_SRC: fa = 1.0f; fb = 0.0f; break;
    case ZOS_BLEND_DST: fa = 0.0f; fb = 1.0f; break;
    case ZOS_BLEND_SRC_OVER: fa = 1.0f; fb = 1.0f - as; break;
    case ZOS_BLEND_DST_OVER: fa = 1.0f - ad; fb = 1.0f; break;
    case ZOS_BLEND_SRC_IN: fa = ad; fb = 0.0f; break;
    case ZOS_BLEND_DST_IN: fa = 0.0f; fb = as; break;
    case ZOS_BLEND_SRC_OUT: fa = 1.0f - ad; fb = 0.0f; break;
    case ZOS_BLEND_DST_OUT: fa = 0.0f; fb = 1.0f - as; break;
    case ZOS_BLEND_SRC_ATOP: fa = ad; fb = 1.0f - as; break;
    case ZOS_BLEND_DST_ATOP: fa = 1.0f - ad; fb = as; break;
    default: fa = 1.0f - ad; fb = 1.0f - as; break;
  }
  float wa = as * fa, wb = ad * fb;
  float ao = wa + wb;
  float4 o;
  // un-premultiply with ONE correctly rounded reciprocal (the oracle does the same: 1.0f / ao)
  float rcp = ao > 0.0f ? __frcp_rn(ao) : 0.0f;
  o.x = fmaf(wb, d.x, wa * s.x) * rcp;
  o.y = fmaf(wb, d.y, wa * s.y) * rcp;
  o.z = fmaf(wb, d.z, wa * s.z) * rcp;
  o.w = ao;
  return o;
}

// inject.frag:20-25: mix(bg, vec4(dot(fg, color)), select); mix(a, b, t) = a*(1-t) + b*t
__device__ __forceinline__ float4 inject_blend(const float* inj, const float4& fg, const float4& bg) {
  const float d = fg.x * inj[4] + fg.y * inj[5] + fg.z * inj[6] + fg.w * inj[7];
  return make_float4(bg.x * (1.0f - inj[0]) + d * inj[0], bg.y * (1.0f - inj[1]) + d * inj[1], bg.z * (1.0f - inj[2]) + d * inj[2],
                     bg.w * (1.0f - inj[3]) + d * inj[3]);
}

}  // namespace zos
