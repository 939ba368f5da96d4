// f32x2.cuh -- packed single-precision pairs (sm_100a `fma/mul/add.rn.f32x2`: FFMA2 / FMUL2 / FADD2).  One instruction does
// the IEEE round-to-nearest operation on both halves of a 64-bit register pair, so the results equal the scalar
// fmaf / * / + bit for bit (no flush-to-zero, no contraction: the kernels are built with --fmad=false and only ever fuse
// through explicit fma).  The streaming kernels here are bound by instruction ISSUE, not by the FMA pipe (26 % busy in
// profiles/r01_c2_blend_lut_kernel_final.txt), so halving the issue slots of the floating-point part is a direct gain.
//
// CAUTION (measured with cuobjdump, CUDA 12.9): ptxas contracts `mul.rn.f32x2` into an `add.rn.f32x2` that is the product's
// ONLY user (one FFMA2, a single rounding) even with explicit .rn modifiers and --fmad=false.  Where the scalar code rounds
// the product and the sum separately and the product has no other user, do the add with scalar instructions.
#pragma once
#include <cuda_runtime.h>

namespace zos {

#ifndef ZOS_F2_SCALAR
struct F2 { unsigned long long v; };

__device__ __forceinline__ F2 f2(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ F2 f2(float both) { return f2(both, both); }
__device__ __forceinline__ float f2_lo(F2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float f2_hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
#else  // the same interface on two scalar registers (A/B measurements of the packed instructions)
struct F2 { float lo, hi; };
__device__ __forceinline__ F2 f2(float lo, float hi) { F2 r; r.lo = lo; r.hi = hi; return r; }
__device__ __forceinline__ F2 f2(float both) { return f2(both, both); }
__device__ __forceinline__ float f2_lo(F2 a) { return a.lo; }
__device__ __forceinline__ float f2_hi(F2 a) { return a.hi; }
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) { return f2(fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)); }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) { return f2(a.lo * b.lo, a.hi * b.hi); }
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) { return f2(a.lo + b.lo, a.hi + b.hi); }
__device__ __forceinline__ F2 f2_sub(F2 a, F2 b) { return f2(a.lo - b.lo, a.hi - b.hi); }
#endif

// mat3_mul of colorops.cuh on two pixels: the same fmaf chain, M[2]*z + (M[1]*y + M[0]*x)
__device__ __forceinline__ void f2_mat3(const float* M, F2& x, F2& y, F2& z) {
  const F2 r = f2_fma(f2(M[2]), z, f2_fma(f2(M[1]), y, f2_mul(f2(M[0]), x)));
  const F2 g = f2_fma(f2(M[5]), z, f2_fma(f2(M[4]), y, f2_mul(f2(M[3]), x)));
  const F2 b = f2_fma(f2(M[8]), z, f2_fma(f2(M[7]), y, f2_mul(f2(M[6]), x)));
  x = r; y = g; z = b;
}

}  // namespace zos
