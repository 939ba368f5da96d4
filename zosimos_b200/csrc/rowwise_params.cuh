// rowwise_params.cuh -- launch parameters shared by the specialised streaming kernels
// (rowwise_fast.cu: float / RGB10 texels; rowwise_lut.cu: 8-bit texels).
#pragma once
#include "zos_internal.h"

namespace zos {

constexpr int REP = 16;
enum Kind { K_SRGB8 = 0, K_UNORM8 = 1, K_F16 = 2, K_F32 = 3, K_RGB10 = 4 /* staged UInt1010102 RgbA, linear or sRGB transfer */ };

struct FastParams {
  const uint8_t* below;
  const uint8_t* above;
  uint8_t* dst;
  uint64_t below_pitch, above_pitch, dst_pitch;
  uint64_t below_bstride, above_bstride, dst_bstride;
  int32_t w, h;
  int32_t has_below;
  int32_t tx, ty, aw, ah;  // placement of `above`
  int32_t src_bgra, dst_bgra;
  int32_t src_tr, dst_tr;  // ZOS_TRANSFER_* of K_RGB10 sources / destinations
  int32_t nmat;
  int32_t linear;          // rowwise_lut: all layers share one contiguous geometry
  float m[2][9];
  uint32_t groups_per_row, total_groups;
  FastDiv div_gpr, div_h;
  int* fault;  // mapped host word (zos_ctx::fault_dev): a kernel that cannot run as built sets it and returns; zos_sync reports it
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}

struct Loc {
  uint64_t ob, oa, od;  // byte offsets of the 4-texel group in below / above / dst
  int npx, ncov;        // valid texels of the group, of which covered by `above` (from the left)
};

template <int MODE>
__device__ __forceinline__ Loc locate(const FastParams& P, uint32_t idx) {
  Loc L;
  uint32_t rowid = fastdiv(idx, P.div_gpr);
  uint32_t g = idx - rowid * P.groups_per_row;
  uint32_t frame = fastdiv(rowid, P.div_h);
  int y = (int)(rowid - frame * (uint32_t)P.h);
  int x0 = (int)g * 4;
  L.npx = min(4, P.w - x0);
  L.ncov = 0; L.oa = 0;
  if (MODE != 0) {
    int ax0 = x0 - P.tx, ay = y - P.ty;
    bool row_in = ay >= 0 && ay < P.ah && ax0 >= 0 && ax0 < P.aw;
    L.ncov = row_in ? min(L.npx, P.aw - ax0) : 0;
    if (row_in) L.oa = frame * P.above_bstride + (uint64_t)ay * P.above_pitch + (uint64_t)ax0 * 4u;
  }
  L.ob = frame * P.below_bstride + (uint64_t)y * P.below_pitch + (uint64_t)x0 * 4u;
  L.od = frame * P.dst_bstride + (uint64_t)y * P.dst_pitch + (uint64_t)x0 * 4u;
  return L;
}

// rowwise_lut.cu: 8-bit -> 8-bit texel pairs (sk, dk in {K_SRGB8, K_UNORM8}); mode 0 = convert, 2 = source-over
cudaError_t launch_rowwise_lut(zos_ctx* ctx, FastParams& P, int sk, int dk, int mode, int nmat);
// rowwise_rgb10.cu: staged RGB10A2 on both sides, convert mode
cudaError_t launch_rowwise_rgb10(zos_ctx* ctx, const FastParams& P, int nmat);
// rowwise_lab.cu: 8-bit -> [Lab encode, 8-bit staged register, Lab decode] -> 8-bit; *handled stays false when not served
cudaError_t launch_rowwise_lab(zos_ctx* ctx, const DevImage& src, const DevImage& dst, const zos_step* steps, uint32_t nsteps, uint32_t batch, bool* handled);

}  // namespace zos
