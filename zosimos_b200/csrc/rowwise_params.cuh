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
};

__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
  uint32_t t = __umulhi(n, f.m);
  return f.l == 0 ? n : (t + ((n - t) >> 1)) >> (f.l - 1);
}

// rowwise_lut.cu: 8-bit -> 8-bit texel pairs (sk, dk in {K_SRGB8, K_UNORM8}); mode 0 = convert, 2 = source-over
cudaError_t launch_rowwise_lut(zos_ctx* ctx, FastParams& P, int sk, int dk, int mode, int nmat);

}  // namespace zos
