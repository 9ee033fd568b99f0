// Device-side helpers shared by all kernels: integer utilities, MV clipping, mvd bit
// counts, QP derivation.  Each mirrors a reference routine (cited).
#ifndef XVCB_DEVICE_CUH_
#define XVCB_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "xvcb_internal.h"

namespace xvcb {

#define XVCB_FULL 0xffffffffu

__host__ __device__ __forceinline__ int clip3i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__host__ __device__ __forceinline__ int ilog2i(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

// InterPrediction::ClipMv, inter_prediction.cc:769-782 (1/16 pel)
__host__ __device__ __forceinline__ void clip_mv(int pos_x, int pos_y, int pic_w, int pic_h, int &mx, int &my) {
  mx = clip3i(mx, -((64 + 8 + pos_x - 1) << 4), (pic_w + 8 - pos_x - 1) << 4);
  my = clip3i(my, -((64 + 8 + pos_y - 1) << 4), (pic_h + 8 - pos_y - 1) << 4);
}

// InterPrediction::DetermineMinMaxMv, inter_prediction.cc:801-817 (results full-pel)
__host__ __device__ __forceinline__ void min_max_mv(int pos_x, int pos_y, int pic_w, int pic_h, int cx, int cy,
                                                    int range, int lo[2], int hi[2]) {
  clip_mv(pos_x, pos_y, pic_w, pic_h, cx, cy);
  int lx = cx - (range << 4), ly = cy - (range << 4), hx = cx + (range << 4), hy = cy + (range << 4);
  clip_mv(pos_x, pos_y, pic_w, pic_h, lx, ly);
  clip_mv(pos_x, pos_y, pic_w, pic_h, hx, hy);
  lo[0] = lx >> 4; lo[1] = ly >> 4; hi[0] = hx >> 4; hi[1] = hy >> 4;
}

// InterSearch::GetNumExpGolombBits, inter_search.cc:1176-1185: 1 + 2*floor(log2(u))
__host__ __device__ __forceinline__ uint32_t exp_golomb_bits(int v) {
  uint32_t u = v <= 0 ? ((uint32_t)(-v) << 1) + 1 : (uint32_t)v << 1;
#ifdef __CUDA_ARCH__
  return 1u + 2u * (31u - (uint32_t)__clz(u));
#else
  uint32_t len = 1; while (u != 1) { u >>= 1; len += 2; } return len;
#endif
}
// GetMvdBitsFullpel, inter_search.cc:1162-1174
__host__ __device__ __forceinline__ uint32_t mvd_bits_fullpel(int mvpx, int mvpy, int x, int y, int down) {
  const int sh = down + 2;
  return exp_golomb_bits((x * 16 - mvpx) >> sh) + exp_golomb_bits((y * 16 - mvpy) >> sh);
}
// GetMvdBits, inter_search.cc:1144-1154
__host__ __device__ __forceinline__ uint32_t mvd_bits(int mvpx, int mvpy, int mx, int my) {
  return exp_golomb_bits((mx - mvpx) >> 2) + exp_golomb_bits((my - mvpy) >> 2);
}

// Quantize::GetTransformShift, quantize.cc:127-131
__host__ __device__ __forceinline__ int transform_shift(int lw, int lh, int bitdepth) {
  return 15 - bitdepth - ((lw + lh) >> 1);
}

// Qp::ScaleChromaQp (quantize.cc:74-81): raw chroma qp for 4:2:0.
__device__ __forceinline__ int chroma_qp_raw(int qp, int offset, int table, const uint8_t *chroma_scale) {
  int q = clip3i(qp + offset, 0, 57);
  return table == 1 ? (int)chroma_scale[q] : q;
}

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(XVCB_FULL, v, o);
  return v;
}

}  // namespace xvcb
#endif
