// Warp-cooperative Hadamard SATD (SampleMetric::ComputeSatd, sample_metric.cc:316-668).
//
// Mapping: one lane per tile ROW.  The lane keeps the TW differences of its row in
// registers, runs the horizontal Walsh-Hadamard there, then the vertical one as
// xor-shuffle butterflies across the TH lanes that hold the same tile; |.| is summed per
// lane and across the TH lanes, the per-shape normalisation (sample_metric.cc:633-639)
// is applied once per tile.  The sum of magnitudes does not depend on the order of the
// Hadamard outputs, so the butterfly schedule is free to differ from the reference's.
#ifndef XVCB_SATD_CUH_
#define XVCB_SATD_CUH_

#include "xvcb_device.cuh"

namespace xvcb {

template <int TW, int TH> __device__ __forceinline__ int satd_norm(int s) {
  if (TW == 2 && TH == 2) return s;
  if (TW == 4 && TH == 4) return (s + 1) >> 1;
  if (TW == TH) return (s + 2) >> 2;
  // static_cast<int>(2.0 * sum / std::sqrt(W*H)): IEEE double, sqrt(32) / sqrt(128) as the
  // correctly rounded doubles std::sqrt returns
  const double root = __longlong_as_double(TW * TH == 32 ? 0x4016A09E667F3BCDLL : 0x4026A09E667F3BCDLL);
  return (int)(__ddiv_rn(__dmul_rn(2.0, (double)s), root));
}

// diff(x, y) -> int difference at block position (x, y).  All `nthreads` threads (a multiple
// of 32, whole warps) must call this together.  Returns this thread's share of the block
// SATD (before the >> (bitdepth-8)); the caller reduces over threads.
template <int TW, int TH, class Diff>
__device__ __forceinline__ unsigned satd_partial(Diff diff, int w, int h, int tid, int nthreads) {
  const int ltx = 31 - __clz(w / TW);   // log2(tiles per block row); all dimensions are powers of two
  const int total = (w / TW) * h;       // tile rows in the block
  const int lane = tid & 31;
  unsigned acc = 0;
  for (int g0 = 0; g0 < total; g0 += nthreads) {
    const int g = g0 + tid;
    const bool active = g < total;
    const int tile = g / TH, r = g % TH;   // TH is a compile-time power of two
    const int tx = (tile & ((1 << ltx) - 1)) * TW, ty = (tile >> ltx) * TH + r;
    int v[TW];
#pragma unroll
    for (int i = 0; i < TW; i++) v[i] = active ? diff(tx + i, ty) : 0;
#pragma unroll
    for (int len = 1; len < TW; len <<= 1)
#pragma unroll
      for (int i = 0; i < TW; i += 2 * len)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const int p = v[j], q = v[j + len];
          v[j] = p + q;
          v[j + len] = p - q;
        }
#pragma unroll
    for (int o = 1; o < TH; o <<= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < TW; i++) {
        const int pv = __shfl_xor_sync(XVCB_FULL, v[i], o);
        v[i] = upper ? pv - v[i] : v[i] + pv;
      }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < TW; i++) s += abs(v[i]);
#pragma unroll
    for (int o = 1; o < TH; o <<= 1) s += __shfl_xor_sync(XVCB_FULL, s, o);
    if (active && r == 0) acc += (unsigned)satd_norm<TW, TH>(s);
  }
  return acc;
}

// Tile choice by block shape, sample_metric.cc:322-387.
template <class Diff>
__device__ __forceinline__ unsigned satd_block_partial(Diff diff, int w, int h, int tid, int nthreads) {
  if (w == 2 || h == 2) return satd_partial<2, 2>(diff, w, h, tid, nthreads);
  if (w == 4 && h == 4) return satd_partial<4, 4>(diff, w, h, tid, nthreads);
  if (h == 4 && w > h) return satd_partial<8, 4>(diff, w, h, tid, nthreads);
  if (w == 4 && h > w) return satd_partial<4, 8>(diff, w, h, tid, nthreads);
  if (w > h) return satd_partial<16, 8>(diff, w, h, tid, nthreads);
  if (w < h) return satd_partial<8, 16>(diff, w, h, tid, nthreads);
  return satd_partial<8, 8>(diff, w, h, tid, nthreads);
}

}  // namespace xvcb
#endif
