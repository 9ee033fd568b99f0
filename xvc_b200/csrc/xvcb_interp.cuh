// Sub-pel interpolation device routines (InterPrediction filters, inter_prediction.cc:1207-1538).
#ifndef XVCB_INTERP_CUH_
#define XVCB_INTERP_CUH_

#include "xvcb_device.cuh"

namespace xvcb {

// kLumaFilterHighPrec / kChromaFilterHighPrec, inter_prediction.cc:55-73, 91-126
static __constant__ int16_t c_luma_taps[16][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},      {-1, 2, -5, 62, 8, -3, 1, 0},
    {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},   {-1, 4, -11, 52, 26, -8, 3, -1},
    {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
    {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
    {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},    {0, 1, -3, 8, 62, -5, 2, -1},
    {0, 1, -2, 4, 63, -3, 1, 0}};
static __constant__ int16_t c_chroma_taps[32][4] = {
    {0, 64, 0, 0},    {-1, 63, 2, 0},   {-2, 62, 4, 0},   {-2, 60, 7, -1},  {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},   {0, 2, 63, -1}};

struct Taps { int t[8]; };

__device__ __forceinline__ Taps load_taps(int chroma, int frac) {
  Taps r;
#pragma unroll
  for (int k = 0; k < 8; k++) r.t[k] = chroma ? (k < 4 ? (int)c_chroma_taps[frac][k] : 0) : (int)c_luma_taps[frac][k];
  return r;
}

// Shift / offset rules of inter_prediction.h:218-255 (internal precision 14, filter precision 6,
// internal offset 8192).  src_short: input is the 14-bit intermediate; dst_sample: output is a
// clipped Sample.
__device__ __forceinline__ void filter_shift_offset(bool src_short, bool dst_sample, int bitdepth, int &shift, int &offset) {
  const int head = 14 - bitdepth;
  if (!src_short && dst_sample) { shift = 6; offset = 32; }
  else if (!src_short) { shift = 6 - head; offset = -(8192 << shift); }
  else if (dst_sample) { shift = 6 + head; offset = (8192 << 6) + (1 << (shift - 1)); }
  else { shift = 6; offset = 0; }
}

// One FIR pass over a w x h block by `nthreads` threads.  KIND as in the reference table:
// 0 H u16->u16, 1 H u16->i16, 2 V u16->u16, 3 V u16->i16, 4 V i16->u16, 5 V i16->i16.
// `src` points at the centre sample of output (0,0) (inter_prediction.cc:1215, 1275).
template <int KIND, int NTAPS, typename ST, typename DT>
__device__ __forceinline__ void fir_pass(int w, int h, int bitdepth, const Taps &taps, const ST *src, int ss, DT *dst,
                                         int ds, int tid, int nthreads) {
  constexpr bool kSrcShort = KIND >= 4;
  constexpr bool kDstSample = KIND == 0 || KIND == 2 || KIND == 4;
  const int step = KIND <= 1 ? 1 : ss;
  int shift, offset;
  filter_shift_offset(kSrcShort, kDstSample, bitdepth, shift, offset);
  const int maxv = (1 << bitdepth) - 1;
  const int lw = 31 - __clz(w);                      // block widths are powers of two
  for (int i = tid; i < w * h; i += nthreads) {
    const int y = i >> lw, x = i & (w - 1);
    const ST *p = src + y * ss + x - (NTAPS / 2 - 1) * step;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < NTAPS; k++) sum += (int)p[k * step] * taps.t[k];
    int val = (sum + offset) >> shift;
    if (kDstSample) {
      if (KIND != 0) val = (int)(int16_t)val;   // vertical variants narrow before ClipBD (cc:1290, 1350)
      dst[y * ds + x] = (DT)clip3i(val, 0, maxv);
    } else {
      dst[y * ds + x] = (DT)(int16_t)val;
    }
  }
}

// InterPrediction::MotionCompUniPred + FilterLuma/FilterChroma (+Bipred variants),
// inter_prediction.cc:1138-1172, 1387-1538, for one block by one CTA.
//   BIPRED = false: pred is Sample (u16);  true: pred is the 14-bit int16 intermediate.
//   tmp: shared scratch of (h + NTAPS - 1) * w int16 for the 2-D case (stride = w, as the
//   reference's filter_buffer_).  Contains __syncthreads(): call from all threads.
template <bool BIPRED, int NTAPS, typename PT>
__device__ __forceinline__ void interp_cta(int w, int h, int bitdepth, int fx, int fy, const Sample *ref, int rs,
                                           PT *pred, int ps, int16_t *tmp, int tid, int nthreads) {
  const int chroma = NTAPS == 4;
  if (fx == 0 && fy == 0) {
    const int shift = 14 - bitdepth;
    const int lw = 31 - __clz(w);
    for (int i = tid; i < w * h; i += nthreads) {
      const int y = i >> lw, x = i & (w - 1);
      const Sample s = ref[y * rs + x];
      if (BIPRED) pred[y * ps + x] = (PT)(int16_t)((int16_t)(s << shift) - (int16_t)8192);   // FilterCopyBipred_c, cc:1462-1473
      else pred[y * ps + x] = (PT)s;
    }
  } else if (fy == 0) {
    fir_pass<BIPRED ? 1 : 0, NTAPS>(w, h, bitdepth, load_taps(chroma, fx), ref, rs, pred, ps, tid, nthreads);
  } else if (fx == 0) {
    fir_pass<BIPRED ? 3 : 2, NTAPS>(w, h, bitdepth, load_taps(chroma, fy), ref, rs, pred, ps, tid, nthreads);
  } else {
    fir_pass<1, NTAPS>(w, h + NTAPS - 1, bitdepth, load_taps(chroma, fx), ref - (NTAPS / 2 - 1) * rs, rs, tmp, w, tid, nthreads);
    __syncthreads();
    fir_pass<BIPRED ? 5 : 4, NTAPS>(w, h, bitdepth, load_taps(chroma, fy), tmp + (NTAPS / 2 - 1) * w, w, pred, ps, tid, nthreads);
  }
}

}  // namespace xvcb
#endif
