// Interpolation kernels: the table-shaped single-block entries of InterPrediction::SimdFunc
// and the batched motion compensation (InterPrediction::MotionCompensation,
// inter_prediction.cc:710-738) over every CU of a picture.
#include "xvcb_interp.cuh"

namespace xvcb {

// ---------------------------------------------------------------- single block (table ABI)
template <int KIND, int NTAPS, typename ST, typename DT>
__global__ void __launch_bounds__(128) block_filter_kernel(int w, int h, int bitdepth, Taps taps, const ST *src, int ss,
                                                           DT *dst, int ds) {
  fir_pass<KIND, NTAPS>(w, h, bitdepth, taps, src, ss, dst, ds, threadIdx.x, 128);
}

cudaError_t launch_block_filter(cudaStream_t s, int kind, int chroma, int w, int h, int bitdepth, const int16_t taps[8],
                                const void *src, int ss, void *dst, int ds) {
  Taps t;
  for (int k = 0; k < 8; k++) t.t[k] = (chroma && k >= 4) ? 0 : taps[k];
  g_launch_count++;
#define XVCB_F(K, ST, DT)                                                                                         \
  case K:                                                                                                         \
    if (chroma) block_filter_kernel<K, 4, ST, DT><<<1, 128, 0, s>>>(w, h, bitdepth, t, (const ST *)src, ss, (DT *)dst, ds); \
    else block_filter_kernel<K, 8, ST, DT><<<1, 128, 0, s>>>(w, h, bitdepth, t, (const ST *)src, ss, (DT *)dst, ds);        \
    break;
  switch (kind) {
    XVCB_F(0, uint16_t, uint16_t)
    XVCB_F(1, uint16_t, int16_t)
    XVCB_F(2, uint16_t, uint16_t)
    XVCB_F(3, uint16_t, int16_t)
    XVCB_F(4, int16_t, uint16_t)
    XVCB_F(5, int16_t, int16_t)
    default: return cudaErrorInvalidValue;
  }
#undef XVCB_F
  return cudaGetLastError();
}

// SampleBuffer::AddAvg, sample_buffer.h:89-106
__device__ __forceinline__ Sample add_avg_one(int a, int b, int offset, int shift, int maxv) {
  return (Sample)clip3i((a + b + offset) >> shift, 0, maxv);
}

__global__ void __launch_bounds__(128) block_add_avg_kernel(int w, int h, int offset, int shift, int bitdepth,
                                                            const int16_t *a, int sa, const int16_t *b, int sb,
                                                            Sample *dst, int ds) {
  const int maxv = (1 << bitdepth) - 1;
  for (int i = threadIdx.x; i < w * h; i += 128) {
    const int y = i / w, x = i - y * w;
    dst[y * ds + x] = add_avg_one(a[y * sa + x], b[y * sb + x], offset, shift, maxv);
  }
}

cudaError_t launch_block_add_avg(cudaStream_t s, int w, int h, int offset, int shift, int bitdepth, const int16_t *a,
                                 int sa, const int16_t *b, int sb, Sample *dst, int ds) {
  g_launch_count++;
  block_add_avg_kernel<<<1, 128, 0, s>>>(w, h, offset, shift, bitdepth, a, sa, b, sb, dst, ds);
  return cudaGetLastError();
}

// FilterCopyBipred_c, inter_prediction.cc:1462-1473
__global__ void __launch_bounds__(128) block_copy_bipred_kernel(int w, int h, int offset, int shift, const Sample *ref,
                                                                int rs, int16_t *pred, int ps) {
  for (int i = threadIdx.x; i < w * h; i += 128) {
    const int y = i / w, x = i - y * w;
    const int16_t v = (int16_t)(ref[y * rs + x] << shift);
    pred[y * ps + x] = (int16_t)(v - (int16_t)offset);
  }
}

cudaError_t launch_block_copy_bipred(cudaStream_t s, int w, int h, int offset, int shift, const Sample *ref, int rs,
                                     int16_t *pred, int ps) {
  g_launch_count++;
  block_copy_bipred_kernel<<<1, 128, 0, s>>>(w, h, offset, shift, ref, rs, pred, ps);
  return cudaGetLastError();
}

template <bool BIPRED, int NTAPS, typename PT>
__global__ void __launch_bounds__(128) block_interp_kernel(int w, int h, int bitdepth, int fx, int fy, const Sample *ref,
                                                           int rs, PT *pred, int ps) {
  __shared__ int16_t tmp[64 * 71];
  interp_cta<BIPRED, NTAPS>(w, h, bitdepth, fx, fy, ref, rs, pred, ps, tmp, threadIdx.x, 128);
}

cudaError_t launch_block_interp(cudaStream_t s, int chroma, int bipred, int w, int h, int bitdepth, int fx, int fy,
                                const Sample *ref, int rs, void *pred, int ps) {
  g_launch_count++;
  if (!bipred) {
    if (chroma) block_interp_kernel<false, 4><<<1, 128, 0, s>>>(w, h, bitdepth, fx, fy, ref, rs, (Sample *)pred, ps);
    else block_interp_kernel<false, 8><<<1, 128, 0, s>>>(w, h, bitdepth, fx, fy, ref, rs, (Sample *)pred, ps);
  } else {
    if (chroma) block_interp_kernel<true, 4><<<1, 128, 0, s>>>(w, h, bitdepth, fx, fy, ref, rs, (int16_t *)pred, ps);
    else block_interp_kernel<true, 8><<<1, 128, 0, s>>>(w, h, bitdepth, fx, fy, ref, rs, (int16_t *)pred, ps);
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- batched motion compensation
struct McRefs { Pic3 r[2][5]; };

// One CTA per (CU, component).  MotionCompensation -> MotionCompRefList -> ClipMv ->
// GetFullpelRef -> MotionCompUniPred (inter_prediction.cc:710-738, 1011-1042, 1174-1205):
// luma pel = mv >> 4, frac = mv & 15; 4:2:0 chroma pel = mv >> 5, frac = mv & 31.
// Bi-prediction: both lists into 14-bit intermediates, then AddAvgBi (:1540-1553).
template <int NTAPS>
__device__ __forceinline__ void mc_cu(const xvcb200_cu &cu, int comp, int bitdepth, const McRefs &refs, PlaneView pred,
                                      int16_t *tmp, int16_t *bi0, int16_t *bi1) {
  const int cs = comp ? 1 : 0;
  const int x = cu.x >> cs, y = cu.y >> cs, w = cu.w >> cs, h = cu.h >> cs;
  const int sh = 4 + cs, mask = (1 << sh) - 1;
  const bool l0 = cu.ref_idx[0] >= 0, l1 = cu.ref_idx[1] >= 0;
  Sample *dst = pred.base + y * pred.pitch + x;
  const int tid = threadIdx.x;
  if (l0 && l1) {
#pragma unroll
    for (int l = 0; l < 2; l++) {
      const PlaneView rp = refs.r[l][cu.ref_idx[l]].p[comp];
      const PlaneView rl = refs.r[l][cu.ref_idx[l]].p[0];
      int mx = cu.mv[l][0], my = cu.mv[l][1];
      clip_mv(cu.x, cu.y, rl.width, rl.height, mx, my);
      const Sample *r = rp.base + (y + (my >> sh)) * rp.pitch + x + (mx >> sh);
      interp_cta<true, NTAPS>(w, h, bitdepth, mx & mask, my & mask, r, rp.pitch, l ? bi1 : bi0, 64, tmp, tid, 128);
      __syncthreads();
    }
    const int head = 14 - bitdepth;
    const int shift = (head > 2 ? head : 2) + 1;
    const int offset = (1 << (shift - 1)) + 2 * 8192;
    const int maxv = (1 << bitdepth) - 1;
    const int lw = 31 - __clz(w);
    for (int i = tid; i < w * h; i += 128) {
      const int yy = i >> lw, xx = i & (w - 1);
      dst[yy * pred.pitch + xx] = add_avg_one(bi0[yy * 64 + xx], bi1[yy * 64 + xx], offset, shift, maxv);
    }
  } else {
    const int l = l1 ? 1 : 0;
    const PlaneView rp = refs.r[l][cu.ref_idx[l]].p[comp];
    const PlaneView rl = refs.r[l][cu.ref_idx[l]].p[0];
    int mx = cu.mv[l][0], my = cu.mv[l][1];
    clip_mv(cu.x, cu.y, rl.width, rl.height, mx, my);
    const Sample *r = rp.base + (y + (my >> sh)) * rp.pitch + x + (mx >> sh);
    interp_cta<false, NTAPS>(w, h, bitdepth, mx & mask, my & mask, r, rp.pitch, dst, pred.pitch, tmp, tid, 128);
  }
}

__global__ void __launch_bounds__(128) motion_compensate_kernel(const xvcb200_cu *__restrict__ cus, int n, int bitdepth,
                                                                const __grid_constant__ McRefs refs, Pic3 pred) {
  __shared__ int16_t tmp[64 * 71];
  __shared__ int16_t bi0[64 * 64];
  __shared__ int16_t bi1[64 * 64];
  const int i = blockIdx.x / 3, comp = blockIdx.x % 3;
  const xvcb200_cu cu = cus[i];
  if (cu.flags & XVCB200_CU_INTRA) return;
  if (cu.ref_idx[0] < 0 && cu.ref_idx[1] < 0) return;
  if (comp == 0) mc_cu<8>(cu, comp, bitdepth, refs, pred.p[0], tmp, bi0, bi1);
  else mc_cu<4>(cu, comp, bitdepth, refs, pred.p[comp], tmp, bi0, bi1);
}

cudaError_t launch_motion_compensate(cudaStream_t s, const xvcb200_cu *d_cus, int n, int bitdepth,
                                     const Pic3 refs[2][5], Pic3 pred) {
  if (n <= 0) return cudaSuccess;
  McRefs r;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) r.r[l][i] = refs[l][i];
  g_launch_count++;
  motion_compensate_kernel<<<3 * n, 128, 0, s>>>(d_cus, n, bitdepth, r, pred);
  return cudaGetLastError();
}

}  // namespace xvcb
