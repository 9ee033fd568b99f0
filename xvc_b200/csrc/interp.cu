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

// ---------------------------------------------------------------- affine motion compensation
// InterPrediction::MotionCompAffine (inter_prediction.cc:1044-1136).  The reference walks the
// sub-blocks of a CU serially, accumulating the model MV, and runs MotionCompUniPred on each.
// Here a CTA takes one (CU, component) and every thread takes samples: the MV of the sub-block a
// sample lies in has a closed form (the accumulations are plain integer sums), and the sample is
// filtered directly from the reference picture with the arithmetic of the separable passes
// (first pass stored as int16, inter_prediction.cc:1387-1448) -- sub-blocks of 4x4 luma / 2x2
// chroma samples with their own fractional phases leave nothing to share between neighbours
// but reference samples, which the L1 holds.
struct AffineModel {
  int sbw, sbh;                      // sub-block size in samples of the component
  int dhx, dhy;                      // MV change per sample along x, 1/256 MV units; along y: (-dhy, dhx)
  int mv0x, mv0y;                    // clipped top-left control point
  int min_x, max_x, min_y, max_y;    // ClipMv bounds of the CU
};

__device__ __forceinline__ int affine_subblock_size(int rx, int ry, int ux, int uy, int size, int scale) {
  const int max_len = max(abs(ux - rx), abs(uy - ry));
  if (!max_len) return size;
  int sub = max(1, (size >> 2) / max_len);
  while (size % sub) sub--;
  return max(4, sub) >> scale;
}

__device__ __forceinline__ AffineModel affine_model(const xvcb200_cu &cu, int cs, int pic_w, int pic_h, const int32_t (*mv_raw)[2]) {
  AffineModel m;
  m.min_x = -((64 + 8 + cu.x - 1) * 16); m.max_x = (pic_w + 8 - cu.x - 1) * 16;
  m.min_y = -((64 + 8 + cu.y - 1) * 16); m.max_y = (pic_h + 8 - cu.y - 1) * 16;
  int mv[3][2];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    mv[i][0] = clip3i(mv_raw[i][0], m.min_x, m.max_x);
    mv[i][1] = clip3i(mv_raw[i][1], m.min_y, m.max_y);
  }
  const int w = cu.w >> cs, h = cu.h >> cs;
  m.sbw = affine_subblock_size(mv[0][0], mv[0][1], mv[1][0], mv[1][1], w, cs);
  m.sbh = affine_subblock_size(mv[0][0], mv[0][1], mv[2][0], mv[2][1], h, cs);
  m.dhx = ((mv[1][0] - mv[0][0]) * 256) / w;       // C++ division, toward zero
  m.dhy = ((mv[1][1] - mv[0][1]) * 256) / w;
  m.mv0x = mv[0][0]; m.mv0y = mv[0][1];
  return m;
}

// One sample of MotionCompUniPred at reference position r (centre sample), phases (fx, fy).
template <bool BIPRED, int NTAPS>
__device__ __forceinline__ int mc_sample(const Sample *r, int rs, int fx, int fy, int bitdepth) {
  constexpr int kBack = NTAPS / 2 - 1;
  const int maxv = (1 << bitdepth) - 1;
  const int head = 14 - bitdepth;
  if (fx == 0 && fy == 0) {
    const int s = *r;
    return BIPRED ? (int)(int16_t)((int16_t)(s << head) - (int16_t)8192) : s;
  }
  const int16_t *tx = NTAPS == 4 ? c_chroma_taps[fx] : c_luma_taps[fx];
  const int16_t *ty = NTAPS == 4 ? c_chroma_taps[fy] : c_luma_taps[fy];
  int shift, offset;
  if (fy == 0 || fx == 0) {
    const int16_t *t = fy == 0 ? tx : ty;
    const int step = fy == 0 ? 1 : rs;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < NTAPS; k++) sum += (int)r[(k - kBack) * step] * t[k];
    filter_shift_offset(false, !BIPRED, bitdepth, shift, offset);
    int val = (sum + offset) >> shift;
    if (BIPRED) return (int)(int16_t)val;
    if (fy != 0) val = (int)(int16_t)val;
    return clip3i(val, 0, maxv);
  }
  int sh1, off1;
  filter_shift_offset(false, false, bitdepth, sh1, off1);
  int acc = 0;
#pragma unroll
  for (int j = 0; j < NTAPS; j++) {
    const Sample *row = r + (j - kBack) * rs;
    int sum = 0;
#pragma unroll
    for (int k = 0; k < NTAPS; k++) sum += (int)row[k - kBack] * tx[k];
    acc += (int)(int16_t)((sum + off1) >> sh1) * ty[j];
  }
  filter_shift_offset(true, !BIPRED, bitdepth, shift, offset);
  int val = (acc + offset) >> shift;
  if (BIPRED) return (int)(int16_t)val;
  return clip3i((int)(int16_t)val, 0, maxv);
}

template <bool BIPRED, int NTAPS>
__device__ __forceinline__ int affine_sample(const AffineModel &m, PlaneView rp, int cs, int px, int py, int x, int y, int bitdepth) {
  const int bx = x / m.sbw, by = y / m.sbh;
  const int dvx = -m.dhy, dvy = m.dhx;
  const int hx = m.mv0x * 256 + by * (dvx * m.sbh) + bx * (m.dhx * m.sbw);
  const int hy = m.mv0y * 256 + by * (dvy * m.sbh) + bx * (m.dhy * m.sbw);
  const int mx = clip3i((hx + m.dhx * (m.sbw >> 1) + dvx * (m.sbh >> 1)) >> 8, m.min_x, m.max_x);
  const int my = clip3i((hy + m.dhy * (m.sbw >> 1) + dvy * (m.sbh >> 1)) >> 8, m.min_y, m.max_y);
  const int sh = 4 + cs, mask = (1 << sh) - 1;
  const Sample *r = rp.base + (py + y + (my >> sh)) * rp.pitch + px + x + (mx >> sh);
  return mc_sample<BIPRED, NTAPS>(r, rp.pitch, mx & mask, my & mask, bitdepth);
}

template <int NTAPS>
__device__ __forceinline__ void affine_cu(const xvcb200_cu &cu, const xvcb200_affine_cu &a, int comp, int bitdepth,
                                          const McRefs &refs, PlaneView pred) {
  const int cs = comp ? 1 : 0;
  const int px = cu.x >> cs, py = cu.y >> cs, w = cu.w >> cs, h = cu.h >> cs;
  const bool l0 = cu.ref_idx[0] >= 0, l1 = cu.ref_idx[1] >= 0;
  const int lw = 31 - __clz(w);
  Sample *dst = pred.base + py * pred.pitch + px;
  if (l0 && l1) {
    const PlaneView r0 = refs.r[0][cu.ref_idx[0]].p[comp], r1 = refs.r[1][cu.ref_idx[1]].p[comp];
    const AffineModel m0 = affine_model(cu, cs, refs.r[0][cu.ref_idx[0]].p[0].width, refs.r[0][cu.ref_idx[0]].p[0].height, a.mv[0]);
    const AffineModel m1 = affine_model(cu, cs, refs.r[1][cu.ref_idx[1]].p[0].width, refs.r[1][cu.ref_idx[1]].p[0].height, a.mv[1]);
    const int head = 14 - bitdepth;
    const int shift = (head > 2 ? head : 2) + 1;
    const int offset = (1 << (shift - 1)) + 2 * 8192;
    const int maxv = (1 << bitdepth) - 1;
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
      const int y = i >> lw, x = i & (w - 1);
      const int p0 = affine_sample<true, NTAPS>(m0, r0, cs, px, py, x, y, bitdepth);
      const int p1 = affine_sample<true, NTAPS>(m1, r1, cs, px, py, x, y, bitdepth);
      dst[y * pred.pitch + x] = add_avg_one(p0, p1, offset, shift, maxv);
    }
  } else {
    const int l = l1 ? 1 : 0;
    const PlaneView rp = refs.r[l][cu.ref_idx[l]].p[comp];
    const AffineModel m = affine_model(cu, cs, refs.r[l][cu.ref_idx[l]].p[0].width, refs.r[l][cu.ref_idx[l]].p[0].height, a.mv[l]);
    for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
      const int y = i >> lw, x = i & (w - 1);
      dst[y * pred.pitch + x] = (Sample)affine_sample<false, NTAPS>(m, rp, cs, px, py, x, y, bitdepth);
    }
  }
}

__global__ void __launch_bounds__(128) affine_mc_kernel(const xvcb200_cu *__restrict__ cus, int n_cus,
                                                        const xvcb200_affine_cu *__restrict__ aff, int bitdepth,
                                                        const __grid_constant__ McRefs refs, Pic3 pred) {
  const int i = blockIdx.x / 3, comp = blockIdx.x % 3;
  const xvcb200_affine_cu a = aff[i];
  if (a.cu < 0 || a.cu >= n_cus) return;
  const xvcb200_cu cu = cus[a.cu];
  if (cu.flags & XVCB200_CU_INTRA) return;
  if (cu.ref_idx[0] < 0 && cu.ref_idx[1] < 0) return;
  if (comp == 0) affine_cu<8>(cu, a, comp, bitdepth, refs, pred.p[0]);
  else affine_cu<4>(cu, a, comp, bitdepth, refs, pred.p[comp]);
}

cudaError_t launch_motion_compensate_affine(cudaStream_t s, const xvcb200_cu *d_cus, int n_cus, const xvcb200_affine_cu *d_aff,
                                            int n, int bitdepth, const Pic3 refs[2][5], Pic3 pred) {
  if (n <= 0) return cudaSuccess;
  McRefs r;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) r.r[l][i] = refs[l][i];
  g_launch_count++;
  affine_mc_kernel<<<3 * n, 128, 0, s>>>(d_cus, n_cus, d_aff, bitdepth, r, pred);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- motion compensation with LIC
// InterPrediction::LocalIlluminationComp / DeriveLicParams (inter_prediction.cc:1555-1673).  One CTA
// per (entry, component).  Warp 0 gathers the <= 64 neighbour pairs of each list in use (row above,
// column left: reference picture at the rounded full-pel MV against the current reconstruction),
// reduces the four sums with shuffles, lane 0 derives (scale, offset) with the reference's integer
// recipe; then every thread predicts its samples (mc_sample, as the affine kernel) and applies
// the model -- bi-prediction through FilterCopyBipred + AddAvgBi on the two compensated samples.
struct LicModel { int scale, offset; };

__device__ __forceinline__ int size_to_log2_dev(int size) { int l = 1; while ((1 << l) < size) l++; return l; }

// all 32 lanes of one warp
__device__ __forceinline__ LicModel lic_derive(const xvcb200_cu &cu, const xvcb200_lic_cu &nb, int cs, int bitdepth, PlaneView rp,
                                               int pic_w, int pic_h, PlaneView rec, int mvx, int mvy, int lane) {
  LicModel m; m.scale = 32; m.offset = 0;
  const bool has_above = nb.above_x >= 0, has_left = nb.left_x >= 0;
  if (!has_above && !has_left) return m;
  const int sh = 4 + cs;
  const int px = cu.x >> cs, py = cu.y >> cs, w = cu.w >> cs, h = cu.h >> cs;
  const int fx = (mvx + (1 << (sh - 1))) >> sh, fy = (mvy + (1 << (sh - 1))) >> sh;
  const int step = min(w, h) > 8 ? 2 : 1;
  const int per_side = min(w, h) / step;                 // w / dx == h / dy == min(w, h) / step
  const Sample *rbase = rp.base + py * rp.pitch + px;
  const Sample *sbase = rec.base + py * rec.pitch + px;
  int sum_x = 0, sum_y = 0, sum_xx = 0, sum_xy = 0, nbr = 0;
  if (has_above) {
    int cx = fx, cy = fy;
    clip_mv(nb.above_x, nb.above_y, pic_w, pic_h, cx, cy);       // bounds in 1/16 pel on a full-sample value, as the reference
    const Sample *r = rbase + cx + (cy - 1) * rp.pitch, *q = sbase - rec.pitch;
    const int dx = step * max(1, w / h);
    for (int k = lane; k < per_side; k += 32) {
      const int a = r[k * dx], b = q[k * dx];
      sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b;
    }
    nbr += per_side;
  }
  if (has_left) {
    int cx = fx, cy = fy;
    clip_mv(nb.left_x, nb.left_y, pic_w, pic_h, cx, cy);
    const Sample *r = rbase + cx + cy * rp.pitch - 1, *q = sbase - 1;
    const int dy = step * max(1, h / w);
    for (int k = lane; k < per_side; k += 32) {
      const int a = r[k * dy * rp.pitch], b = q[k * dy * rec.pitch];
      sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b;
    }
    nbr += per_side;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sum_x += __shfl_xor_sync(XVCB_FULL, sum_x, o); sum_y += __shfl_xor_sync(XVCB_FULL, sum_y, o);
    sum_xx += __shfl_xor_sync(XVCB_FULL, sum_xx, o); sum_xy += __shfl_xor_sync(XVCB_FULL, sum_xy, o);
  }
  const int size_shift = size_to_log2_dev(nbr);
  const int base_shift = max(0, bitdepth + size_shift - 15);
  const int avg_x = sum_x >> base_shift, avg_y = sum_y >> base_shift;
  const int xx_offset = sum_xx >> 7;
  const int avg_xy = ((sum_xy + xx_offset) >> (2 * base_shift)) << size_shift;
  const int avg_xx = ((sum_xx + xx_offset) >> (2 * base_shift)) << size_shift;
  const int sd_xy = avg_xy - avg_x * avg_y, sd_xx = avg_xx - avg_x * avg_x;
  const int shift_xx = max(0, (32 - __clz(abs(sd_xx))) - 6);
  const int shift_xy = max(0, shift_xx - 12);
  const int total_shift = 15 - 5 + shift_xx - shift_xy;
  const int sd_xy_s = sd_xy >> shift_xy;
  const int sd_xx_s = clip3i(sd_xx >> shift_xx, 0, 63);
  if (sd_xx_s == 0) return m;
  const int sd_xx_scaled = ((1 << 15) + sd_xx_s / 2) / sd_xx_s;
  m.scale = clip3i((sd_xy_s * sd_xx_scaled) >> total_shift, 0, 128);
  const int offset = (sum_y - ((m.scale * sum_x) >> 5) + (1 << (size_shift - 1))) >> size_shift;
  m.offset = clip3i(offset, -(1 << (bitdepth - 1)), (1 << (bitdepth - 1)) - 1);
  return m;
}

template <int NTAPS>
__device__ __forceinline__ void lic_cu(const xvcb200_cu &cu, const xvcb200_lic_cu &nb, int comp, int bitdepth, const McRefs &refs,
                                       PlaneView rec, PlaneView pred, LicModel *s_model) {
  const int cs = comp ? 1 : 0, sh = 4 + cs, mask = (1 << sh) - 1;
  const int px = cu.x >> cs, py = cu.y >> cs, w = cu.w >> cs, h = cu.h >> cs;
  const int lw = 31 - __clz(w);
  const int maxv = (1 << bitdepth) - 1;
  int mvx[2], mvy[2];
  PlaneView rp[2];
#pragma unroll
  for (int l = 0; l < 2; l++) {
    if (cu.ref_idx[l] < 0) continue;
    rp[l] = refs.r[l][cu.ref_idx[l]].p[comp];
    const PlaneView rl = refs.r[l][cu.ref_idx[l]].p[0];
    mvx[l] = cu.mv[l][0]; mvy[l] = cu.mv[l][1];
    clip_mv(cu.x, cu.y, rl.width, rl.height, mvx[l], mvy[l]);
    if (threadIdx.x < 32) {
      const LicModel m = lic_derive(cu, nb, cs, bitdepth, rp[l], rl.width, rl.height, rec, mvx[l], mvy[l], threadIdx.x);
      if (threadIdx.x == 0) s_model[l] = m;
    }
  }
  __syncthreads();
  Sample *dst = pred.base + py * pred.pitch + px;
  const bool l0 = cu.ref_idx[0] >= 0, l1 = cu.ref_idx[1] >= 0;
  const int head = 14 - bitdepth;
  for (int i = threadIdx.x; i < w * h; i += blockDim.x) {
    const int y = i >> lw, x = i & (w - 1);
    int v[2] = {0, 0};
#pragma unroll
    for (int l = 0; l < 2; l++) {
      if (cu.ref_idx[l] < 0) continue;
      const Sample *r = rp[l].base + (py + y + (mvy[l] >> sh)) * rp[l].pitch + px + x + (mvx[l] >> sh);
      const int p = mc_sample<false, NTAPS>(r, rp[l].pitch, mvx[l] & mask, mvy[l] & mask, bitdepth);
      v[l] = clip3i(((s_model[l].scale * p) >> 5) + s_model[l].offset, 0, maxv);     // AddLinearModel, sample_buffer.h:108-122
    }
    if (l0 && l1) {
      const int a = (int)(int16_t)((int16_t)(v[0] << head) - (int16_t)8192);         // FilterCopyBipred_c, cc:1462-1473
      const int b = (int)(int16_t)((int16_t)(v[1] << head) - (int16_t)8192);
      const int shift = (head > 2 ? head : 2) + 1;
      dst[y * pred.pitch + x] = add_avg_one(a, b, (1 << (shift - 1)) + 2 * 8192, shift, maxv);
    } else {
      dst[y * pred.pitch + x] = (Sample)v[l1 ? 1 : 0];
    }
  }
}

__global__ void __launch_bounds__(128) lic_mc_kernel(const xvcb200_cu *__restrict__ cus, int n_cus, const xvcb200_lic_cu *__restrict__ lic,
                                                     int bitdepth, const __grid_constant__ McRefs refs, Pic3 rec, Pic3 pred) {
  __shared__ LicModel s_model[2];
  const int i = blockIdx.x / 3, comp = blockIdx.x % 3;
  const xvcb200_lic_cu nb = lic[i];
  if (nb.cu < 0 || nb.cu >= n_cus) return;
  const xvcb200_cu cu = cus[nb.cu];
  if (cu.flags & XVCB200_CU_INTRA) return;
  if (cu.ref_idx[0] < 0 && cu.ref_idx[1] < 0) return;
  if (comp == 0) lic_cu<8>(cu, nb, comp, bitdepth, refs, rec.p[0], pred.p[0], s_model);
  else lic_cu<4>(cu, nb, comp, bitdepth, refs, rec.p[comp], pred.p[comp], s_model);
}

cudaError_t launch_motion_compensate_lic(cudaStream_t s, const xvcb200_cu *d_cus, int n_cus, const xvcb200_lic_cu *d_lic, int n,
                                         int bitdepth, const Pic3 refs[2][5], Pic3 rec, Pic3 pred) {
  if (n <= 0) return cudaSuccess;
  McRefs r;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) r.r[l][i] = refs[l][i];
  g_launch_count++;
  lic_mc_kernel<<<3 * n, 128, 0, s>>>(d_cus, n_cus, d_lic, bitdepth, r, rec, pred);
  return cudaGetLastError();
}

}  // namespace xvcb
