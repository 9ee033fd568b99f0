// Intra prediction (xvc_common_lib/intra_prediction.cc), unrestricted 67-mode configuration:
//
//   intra_ref_kernel        IntraPrediction::FillReferenceState (:128-147) for one block: reference
//                           samples from the neighbourhood with the reference's substitution rules
//                           (ComputeRefSamples :709-851) + the [1 2 1] smoothing (FilterRefSamples
//                           :853-876)
//   intra_predict_kernel    IntraPrediction::Predict (:81-126): planar, DC, angular
//   intra_satd_scan_kernel  the first loop of IntraSearch::DetermineSlowIntraModes
//                           (xvc_enc_lib/intra_search.cc:185-216) for a batch of luma blocks: SATD
//                           of every mode's prediction against the original
//
// Every prediction sample has a closed form in the reference samples (the reference's flip /
// project / transpose passes only move data), so nothing is materialised: a warp evaluates a mode
// by computing the prediction sample inside the SATD's difference functor.
#include "xvcb_satd.cuh"

namespace xvcb {

constexpr int kRS = XVCB200_INTRA_REF_STRIDE;

// kAngleTableExt / kInvAngleTableExt, intra_prediction.cc:38-50
__constant__ int8_t c_intra_angle[33] = {-32, -29, -26, -23, -21, -19, -17, -15, -13, -11, -9, -7, -5, -3, -2, -1, 0,
                                         1,   2,   3,   5,   7,   9,   11,  13,  15,  17,  19, 21, 23, 26, 29, 32};
__constant__ int16_t c_intra_inv_angle[16] = {8192, 4096, 2731, 1638, 1170, 910, 745, 630, 546, 482, 431, 390, 356, 315, 282, 256};
// kFilterRefThresholdExt, intra_prediction.cc:351-353 (index = mean log2 size)
__constant__ int8_t c_intra_filter_thr[8] = {0, 20, 20, 14, 2, 0, 20, 0};

struct IntraNeighbours { int above_left, above, above_right, left, below_left; };

// Reference samples of one block, cooperatively by the CTA.  above[0] = corner sample,
// above[1 + i] = row above; left[i * left_stride] = column to the left.  Only samples the
// flags declare available are read.  line: 5 * 64 samples of shared memory scratch.
__device__ void intra_build_refs(int w, int h, int bitdepth, IntraNeighbours nb, const Sample *above, const Sample *left,
                                 int left_stride, Sample *ref, Sample *line, int tid, int nthreads) {
  const Sample dc = (Sample)(1 << (bitdepth - 1));
  const int n = w + h;
  if (!nb.above_left && !nb.above && !nb.left && nb.above_right <= 0 && nb.below_left <= 0) {
    for (int i = tid; i <= n; i += nthreads) ref[i] = dc;
    for (int i = tid; i < n; i += nthreads) ref[kRS + i] = dc;
    __syncthreads();
    return;
  }
  // line[n-1-y] = left sample y (below-left beyond y = h), line[n .. n+w) = the corner,
  // line[n+w+x] = above sample x (above-right beyond x = w)
  for (int i = tid; i < 2 * n + w; i += nthreads) {
    Sample v = dc;
    if (i < n) {
      const int y = n - 1 - i;
      if (nb.left) {
        if (y < h) v = left[y * left_stride];
        else if (nb.below_left > 0) v = left[min(y, h + nb.below_left - 1) * left_stride];   // beyond the picture: last one
      }
    } else if (i < n + w) {
      if (nb.above_left) v = above[0];
    } else {
      const int x = i - n - w;
      if (nb.above) {
        if (x < w) v = above[1 + x];
        else if (nb.above_right > 0) v = above[1 + min(x, w + nb.above_right - 1)];
      }
    }
    line[i] = v;
  }
  __syncthreads();
  if (tid == 0) {      // substitution of what is missing, from the bottom-left end upwards (:806-839)
    if (nb.below_left <= 0) {
      const Sample v = nb.left ? line[w] : (nb.above_left ? line[n] : (nb.above ? line[n + w] : line[n + 2 * w]));
      for (int i = 0; i < w; i++) line[i] = v;
    }
    if (!nb.left)
      for (int i = 0; i < h; i++) line[w + i] = line[w - 1];
    if (!nb.above_left)
      for (int i = 0; i < w; i++) line[n + i] = line[n - 1];
    if (!nb.above)
      for (int i = 0; i < w; i++) line[n + w + i] = line[n + w - 1];
    if (nb.above_right <= 0)
      for (int i = 0; i < h; i++) line[n + 2 * w + i] = line[n + 2 * w - 1];
  }
  __syncthreads();
  for (int x = tid; x <= n; x += nthreads) ref[x] = line[n + w - 1 + x];
  for (int y = tid; y < n; y += nthreads) ref[kRS + y] = line[n - 1 - y];
  __syncthreads();
}

__device__ void intra_filter_refs(int w, int h, const Sample *src, Sample *dst, int tid, int nthreads) {
  const int n = w + h;
  for (int i = tid; i <= n; i += nthreads) {
    int v;
    if (i == 0) v = (2 * src[0] + src[1] + src[kRS] + 2) >> 2;
    else if (i == n) v = src[n];
    else v = (2 * src[i] + src[i - 1] + src[i + 1] + 2) >> 2;
    dst[i] = (Sample)v;
  }
  for (int i = tid; i < n; i += nthreads) {
    int v;
    if (i == 0) v = (2 * src[kRS] + src[0] + src[kRS + 1] + 2) >> 2;
    else if (i == n - 1) v = src[kRS + n - 1];
    else v = (2 * src[kRS + i] + src[kRS + i - 1] + src[kRS + i + 1] + 2) >> 2;
    dst[kRS + i] = (Sample)v;
  }
  __syncthreads();
}

// One mode of one block, prepared once per warp (all lanes call prepare together).
struct IntraMode {
  const Sample *ref;           // reference samples this mode reads (smoothed or not)
  int mode, w, h, maxv, post;
  int dc;                      // mode 1
  int horizontal, angle, inv_angle;
  int lw, lh;

  __device__ __forceinline__ void prepare(int mode_, int w_, int h_, int bitdepth, int luma, const Sample *ref_samples,
                                          const Sample *ref_filtered, int lane) {
    mode = mode_; w = w_; h = h_; maxv = (1 << bitdepth) - 1;
    lw = 31 - __clz(w); lh = 31 - __clz(h);
    post = luma && w <= 16 && h <= 16;
    // UseFilteredRefSamples, :342-364
    const int dist = min(abs(mode - 18), abs(mode - 50));
    const bool filtered = luma && ref_filtered != nullptr && dist > c_intra_filter_thr[(lw + lh) >> 1];
    ref = filtered ? ref_filtered : ref_samples;
    dc = 0; horizontal = 0; angle = 0; inv_angle = 0;
    if (mode == 1) {           // PredIntraDC always reads the unfiltered samples (:113-116)
      ref = ref_samples;
      int sum = 0;
      for (int i = lane; i < w; i += 32) sum += ref[1 + i];
      for (int i = lane; i < h; i += 32) sum += ref[kRS + i];
      sum = warp_sum(sum);
      dc = (sum + ((w + h) >> 1)) / (w + h);
    } else if (mode >= 2) {
      horizontal = mode < 34;
      const int angle_offset = horizontal ? 18 - mode : mode - 50;
      angle = c_intra_angle[16 + angle_offset];
      if (angle < 0) inv_angle = c_intra_inv_angle[-angle_offset - 1];
    }
  }

  __device__ __forceinline__ int sample(int x, int y) const {
    if (mode == 0) {           // PlanarPred, :402-424
      const int shift = lw + lh + 1;
      const int hor = (h - 1 - y) * ref[1 + x] + (y + 1) * ref[kRS + h];
      const int ver = (w - 1 - x) * ref[kRS + y] + (x + 1) * ref[1 + w];
      return ((hor << lw) + (ver << lh) + (1 << (shift - 1))) >> shift;
    }
    if (mode == 1) {           // PredIntraDC, :366-400
      if (!post || (x > 0 && y > 0)) return dc;
      if (x == 0 && y == 0) return (ref[1] + ref[kRS] + 2 * dc + 2) >> 2;
      return ((x == 0 ? ref[kRS + y] : ref[1 + x]) + 3 * dc + 2) >> 2;
    }
    // AngularPred, :426-558: horizontal-class modes = vertical-class prediction of the transposed
    // block with the two reference edges exchanged
    const int px = horizontal ? y : x, py = horizontal ? x : y;
    const Sample *main_edge = horizontal ? ref + kRS : ref + 1;
    const Sample *side_edge = horizontal ? ref + 1 : ref + kRS;
    const int corner = ref[0];
    if (angle == 0) {
      int v = main_edge[px];
      if (post && px == 0) v = clip3i((int)(int16_t)(v + ((side_edge[py] - corner) >> 1)), 0, maxv);
      return v;
    }
    const int sum = (py + 1) * angle, off = sum >> 5, wgt = sum & 31;
    int s[2];
#pragma unroll
    for (int t = 0; t < 2; t++) {
      const int j = off + px + t;          // position on the prediction line, -1 = the corner
      if (j >= 0) s[t] = main_edge[j];
      else if (j == -1) s[t] = corner;
      else s[t] = side_edge[((128 + (-1 - j) * inv_angle) >> 8) - 1];     // projected from the other edge, :478-487
    }
    int v = wgt ? ((32 - wgt) * s[0] + wgt * s[1] + 16) >> 5 : s[0];
    if (post && px == 0 && (angle == 1 || angle == -1)) v = clip3i((int)(int16_t)(v + ((side_edge[py] - corner) >> 2)), 0, maxv);
    return v;
  }
};

// ---------------------------------------------------------------- single block (table-shaped ABI)
// edges: [0 .. w+h] corner + row above, [w+h+1 ..] the column to the left (tightly staged by the host)
__global__ void __launch_bounds__(128) intra_ref_kernel(int w, int h, int bitdepth, IntraNeighbours nb, const Sample *edges,
                                                        Sample *ref, Sample *filt) {
  __shared__ Sample s_line[5 * 64];
  __shared__ Sample s_ref[2 * kRS];
  intra_build_refs(w, h, bitdepth, nb, edges, edges + w + h + 1, 1, s_ref, s_line, threadIdx.x, 128);
  for (int i = threadIdx.x; i < 2 * kRS; i += 128) {
    const bool used = i <= w + h || (i >= kRS && i < kRS + w + h);
    ref[i] = used ? s_ref[i] : (Sample)0;
  }
  if (filt != nullptr) {
    __shared__ Sample s_filt[2 * kRS];
    intra_filter_refs(w, h, s_ref, s_filt, threadIdx.x, 128);
    for (int i = threadIdx.x; i < 2 * kRS; i += 128) {
      const bool used = i <= w + h || (i >= kRS && i < kRS + w + h);
      filt[i] = used ? s_filt[i] : (Sample)0;
    }
  }
}

__global__ void __launch_bounds__(128) intra_predict_kernel(int mode, int w, int h, int bitdepth, int luma, const Sample *ref,
                                                            const Sample *filt, Sample *out, int os) {
  __shared__ Sample s_ref[2 * kRS], s_filt[2 * kRS];
  for (int i = threadIdx.x; i < 2 * kRS; i += 128) {
    s_ref[i] = ref[i];
    s_filt[i] = filt ? filt[i] : (Sample)0;
  }
  __syncthreads();
  IntraMode m;
  m.prepare(mode, w, h, bitdepth, luma, s_ref, filt ? s_filt : nullptr, threadIdx.x & 31);
  const int lw = 31 - __clz(w);
  for (int i = threadIdx.x; i < w * h; i += 128) {
    const int y = i >> lw, x = i & (w - 1);
    out[y * os + x] = (Sample)m.sample(x, y);
  }
}

cudaError_t launch_intra_ref(cudaStream_t s, int w, int h, int bitdepth, const int nb[5], const Sample *d_edges, Sample *d_ref,
                             Sample *d_filt) {
  g_launch_count++;
  IntraNeighbours n{nb[0], nb[1], nb[2], nb[3], nb[4]};
  intra_ref_kernel<<<1, 128, 0, s>>>(w, h, bitdepth, n, d_edges, d_ref, d_filt);
  return cudaGetLastError();
}

cudaError_t launch_intra_predict(cudaStream_t s, int mode, int w, int h, int bitdepth, int luma, const Sample *d_ref,
                                 const Sample *d_filt, Sample *d_out, int os) {
  g_launch_count++;
  intra_predict_kernel<<<1, 128, 0, s>>>(mode, w, h, bitdepth, luma, d_ref, d_filt, d_out, os);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- batched mode scan
// One CTA per (job, quarter of the modes); a warp per mode.  Reference samples come from the
// luma plane of `src` (the reconstruction in coding order, or any picture for a pre-analysis),
// the original block is staged in shared memory once.
constexpr int kScanSplit = 4;
__global__ void __launch_bounds__(128) intra_satd_scan_kernel(const xvcb200_intra_job *__restrict__ jobs, int bitdepth,
                                                              PlaneView orig, PlaneView src, uint32_t *__restrict__ satd) {
  __shared__ Sample s_line[5 * 64];
  __shared__ Sample s_ref[2 * kRS], s_filt[2 * kRS];
  __shared__ Sample s_org[64 * 64];
  const xvcb200_intra_job job = jobs[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w = job.w, h = job.h;
  const IntraNeighbours nb{job.has_above_left, job.has_above, job.above_right, job.has_left, job.below_left};
  const Sample *blk = src.base + job.y * src.pitch + job.x;
  intra_build_refs(w, h, bitdepth, nb, blk - src.pitch - 1, blk - 1, src.pitch, s_ref, s_line, tid, 128);
  intra_filter_refs(w, h, s_ref, s_filt, tid, 128);
  const int lw = 31 - __clz(w);
  for (int i = tid; i < w * h; i += 128) {
    const int y = i >> lw, x = i & (w - 1);
    s_org[i] = orig.base[(job.y + y) * orig.pitch + job.x + x];
  }
  __syncthreads();
  for (int mode = blockIdx.y * 4 + warp; mode < XVCB200_INTRA_NUM_MODES; mode += 4 * kScanSplit) {
    IntraMode m;
    m.prepare(mode, w, h, bitdepth, 1, s_ref, s_filt, lane);
    auto diff = [&](int x, int y) { return (int)s_org[(y << lw) + x] - m.sample(x, y); };
    unsigned v = satd_block_partial(diff, w, h, lane, 32);
    v = warp_sum(v);
    if (lane == 0) satd[(size_t)blockIdx.x * XVCB200_INTRA_NUM_MODES + mode] = v >> (bitdepth - 8);
  }
}

cudaError_t launch_intra_satd_scan(cudaStream_t s, const xvcb200_intra_job *d_jobs, int n, int bitdepth, PlaneView orig,
                                   PlaneView src, uint32_t *d_satd) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  intra_satd_scan_kernel<<<dim3(n, kScanSplit), 128, 0, s>>>(d_jobs, bitdepth, orig, src, d_satd);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- chroma from luma (LM chroma)
// IntraPrediction::PredLmChroma / RescaleLuma (4:2:0) / DeriveLmParams (intra_prediction.cc:560-686,
// 873-913).  One CTA per (block, chroma component): all threads reduce the CU's reconstructed luma
// (plus the row above / column left when the CU is not at the picture border) to chroma resolution
// into shared memory, warp 0 gathers the <= 64 (reduced luma, reconstructed chroma) neighbour pairs,
// reduces the four sums with shuffles and lane 0 runs the reference's integer recipe, then all
// threads apply the model to the reduced luma of the block.  Shift counts are masked to 5 bits where
// the reference shifts by a computed count (the x86 behaviour the oracle is pinned to).
__device__ __forceinline__ int log2_floor_dev(int x) { return x > 1 ? 31 - __clz(x) : 0; }

__global__ void __launch_bounds__(128) intra_lm_chroma_kernel(const xvcb200_intra_job *__restrict__ jobs, int bitdepth, PlaneView luma,
                                                              PlaneView cu_plane, PlaneView cv_plane, PlaneView pu_plane, PlaneView pv_plane) {
  constexpr int SS = 33;
  __shared__ uint16_t sub[SS * SS];
  __shared__ int s_model[3];
  const xvcb200_intra_job j = jobs[blockIdx.x];
  const int comp = 1 + blockIdx.y;
  const PlaneView cp = comp == 1 ? cu_plane : cv_plane, pp = comp == 1 ? pu_plane : pv_plane;
  const int cw = j.w >> 1, ch = j.h >> 1;
  const bool has_above = j.y > 0, has_left = j.x > 0;
  const Sample *lbase = luma.base + j.y * luma.pitch + j.x;
  const Sample *cbase = cp.base + (j.y >> 1) * cp.pitch + (j.x >> 1);
  const int tid = threadIdx.x;
  for (int i = tid; i < (ch + 1) * (cw + 1); i += blockDim.x) {
    const int y = i / (cw + 1) - 1, x = i - (y + 1) * (cw + 1) - 1;
    if ((y < 0 && !has_above) || (x < 0 && !has_left)) continue;
    const Sample *s0 = lbase + (2 * y) * luma.pitch, *s1 = s0 + luma.pitch;
    int v;
    if (!has_left && x == 0) v = (s0[0] + s1[0] + 1) >> 1;
    else v = (s0[2 * x - 1] + 2 * s0[2 * x] + s0[2 * x + 1] + s1[2 * x - 1] + 2 * s1[2 * x] + s1[2 * x + 1] + 4) >> 3;
    sub[(y + 1) * SS + x + 1] = (uint16_t)v;
  }
  __syncthreads();
  if (tid < 32) {
    int scale = 0, offset = 1 << (bitdepth - 1), shift = 0;
    if (has_above || has_left) {
      int sum_x = 0, sum_y = 0, sum_xx = 0, sum_xy = 0, nbr = 0;
      if (has_above) {
        const int dx = has_left ? max(1, cw / ch) : 1, cnt = (cw + dx - 1) / dx;
        for (int k = tid; k < cnt; k += 32) {
          const int a = sub[k * dx + 1], b = cbase[-cp.pitch + k * dx];
          sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b;
        }
        nbr += cnt;
      }
      if (has_left) {
        const int dy = has_above ? max(1, ch / cw) : 1, cnt = (ch + dy - 1) / dy;
        for (int k = tid; k < cnt; k += 32) {
          const int a = sub[(k * dy + 1) * SS], b = cbase[k * dy * cp.pitch - 1];
          sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b;
        }
        nbr += cnt;
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        sum_x += __shfl_xor_sync(XVCB_FULL, sum_x, o); sum_y += __shfl_xor_sync(XVCB_FULL, sum_y, o);
        sum_xx += __shfl_xor_sync(XVCB_FULL, sum_xx, o); sum_xy += __shfl_xor_sync(XVCB_FULL, sum_xy, o);
      }
      int size_shift = 1;
      while ((1 << size_shift) < nbr) size_shift++;
      if (size_shift > 15 - bitdepth) {
        const int sh = size_shift + bitdepth - 15, r = 1 << (sh - 1);
        sum_x = (sum_x + r) >> sh; sum_y = (sum_y + r) >> sh; sum_xx = (sum_xx + r) >> sh; sum_xy = (sum_xy + r) >> sh;
        size_shift -= sh;
      }
      const int avg_x = sum_x >> size_shift, avg_y = sum_y >> size_shift;
      const int x_frac = sum_x & ((1 << size_shift) - 1), y_frac = sum_y & ((1 << size_shift) - 1);
      const int sd_xy = sum_xy - ((avg_x * avg_y) << size_shift) - avg_x * y_frac - avg_y * x_frac;
      const int sd_xx = sum_xx - ((avg_x * avg_x) << size_shift) - 2 * avg_x * x_frac;
      const int shift_xy = sd_xy == 0 ? 0 : max(0, log2_floor_dev(abs(sd_xy)) - bitdepth + 2);
      const int shift_xx = sd_xx == 0 ? 0 : max(0, log2_floor_dev(abs(sd_xx)) - 5);
      const int sd_xy_s = sd_xy >> shift_xy, sd_xx_s = sd_xx >> shift_xx;
      const int total_shift = bitdepth + shift_xx + 4 + 7 - 13 - shift_xy;
      if (sd_xx_s < 32) {
        offset = avg_y;
      } else {
        int sc = (int)((uint32_t)sd_xy_s * (uint32_t)(((1 << (bitdepth + 4)) + sd_xx_s / 2) / sd_xx_s));
        sc = sc >> (total_shift & 31);
        sc = clip3i(sc, -256, 255) * 128;
        const int base_shift = log2_floor_dev(abs(sc) + (sc < 0 ? -1 : 0)) - (sc ? 5 : 0);
        shift = 13 - base_shift;
        scale = sc >> base_shift;
        offset = avg_y - ((scale * avg_x) >> (shift & 31));
      }
    }
    if (tid == 0) { s_model[0] = scale; s_model[1] = offset; s_model[2] = shift; }
  }
  __syncthreads();
  const int scale = s_model[0], offset = s_model[1], shift = s_model[2] & 31, maxv = (1 << bitdepth) - 1;
  Sample *dst = pp.base + (j.y >> 1) * pp.pitch + (j.x >> 1);
  for (int i = tid; i < cw * ch; i += blockDim.x) {
    const int y = i / cw, x = i - y * cw;
    dst[y * pp.pitch + x] = (Sample)clip3i(((scale * (int)sub[(y + 1) * SS + x + 1]) >> shift) + offset, 0, maxv);
  }
}

cudaError_t launch_intra_lm_chroma(cudaStream_t s, const xvcb200_intra_job *d_jobs, int n, int bitdepth, PlaneView luma, PlaneView rec_u,
                                   PlaneView rec_v, PlaneView pred_u, PlaneView pred_v) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  intra_lm_chroma_kernel<<<dim3(n, 2), 128, 0, s>>>(d_jobs, bitdepth, luma, rec_u, rec_v, pred_u, pred_v);
  return cudaGetLastError();
}

}  // namespace xvcb
