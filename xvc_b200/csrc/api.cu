// C-ABI layer of libxvc_b200.so (include/xvc_b200.h): context and picture-slot management,
// host<->device staging for the table-shaped entry points, and the per-picture pipeline.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <utility>

#include "xvcb_internal.h"

namespace xvcb {

cudaError_t launch_tq_reconstruct_classes(cudaStream_t s, xvcb200_cu *d_cus, const int *d_tu_list,
                                          const int class_count[7][7], const int class_offset[7][7],
                                          const TqParams &p, Pic3 orig, Pic3 pred, Pic3 rec, int16_t *const lev[3],
                                          const int lev_pitch[3], xvcb200_tu_result *d_res, cudaStream_t *side,
                                          cudaEvent_t *side_ev, int n_side, cudaEvent_t fork_ev);

static thread_local int t_last_error = XVCB200_OK;
static thread_local char t_last_error_str[256] = "";

void set_last_error(int code, const char *what) {
  if (t_last_error == XVCB200_OK) {
    t_last_error = code;
    snprintf(t_last_error_str, sizeof(t_last_error_str), "%s", what);
  }
}

void *xvcb_ctx_impl::scratch(size_t bytes) {
  if (bytes > scratch_bytes) {
    if (d_scratch) { cudaStreamSynchronize(stream); cudaFree(d_scratch); d_scratch = nullptr; }
    size_t want = bytes + bytes / 4 + 4096;
    if (!check(cudaMalloc(&d_scratch, want), "cudaMalloc(scratch)")) { scratch_bytes = 0; return nullptr; }
    scratch_bytes = want;
  }
  return d_scratch;
}
void *xvcb_ctx_impl::scratch2(size_t bytes) {
  if (bytes > scratch2_bytes) {
    if (d_scratch2) { cudaStreamSynchronize(stream); cudaFree(d_scratch2); d_scratch2 = nullptr; }
    size_t want = bytes + bytes / 4 + 4096;
    if (!check(cudaMalloc(&d_scratch2, want), "cudaMalloc(scratch2)")) { scratch2_bytes = 0; return nullptr; }
    scratch2_bytes = want;
  }
  return d_scratch2;
}
void *xvcb_ctx_impl::pinned(size_t bytes) {
  if (bytes > pinned_bytes) {
    if (h_pinned) { cudaStreamSynchronize(stream); cudaFreeHost(h_pinned); h_pinned = nullptr; }
    size_t want = bytes + bytes / 4 + 4096;
    if (!check(cudaMallocHost(&h_pinned, want), "cudaMallocHost")) { pinned_bytes = 0; return nullptr; }
    pinned_bytes = want;
  }
  return h_pinned;
}

// ------------------------------------------------------------------------------------------
// per-thread staging context for the table-shaped (host pointer, one block) entry points
// ------------------------------------------------------------------------------------------
struct Leaf {
  cudaStream_t stream = nullptr;
  uint8_t *d = nullptr;       // device staging
  uint8_t *h = nullptr;       // pinned host staging
  static constexpr size_t kBytes = 1 << 20;
  bool ok = false;
  bool init() {
    if (ok) return true;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
      set_last_error(XVCB200_NO_DEVICE, "no CUDA device: xvc_b200 has no CPU fallback");
      return false;
    }
    if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&d, kBytes) != cudaSuccess || cudaMallocHost(&h, kBytes) != cudaSuccess) {
      set_last_error(XVCB200_CUDA_ERROR, cudaGetErrorString(cudaGetLastError()));
      return false;
    }
    ok = true;
    return true;
  }
  bool done(cudaError_t e) {
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { set_last_error(XVCB200_CUDA_ERROR, cudaGetErrorString(e)); return false; }
    return true;
  }
};
static thread_local Leaf t_leaf;

// copies a strided host block (rows x cols elements of esz bytes) into pinned staging at byte
// offset `off`, tightly packed, and returns the offset after it (16-byte aligned)
static size_t stage_in(Leaf &L, size_t off, const void *src, ptrdiff_t stride, int rows, int cols, int esz) {
  const uint8_t *s = static_cast<const uint8_t *>(src);
  for (int y = 0; y < rows; y++) memcpy(L.h + off + (size_t)y * cols * esz, s + (ptrdiff_t)y * stride * esz, (size_t)cols * esz);
  return (off + (size_t)rows * cols * esz + 15) & ~(size_t)15;
}
static void stage_out(Leaf &L, size_t off, void *dst, ptrdiff_t stride, int rows, int cols, int esz) {
  uint8_t *d = static_cast<uint8_t *>(dst);
  for (int y = 0; y < rows; y++) memcpy(d + (ptrdiff_t)y * stride * esz, L.h + off + (size_t)y * cols * esz, (size_t)cols * esz);
}

static uint64_t leaf_metric(int metric, int bitdepth, int w, int h, int a_short, int b_short, const void *a,
                            ptrdiff_t sa, const void *b, ptrdiff_t sb, int rows_used_step) {
  Leaf &L = t_leaf;
  if (!L.init()) return ~0ull;
  (void)rows_used_step;
  size_t off_a = 0;
  size_t off_b = stage_in(L, off_a, a, sa, h, w, 2);
  size_t off_o = stage_in(L, off_b, b, sb, h, w, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_metric(L.stream, metric, bitdepth, w, h, a_short, b_short, L.d + off_a, w, L.d + off_b, w,
                                      reinterpret_cast<unsigned long long *>(L.d + off_o));
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, 8, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return ~0ull;
  uint64_t r;
  memcpy(&r, L.h + off_o, 8);
  return r;
}

}  // namespace xvcb

using namespace xvcb;  // NOLINT

extern "C" {

int xvcb200_last_error(void) { return t_last_error; }
const char *xvcb200_last_error_string(void) { return t_last_error_str; }
void xvcb200_clear_error(void) { t_last_error = XVCB200_OK; t_last_error_str[0] = 0; }
uint64_t xvcb200_launch_count(void) { return g_launch_count.load(); }
const char *xvcb200_version(void) { return "xvc_b200 0.2 (sm_100a)"; }
int xvcb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int xvcb200_abi_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(xvcb200_cu);
    case 1: return (int)sizeof(xvcb200_me_job);
    case 2: return (int)sizeof(xvcb200_me_result);
    case 3: return (int)sizeof(xvcb200_fullsearch_job);
    case 4: return (int)sizeof(xvcb200_tu_result);
    case 5: return (int)sizeof(xvcb200_picture_params);
    case 6: return (int)sizeof(xvcb200_plane_geom);
    case 7: return (int)sizeof(xvcb200_qp);
    case 8: return (int)sizeof(xvcb200_intra_job);
    case 9: return (int)sizeof(xvcb200_affine_cu);
    case 10: return (int)sizeof(xvcb200_lic_cu);
    case 11: return (int)sizeof(xvcb200_tu_mode);
    case 12: return (int)sizeof(xvcb200_partition_params);
    default: return -1;
  }
}

// ---------------------------------------------------------------- (A) metrics
int xvcb200_sad_sample_sample(int w, int h, const uint16_t *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  return (int)leaf_metric(-1, 8, w, h, 0, 0, a, sa, b, sb, 1);
}
int xvcb200_sad_short_sample(int w, int h, const int16_t *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  return (int)leaf_metric(-1, 8, w, h, 1, 0, a, sa, b, sb, 1);
}
uint64_t xvcb200_ssd_sample_sample(int w, int h, const uint16_t *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  return leaf_metric(-2, 8, w, h, 0, 0, a, sa, b, sb, 1);
}
uint64_t xvcb200_ssd_short_sample(int w, int h, const int16_t *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  return leaf_metric(-2, 8, w, h, 1, 0, a, sa, b, sb, 1);
}
uint64_t xvcb200_ssd_short_short(int w, int h, const int16_t *a, ptrdiff_t sa, const int16_t *b, ptrdiff_t sb) {
  return leaf_metric(-2, 8, w, h, 1, 1, a, sa, b, sb, 1);
}
uint64_t xvcb200_compare_sample_sample(int metric, int bitdepth, int w, int h, const uint16_t *a, ptrdiff_t sa,
                                       const uint16_t *b, ptrdiff_t sb) {
  return leaf_metric(metric, bitdepth, w, h, 0, 0, a, sa, b, sb, 1);
}
uint64_t xvcb200_compare_short_sample(int metric, int bitdepth, int w, int h, const int16_t *a, ptrdiff_t sa,
                                      const uint16_t *b, ptrdiff_t sb) {
  return leaf_metric(metric, bitdepth, w, h, 1, 0, a, sa, b, sb, 1);
}

// ---------------------------------------------------------------- (A) filters
static void leaf_filter(int kind, int chroma, int w, int h, int bitdepth, const int16_t *filter, const void *src,
                        ptrdiff_t ss, void *dst, ptrdiff_t ds) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  const int n = chroma ? 4 : 8, before = n / 2 - 1, after = n / 2;
  const bool hor = kind <= 1;
  // stage exactly the apron the reference kernel reads
  const int x0 = hor ? -before : 0, y0 = hor ? 0 : -before;
  const int cols = w + (hor ? before + after : 0), rows = h + (hor ? 0 : before + after);
  const uint8_t *s = static_cast<const uint8_t *>(src) + ((ptrdiff_t)y0 * ss + x0) * 2;
  size_t off_out = stage_in(L, 0, s, ss, rows, cols, 2);
  cudaMemcpyAsync(L.d, L.h, off_out, cudaMemcpyHostToDevice, L.stream);
  const uint8_t *dsrc = L.d + ((size_t)(-y0) * cols + (-x0)) * 2;
  cudaError_t e = launch_block_filter(L.stream, kind, chroma, w, h, bitdepth, filter, dsrc, cols, L.d + off_out, w);
  cudaMemcpyAsync(L.h + off_out, L.d + off_out, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_out, dst, ds, h, w, 2);
}
void xvcb200_filter_h_sample_sample(int chroma, int w, int h, int bd, const int16_t *f, const uint16_t *s, ptrdiff_t ss, uint16_t *d, ptrdiff_t ds) { leaf_filter(0, chroma, w, h, bd, f, s, ss, d, ds); }
void xvcb200_filter_h_sample_short(int chroma, int w, int h, int bd, const int16_t *f, const uint16_t *s, ptrdiff_t ss, int16_t *d, ptrdiff_t ds) { leaf_filter(1, chroma, w, h, bd, f, s, ss, d, ds); }
void xvcb200_filter_v_sample_sample(int chroma, int w, int h, int bd, const int16_t *f, const uint16_t *s, ptrdiff_t ss, uint16_t *d, ptrdiff_t ds) { leaf_filter(2, chroma, w, h, bd, f, s, ss, d, ds); }
void xvcb200_filter_v_sample_short(int chroma, int w, int h, int bd, const int16_t *f, const uint16_t *s, ptrdiff_t ss, int16_t *d, ptrdiff_t ds) { leaf_filter(3, chroma, w, h, bd, f, s, ss, d, ds); }
void xvcb200_filter_v_short_sample(int chroma, int w, int h, int bd, const int16_t *f, const int16_t *s, ptrdiff_t ss, uint16_t *d, ptrdiff_t ds) { leaf_filter(4, chroma, w, h, bd, f, s, ss, d, ds); }
void xvcb200_filter_v_short_short(int chroma, int w, int h, int bd, const int16_t *f, const int16_t *s, ptrdiff_t ss, int16_t *d, ptrdiff_t ds) { leaf_filter(5, chroma, w, h, bd, f, s, ss, d, ds); }

void xvcb200_add_avg(int w, int h, int offset, int shift, int bitdepth, const int16_t *a, intptr_t sa, const int16_t *b,
                     intptr_t sb, uint16_t *dst, intptr_t ds) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  size_t off_b = stage_in(L, 0, a, sa, h, w, 2);
  size_t off_o = stage_in(L, off_b, b, sb, h, w, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_add_avg(L.stream, w, h, offset, shift, bitdepth, (const int16_t *)L.d, w,
                                       (const int16_t *)(L.d + off_b), w, (Sample *)(L.d + off_o), w);
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_o, dst, ds, h, w, 2);
}

void xvcb200_filter_copy_bipred(int w, int h, int16_t offset, int shift, const uint16_t *ref, ptrdiff_t rs, int16_t *pred,
                                ptrdiff_t ps) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  size_t off_o = stage_in(L, 0, ref, rs, h, w, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_copy_bipred(L.stream, w, h, offset, shift, (const Sample *)L.d, w, (int16_t *)(L.d + off_o), w);
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_o, pred, ps, h, w, 2);
}

static void leaf_interp(int chroma, int bipred, int w, int h, int bitdepth, int fx, int fy, const uint16_t *ref,
                        ptrdiff_t rs, void *pred, ptrdiff_t ps) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  const int n = chroma ? 4 : 8, before = n / 2 - 1, after = n / 2;
  const int x0 = fx ? -before : 0, y0 = fy ? -before : 0;
  const int cols = w + (fx ? before + after : 0), rows = h + (fy ? before + after : 0);
  size_t off_o = stage_in(L, 0, ref + (ptrdiff_t)y0 * rs + x0, rs, rows, cols, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  const Sample *dref = (const Sample *)L.d + (size_t)(-y0) * cols + (-x0);
  cudaError_t e = launch_block_interp(L.stream, chroma, bipred, w, h, bitdepth, fx, fy, dref, cols, L.d + off_o, w);
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_o, pred, ps, h, w, 2);
}
void xvcb200_interp_block(int chroma, int w, int h, int bd, int fx, int fy, const uint16_t *ref, ptrdiff_t rs, uint16_t *pred, ptrdiff_t ps) { leaf_interp(chroma, 0, w, h, bd, fx, fy, ref, rs, pred, ps); }
void xvcb200_interp_block_bipred(int chroma, int w, int h, int bd, int fx, int fy, const uint16_t *ref, ptrdiff_t rs, int16_t *pred, ptrdiff_t ps) { leaf_interp(chroma, 1, w, h, bd, fx, fy, ref, rs, pred, ps); }

// table registration: the reference's tables carry one pointer per [size class] / [luma, chroma];
// thin adaptors bind the `chroma` argument
#define XVCB_ADAPT(name, ST, DT)                                                                                         \
  static void name##_luma(int w, int h, int bd, const int16_t *f, const ST *s, ptrdiff_t ss, DT *d, ptrdiff_t ds) { xvcb200_##name(0, w, h, bd, f, s, ss, d, ds); } \
  static void name##_chroma(int w, int h, int bd, const int16_t *f, const ST *s, ptrdiff_t ss, DT *d, ptrdiff_t ds) { xvcb200_##name(1, w, h, bd, f, s, ss, d, ds); }
XVCB_ADAPT(filter_h_sample_sample, uint16_t, uint16_t)
XVCB_ADAPT(filter_h_sample_short, uint16_t, int16_t)
XVCB_ADAPT(filter_v_sample_sample, uint16_t, uint16_t)
XVCB_ADAPT(filter_v_sample_short, uint16_t, int16_t)
XVCB_ADAPT(filter_v_short_sample, int16_t, uint16_t)
XVCB_ADAPT(filter_v_short_short, int16_t, int16_t)
#undef XVCB_ADAPT

void xvcb200_register_inter_prediction(xvcb200_inter_prediction_simd_func *t) {
  for (int i = 0; i < 2; i++) { t->add_avg[i] = &xvcb200_add_avg; t->filter_copy_bipred[i] = &xvcb200_filter_copy_bipred; }
#define XVCB_SET(name) t->name[0] = &name##_luma; t->name[1] = &name##_chroma;
  XVCB_SET(filter_h_sample_sample) XVCB_SET(filter_h_sample_short) XVCB_SET(filter_v_sample_sample)
  XVCB_SET(filter_v_sample_short) XVCB_SET(filter_v_short_sample) XVCB_SET(filter_v_short_short)
#undef XVCB_SET
}

void xvcb200_register_sample_metric(int bitdepth, xvcb200_sample_metric_simd_func *t) {
  (void)bitdepth;
  for (int i = 0; i < 7; i++) {   // index = log2(width); the reference leaves [0] null (sample_metric.cc:785-823)
    t->sad_sample_sample[i] = i ? &xvcb200_sad_sample_sample : nullptr;
    t->sad_short_sample[i] = i ? &xvcb200_sad_short_sample : nullptr;
    t->ssd_sample_sample[i] = i ? &xvcb200_ssd_sample_sample : nullptr;
    t->ssd_short_sample[i] = i ? &xvcb200_ssd_short_sample : nullptr;
    t->ssd_short_short[i] = i ? &xvcb200_ssd_short_short : nullptr;
  }
}

// ---------------------------------------------------------------- (A) transform / quant
static void leaf_transform(int forward, int w, int h, int bitdepth, int tx_hor, int tx_ver, int dst4x4, int dc_only,
                           int skip, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  size_t off_o = stage_in(L, 0, in, is, h, w, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_transform(L.stream, forward, w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, skip,
                                         (const int16_t *)L.d, w, (int16_t *)(L.d + off_o), w);
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_o, out, os, h, w, 2);
}
void xvcb200_fwd_transform(int w, int h, int bd, int tx_hor, int tx_ver, int dst4x4, const int16_t *resi, ptrdiff_t rs, int16_t *coeff, ptrdiff_t cs) { leaf_transform(1, w, h, bd, tx_hor, tx_ver, dst4x4, 0, 0, resi, rs, coeff, cs); }
void xvcb200_fwd_transform_skip(int w, int h, int bd, const int16_t *resi, ptrdiff_t rs, int16_t *coeff, ptrdiff_t cs) { leaf_transform(1, w, h, bd, 0, 0, 0, 0, 1, resi, rs, coeff, cs); }
void xvcb200_inv_transform(int w, int h, int bd, int tx_hor, int tx_ver, int dst4x4, int dc_only, const int16_t *coeff, ptrdiff_t cs, int16_t *resi, ptrdiff_t rs) { leaf_transform(0, w, h, bd, tx_hor, tx_ver, dst4x4, dc_only, 0, coeff, cs, resi, rs); }
void xvcb200_inv_transform_skip(int w, int h, int bd, const int16_t *coeff, ptrdiff_t cs, int16_t *resi, ptrdiff_t rs) { leaf_transform(0, w, h, bd, 0, 0, 0, 0, 1, coeff, cs, resi, rs); }

// ---------------------------------------------------------------- (A) intra prediction, one block
void xvcb200_intra_ref_samples(int w, int h, int bitdepth, int has_above_left, int has_above, int above_right, int has_left,
                               int below_left, const uint16_t *block, ptrdiff_t stride, uint16_t *ref_samples,
                               uint16_t *ref_filtered) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  // the available neighbours, tightly: [0] corner, [1 .. w+h] row above, then the column to the left
  const int n = w + h;
  uint16_t *e = reinterpret_cast<uint16_t *>(L.h);
  memset(e, 0, sizeof(uint16_t) * (2 * n + 1));
  if (has_above_left) e[0] = block[-stride - 1];
  if (has_above) {
    memcpy(e + 1, block - stride, sizeof(uint16_t) * w);
    if (above_right > 0) memcpy(e + 1 + w, block - stride + w, sizeof(uint16_t) * above_right);
  }
  if (has_left)
    for (int y = 0; y < h + (below_left > 0 ? below_left : 0); y++) e[n + 1 + y] = block[y * stride - 1];
  const size_t off_ref = (sizeof(uint16_t) * (2 * n + 1) + 15) & ~(size_t)15, ref_bytes = sizeof(uint16_t) * 2 * XVCB200_INTRA_REF_STRIDE;
  const size_t off_filt = (off_ref + ref_bytes + 15) & ~(size_t)15;
  cudaMemcpyAsync(L.d, L.h, off_ref, cudaMemcpyHostToDevice, L.stream);
  const int nb[5] = {has_above_left, has_above, above_right, has_left, below_left};
  cudaError_t err = launch_intra_ref(L.stream, w, h, bitdepth, nb, reinterpret_cast<const Sample *>(L.d),
                                     reinterpret_cast<Sample *>(L.d + off_ref),
                                     ref_filtered ? reinterpret_cast<Sample *>(L.d + off_filt) : nullptr);
  cudaMemcpyAsync(L.h + off_ref, L.d + off_ref, off_filt + ref_bytes - off_ref, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(err)) return;
  memcpy(ref_samples, L.h + off_ref, ref_bytes);
  if (ref_filtered) memcpy(ref_filtered, L.h + off_filt, ref_bytes);
}

void xvcb200_intra_predict(int mode, int w, int h, int bitdepth, int is_luma, const uint16_t *ref_samples,
                           const uint16_t *ref_filtered, uint16_t *pred, ptrdiff_t stride) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  const size_t ref_bytes = sizeof(uint16_t) * 2 * XVCB200_INTRA_REF_STRIDE;
  const size_t off_filt = (ref_bytes + 15) & ~(size_t)15, off_out = (off_filt + ref_bytes + 15) & ~(size_t)15;
  memcpy(L.h, ref_samples, ref_bytes);
  if (ref_filtered) memcpy(L.h + off_filt, ref_filtered, ref_bytes);
  cudaMemcpyAsync(L.d, L.h, off_out, cudaMemcpyHostToDevice, L.stream);
  cudaError_t err = launch_intra_predict(L.stream, mode, w, h, bitdepth, is_luma, reinterpret_cast<const Sample *>(L.d),
                                         ref_filtered ? reinterpret_cast<const Sample *>(L.d + off_filt) : nullptr,
                                         reinterpret_cast<Sample *>(L.d + off_out), w);
  cudaMemcpyAsync(L.h + off_out, L.d + off_out, sizeof(uint16_t) * (size_t)w * h, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(err)) return;
  stage_out(L, off_out, pred, stride, h, w, 2);
}

// Qp::Qp (quantize.cc:48-92): host-side scalar set-up, no device work
void xvcb200_qp_init(xvcb200_qp *out, int qp, int chroma_format, int bitdepth, double lambda, int table, int off_u, int off_v) {
  static const uint8_t chroma_scale[58] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
                                           20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35,
                                           36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};
  const int offs[3] = {0, off_u, off_v};
  out->lambda_sqrt = std::sqrt(lambda);
  for (int c = 0; c < 3; c++) {
    int raw = qp;
    double weight = 1.0;
    if (c > 0) {
      const int base = qp < 0 ? 0 : (qp > 57 ? 57 : qp);
      int with_off = qp + offs[c];
      with_off = with_off < 0 ? 0 : (with_off > 57 ? 57 : with_off);
      raw = with_off;
      int delta = with_off - base;
      if (chroma_format == 1 && table == 1) { raw = chroma_scale[with_off]; delta = chroma_scale[with_off] - base; }
      weight = std::pow(2.0, -delta / 3.0);
    }
    out->qp_raw[c] = raw;
    const int qbd = raw + 6 * (bitdepth - 8);
    out->qp_bitdepth[c] = qbd < 0 ? 0 : qbd;
    out->distortion_weight[c] = weight;
    out->lambda[c] = c == 0 ? lambda : lambda / weight;
  }
}

int xvcb200_quant_fast(int w, int h, int bitdepth, int qp_bitdepth, int intra_picture, int sign_hiding, int scan_order,
                       const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  Leaf &L = t_leaf;
  if (!L.init()) return -1;
  size_t off_o = stage_in(L, 0, in, is, h, w, 2);
  size_t off_n = (off_o + (size_t)w * h * 2 + 15) & ~(size_t)15;
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_quant(L.stream, w, h, bitdepth, qp_bitdepth, intra_picture, sign_hiding, scan_order,
                                     (const int16_t *)L.d, w, (int16_t *)(L.d + off_o), w, (int *)(L.d + off_n));
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, off_n + 4 - off_o, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return -1;
  stage_out(L, off_o, out, os, h, w, 2);
  int nnz;
  memcpy(&nnz, L.h + off_n, 4);
  return nnz;
}

void xvcb200_dequant(int w, int h, int bitdepth, int qp_bitdepth, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  Leaf &L = t_leaf;
  if (!L.init()) return;
  size_t off_o = stage_in(L, 0, in, is, h, w, 2);
  cudaMemcpyAsync(L.d, L.h, off_o, cudaMemcpyHostToDevice, L.stream);
  cudaError_t e = launch_block_dequant(L.stream, w, h, bitdepth, qp_bitdepth, (const int16_t *)L.d, w, (int16_t *)(L.d + off_o), w);
  cudaMemcpyAsync(L.h + off_o, L.d + off_o, (size_t)w * h * 2, cudaMemcpyDeviceToHost, L.stream);
  if (!L.done(e)) return;
  stage_out(L, off_o, out, os, h, w, 2);
}

// ---------------------------------------------------------------- (B) context
struct CtxExtra {           // host-side state that is not needed by kernels
  // peer arenas opened through CUDA IPC + the DMA streams that push slots into them
  std::vector<uint8_t *> peer_arena;
  std::vector<cudaStream_t> push_stream;
  cudaEvent_t push_ready = nullptr, push_done = nullptr;
  std::vector<cudaEvent_t> slot_pushed;       // per slot: its last push has left (all peers)
  PlaneView *d_luma_views = nullptr;          // luma PlaneView of every slot
  int *d_tu_list = nullptr;                   // TU ids (3*cu + comp) grouped by shape class
  int tu_cap = 0;
  int class_count[7][7];
  int class_offset[7][7];
  xvcb200_me_job *d_jobs = nullptr; xvcb200_me_result *d_me = nullptr; xvcb200_tu_result *d_tu = nullptr;
  int jobs_cap = 0, me_cap = 0, tu_res_cap = 0;
  xvcb200_affine_cu *d_affine = nullptr; int affine_cap = 0;
  xvcb200_lic_cu *d_lic = nullptr; int lic_cap = 0;
  // TZ search job groups (jobs sharing a reference picture and a CTU share one staged window)
  std::vector<xvcb200_cu> h_cus;              // host copy of the CU array (set_cus)
  int *d_job_index = nullptr; int job_index_cap = 0;
  int *d_groups = nullptr; int groups_cap = 0;       // pairs {first, count}
  uint8_t *d_tz_states = nullptr; int tz_states_cap = 0;
  int *d_counter = nullptr;
  int *d_subpel_lists = nullptr; int subpel_lists_cap = 0;   // job lists by block-size class + their counters
  // the picture pipeline's own grouping, built by set_cus: the searched CUs by CTU (jobs are cu * J + column)
  int *d_pipe_index = nullptr;
  int *d_pipe_groups = nullptr;
  int pipe_n_groups = 0;
  // motion decision of the picture pipeline (me_pipe.cu): per-CU state, the bi-prediction passes' jobs / results,
  // the int16 weighted original (luma, laid out like a slot's luma plane)
  uint8_t *d_me_state = nullptr; int me_state_cap = 0;
  xvcb200_me_job *d_bi_jobs = nullptr; int bi_jobs_cap = 0;
  xvcb200_me_result *d_bi_res = nullptr; int bi_res_cap = 0;
  int16_t *d_worig = nullptr;
  std::vector<CUtensorMap> luma_tmaps;         // per slot: [2 * slot + 0 / 1] = narrow / wide box over the padded luma plane (full search)
  xvcb200_tu_mode *d_tu_modes = nullptr; int tu_modes_cap = 0; bool tu_modes_set = false;   // xvcb200_set_tu_modes
  uint8_t *d_part = nullptr, *h_part = nullptr; cudaEvent_t part_ev = nullptr, part_k_ev = nullptr; cudaStream_t part_stream = nullptr; bool part_pending = false;   // xvcb200_decide_partition_begin / _end
  int32_t *d_cu_map2 = nullptr;                // 4x4 CU map of the secondary (chroma) tree (xvcb200_deblock_picture_ext)
  int32_t *d_mvp = nullptr; int mvp_cap = 0;   // xvcb200_set_mv_predictors: [cu][column][2]
  int mvp_cols = 0;                            // 0: none given for the current CU array
  int32_t *h_mvp[2] = {nullptr, nullptr}; size_t h_mvp_cap[2] = {0, 0}; cudaEvent_t mvp_ev[2] = {nullptr, nullptr};   // page-locked staging, alternating
  unsigned mvp_next = 0;
  // Everything set_cus sends ([cus][tu list][CU index by CTU][groups]) is ONE blob,
  // double-buffered: the upload for the next picture goes to the other blob on the upload stream
  // while the kernels of the current picture still use theirs.  d_cus / d_tu_list / d_pipe_* point
  // into the current blob.
  struct CuBlob {
    uint8_t *d = nullptr; size_t cap = 0;       // device
    void *h = nullptr; size_t hcap = 0;         // page-locked host image
    cudaEvent_t up_ev = nullptr;                // upload done (also: host image free again)
    cudaEvent_t free_ev = nullptr;              // compute enqueued while it was current is done
    cudaEvent_t dl_ev = nullptr; bool dl_pending = false;   // xvcb200_get_cus_async reading it
  } blob[2];
  int cur_blob = -1;
  // successive-elimination support: 8-sample segment sums of every slot's luma plane, survivor pool
  std::vector<PlaneView> h_luma_views;
  uint32_t *d_pool = nullptr;
  int pool_cap = 1 << 16, pool_ctas = 0;
  // side streams for independent launches inside one stage (T/Q shape classes)
  static constexpr int kSide = 36;
  cudaStream_t side[kSide] = {nullptr};
  cudaEvent_t side_ev[kSide] = {nullptr};
  cudaEvent_t fork_ev = nullptr;
  int n_side = 0;
  // tight device staging of whole pictures for PCIe transfers: [stream set 0 = context stream, 1 = copy stream][0 up, 1 down]
  uint16_t *d_stage[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  // second stream for transfers that overlap compute (xvcb200_*_async)
  cudaStream_t copy_stream = nullptr;         // device -> host
  cudaStream_t up_stream = nullptr;           // host -> device (own stream: uploads never queue behind downloads)
  cudaEvent_t mark_ev = nullptr;              // "compute enqueued so far", re-recorded per transfer
  std::vector<cudaEvent_t> up_ev; std::vector<char> up_pending;   // per slot; pending = 1 + staging index (tight) or 1 (direct)
  // staging rings of the async transfers: the copy streams only run DMA, the pack / unpack
  // kernels run on the context stream in program order (a small kernel on another stream would
  // wait for the persistent search kernels to leave the SMs)
  static constexpr int kUpRing = 2, kDownRing = 4;
  uint16_t *d_up_ring[kUpRing] = {nullptr}; cudaEvent_t up_ring_ev[kUpRing] = {nullptr};     // event: unpacked, buffer free
  uint16_t *d_down_ring[kDownRing] = {nullptr}; cudaEvent_t down_ring_ev[kDownRing] = {nullptr}; // event: copied out, buffer free
  unsigned up_ring_next = 0, down_ring_next = 0;
  std::vector<char> up_tight;                 // per slot: staging index of the pending upload, -1 = written directly
  std::vector<cudaEvent_t> dl_ev; std::vector<char> dl_pending;   // per slot, last entry = the CU array
  // optional per-stage timing of xvcb200_encode_picture (CUDA events on the context stream)
  // highest ref_idx per list any inter CU of the current array may carry (set_cus; raised by encode_picture's decisions):
  // every (list, ref_idx) up to it must name a valid slot in the calls that dereference reference pictures
  int max_ref_idx[2] = {-1, -1};
  bool profile = false;
  cudaEvent_t ev[9] = {nullptr};
  bool ev_valid = false;
};

}  // extern "C"

struct CtxFull : public xvcb200_ctx { CtxExtra ex; };
static CtxFull *full(xvcb200_ctx *c) { return static_cast<CtxFull *>(c); }
static void join_uploads(CtxFull *c);
static void join_upload_slot(CtxFull *c, int slot);
static void join_downloads(CtxFull *c, int slot);

template <typename T> static bool ensure(xvcb200_ctx *c, T **ptr, int *cap, int n) {
  if (n <= *cap) return true;
  if (*ptr) { cudaStreamSynchronize(c->stream); cudaFree(*ptr); *ptr = nullptr; }
  const int want = n + n / 4 + 64;
  if (!c->check(cudaMalloc(ptr, sizeof(T) * (size_t)want), "cudaMalloc")) { *cap = 0; return false; }
  *cap = want;
  return true;
}

// TMA descriptors of every slot's luma plane (see FsTensorMaps).  cuTensorMapEncodeTiled comes from the
// driver through the runtime's entry-point query, so the library does not link libcuda.
static bool encode_luma_tensor_maps(CtxFull *c) {
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
  if (!c->check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q), "cudaGetDriverEntryPoint") || !fn ||
      q != cudaDriverEntryPointSuccess)
    return c->fail(XVCB200_CUDA_ERROR, "cuTensorMapEncodeTiled is not available from this driver");
  EncodeTiled encode = reinterpret_cast<EncodeTiled>(fn);
  const int pitch = c->geom.pitch[0], rows = c->geom.height[0] + 2 * c->geom.margin_y[0];
  c->ex.luma_tmaps.resize(2 * c->slots.size());
  for (size_t s = 0; s < c->slots.size(); s++)
    for (int k = 0; k < 2; k++) {
      void *base = c->slots[s].base[0] - ((size_t)c->geom.margin_y[0] * pitch + c->geom.margin_x[0]);    // allocation start of the plane
      const cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows};
      const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(Sample)};
      const cuuint32_t box[2] = {(cuuint32_t)(k ? kFsBoxWide : kFsBoxNarrow), (cuuint32_t)kFsBoxRows};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = encode(&c->ex.luma_tmaps[2 * s + k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, base, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return c->fail(XVCB200_CUDA_ERROR, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    }
  return true;
}

extern "C" {

int xvcb200_ctx_create(xvcb200_ctx **out, int device, int width, int height, int bitdepth, int chroma_format,
                       int num_slots) {
  if (!out) return XVCB200_INVALID_ARGUMENT;
  *out = nullptr;
  if (width <= 0 || height <= 0 || width >= 65536 || height >= 65536 || (width & 7) || (height & 7) || bitdepth < 8 ||
      bitdepth > 12 || num_slots < 1 || num_slots > 64)
    return XVCB200_INVALID_ARGUMENT;
  if (chroma_format != 1) return XVCB200_UNSUPPORTED;    // 4:2:0 only on the batched path
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    set_last_error(XVCB200_NO_DEVICE, "no CUDA device: xvc_b200 has no CPU fallback");
    return XVCB200_NO_DEVICE;
  }
  CtxFull *c = new CtxFull();
  memset(c->ex.class_count, 0, sizeof(c->ex.class_count));
  memset(c->ex.class_offset, 0, sizeof(c->ex.class_offset));
  c->device = device; c->width = width; c->height = height; c->bitdepth = bitdepth; c->chroma_format = chroma_format;
  if (!c->check(cudaSetDevice(device), "cudaSetDevice") ||
      !c->check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) {
    int st = c->status; delete c; return st;
  }
  c->own_stream = true;
  size_t plane_bytes[3], total = 0;
  for (int p = 0; p < 3; p++) {
    const int cs = p ? 1 : 0;
    c->geom.width[p] = width >> cs; c->geom.height[p] = height >> cs;
    c->geom.margin_x[p] = 128 >> cs;            // >= 80/40 (yuv_pic.cc:39-41) and keeps x = 0 on a 128-byte boundary
    c->geom.margin_y[p] = 80 >> cs;
    c->geom.pitch[p] = ((c->geom.width[p] + 2 * c->geom.margin_x[p]) + 63) & ~63;
    plane_bytes[p] = ((size_t)c->geom.pitch[p] * (c->geom.height[p] + 2 * c->geom.margin_y[p]) * 2 + 255) & ~(size_t)255;
    total += plane_bytes[p];
  }
  // all slots in ONE allocation, back to back: a run of slots is a contiguous buffer, so a
  // collective (NCCL all-gather of finished reconstructions) can land directly in reference slots
  c->slots.resize(num_slots);
  std::vector<PlaneView> views(num_slots);
  uint8_t *arena = nullptr;
  // (+ a tail of one 32-bit arrival tag per slot, inside the allocation a peer maps: xvcb200_push_slot_tagged)
  const size_t tag_bytes = (sizeof(uint32_t) * (size_t)num_slots + 255) & ~(size_t)255;
  if (!c->check(cudaMalloc(&arena, total * num_slots + tag_bytes), "cudaMalloc(slots)")) { int st = c->status; delete c; return st; }
  cudaMemsetAsync(arena, 0, total * num_slots + tag_bytes, c->stream);
  c->slot_stride = total;
  for (int s = 0; s < num_slots; s++) {
    DevPicture &d = c->slots[s];
    d.alloc = arena + (size_t)s * total;
    d.bytes = total;
    size_t off = 0;
    for (int p = 0; p < 3; p++) {
      d.base[p] = reinterpret_cast<Sample *>(d.alloc + off) + (size_t)c->geom.margin_y[p] * c->geom.pitch[p] + c->geom.margin_x[p];
      off += plane_bytes[p];
    }
    views[s] = c->plane(s, 0);
  }
  if (cudaEventCreateWithFlags(&c->ex.fork_ev, cudaEventDisableTiming) == cudaSuccess) {
    for (int i = 0; i < CtxExtra::kSide; i++) {
      if (cudaStreamCreateWithFlags(&c->ex.side[i], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&c->ex.side_ev[i], cudaEventDisableTiming) != cudaSuccess) break;
      c->ex.n_side = i + 1;
    }
  }
  c->ex.h_luma_views = views;
  if (!encode_luma_tensor_maps(c)) { int st = c->status; xvcb200_ctx_destroy(c); return st; }
  c->map_w = width >> 2; c->map_h = height >> 2;
  const size_t cells = (size_t)c->map_w * c->map_h;
  if (!c->check(cudaMalloc(&c->ex.d_luma_views, sizeof(PlaneView) * num_slots), "cudaMalloc(views)") ||
      !c->check(cudaMemcpyAsync(c->ex.d_luma_views, views.data(), sizeof(PlaneView) * num_slots, cudaMemcpyHostToDevice, c->stream), "memcpy(views)") ||
      !c->check(cudaMalloc(&c->d_cu_map, sizeof(int32_t) * cells), "cudaMalloc(map)") ||
      !c->check(cudaMalloc(&c->d_edge_bs[0], cells), "cudaMalloc(bs)") ||
      !c->check(cudaMalloc(&c->d_edge_bs[1], cells), "cudaMalloc(bs)") ||
      !c->check(cudaStreamSynchronize(c->stream), "sync")) {
    int st = c->status; xvcb200_ctx_destroy(c); return st;
  }
  *out = c;
  return XVCB200_OK;
}

void xvcb200_ctx_destroy(xvcb200_ctx *ctx) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return;
  CtxFull *c = full(ctx);
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (!c->slots.empty() && c->slots[0].alloc) cudaFree(c->slots[0].alloc);   // one arena for all slots
  cudaFree(c->d_cu_map); cudaFree(c->d_edge_bs[0]); cudaFree(c->d_edge_bs[1]);
  cudaFree(c->d_scratch); cudaFree(c->d_scratch2);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto &b : c->ex.blob) {
    cudaFree(b.d);
    if (b.h) cudaFreeHost(b.h);
    if (b.up_ev) cudaEventDestroy(b.up_ev);
    if (b.free_ev) cudaEventDestroy(b.free_ev);
    if (b.dl_ev) cudaEventDestroy(b.dl_ev);
  }
  for (size_t p = 0; p < c->ex.peer_arena.size(); p++) {
    cudaStreamSynchronize(c->ex.push_stream[p]);
    cudaStreamDestroy(c->ex.push_stream[p]);
    cudaIpcCloseMemHandle(c->ex.peer_arena[p]);
  }
  if (c->ex.push_ready) cudaEventDestroy(c->ex.push_ready);
  if (c->ex.push_done) cudaEventDestroy(c->ex.push_done);
  for (cudaEvent_t e : c->ex.slot_pushed) if (e) cudaEventDestroy(e);
  if (c->ex.copy_stream) { cudaStreamSynchronize(c->ex.copy_stream); cudaStreamDestroy(c->ex.copy_stream); }
  for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) cudaFree(c->ex.d_stage[a][b]);
  if (c->ex.up_stream) { cudaStreamSynchronize(c->ex.up_stream); cudaStreamDestroy(c->ex.up_stream); }
  if (c->ex.mark_ev) cudaEventDestroy(c->ex.mark_ev);
  for (cudaEvent_t e : c->ex.up_ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < CtxExtra::kUpRing; i++) { cudaFree(c->ex.d_up_ring[i]); if (c->ex.up_ring_ev[i]) cudaEventDestroy(c->ex.up_ring_ev[i]); }
  for (int i = 0; i < CtxExtra::kDownRing; i++) { cudaFree(c->ex.d_down_ring[i]); if (c->ex.down_ring_ev[i]) cudaEventDestroy(c->ex.down_ring_ev[i]); }
  for (cudaEvent_t e : c->ex.dl_ev) if (e) cudaEventDestroy(e);
  cudaFree(c->ex.d_subpel_lists); cudaFree(c->ex.d_pool);
  cudaFree(c->ex.d_job_index); cudaFree(c->ex.d_groups); cudaFree(c->ex.d_tz_states); cudaFree(c->ex.d_counter);
  for (int b = 0; b < 2; b++) { if (c->ex.h_mvp[b]) cudaFreeHost(c->ex.h_mvp[b]); if (c->ex.mvp_ev[b]) cudaEventDestroy(c->ex.mvp_ev[b]); }
  cudaFree(c->ex.d_part); if (c->ex.h_part) cudaFreeHost(c->ex.h_part); if (c->ex.part_ev) cudaEventDestroy(c->ex.part_ev); if (c->ex.part_k_ev) cudaEventDestroy(c->ex.part_k_ev);
  if (c->ex.part_stream) { cudaStreamSynchronize(c->ex.part_stream); cudaStreamDestroy(c->ex.part_stream); }
  cudaFree(c->ex.d_cu_map2); cudaFree(c->ex.d_tu_modes); cudaFree(c->ex.d_mvp); cudaFree(c->ex.d_me_state); cudaFree(c->ex.d_bi_jobs); cudaFree(c->ex.d_bi_res); cudaFree(c->ex.d_worig);
  cudaFree(c->ex.d_luma_views); cudaFree(c->ex.d_jobs); cudaFree(c->ex.d_me); cudaFree(c->ex.d_tu); cudaFree(c->ex.d_affine); cudaFree(c->ex.d_lic);
  for (auto &e : c->ex.ev) if (e) cudaEventDestroy(e);
  for (int i = 0; i < c->ex.n_side; i++) { cudaStreamDestroy(c->ex.side[i]); cudaEventDestroy(c->ex.side_ev[i]); }
  if (c->ex.fork_ev) cudaEventDestroy(c->ex.fork_ev);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int xvcb200_ctx_set_stream(xvcb200_ctx *c, void *cuda_stream) {
  xvcb::DevGuard dev_guard(c);
  if (!c) return XVCB200_INVALID_ARGUMENT;
  cudaStreamSynchronize(c->stream);
  if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
  if (cuda_stream == nullptr) {
    if (!c->check(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate")) return c->status;
    c->own_stream = true;
  } else {
    c->stream = static_cast<cudaStream_t>(cuda_stream);
  }
  return XVCB200_OK;
}

int xvcb200_sync(xvcb200_ctx *c) {
  xvcb::DevGuard dev_guard(c);
  if (!c) return XVCB200_INVALID_ARGUMENT;
  c->check(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
  return c->status;
}
const char *xvcb200_ctx_error_string(xvcb200_ctx *c) { return c ? c->error.c_str() : "null context"; }

int xvcb200_get_geometry(xvcb200_ctx *c, xvcb200_plane_geom *g) {
  xvcb::DevGuard dev_guard(c);
  if (!c || !g) return XVCB200_INVALID_ARGUMENT;
  *g = c->geom;
  return XVCB200_OK;
}
int xvcb200_slot_ptr(xvcb200_ctx *c, int slot, int comp, void **p) {
  xvcb::DevGuard dev_guard(c);
  if (!c || !p || slot < 0 || slot >= (int)c->slots.size() || comp < 0 || comp > 2) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);          // raw access: a deferred upload is unpacked (on the context stream) first
  *p = c->slots[slot].base[comp];
  return XVCB200_OK;
}
void *xvcb200_stream(xvcb200_ctx *c) { return c ? c->stream : nullptr; }
// whole allocation of a slot (three padded planes); consecutive slots are contiguous
int xvcb200_slot_region(xvcb200_ctx *c, int slot, void **base, uint64_t *bytes) {
  xvcb::DevGuard dev_guard(c);
  if (!c || !base || !bytes || slot < 0 || slot >= (int)c->slots.size()) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);          // raw access: a deferred upload is unpacked (on the context stream) first
  *base = c->slots[slot].alloc;
  *bytes = c->slot_stride;
  return XVCB200_OK;
}

// handle = [cudaIpcMemHandle_t of the slot arena (64 bytes)][slot stride in bytes (u64)][number of slots (u32)][magic (u32)]:
// the layout travels with the handle, so a peer whose arena is laid out differently is refused
// instead of being written out of bounds by xvcb200_push_slot.
static const uint32_t kIpcMagic = 0x58564342u;   // "XVCB"
int xvcb200_ipc_export(xvcb200_ctx *ctx, void *handle) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !handle || ctx->slots.empty()) return XVCB200_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64 && XVCB200_IPC_HANDLE_BYTES == 80, "handle size");
  cudaIpcMemHandle_t h;
  if (!ctx->check(cudaIpcGetMemHandle(&h, ctx->slots[0].alloc), "cudaIpcGetMemHandle")) return ctx->status;
  uint8_t *out = static_cast<uint8_t *>(handle);
  const uint64_t stride = ctx->slot_stride;
  const uint32_t nslots = (uint32_t)ctx->slots.size();
  memcpy(out, &h, sizeof(h));
  memcpy(out + 64, &stride, 8);
  memcpy(out + 72, &nslots, 4);
  memcpy(out + 76, &kIpcMagic, 4);
  return XVCB200_OK;
}
int xvcb200_ipc_open_peer(xvcb200_ctx *ctx, const void *handle, int *peer_index) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !handle) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  {
    const uint8_t *in = static_cast<const uint8_t *>(handle);
    uint64_t stride; uint32_t nslots, magic;
    memcpy(&stride, in + 64, 8); memcpy(&nslots, in + 72, 4); memcpy(&magic, in + 76, 4);
    if (magic != kIpcMagic || stride != (uint64_t)c->slot_stride || nslots != (uint32_t)c->slots.size()) {
      set_last_error(XVCB200_INVALID_ARGUMENT, "xvcb200_ipc_open_peer: the peer's slot arena has another layout (slot stride / slot count)");
      return XVCB200_INVALID_ARGUMENT;
    }
  }
  void *base = nullptr;
  cudaError_t oe = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
  if (oe != cudaSuccess) {      // not sticky: the caller may fall back to another exchange (NCCL)
    cudaGetLastError();
    set_last_error(XVCB200_CUDA_ERROR, (std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(oe)).c_str());
    return XVCB200_CUDA_ERROR;
  }
  cudaStream_t st = nullptr;
  if (!c->check(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate")) return c->status;
  if (!c->ex.push_ready && (!c->check(cudaEventCreateWithFlags(&c->ex.push_ready, cudaEventDisableTiming), "cudaEventCreate") ||
                            !c->check(cudaEventCreateWithFlags(&c->ex.push_done, cudaEventDisableTiming), "cudaEventCreate")))
    return c->status;
  c->ex.peer_arena.push_back(static_cast<uint8_t *>(base));
  c->ex.push_stream.push_back(st);
  if (peer_index) *peer_index = (int)c->ex.peer_arena.size() - 1;
  return XVCB200_OK;
}
// Driver entry points of the arrival tags (stream memory operations), through the runtime's query like
// cuTensorMapEncodeTiled: the library does not link libcuda.
typedef CUresult (*StreamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*MemsetD32Async)(CUdeviceptr, unsigned int, size_t, CUstream);
static bool tag_entry_points(CtxFull *c, StreamWaitValue32 *wait, MemsetD32Async *set) {
  static void *fn_wait = nullptr, *fn_set = nullptr;
  if (!fn_wait || !fn_set) {
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (!c->check(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn_wait, cudaEnableDefault, &q), "cudaGetDriverEntryPoint") || !fn_wait ||
        q != cudaDriverEntryPointSuccess)
      return c->fail(XVCB200_CUDA_ERROR, "cuStreamWaitValue32 is not available from this driver");
    if (!c->check(cudaGetDriverEntryPoint("cuMemsetD32Async", &fn_set, cudaEnableDefault, &q), "cudaGetDriverEntryPoint") || !fn_set ||
        q != cudaDriverEntryPointSuccess)
      return c->fail(XVCB200_CUDA_ERROR, "cuMemsetD32Async is not available from this driver");
  }
  *wait = reinterpret_cast<StreamWaitValue32>(fn_wait);
  *set = reinterpret_cast<MemsetD32Async>(fn_set);
  return true;
}
static size_t tag_offset(const xvcb200_ctx *c, int slot) { return c->slot_stride * c->slots.size() + sizeof(uint32_t) * (size_t)slot; }

// The slot into the same slot of every opened peer, one DMA stream per peer (the copies run side by side), ordered after
// the work enqueued on the context stream so far; with `tagged` the arrival tag follows the slot on the same stream.
static int push_slot_impl(xvcb200_ctx *ctx, int slot, bool tagged, uint32_t tag) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  StreamWaitValue32 wait = nullptr; MemsetD32Async set = nullptr;
  if (tagged && !tag_entry_points(c, &wait, &set)) return c->status;
  if (c->ex.peer_arena.empty()) return XVCB200_OK;
  join_upload_slot(c, slot);
  const size_t off = (size_t)slot * c->slot_stride;
  c->check(cudaEventRecord(c->ex.push_ready, c->stream), "cudaEventRecord");
  for (size_t p = 0; p < c->ex.peer_arena.size(); p++) {
    cudaStream_t st = c->ex.push_stream[p];
    c->check(cudaStreamWaitEvent(st, c->ex.push_ready, 0), "cudaStreamWaitEvent");
    c->check(cudaMemcpyAsync(c->ex.peer_arena[p] + off, c->slots[0].alloc + off, c->slot_stride, cudaMemcpyDeviceToDevice, st), "push slot");
    if (tagged && set(reinterpret_cast<CUdeviceptr>(c->ex.peer_arena[p] + tag_offset(c, slot)), tag, 1, st) != CUDA_SUCCESS)
      return c->fail(XVCB200_CUDA_ERROR, "cuMemsetD32Async (arrival tag) failed");
    if (p > 0) {      // funnel completion into the first push stream
      c->check(cudaEventRecord(c->ex.push_done, st), "cudaEventRecord");
      c->check(cudaStreamWaitEvent(c->ex.push_stream[0], c->ex.push_done, 0), "cudaStreamWaitEvent");
    }
  }
  c->ex.slot_pushed.resize(c->slots.size(), nullptr);
  cudaEvent_t &ev = c->ex.slot_pushed[slot];
  if (!ev && !c->check(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate")) return c->status;
  c->check(cudaEventRecord(ev, c->ex.push_stream[0]), "cudaEventRecord");
  return c->status;
}
int xvcb200_push_slot(xvcb200_ctx *ctx, int slot) { return push_slot_impl(ctx, slot, false, 0); }
int xvcb200_push_slot_tagged(xvcb200_ctx *ctx, int slot, uint32_t tag) { return push_slot_impl(ctx, slot, true, tag); }
int xvcb200_wait_slot_tag(xvcb200_ctx *ctx, int slot, uint32_t tag) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  StreamWaitValue32 wait = nullptr; MemsetD32Async unused = nullptr;
  if (!tag_entry_points(c, &wait, &unused)) return c->status;
  // (tag - value) as a signed difference >= 0: CU_STREAM_WAIT_VALUE_GEQ compares cyclically, tags may wrap
  if (wait(c->stream, reinterpret_cast<CUdeviceptr>(c->slots[0].alloc + tag_offset(c, slot)), tag, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return c->fail(XVCB200_CUDA_ERROR, "cuStreamWaitValue32 (arrival tag) failed");
  return c->status;
}
int xvcb200_wait_pushes(xvcb200_ctx *ctx, int slot) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || slot >= (int)ctx->slots.size()) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int s = 0; s < (int)c->ex.slot_pushed.size(); s++)
    if ((slot < 0 || s == slot) && c->ex.slot_pushed[s])
      c->check(cudaStreamWaitEvent(c->stream, c->ex.slot_pushed[s], 0), "cudaStreamWaitEvent");
  return c->status;
}

static bool slot_ok(xvcb200_ctx *c, int slot) { return c && slot >= 0 && slot < (int)c->slots.size(); }

// One picture between host planes and a slot, enqueued on `st`.  Tight host planes (stride ==
// width) travel as one contiguous copy per plane through a tight device staging buffer plus a
// pack / unpack kernel; a row-by-row 2-D copy of a padded plane runs at a fraction of the PCIe
// rate.  Strided host planes take the 2-D copy.  set: 0 = context stream, 1 = copy stream.
static int transfer_picture(xvcb200_ctx *ctx, int slot, void *const planes[3], const ptrdiff_t strides[3], bool to_device,
                            cudaStream_t st, int set) {
  xvcb::DevGuard dev_guard(ctx);
  CtxFull *c = full(ctx);
  bool tight = (c->geom.width[0] & 7) == 0;
  size_t samples = 0, off[3];
  for (int p = 0; p < 3; p++) {
    tight = tight && strides[p] == c->geom.width[p];
    off[p] = samples;
    samples += (size_t)c->geom.width[p] * c->geom.height[p];
  }
  if (tight) {
    uint16_t *&stage = c->ex.d_stage[set][to_device ? 0 : 1];
    if (!stage && !c->check(cudaMalloc(&stage, samples * 2), "cudaMalloc(staging)")) return c->status;
    if (to_device) {
      for (int p = 0; p < 3; p++)
        if (!c->check(cudaMemcpyAsync(stage + off[p], planes[p], (size_t)c->geom.width[p] * c->geom.height[p] * 2,
                                      cudaMemcpyHostToDevice, st), "upload_picture"))
          return c->status;
      c->check(launch_plane_pack(st, pic3(c, slot), stage, 0), "plane_unpack");
    } else {
      if (!c->check(launch_plane_pack(st, pic3(c, slot), stage, 1), "plane_pack")) return c->status;
      for (int p = 0; p < 3; p++)
        if (!c->check(cudaMemcpyAsync(planes[p], stage + off[p], (size_t)c->geom.width[p] * c->geom.height[p] * 2,
                                      cudaMemcpyDeviceToHost, st), "download_picture"))
          return c->status;
    }
    return c->status;
  }
  for (int p = 0; p < 3; p++) {
    const cudaError_t e = to_device
        ? cudaMemcpy2DAsync(c->slots[slot].base[p], (size_t)c->geom.pitch[p] * 2, planes[p], (size_t)strides[p] * 2,
                            (size_t)c->geom.width[p] * 2, c->geom.height[p], cudaMemcpyHostToDevice, st)
        : cudaMemcpy2DAsync(planes[p], (size_t)strides[p] * 2, c->slots[slot].base[p], (size_t)c->geom.pitch[p] * 2,
                            (size_t)c->geom.width[p] * 2, c->geom.height[p], cudaMemcpyDeviceToHost, st);
    if (!c->check(e, to_device ? "upload_picture" : "download_picture")) return c->status;
  }
  return c->status;
}

int xvcb200_upload_picture(xvcb200_ctx *c, int slot, const uint16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, slot) || !planes || !strides) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);          // an async upload nobody consumed must not land on top of this one
  join_downloads(full(c), slot);
  // pageable host memory is consumed before cudaMemcpyAsync returns; page-locked memory must stay valid until the stream reaches the copy
  return transfer_picture(c, slot, reinterpret_cast<void *const *>(const_cast<uint16_t *const *>(planes)), strides, true, c->stream, 0);
}
static int download_planes(xvcb200_ctx *c, int slot, void *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, slot) || !planes || !strides) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);          // a deferred (tight) async upload is unpacked into the slot first
  const int st = transfer_picture(c, slot, planes, strides, false, c->stream, 0);
  if (st != XVCB200_OK) return st;
  return xvcb200_sync(c);
}
int xvcb200_download_picture(xvcb200_ctx *c, int slot, uint16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(c);
  return download_planes(c, slot, reinterpret_cast<void *const *>(planes), strides);
}
int xvcb200_download_coeff(xvcb200_ctx *c, int slot, int16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(c);
  return download_planes(c, slot, reinterpret_cast<void *const *>(planes), strides);
}
// full padded allocation of one plane (tests of PadBorder): rows x cols = (h+2*pad) x (w+2*pad), pad = 80 / 40
int xvcb200_download_padded(xvcb200_ctx *c, int slot, int comp, uint16_t *dst) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, slot) || comp < 0 || comp > 2 || !dst) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);
  const int pad = 80 >> (comp ? 1 : 0), w = c->geom.width[comp] + 2 * pad, h = c->geom.height[comp] + 2 * pad;
  const Sample *src = c->slots[slot].base[comp] - (ptrdiff_t)pad * c->geom.pitch[comp] - pad;
  if (!c->check(cudaMemcpy2DAsync(dst, (size_t)w * 2, src, (size_t)c->geom.pitch[comp] * 2, (size_t)w * 2, h,
                                  cudaMemcpyDeviceToHost, c->stream), "download_padded"))
    return c->status;
  return xvcb200_sync(c);
}

}  // extern "C"

// ---- transfers on the copy stream.  Ordering is enforced with events, never by the host:
//  * an async upload starts after all compute enqueued BEFORE the call (nobody still reads the
//    slot) and compute enqueued after it waits for it (join_uploads);
//  * an async download starts after all compute enqueued before the call; compute that later
//    WRITES the same slot waits for it (join_downloads).
static bool copy_setup(CtxFull *c) {
  if (c->ex.copy_stream) return true;
  if (!c->check(cudaStreamCreateWithFlags(&c->ex.copy_stream, cudaStreamNonBlocking), "cudaStreamCreate(copy)") ||
      !c->check(cudaStreamCreateWithFlags(&c->ex.up_stream, cudaStreamNonBlocking), "cudaStreamCreate(upload)") ||
      !c->check(cudaEventCreateWithFlags(&c->ex.mark_ev, cudaEventDisableTiming), "cudaEventCreate"))
    return false;
  c->ex.dl_ev.assign(c->slots.size() + 1, nullptr);
  c->ex.dl_pending.assign(c->slots.size() + 1, 0);
  c->ex.up_ev.assign(c->slots.size(), nullptr);
  c->ex.up_pending.assign(c->slots.size(), 0);
  c->ex.up_tight.assign(c->slots.size(), -1);
  size_t samples = 0;
  for (int p = 0; p < 3; p++) samples += (size_t)c->geom.width[p] * c->geom.height[p];
  for (int i = 0; i < CtxExtra::kUpRing; i++)
    if (!c->check(cudaMalloc(&c->ex.d_up_ring[i], samples * 2), "cudaMalloc(staging)") ||
        !c->check(cudaEventCreateWithFlags(&c->ex.up_ring_ev[i], cudaEventDisableTiming), "cudaEventCreate"))
      return false;
  for (int i = 0; i < CtxExtra::kDownRing; i++)
    if (!c->check(cudaMalloc(&c->ex.d_down_ring[i], samples * 2), "cudaMalloc(staging)") ||
        !c->check(cudaEventCreateWithFlags(&c->ex.down_ring_ev[i], cudaEventDisableTiming), "cudaEventCreate"))
      return false;
  for (auto &e : c->ex.dl_ev)
    if (!c->check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return false;
  for (auto &e : c->ex.up_ev)
    if (!c->check(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate")) return false;
  return true;
}
static bool copy_after_compute(CtxFull *c, cudaStream_t st) {
  return c->check(cudaEventRecord(c->ex.mark_ev, c->stream), "cudaEventRecord") &&
         c->check(cudaStreamWaitEvent(st, c->ex.mark_ev, 0), "cudaStreamWaitEvent");
}
// Compute waits only for pending uploads of the slots it reads (join_upload_slot): an upload
// made ahead for the next picture must not hold back the kernels of the current one.
static void join_upload_slot(CtxFull *c, int slot) {
  if (c->ex.up_pending.empty() || slot < 0 || slot >= (int)c->ex.up_pending.size() || !c->ex.up_pending[slot]) return;
  c->check(cudaStreamWaitEvent(c->stream, c->ex.up_ev[slot], 0), "cudaStreamWaitEvent");
  c->ex.up_pending[slot] = 0;
  const int k = c->ex.up_tight[slot];
  if (k >= 0) {                 // the picture sits in a tight staging buffer: unpack it into the slot, here, in program order
    if (slot < (int)c->ex.dl_pending.size() && c->ex.dl_pending[slot]) {     // a strided async download still reads the slot
      c->check(cudaStreamWaitEvent(c->stream, c->ex.dl_ev[slot], 0), "cudaStreamWaitEvent");
      c->ex.dl_pending[slot] = 0;
    }
    c->check(launch_plane_pack(c->stream, pic3(c, slot), c->ex.d_up_ring[k], 0), "plane_unpack");
    c->check(cudaEventRecord(c->ex.up_ring_ev[k], c->stream), "cudaEventRecord");
    c->ex.up_tight[slot] = -1;
  }
}
static void join_uploads(CtxFull *c) {          // entry points without an explicit read set: all pending uploads
  for (int s = 0; s < (int)c->ex.up_pending.size(); s++) join_upload_slot(c, s);
}
static void join_downloads(CtxFull *c, int slot) {     // slot == -1: the CU array
  if (slot < 0) {
    if (c->ex.cur_blob >= 0 && c->ex.blob[c->ex.cur_blob].dl_pending) {
      c->check(cudaStreamWaitEvent(c->stream, c->ex.blob[c->ex.cur_blob].dl_ev, 0), "cudaStreamWaitEvent");
      c->ex.blob[c->ex.cur_blob].dl_pending = false;
    }
    return;
  }
  if (c->ex.dl_pending.empty()) return;
  const size_t i = (size_t)slot;
  if (!c->ex.dl_pending[i]) return;
  c->check(cudaStreamWaitEvent(c->stream, c->ex.dl_ev[i], 0), "cudaStreamWaitEvent");
  c->ex.dl_pending[i] = 0;
}
static bool host_planes_tight(const CtxFull *c, const ptrdiff_t strides[3]) {
  bool tight = (c->geom.width[0] & 7) == 0;
  for (int p = 0; p < 3; p++) tight = tight && strides[p] == c->geom.width[p];
  return tight;
}
static int download_planes_async(CtxFull *c, int slot, void *const planes[3], const ptrdiff_t strides[3]) {
  if (!slot_ok(c, slot) || !planes || !strides) return XVCB200_INVALID_ARGUMENT;
  if (!copy_setup(c)) return c->status;
  join_upload_slot(c, slot);
  if (host_planes_tight(c, strides)) {
    // pack on the context stream (program order: the slot may be rewritten right after), DMA on the copy stream
    const int k = (int)(c->ex.down_ring_next++ % CtxExtra::kDownRing);
    c->check(cudaStreamWaitEvent(c->stream, c->ex.down_ring_ev[k], 0), "cudaStreamWaitEvent");
    if (!c->check(launch_plane_pack(c->stream, pic3(c, slot), c->ex.d_down_ring[k], 1), "plane_pack") ||
        !copy_after_compute(c, c->ex.copy_stream))
      return c->status;
    size_t off = 0;
    for (int p = 0; p < 3; p++) {
      const size_t cnt = (size_t)c->geom.width[p] * c->geom.height[p];
      if (!c->check(cudaMemcpyAsync(planes[p], c->ex.d_down_ring[k] + off, cnt * 2, cudaMemcpyDeviceToHost, c->ex.copy_stream),
                    "download_async"))
        return c->status;
      off += cnt;
    }
    c->check(cudaEventRecord(c->ex.down_ring_ev[k], c->ex.copy_stream), "cudaEventRecord");
    c->check(cudaEventRecord(c->ex.dl_ev[slot], c->ex.copy_stream), "cudaEventRecord");   // for xvcb200_wait_download only
    return c->status;
  }
  if (!copy_after_compute(c, c->ex.copy_stream)) return c->status;
  if (transfer_picture(c, slot, planes, strides, false, c->ex.copy_stream, 1) != XVCB200_OK) return c->status;
  c->check(cudaEventRecord(c->ex.dl_ev[slot], c->ex.copy_stream), "cudaEventRecord");
  c->ex.dl_pending[slot] = 1;
  return c->status;
}

extern "C" {

int xvcb200_upload_picture_async(xvcb200_ctx *ctx, int slot, const uint16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, slot) || !planes || !strides) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (!copy_setup(c)) return c->status;
  if (host_planes_tight(c, strides)) {
    // DMA into a tight staging buffer now; the unpack kernel runs on the context stream when the
    // first call that reads the slot is enqueued (join_upload_slot), so the slot itself is only
    // ever written in program order
    if (c->ex.up_pending[slot]) join_upload_slot(c, slot);      // an earlier upload of this slot nobody consumed
    const int k = (int)(c->ex.up_ring_next++ % CtxExtra::kUpRing);
    for (size_t s = 0; s < c->ex.up_tight.size(); s++)          // a pending upload still owns this staging buffer
      if (c->ex.up_pending[s] && c->ex.up_tight[s] == k) join_upload_slot(c, (int)s);
    c->check(cudaStreamWaitEvent(c->ex.up_stream, c->ex.up_ring_ev[k], 0), "cudaStreamWaitEvent");
    size_t off = 0;
    for (int p = 0; p < 3; p++) {
      const size_t cnt = (size_t)c->geom.width[p] * c->geom.height[p];
      if (!c->check(cudaMemcpyAsync(c->ex.d_up_ring[k] + off, planes[p], cnt * 2, cudaMemcpyHostToDevice, c->ex.up_stream),
                    "upload_picture_async"))
        return c->status;
      off += cnt;
    }
    c->check(cudaEventRecord(c->ex.up_ev[slot], c->ex.up_stream), "cudaEventRecord");
    c->ex.up_pending[slot] = 1;
    c->ex.up_tight[slot] = (char)k;
    return c->status;
  }
  if (!copy_after_compute(c, c->ex.up_stream)) return c->status;
  // a download of the same slot still in flight reads what this upload overwrites
  if (c->ex.dl_pending[slot]) c->check(cudaStreamWaitEvent(c->ex.up_stream, c->ex.dl_ev[slot], 0), "cudaStreamWaitEvent");
  if (transfer_picture(c, slot, reinterpret_cast<void *const *>(const_cast<uint16_t *const *>(planes)), strides, true,
                       c->ex.up_stream, 1) != XVCB200_OK)
    return c->status;
  c->check(cudaEventRecord(c->ex.up_ev[slot], c->ex.up_stream), "cudaEventRecord");
  c->ex.up_pending[slot] = 1;
  return c->status;
}
int xvcb200_download_picture_async(xvcb200_ctx *ctx, int slot, uint16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return XVCB200_INVALID_ARGUMENT;
  return download_planes_async(full(ctx), slot, reinterpret_cast<void *const *>(planes), strides);
}
int xvcb200_download_coeff_async(xvcb200_ctx *ctx, int slot, int16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return XVCB200_INVALID_ARGUMENT;
  return download_planes_async(full(ctx), slot, reinterpret_cast<void *const *>(planes), strides);
}
int xvcb200_get_cus_async(xvcb200_ctx *ctx, xvcb200_cu *cus, int n) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !cus || n < 0 || n > ctx->n_cus) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (!copy_setup(c) || !copy_after_compute(c, c->ex.copy_stream)) return c->status;
  if (!c->check(cudaMemcpyAsync(cus, c->d_cus, sizeof(xvcb200_cu) * (size_t)n, cudaMemcpyDeviceToHost, c->ex.copy_stream), "get_cus_async"))
    return c->status;
  if (c->ex.cur_blob >= 0) {
    c->check(cudaEventRecord(c->ex.blob[c->ex.cur_blob].dl_ev, c->ex.copy_stream), "cudaEventRecord");
    c->ex.blob[c->ex.cur_blob].dl_pending = true;
  }
  return c->status;
}
// host waits for the last async download of one slot (slot < 0: the CU array)
int xvcb200_wait_download(xvcb200_ctx *ctx, int slot) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || slot >= (int)ctx->slots.size()) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (slot < 0) {
    for (auto &b : c->ex.blob)
      if (b.dl_ev) c->check(cudaEventSynchronize(b.dl_ev), "cudaEventSynchronize");
    return c->status;
  }
  if (c->ex.dl_ev.empty()) return c->status;
  c->check(cudaEventSynchronize(c->ex.dl_ev[(size_t)slot]), "cudaEventSynchronize");
  return c->status;
}
int xvcb200_sync_copies(xvcb200_ctx *ctx) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (c->ex.copy_stream) c->check(cudaStreamSynchronize(c->ex.copy_stream), "cudaStreamSynchronize(copy)");
  if (c->ex.up_stream) c->check(cudaStreamSynchronize(c->ex.up_stream), "cudaStreamSynchronize(upload)");
  return c->status;
}

int xvcb200_pad_border(xvcb200_ctx *c, int slot) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, slot)) return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), slot);
  join_downloads(full(c), slot);
  const int pad[3] = {80, 40, 40};
  c->check(launch_pad_border(c->stream, pic3(c, slot), pad), "pad_border");
  return c->status;
}

// CU array -> device, plus the shape-class lists of the transform units (host bucketing:
// shapes are fixed for the picture, only mv / flags change on the device afterwards)
int xvcb200_set_cus(xvcb200_ctx *ctx, const xvcb200_cu *cus, int n) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || (!cus && n > 0) || n < 0) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++) {
    const xvcb200_cu &u = cus[i];
    const bool pow2 = (u.w & (u.w - 1)) == 0 && (u.h & (u.h - 1)) == 0;
    if (!pow2 || u.w < 4 || u.h < 4 || u.w > 64 || u.h > 64 || u.x < 0 || u.y < 0 || u.x + u.w > c->width ||
        u.y + u.h > c->height || (u.x & 3) || (u.y & 3) || u.ref_idx[0] < -1 || u.ref_idx[0] > 4 || u.ref_idx[1] < -1 ||
        u.ref_idx[1] > 4)
      return XVCB200_INVALID_ARGUMENT;
  }
  c->ex.max_ref_idx[0] = c->ex.max_ref_idx[1] = -1;
  for (int i = 0; i < n; i++)
    for (int l = 0; l < 2; l++)
      if (!(cus[i].flags & XVCB200_CU_INTRA) && cus[i].ref_idx[l] > c->ex.max_ref_idx[l]) c->ex.max_ref_idx[l] = cus[i].ref_idx[l];
  c->n_cus = n;
  c->ex.mvp_cols = 0;                       // predictors and transform modes belong to a CU array
  c->ex.tu_modes_set = false;
  if (n == 0) { c->ex.h_cus.clear(); return XVCB200_OK; }
  if (!copy_setup(c)) return c->status;
  // CU groups by CTU (counting sort: coding order inside a CTU is kept).  Only CUs that take part in the
  // motion search: XVCB200_CU_INTRA / XVCB200_CU_SKIP_ME are set by the host and never searched.
  auto searched = [&](int i) { return !(cus[i].flags & (XVCB200_CU_INTRA | XVCB200_CU_SKIP_ME)); };
  const int ctus_x = (c->width + 63) >> 6, n_ctus = ctus_x * ((c->height + 63) >> 6);
  std::vector<int> ctu_first((size_t)n_ctus + 1, 0), by_ctu((size_t)n);
  for (int i = 0; i < n; i++)
    if (searched(i)) ctu_first[(size_t)((cus[i].y >> 6) * ctus_x + (cus[i].x >> 6)) + 1]++;
  // A search group = a run of at most kGroupCus (48: CTUs of more are halved) searched CUs of one CTU (coding order): the unit the persistent search
  // CTAs take from their counter.  Whole CTUs of 64 small CUs next to CTUs of one CU balance badly across 148 CTAs.
  static const int kGroupCus = getenv("XVCB_GROUP_CUS") ? std::max(1, atoi(getenv("XVCB_GROUP_CUS"))) : 48;   // (the variable: experiments)
  int n_groups = 0;
  for (int t = 0; t < n_ctus; t++) {
    n_groups += (ctu_first[(size_t)t + 1] + kGroupCus - 1) / kGroupCus;
    ctu_first[(size_t)t + 1] += ctu_first[(size_t)t];
  }
  {
    std::vector<int> fill(ctu_first.begin(), ctu_first.end() - 1);
    for (int i = 0; i < n; i++)
      if (searched(i)) by_ctu[(size_t)fill[(size_t)((cus[i].y >> 6) * ctus_x + (cus[i].x >> 6))]++] = i;
  }
  const int n_search = ctu_first[(size_t)n_ctus];
  // blob layout: [cus][tu list 3n][CU index by CTU: n][groups: 2g]
  const size_t ints = 3 * (size_t)n + (size_t)n + 2 * (size_t)n_groups;
  const size_t need = sizeof(xvcb200_cu) * (size_t)n + sizeof(int) * ints;
  const int nb = c->ex.cur_blob < 0 ? 0 : c->ex.cur_blob ^ 1;
  CtxExtra::CuBlob &B = c->ex.blob[nb];
  if (!B.up_ev && (!c->check(cudaEventCreateWithFlags(&B.up_ev, cudaEventDisableTiming), "cudaEventCreate") ||
                   !c->check(cudaEventCreateWithFlags(&B.free_ev, cudaEventDisableTiming), "cudaEventCreate") ||
                   !c->check(cudaEventCreateWithFlags(&B.dl_ev, cudaEventDisableTiming), "cudaEventCreate")))
    return c->status;
  if (!c->check(cudaEventSynchronize(B.up_ev), "set_cus staging")) return c->status;   // earlier copy out of its host image
  if (need > B.hcap) {
    if (B.h) cudaFreeHost(B.h);
    B.h = nullptr; B.hcap = 0;
    if (!c->check(cudaHostAlloc(&B.h, need + need / 4, cudaHostAllocDefault), "cudaHostAlloc")) return c->status;
    B.hcap = need + need / 4;
  }
  if (need > B.cap) {                       // rare: the picture before the current one may still be running on it
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->ex.copy_stream);
    cudaFree(B.d); B.d = nullptr; B.cap = 0;
    if (!c->check(cudaMalloc(&B.d, need + need / 4), "cudaMalloc(cu blob)")) return c->status;
    B.cap = need + need / 4;
  }
  xvcb200_cu *hc = static_cast<xvcb200_cu *>(B.h);
  int *list = reinterpret_cast<int *>(hc + n);
  int *idx1 = list + 3 * (size_t)n, *grp1 = idx1 + n;
  memcpy(hc, cus, sizeof(xvcb200_cu) * (size_t)n);
  c->ex.h_cus.assign(cus, cus + n);
  for (int i = 0; i < n; i++) idx1[i] = i < n_search ? by_ctu[(size_t)i] : 0;
  // Groups are handed to the persistent search CTAs in this order: longest first (cost grows with
  // the number of jobs), so that the groups still running when the counter runs out are the short ones.
  std::vector<std::pair<int, int>> runs;       // (first, count)
  runs.reserve((size_t)n_groups);
  for (int t = 0; t < n_ctus; t++) {
    const int first = ctu_first[(size_t)t], cnt = ctu_first[(size_t)t + 1] - first;
    const int pieces = (cnt + kGroupCus - 1) / kGroupCus;
    for (int k = 0; k < pieces; k++) {
      const int a = (int)((long long)cnt * k / pieces), b = (int)((long long)cnt * (k + 1) / pieces);
      runs.emplace_back(first + a, b - a);
    }
  }
  std::stable_sort(runs.begin(), runs.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
  for (int g = 0; g < n_groups; g++) { grp1[2 * g] = runs[(size_t)g].first; grp1[2 * g + 1] = runs[(size_t)g].second; }
  c->ex.pipe_n_groups = n_groups;
  auto lg = [](int v) { int l = 0; while ((1 << l) < v) l++; return l; };
  memset(c->ex.class_count, 0, sizeof(c->ex.class_count));
  for (int i = 0; i < n; i++)
    for (int comp = 0; comp < 3; comp++) c->ex.class_count[lg(cus[i].w >> (comp ? 1 : 0))][lg(cus[i].h >> (comp ? 1 : 0))]++;
  int off = 0, fill[7][7];
  for (int a = 0; a < 7; a++)
    for (int b = 0; b < 7; b++) { c->ex.class_offset[a][b] = off; fill[a][b] = off; off += c->ex.class_count[a][b]; }
  for (int i = 0; i < n; i++)
    for (int comp = 0; comp < 3; comp++) list[fill[lg(cus[i].w >> (comp ? 1 : 0))][lg(cus[i].h >> (comp ? 1 : 0))]++] = 3 * i + comp;
  // the blob being left: free once the work enqueued so far is done
  if (c->ex.cur_blob >= 0) c->check(cudaEventRecord(c->ex.blob[c->ex.cur_blob].free_ev, c->stream), "cudaEventRecord");
  // upload on the upload stream, after the last users of this blob; the context stream waits for it
  c->check(cudaStreamWaitEvent(c->ex.up_stream, B.free_ev, 0), "cudaStreamWaitEvent");
  if (B.dl_pending) { c->check(cudaStreamWaitEvent(c->ex.up_stream, B.dl_ev, 0), "cudaStreamWaitEvent"); B.dl_pending = false; }
  if (!c->check(cudaMemcpyAsync(B.d, B.h, need, cudaMemcpyHostToDevice, c->ex.up_stream), "set_cus") ||
      !c->check(cudaEventRecord(B.up_ev, c->ex.up_stream), "set_cus") ||
      !c->check(cudaStreamWaitEvent(c->stream, B.up_ev, 0), "set_cus"))
    return c->status;
  c->ex.cur_blob = nb;
  c->d_cus = reinterpret_cast<xvcb200_cu *>(B.d);
  int *dl = reinterpret_cast<int *>(c->d_cus + n);
  c->ex.d_tu_list = dl;
  c->ex.d_pipe_index = dl + 3 * (size_t)n;
  c->ex.d_pipe_groups = c->ex.d_pipe_index + n;
  return XVCB200_OK;
}

int xvcb200_set_tu_modes(xvcb200_ctx *ctx, const xvcb200_tu_mode *modes) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  c->ex.tu_modes_set = false;
  if (!modes || c->n_cus == 0) return XVCB200_OK;
  for (int i = 0; i < c->n_cus; i++)
    if (modes[i].tx_ver > XVCB200_TX_DST7 || modes[i].tx_hor > XVCB200_TX_DST7 || modes[i].tskip > 7 || modes[i].scan[0] > 2 ||
        modes[i].scan[1] > 2 || modes[i].scan[2] > 2)
      return XVCB200_INVALID_ARGUMENT;
  if (!ensure(c, &c->ex.d_tu_modes, &c->ex.tu_modes_cap, c->n_cus)) return c->status;
  // pageable host array: staged by the runtime before the call returns (the caller owns `modes`)
  if (!c->check(cudaMemcpyAsync(c->ex.d_tu_modes, modes, sizeof(*modes) * (size_t)c->n_cus, cudaMemcpyHostToDevice, c->stream), "tu modes"))
    return c->status;
  c->ex.tu_modes_set = true;
  return XVCB200_OK;
}

int xvcb200_set_mv_predictors(xvcb200_ctx *ctx, const int32_t *mvp, int n_cols) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || n_cols < 0 || n_cols > 10 || (n_cols > 0 && !mvp)) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  c->ex.mvp_cols = 0;
  if (n_cols == 0 || c->n_cus == 0) return XVCB200_OK;
  const int count = 2 * c->n_cus * n_cols;
  if (!ensure(c, &c->ex.d_mvp, &c->ex.mvp_cap, count)) return c->status;
  // through page-locked staging (two buffers, alternating): the call never waits for the kernels in flight
  const int b = (int)(c->ex.mvp_next++ & 1);
  const size_t bytes = sizeof(int32_t) * (size_t)count;
  if (!c->ex.mvp_ev[b] && !c->check(cudaEventCreateWithFlags(&c->ex.mvp_ev[b], cudaEventDisableTiming), "cudaEventCreate")) return c->status;
  if (!c->check(cudaEventSynchronize(c->ex.mvp_ev[b]), "mv predictors staging")) return c->status;
  if (bytes > c->ex.h_mvp_cap[b]) {
    if (c->ex.h_mvp[b]) cudaFreeHost(c->ex.h_mvp[b]);
    c->ex.h_mvp[b] = nullptr; c->ex.h_mvp_cap[b] = 0;
    if (!c->check(cudaHostAlloc(&c->ex.h_mvp[b], bytes + bytes / 4, cudaHostAllocDefault), "cudaHostAlloc")) return c->status;
    c->ex.h_mvp_cap[b] = bytes + bytes / 4;
  }
  memcpy(c->ex.h_mvp[b], mvp, bytes);
  if (!c->check(cudaMemcpyAsync(c->ex.d_mvp, c->ex.h_mvp[b], bytes, cudaMemcpyHostToDevice, c->stream), "mv predictors") ||
      !c->check(cudaEventRecord(c->ex.mvp_ev[b], c->stream), "mv predictors"))
    return c->status;
  c->ex.mvp_cols = n_cols;
  return XVCB200_OK;
}

int xvcb200_get_cus(xvcb200_ctx *c, xvcb200_cu *cus, int n) {
  xvcb::DevGuard dev_guard(c);
  if (!c || !cus || n < 0 || n > c->n_cus) return XVCB200_INVALID_ARGUMENT;
  if (!c->check(cudaMemcpyAsync(cus, c->d_cus, sizeof(xvcb200_cu) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "get_cus"))
    return c->status;
  return xvcb200_sync(c);
}

}  // extern "C"

namespace xvcb { size_t tz_state_bytes(); int tz_max_ctas(); }

// Per-job search state, the survivor pools and the group counter of tz_search_kernel.
static bool ensure_tz_scratch(CtxFull *c, int n_jobs) {
  if (!c->ex.d_pool) {
    c->ex.pool_ctas = xvcb::tz_max_ctas();
    if (!c->check(cudaMalloc(&c->ex.d_pool, sizeof(uint32_t) * (size_t)c->ex.pool_cap * c->ex.pool_ctas), "cudaMalloc(pool)")) return false;
  }
  int st_cap_elems = c->ex.tz_states_cap;
  if (!ensure(c, &c->ex.d_tz_states, &st_cap_elems, (int)(n_jobs * xvcb::tz_state_bytes()))) return false;
  c->ex.tz_states_cap = st_cap_elems;
  if (!c->ex.d_counter && !c->check(cudaMalloc(&c->ex.d_counter, sizeof(int)), "cudaMalloc(counter)")) return false;
  if (!ensure(c, &c->ex.d_subpel_lists, &c->ex.subpel_lists_cap, 15 * n_jobs + 32)) return false;
  return true;
}

// Groups search jobs by (reference slot, CTU of the CU) and uploads the index + group arrays.
// key(i) must be cheap; jobs of one group end up adjacent in job_index, in their original order.
template <class KeyOf>
static int upload_tz_groups(CtxFull *c, int n_jobs, KeyOf key_of) {
  std::vector<std::pair<long long, int>> order((size_t)n_jobs);
  for (int i = 0; i < n_jobs; i++) order[(size_t)i] = std::make_pair(key_of(i), i);
  std::stable_sort(order.begin(), order.end(), [](const std::pair<long long, int> &a, const std::pair<long long, int> &b) { return a.first < b.first; });
  std::vector<int> index((size_t)n_jobs), groups;
  for (int i = 0; i < n_jobs; i++) {
    index[(size_t)i] = order[(size_t)i].second;
    if (i == 0 || order[(size_t)i].first != order[(size_t)i - 1].first) { groups.push_back(i); groups.push_back(0); }
    groups.back()++;
  }
  const int n_groups = (int)groups.size() / 2;
  {   // longest groups first (see set_cus)
    std::vector<std::pair<int, int>> g2((size_t)n_groups);
    for (int g = 0; g < n_groups; g++) g2[(size_t)g] = std::make_pair(groups[2 * (size_t)g], groups[2 * (size_t)g + 1]);
    std::stable_sort(g2.begin(), g2.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
    for (int g = 0; g < n_groups; g++) { groups[2 * (size_t)g] = g2[(size_t)g].first; groups[2 * (size_t)g + 1] = g2[(size_t)g].second; }
  }
  if (!ensure(c, &c->ex.d_job_index, &c->ex.job_index_cap, n_jobs) || !ensure(c, &c->ex.d_groups, &c->ex.groups_cap, 2 * n_groups))
    return -1;
  if (!ensure_tz_scratch(c, n_jobs)) return -1;
  // synchronous copies: the vectors die at return
  if (!c->check(cudaMemcpyAsync(c->ex.d_job_index, index.data(), sizeof(int) * (size_t)n_jobs, cudaMemcpyHostToDevice, c->stream), "tz index") ||
      !c->check(cudaMemcpyAsync(c->ex.d_groups, groups.data(), sizeof(int) * groups.size(), cudaMemcpyHostToDevice, c->stream), "tz groups") ||
      !c->check(cudaStreamSynchronize(c->stream), "tz groups sync"))
    return -1;
  return n_groups;
}

extern "C" {

static uint32_t lambda_me_of(double lambda_sqrt) { return (uint32_t)std::floor(65536.0 * lambda_sqrt); }

int xvcb200_me_search(xvcb200_ctx *ctx, int orig_slot, const xvcb200_me_job *jobs, int n, double lambda_sqrt,
                      xvcb200_me_result *results) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, orig_slot) || !jobs || !results || n < 0) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++)
    if (jobs[i].cu < 0 || jobs[i].cu >= c->n_cus || !slot_ok(c, jobs[i].ref_slot) || jobs[i].search_range < 1 ||
        jobs[i].search_range > 256)
      return XVCB200_INVALID_ARGUMENT;
  if (n == 0) return XVCB200_OK;
  if (!ensure(c, &c->ex.d_jobs, &c->ex.jobs_cap, n) || !ensure(c, &c->ex.d_me, &c->ex.me_cap, n)) return c->status;
  join_upload_slot(c, orig_slot);
  for (int i = 0; i < n; i++) join_upload_slot(c, jobs[i].ref_slot);
  c->check(cudaMemcpyAsync(c->ex.d_jobs, jobs, sizeof(*jobs) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "me jobs");
  const int ctus_x = (c->width + 63) >> 6;
  const int n_groups = upload_tz_groups(c, n, [&](int i) {
    const xvcb200_cu &u = c->ex.h_cus[(size_t)jobs[i].cu];
    return ((long long)jobs[i].ref_slot << 40) | ((long long)jobs[i].search_range << 28) | (long long)((u.y >> 6) * ctus_x + (u.x >> 6));
  });
  if (n_groups < 0) return c->status;
  c->check(launch_tz_search(c->stream, c->d_cus, c->ex.d_jobs, n, c->bitdepth, lambda_me_of(lambda_sqrt),
                            c->plane(orig_slot, 0), c->ex.d_luma_views, c->ex.d_me, c->ex.d_job_index, c->ex.d_groups,
                            n_groups, c->ex.d_tz_states, c->ex.d_counter, c->ex.d_pool, c->ex.pool_cap), "tz_search");
  c->check(launch_subpel_search(c->stream, c->d_cus, c->ex.d_jobs, n, c->bitdepth, lambda_me_of(lambda_sqrt),
                                c->plane(orig_slot, 0), c->ex.d_luma_views, c->ex.d_me, c->ex.d_subpel_lists, c->ex.side, c->ex.side_ev,
                                c->ex.n_side, c->ex.fork_ev), "subpel_search");
  c->check(cudaMemcpyAsync(results, c->ex.d_me, sizeof(*results) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "me results");
  return xvcb200_sync(c);
}

int xvcb200_full_search(xvcb200_ctx *ctx, int orig_slot, const xvcb200_fullsearch_job *jobs, int n, double lambda_sqrt,
                        xvcb200_me_result *results) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, orig_slot) || !jobs || !results || n < 0) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++)
    if (jobs[i].cu < 0 || jobs[i].cu >= c->n_cus || !slot_ok(c, jobs[i].ref_slot) || !slot_ok(c, jobs[i].other_pred_slot) ||
        jobs[i].range < 0 || jobs[i].range > 64)
      return XVCB200_INVALID_ARGUMENT;
  if (n == 0) return XVCB200_OK;
  xvcb200_fullsearch_job *dj = static_cast<xvcb200_fullsearch_job *>(c->scratch(sizeof(*jobs) * (size_t)n));
  if (!dj || !ensure(c, &c->ex.d_me, &c->ex.me_cap, n)) return c->status;
  join_uploads(c);
  c->check(cudaMemcpyAsync(dj, jobs, sizeof(*jobs) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "fs jobs");
  c->check(launch_full_search(c->stream, c->d_cus, dj, n, c->bitdepth, lambda_me_of(lambda_sqrt), c->plane(orig_slot, 0),
                              c->ex.d_luma_views, c->ex.d_me), "full_search");
  c->check(cudaMemcpyAsync(results, c->ex.d_me, sizeof(*results) * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "fs results");
  return xvcb200_sync(c);
}

int xvcb200_intra_satd_scan(xvcb200_ctx *ctx, int orig_slot, int src_slot, const xvcb200_intra_job *jobs, int n,
                            uint32_t *satd) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, orig_slot) || !slot_ok(ctx, src_slot) || !jobs || !satd || n < 0) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++) {
    const xvcb200_intra_job &j = jobs[i];
    const bool pow2 = j.w >= 4 && j.w <= 64 && j.h >= 4 && j.h <= 64 && !(j.w & (j.w - 1)) && !(j.h & (j.h - 1));
    if (!pow2 || j.x < 0 || j.y < 0 || j.x + j.w > c->width || j.y + j.h > c->height || j.above_right > j.h || j.below_left > j.w ||
        ((j.has_above_left || j.has_left || j.below_left) && j.x == 0) || ((j.has_above_left || j.has_above || j.above_right) && j.y == 0) ||
        j.x + j.w + j.above_right > c->width || j.y + j.h + j.below_left > c->height)
      return XVCB200_INVALID_ARGUMENT;
  }
  if (n == 0) return XVCB200_OK;
  const size_t job_bytes = (sizeof(*jobs) * (size_t)n + 15) & ~(size_t)15, out_bytes = sizeof(uint32_t) * XVCB200_INTRA_NUM_MODES * (size_t)n;
  uint8_t *d = static_cast<uint8_t *>(c->scratch(job_bytes + out_bytes));
  if (!d) return c->status;
  join_upload_slot(c, orig_slot);
  join_upload_slot(c, src_slot);
  c->check(cudaMemcpyAsync(d, jobs, sizeof(*jobs) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "intra jobs");
  c->check(launch_intra_satd_scan(c->stream, reinterpret_cast<const xvcb200_intra_job *>(d), n, c->bitdepth, c->plane(orig_slot, 0),
                                  c->plane(src_slot, 0), reinterpret_cast<uint32_t *>(d + job_bytes)), "intra_satd_scan");
  c->check(cudaMemcpyAsync(satd, d + job_bytes, out_bytes, cudaMemcpyDeviceToHost, c->stream), "intra satd");
  return xvcb200_sync(c);
}

int xvcb200_intra_lm_chroma(xvcb200_ctx *ctx, int rec_slot, const xvcb200_intra_job *jobs, int n, int pred_slot) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, rec_slot) || !slot_ok(ctx, pred_slot) || rec_slot == pred_slot || (!jobs && n > 0) || n < 0) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++) {
    const xvcb200_intra_job &j = jobs[i];
    const bool pow2 = j.w >= 4 && j.w <= 64 && j.h >= 4 && j.h <= 64 && !(j.w & (j.w - 1)) && !(j.h & (j.h - 1));
    if (!pow2 || j.x < 0 || j.y < 0 || (j.x & 1) || (j.y & 1) || j.x + j.w > c->width || j.y + j.h > c->height) return XVCB200_INVALID_ARGUMENT;
  }
  if (n == 0) return XVCB200_OK;
  const size_t job_bytes = sizeof(*jobs) * (size_t)n;
  uint8_t *d = static_cast<uint8_t *>(c->scratch(job_bytes));
  if (!d) return c->status;
  join_upload_slot(c, rec_slot);
  join_downloads(c, pred_slot);
  c->check(cudaMemcpyAsync(d, jobs, job_bytes, cudaMemcpyHostToDevice, c->stream), "lm chroma jobs");
  c->check(launch_intra_lm_chroma(c->stream, reinterpret_cast<const xvcb200_intra_job *>(d), n, c->bitdepth, c->plane(rec_slot, 0),
                                  c->plane(rec_slot, 1), c->plane(rec_slot, 2), c->plane(pred_slot, 1), c->plane(pred_slot, 2)),
           "intra_lm_chroma");
  return c->status;
}

// ref_slots[list][ref_idx] -> plane views.  Every entry a CU of the current array can reference must be
// a valid slot (false otherwise: the kernels would read another picture or a wild pointer); entries
// beyond that may be unset and get a placeholder view that is never dereferenced.
static bool refs_from_slots(xvcb200_ctx *c, const int32_t ref_slots[2][5], Pic3 refs[2][5]) {
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) {
      int s = ref_slots[l][i];
      const bool valid = s >= 0 && s < (int)c->slots.size();
      if (!valid && i <= full(c)->ex.max_ref_idx[l]) return false;
      refs[l][i] = pic3(c, valid ? s : 0);
    }
  return true;
}

int xvcb200_motion_compensate(xvcb200_ctx *c, const int32_t ref_slots[2][5], int pred_slot) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, pred_slot) || !ref_slots) return XVCB200_INVALID_ARGUMENT;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) join_upload_slot(full(c), ref_slots[l][i]);
  join_downloads(full(c), pred_slot);
  Pic3 refs[2][5];
  if (!refs_from_slots(c, ref_slots, refs)) return XVCB200_INVALID_ARGUMENT;
  c->check(launch_motion_compensate(c->stream, c->d_cus, c->n_cus, c->bitdepth, refs, pic3(c, pred_slot)), "motion_compensate");
  return c->status;
}

int xvcb200_motion_compensate_affine(xvcb200_ctx *ctx, const xvcb200_affine_cu *aff, int n, const int32_t ref_slots[2][5],
                                     int pred_slot) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, pred_slot) || !ref_slots || n < 0 || (n > 0 && !aff)) return XVCB200_INVALID_ARGUMENT;
  if (n == 0) return XVCB200_OK;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++)
    if (aff[i].cu < 0 || aff[i].cu >= c->n_cus) return XVCB200_INVALID_ARGUMENT;
  if (!ensure(c, &c->ex.d_affine, &c->ex.affine_cap, n)) return c->status;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) join_upload_slot(c, ref_slots[l][i]);
  join_downloads(c, pred_slot);
  // pageable host array: the copy is staged by the runtime before the call returns (caller owns `aff`)
  if (!c->check(cudaMemcpyAsync(c->ex.d_affine, aff, sizeof(*aff) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "affine upload"))
    return c->status;
  Pic3 refs[2][5];
  if (!refs_from_slots(c, ref_slots, refs)) return XVCB200_INVALID_ARGUMENT;
  c->check(launch_motion_compensate_affine(c->stream, c->d_cus, c->n_cus, c->ex.d_affine, n, c->bitdepth, refs, pic3(c, pred_slot)),
           "motion_compensate_affine");
  return c->status;
}

int xvcb200_motion_compensate_lic(xvcb200_ctx *ctx, const xvcb200_lic_cu *lic, int n, const int32_t ref_slots[2][5], int rec_slot,
                                  int pred_slot) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, pred_slot) || !slot_ok(ctx, rec_slot) || !ref_slots || n < 0 || (n > 0 && !lic)) return XVCB200_INVALID_ARGUMENT;
  if (n == 0) return XVCB200_OK;
  CtxFull *c = full(ctx);
  for (int i = 0; i < n; i++)
    if (lic[i].cu < 0 || lic[i].cu >= c->n_cus) return XVCB200_INVALID_ARGUMENT;
  if (!ensure(c, &c->ex.d_lic, &c->ex.lic_cap, n)) return c->status;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) join_upload_slot(c, ref_slots[l][i]);
  join_upload_slot(c, rec_slot);
  join_downloads(c, pred_slot);
  if (!c->check(cudaMemcpyAsync(c->ex.d_lic, lic, sizeof(*lic) * (size_t)n, cudaMemcpyHostToDevice, c->stream), "lic upload"))
    return c->status;
  Pic3 refs[2][5];
  if (!refs_from_slots(c, ref_slots, refs)) return XVCB200_INVALID_ARGUMENT;
  c->check(launch_motion_compensate_lic(c->stream, c->d_cus, c->n_cus, c->ex.d_lic, n, c->bitdepth, refs, pic3(c, rec_slot),
                                        pic3(c, pred_slot)), "motion_compensate_lic");
  return c->status;
}

static int tq_common(xvcb200_ctx *ctx, int orig_slot, int pred_slot, int rec_slot, int coeff_slot, int intra_picture,
                     int table, int off_u, int off_v, int decode_only, xvcb200_tu_result *results) {
  xvcb::DevGuard dev_guard(ctx);
  if (!slot_ok(ctx, pred_slot) || !slot_ok(ctx, rec_slot) || !slot_ok(ctx, coeff_slot) || (!decode_only && !slot_ok(ctx, orig_slot)))
    return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  const int n = c->n_cus;
  if (n == 0) return XVCB200_OK;
  if (!ensure(c, &c->ex.d_tu, &c->ex.tu_res_cap, 3 * n)) return c->status;
  TqParams p;
  join_upload_slot(c, orig_slot);
  join_upload_slot(c, pred_slot);
  join_upload_slot(c, coeff_slot);
  join_downloads(c, rec_slot);
  join_downloads(c, coeff_slot);
  join_downloads(c, -1);
  p.bitdepth = c->bitdepth; p.intra_picture = intra_picture; p.table = table; p.off_u = off_u; p.off_v = off_v;
  p.decode_only = decode_only;
  p.modes = c->ex.tu_modes_set ? c->ex.d_tu_modes : nullptr;
  int16_t *lev[3]; int pitch[3];
  for (int k = 0; k < 3; k++) { lev[k] = reinterpret_cast<int16_t *>(c->slots[coeff_slot].base[k]); pitch[k] = c->geom.pitch[k]; }
  c->check(launch_tq_reconstruct_classes(c->stream, c->d_cus, c->ex.d_tu_list, c->ex.class_count, c->ex.class_offset, p,
                                         pic3(c, decode_only ? pred_slot : orig_slot), pic3(c, pred_slot), pic3(c, rec_slot),
                                         lev, pitch, decode_only ? nullptr : c->ex.d_tu, c->ex.side, c->ex.side_ev,
                                         c->ex.n_side, c->ex.fork_ev), "tq_reconstruct");
  if (results && !decode_only) {
    c->check(cudaMemcpyAsync(results, c->ex.d_tu, sizeof(*results) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "tu results");
    return xvcb200_sync(c);
  }
  return c->status;
}

int xvcb200_tq_reconstruct(xvcb200_ctx *c, int orig_slot, int pred_slot, int rec_slot, int coeff_slot, int pic_qp_unused,
                           int intra_picture, int table, int off_u, int off_v, xvcb200_tu_result *results) {
  xvcb::DevGuard dev_guard(c);
  (void)pic_qp_unused;
  return tq_common(c, orig_slot, pred_slot, rec_slot, coeff_slot, intra_picture, table, off_u, off_v, 0, results);
}
int xvcb200_dequant_reconstruct(xvcb200_ctx *c, int pred_slot, int rec_slot, int coeff_slot, int table, int off_u, int off_v) {
  xvcb::DevGuard dev_guard(c);
  return tq_common(c, 0, pred_slot, rec_slot, coeff_slot, 0, table, off_u, off_v, 1, nullptr);
}
// levels into a coefficient slot (decoder side input)
int xvcb200_upload_coeff(xvcb200_ctx *c, int slot, const int16_t *const planes[3], const ptrdiff_t strides[3]) {
  xvcb::DevGuard dev_guard(c);
  return xvcb200_upload_picture(c, slot, reinterpret_cast<const uint16_t *const *>(planes), strides);
}

int xvcb200_deblock_band(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table, int off_u,
                         int off_v, const int64_t ref_poc[2][5], int pass_mask, int y_begin, int y_end);
int xvcb200_deblock_picture(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset,
                            const int64_t ref_poc[2][5]) {
  xvcb::DevGuard dev_guard(c);
  return xvcb200_deblock_picture_ex(c, rec_slot, pic_type, beta_offset, tc_offset, 1, 0, 0, ref_poc);
}
int xvcb200_deblock_picture_ex(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table,
                               int off_u, int off_v, const int64_t ref_poc[2][5]) {
  xvcb::DevGuard dev_guard(c);
  if (!c) return XVCB200_INVALID_ARGUMENT;
  return xvcb200_deblock_band(c, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v, ref_poc, 3, 0, c->height);
}
static int deblock_impl(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table, int off_u,
                        int off_v, const int64_t ref_poc[2][5], int pass_mask, int y_begin, int y_end, bool map_ready,
                        const xvcb200_deblock_ext *ext = nullptr);
int xvcb200_deblock_picture_ext(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table,
                                int off_u, int off_v, const int64_t ref_poc[2][5], const xvcb200_deblock_ext *ext) {
  xvcb::DevGuard dev_guard(c);
  if (!c) return XVCB200_INVALID_ARGUMENT;
  return deblock_impl(c, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v, ref_poc, 3, 0, c->height, false, ext);
}
int xvcb200_deblock_band(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table, int off_u,
                         int off_v, const int64_t ref_poc[2][5], int pass_mask, int y_begin, int y_end) {
  xvcb::DevGuard dev_guard(c);
  return deblock_impl(c, rec_slot, pic_type, beta_offset, tc_offset, table, off_u, off_v, ref_poc, pass_mask, y_begin, y_end, false);
}
static int deblock_impl(xvcb200_ctx *c, int rec_slot, int pic_type, int beta_offset, int tc_offset, int table, int off_u,
                        int off_v, const int64_t ref_poc[2][5], int pass_mask, int y_begin, int y_end, bool map_ready,
                        const xvcb200_deblock_ext *ext) {
  xvcb::DevGuard dev_guard(c);
  if (!slot_ok(c, rec_slot) || !ref_poc || pic_type < 0 || pic_type > 2 || y_begin < 0 || y_end > c->height ||
      y_begin > y_end || (y_begin & 3) || (y_end & 3) || (pass_mask & ~3))
    return XVCB200_INVALID_ARGUMENT;
  join_upload_slot(full(c), rec_slot);
  join_downloads(full(c), rec_slot);
  DeblockParams p;
  p.bitdepth = c->bitdepth; p.pic_type = pic_type; p.beta_offset = beta_offset; p.tc_offset = tc_offset;
  p.table = table; p.off_u = off_u; p.off_v = off_v;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) p.ref_poc[l][i] = ref_poc[l][i];
  if (ext && ((ext->n_affine > 0 && ext->affine) || (ext->n_chroma_cus > 0 && ext->chroma_cus))) {
    CtxFull *f = full(c);
    const int na = ext->affine ? ext->n_affine : 0, nc = ext->chroma_cus ? ext->n_chroma_cus : 0;
    if (na < 0 || nc < 0) return XVCB200_INVALID_ARGUMENT;
    for (int i = 0; i < na; i++)
      if (ext->affine[i].cu < 0 || ext->affine[i].cu >= c->n_cus) return XVCB200_INVALID_ARGUMENT;
    for (int i = 0; i < nc; i++) {
      const xvcb200_cu &u = ext->chroma_cus[i];
      if (u.w < 4 || u.h < 4 || u.w > 64 || u.h > 64 || u.x < 0 || u.y < 0 || u.x + u.w > c->width || u.y + u.h > c->height || (u.x & 3) || (u.y & 3))
        return XVCB200_INVALID_ARGUMENT;
    }
    // one staging image: [affine index per CU][affine entries][chroma CUs]
    const size_t off_aff = (sizeof(int) * (size_t)(na ? c->n_cus : 0) + 15) & ~(size_t)15;
    const size_t off_cus = (off_aff + sizeof(xvcb200_affine_cu) * (size_t)na + 15) & ~(size_t)15;
    const size_t bytes = off_cus + sizeof(xvcb200_cu) * (size_t)nc;
    std::vector<uint8_t> img(bytes, 0);
    if (na) {
      int *index = reinterpret_cast<int *>(img.data());
      for (int i = 0; i < c->n_cus; i++) index[i] = -1;
      for (int i = 0; i < na; i++) index[ext->affine[i].cu] = i;
      memcpy(img.data() + off_aff, ext->affine, sizeof(xvcb200_affine_cu) * (size_t)na);
    }
    if (nc) memcpy(img.data() + off_cus, ext->chroma_cus, sizeof(xvcb200_cu) * (size_t)nc);
    uint8_t *d = static_cast<uint8_t *>(c->scratch(bytes));
    if (!d) return c->status;
    if (nc && !f->ex.d_cu_map2 &&
        !c->check(cudaMalloc(&f->ex.d_cu_map2, sizeof(int32_t) * (size_t)c->map_w * c->map_h), "cudaMalloc(chroma map)"))
      return c->status;
    // pageable source: the runtime stages it before the call returns
    if (!c->check(cudaMemcpyAsync(d, img.data(), bytes, cudaMemcpyHostToDevice, c->stream), "deblock ext")) return c->status;
    if (na) { p.aff_index = reinterpret_cast<const int *>(d); p.aff = reinterpret_cast<const xvcb200_affine_cu *>(d + off_aff); }
    if (nc) { p.chroma_cus = reinterpret_cast<const xvcb200_cu *>(d + off_cus); p.n_chroma_cus = nc; p.chroma_map = f->ex.d_cu_map2; }
  }
  c->check(launch_deblock(c->stream, c->d_cus, c->n_cus, p, pic3(c, rec_slot), c->d_cu_map, c->d_edge_bs[0], c->d_edge_bs[1],
                          c->map_w, c->map_h, pass_mask, y_begin, y_end, map_ready), "deblock");
  return c->status;
}

// The pre-analysis in two halves, so that a caller can enqueue other work (the previous picture's kernels) between
// them: _begin enqueues the kernel and the copy of its result into page-locked memory, _end waits for that copy only.
int xvcb200_decide_partition_begin(xvcb200_ctx *ctx, const xvcb200_partition_params *prm) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !prm || !slot_ok(ctx, prm->orig_slot) || !slot_ok(ctx, prm->ref_slot) || prm->header_bits_cu < 0 ||
      prm->header_bits_split < 0 || prm->header_bits_cu > 255 || prm->header_bits_split > 255)
    return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  const int n_ctus = ((c->width + 63) >> 6) * ((c->height + 63) >> 6);
  const size_t b_cus = sizeof(xvcb200_cu) * 64 * (size_t)n_ctus, b_spl = 128 * (size_t)n_ctus, b_cnt = sizeof(int) * (size_t)n_ctus;
  const size_t total = b_cus + b_spl + 2 * b_cnt;
  if (!c->ex.d_part) {
    if (!c->check(cudaMalloc(&c->ex.d_part, total), "cudaMalloc(partition)") ||
        !c->check(cudaHostAlloc(&c->ex.h_part, total, cudaHostAllocDefault), "cudaHostAlloc(partition)") ||
        !c->check(cudaEventCreateWithFlags(&c->ex.part_ev, cudaEventDisableTiming), "cudaEventCreate") ||
        !c->check(cudaEventCreateWithFlags(&c->ex.part_k_ev, cudaEventDisableTiming), "cudaEventCreate") ||
        !c->check(cudaStreamCreateWithFlags(&c->ex.part_stream, cudaStreamNonBlocking), "cudaStreamCreate(partition)"))
      return c->status;
  } else {
    // the kernel rewrites d_part: not before the previous result has left it (a caller that skipped _end)
    c->check(cudaStreamWaitEvent(c->stream, c->ex.part_ev, 0), "cudaStreamWaitEvent");
  }
  join_upload_slot(c, prm->orig_slot);
  join_upload_slot(c, prm->ref_slot);
  const uint32_t lam = lambda_me_of(prm->lambda_sqrt);
  const int bits_cu = prm->header_bits_cu ? prm->header_bits_cu : 8, bits_split = prm->header_bits_split ? prm->header_bits_split : 1;
  uint8_t *d = c->ex.d_part;
  xvcb200_cu *d_cus = reinterpret_cast<xvcb200_cu *>(d);
  uint8_t *d_spl = d + b_cus;
  int *d_ncu = reinterpret_cast<int *>(d + b_cus + b_spl), *d_nsp = d_ncu + n_ctus;
  c->check(launch_partition(c->stream, c->plane(prm->orig_slot, 0), c->plane(prm->ref_slot, 0), prm->center[0], prm->center[1], lam,
                            (int)(((unsigned long long)lam * bits_cu) >> 16), (int)(((unsigned long long)lam * bits_split) >> 16), prm->qp,
                            d_cus, d_ncu, d_spl, d_nsp), "partition");
  // the result travels on a copy stream of its own behind the kernel: on the context stream the 1 MB transfer
  // (~40 us of PCIe) would hold back the kernels of the picture enqueued next
  c->check(cudaEventRecord(c->ex.part_k_ev, c->stream), "cudaEventRecord");
  c->check(cudaStreamWaitEvent(c->ex.part_stream, c->ex.part_k_ev, 0), "cudaStreamWaitEvent");
  c->check(cudaMemcpyAsync(c->ex.h_part, d, total, cudaMemcpyDeviceToHost, c->ex.part_stream), "partition results");
  c->check(cudaEventRecord(c->ex.part_ev, c->ex.part_stream), "cudaEventRecord");
  c->ex.part_pending = true;
  return c->status;
}

int xvcb200_decide_partition_end(xvcb200_ctx *ctx, xvcb200_cu *cus_out, int cus_cap, int *n_cus, uint8_t *splits_out, int splits_cap,
                                 int *n_splits) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !cus_out || !n_cus || !splits_out || !n_splits) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (!c->ex.part_pending) return XVCB200_INVALID_ARGUMENT;
  c->ex.part_pending = false;
  if (!c->check(cudaEventSynchronize(c->ex.part_ev), "partition results")) return c->status;
  const int n_ctus = ((c->width + 63) >> 6) * ((c->height + 63) >> 6);
  const size_t b_cus = sizeof(xvcb200_cu) * 64 * (size_t)n_ctus, b_spl = 128 * (size_t)n_ctus;
  const uint8_t *h = c->ex.h_part;
  const xvcb200_cu *h_cus = reinterpret_cast<const xvcb200_cu *>(h);
  const uint8_t *h_spl = h + b_cus;
  const int *h_ncu = reinterpret_cast<const int *>(h + b_cus + b_spl), *h_nsp = h_ncu + n_ctus;
  int nc = 0, ns = 0;
  for (int t = 0; t < n_ctus; t++) {
    if (nc + h_ncu[t] > cus_cap || ns + h_nsp[t] > splits_cap) return XVCB200_INVALID_ARGUMENT;
    memcpy(cus_out + nc, h_cus + 64 * (size_t)t, sizeof(xvcb200_cu) * (size_t)h_ncu[t]);
    memcpy(splits_out + ns, h_spl + 128 * (size_t)t, (size_t)h_nsp[t]);
    nc += h_ncu[t]; ns += h_nsp[t];
  }
  *n_cus = nc; *n_splits = ns;
  return XVCB200_OK;
}

int xvcb200_decide_partition(xvcb200_ctx *ctx, const xvcb200_partition_params *prm, xvcb200_cu *cus_out, int cus_cap, int *n_cus,
                             uint8_t *splits_out, int splits_cap, int *n_splits) {
  if (!cus_out || !n_cus || !splits_out || !n_splits) return XVCB200_INVALID_ARGUMENT;
  const int st = xvcb200_decide_partition_begin(ctx, prm);
  if (st != XVCB200_OK) return st;
  return xvcb200_decide_partition_end(ctx, cus_out, cus_cap, n_cus, splits_out, splits_cap, n_splits);
}

// per-stage device times of the last xvcb200_encode_picture: ms[0..6] = make jobs, full-pel TZ
// search, sub-pel search + list decision, motion compensation, T/Q/recon, deblocking, padding
int xvcb200_set_profiling(xvcb200_ctx *ctx, int enable) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (enable && !c->ex.ev[0])
    for (auto &e : c->ex.ev)
      if (!c->check(cudaEventCreate(&e), "cudaEventCreate")) return c->status;
  c->ex.profile = enable != 0;
  c->ex.ev_valid = false;
  return XVCB200_OK;
}
int xvcb200_get_stage_times(xvcb200_ctx *ctx, float ms[7]) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !ms) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  if (!c->ex.ev_valid) return XVCB200_INVALID_ARGUMENT;
  if (!c->check(cudaEventSynchronize(c->ex.ev[7]), "cudaEventSynchronize")) return c->status;
  for (int i = 0; i < 7; i++)
    if (!c->check(cudaEventElapsedTime(&ms[i], c->ex.ev[i], c->ex.ev[i + 1]), "cudaEventElapsedTime")) return c->status;
  return XVCB200_OK;
}

// ---------------------------------------------------------------- (C) picture pipeline
int xvcb200_encode_picture(xvcb200_ctx *ctx, const xvcb200_picture_params *prm, xvcb200_me_result *me_results,
                           xvcb200_tu_result *tu_results) {
  xvcb::DevGuard dev_guard(ctx);
  if (!ctx || !prm) return XVCB200_INVALID_ARGUMENT;
  CtxFull *c = full(ctx);
  const int n = c->n_cus;
  if (!slot_ok(c, prm->orig_slot) || !slot_ok(c, prm->pred_slot) || !slot_ok(c, prm->rec_slot) || !slot_ok(c, prm->coeff_slot) ||
      prm->pic_type < 0 || prm->pic_type > 1 || prm->num_ref[0] > 5 || prm->num_ref[1] > 5 || prm->bi_iterations < 0 ||
      prm->bi_iterations > 8 || prm->bits_mode < 0 || prm->bits_mode > 1)
    return XVCB200_INVALID_ARGUMENT;
  // InterSearch::SearchMotion's loops over lists and reference pictures (inter_search.cc:199-259, 456-578)
  MePipe P;
  memset(&P, 0, sizeof(P));
  P.n = n;
  P.R[0] = prm->num_ref[0] > 0 ? prm->num_ref[0] : 1;
  P.R[1] = prm->pic_type == 1 ? 0 : (prm->num_ref[1] > 0 ? prm->num_ref[1] : 1);
  P.J = P.R[0] + P.R[1]; P.Rmax = P.R[0] > P.R[1] ? P.R[0] : P.R[1];
  P.pic_uni = prm->pic_type == 1; P.bits_mode = prm->bits_mode; P.bitdepth = c->bitdepth;
  P.bi_iterations = P.R[1] > 0 ? prm->bi_iterations : 0;
  P.lambda = lambda_me_of(prm->lambda_sqrt);
  if (P.bits_mode == 0 && (P.R[0] > 1 || P.R[1] > 1 || P.bi_iterations > 0)) return XVCB200_INVALID_ARGUMENT;
  if (c->ex.mvp_cols != 0 && c->ex.mvp_cols != P.J) return XVCB200_INVALID_ARGUMENT;    // predictors given for other lists
  P.mvp = c->ex.mvp_cols ? c->ex.d_mvp : nullptr;
  int cols[10], n_cols = 0;
  for (int l = 0; l < 2; l++)
    for (int r = 0; r < P.R[l]; r++) {
      // GetSearchRangeUniPred yields 96..256; the search kernel holds a pass of <= 9 rounds
      if (!slot_ok(c, prm->ref_slots[l][r]) || prm->search_range[l][r] < 1 || prm->search_range[l][r] > 256) return XVCB200_INVALID_ARGUMENT;
      P.ref_slot[l][r] = prm->ref_slots[l][r]; P.range[l][r] = prm->search_range[l][r];
      if (l == 1) {       // ReferencePictureLists::GetSamePocMappingFor (reference_picture_lists.cc:105-122); one picture per list (bits_mode 0): always searched
        P.dup_of[r] = -1;
        for (int q = 0; q < P.R[0] && P.bits_mode; q++)
          if (prm->ref_poc[1][r] == prm->ref_poc[0][q]) { P.dup_of[r] = q; break; }
        if (P.dup_of[r] >= 0 && prm->ref_slots[1][r] != prm->ref_slots[0][P.dup_of[r]]) return XVCB200_INVALID_ARGUMENT;
      }
      if (l == 0 || P.dup_of[r] < 0) cols[n_cols++] = l * P.R[0] + r;
    }
  if (n == 0) return XVCB200_OK;
  const int nj = n * P.J, nbj = n * P.Rmax;
  if (!ensure(c, &c->ex.d_jobs, &c->ex.jobs_cap, nj) || !ensure(c, &c->ex.d_me, &c->ex.me_cap, nj) ||
      !ensure(c, &c->ex.d_tu, &c->ex.tu_res_cap, 3 * n) ||
      !ensure(c, &c->ex.d_me_state, &c->ex.me_state_cap, (int)(n * me_cu_state_bytes())))
    return c->status;
  if (P.bi_iterations > 0) {
    if (!ensure(c, &c->ex.d_bi_jobs, &c->ex.bi_jobs_cap, nbj) || !ensure(c, &c->ex.d_bi_res, &c->ex.bi_res_cap, nbj)) return c->status;
    if (!c->ex.d_worig &&
        !c->check(cudaMalloc(&c->ex.d_worig, sizeof(int16_t) * (size_t)c->geom.pitch[0] * c->height), "cudaMalloc(weighted original)"))
      return c->status;
  }
  if (!ensure_tz_scratch(c, nj)) return c->status;
  join_upload_slot(c, prm->orig_slot);           // explicit read set: an upload made ahead for the NEXT picture is not waited for
  for (int l = 0; l < 2; l++)
    for (int r = 0; r < P.R[l]; r++) join_upload_slot(c, prm->ref_slots[l][r]);
  join_downloads(c, prm->pred_slot);
  join_downloads(c, prm->rec_slot);
  join_downloads(c, prm->coeff_slot);
  join_downloads(c, -1);
  for (int l = 0; l < 2; l++) c->ex.max_ref_idx[l] = std::max(c->ex.max_ref_idx[l], P.R[l] - 1);
  int stage = 0;
  auto mark = [&]() { if (c->ex.profile) cudaEventRecord(c->ex.ev[stage], c->stream); stage++; };
  mark();   // 0: start
  // the deblocking stage's CU map depends on the CU geometry only: built on a side stream behind the
  // search (the T/Q stage's join over all side streams orders it before the deblocking kernels)
  const bool map_ahead = prm->deblock && c->ex.n_side >= 3;
  if (map_ahead) {
    cudaEventRecord(c->ex.fork_ev, c->stream);
    cudaStreamWaitEvent(c->ex.side[2], c->ex.fork_ev, 0);
    c->check(launch_cu_map(c->ex.side[2], c->d_cus, n, c->d_cu_map, c->map_w, c->map_h), "cu_map");
  }
  const PlaneView orig = c->plane(prm->orig_slot, 0);
  c->check(launch_make_me_jobs(c->stream, c->d_cus, P, c->ex.d_jobs), "make_me_jobs");
  mark();   // 1: jobs built
  c->check(launch_tz_search(c->stream, c->d_cus, c->ex.d_jobs, nj, c->bitdepth, P.lambda, orig, c->ex.d_luma_views, c->ex.d_me,
                            c->ex.d_pipe_index, c->ex.d_pipe_groups, c->ex.pipe_n_groups, c->ex.d_tz_states, c->ex.d_counter,
                            c->ex.d_pool, c->ex.pool_cap, P.J, n_cols, cols), "tz_search");
  mark();   // 2: full-pel search done
  c->check(launch_subpel_search(c->stream, c->d_cus, c->ex.d_jobs, nj, c->bitdepth, P.lambda, orig, c->ex.d_luma_views, c->ex.d_me,
                                c->ex.d_subpel_lists, c->ex.side, c->ex.side_ev, c->ex.n_side, c->ex.fork_ev), "subpel_search");
  c->check(launch_me_uni_decide(c->stream, c->d_cus, P, c->ex.d_jobs, c->ex.d_me, c->ex.d_me_state), "me_uni_decide");
  if (P.bi_iterations > 0) {
    PlaneView worig = orig;
    worig.base = reinterpret_cast<Sample *>(c->ex.d_worig);
    FsTensorMaps fs_maps;
    memset(&fs_maps, 0, sizeof(fs_maps));
    for (int l = 0; l < 2; l++)
      for (int r = 0; r < P.R[l]; r++)
        for (int k = 0; k < 2; k++) fs_maps.m[l * P.R[0] + r][k] = c->ex.luma_tmaps[2 * (size_t)P.ref_slot[l][r] + k];
    for (int it = 0; it < P.bi_iterations; it++) {
      c->check(launch_bi_search(c->stream, c->d_cus, P, it, c->ex.d_jobs, c->ex.d_me, c->ex.d_me_state, orig, c->ex.d_luma_views, worig,
                                fs_maps, c->geom.margin_x[0], c->geom.margin_y[0], c->ex.d_bi_jobs, c->ex.d_bi_res,
                                c->ex.n_side >= 1 ? c->ex.side[0] : nullptr, c->ex.n_side >= 1 ? c->ex.side_ev[0] : nullptr, c->ex.fork_ev), "bi_search");
      c->check(launch_subpel_search(c->stream, c->d_cus, c->ex.d_bi_jobs, nbj, c->bitdepth, P.lambda, worig, c->ex.d_luma_views,
                                    c->ex.d_bi_res, c->ex.d_subpel_lists, c->ex.side, c->ex.side_ev, c->ex.n_side, c->ex.fork_ev),
               "subpel_search(bi)");
      c->check(launch_me_bi_decide(c->stream, c->d_cus, P, c->ex.d_jobs, c->ex.d_bi_res, c->ex.d_me, c->ex.d_me_state), "me_bi_decide");
    }
  }
  c->check(launch_me_final_decide(c->stream, c->d_cus, P, c->ex.d_me_state), "me_final_decide");
  mark();   // 3: sub-pel search + decision done
  Pic3 refs[2][5];
  if (!refs_from_slots(c, prm->ref_slots, refs)) return XVCB200_INVALID_ARGUMENT;
  c->check(launch_motion_compensate(c->stream, c->d_cus, n, c->bitdepth, refs, pic3(c, prm->pred_slot)), "motion_compensate");
  mark();   // 4: prediction done
  int st = tq_common(c, prm->orig_slot, prm->pred_slot, prm->rec_slot, prm->coeff_slot, 0, prm->chroma_offset_table,
                     prm->chroma_offset_u, prm->chroma_offset_v, 0, nullptr);
  if (st != XVCB200_OK) return st;
  mark();   // 5: T/Q/recon done
  if (prm->deblock) {
    st = deblock_impl(c, prm->rec_slot, prm->pic_type, prm->beta_offset, prm->tc_offset, prm->chroma_offset_table,
                      prm->chroma_offset_u, prm->chroma_offset_v, prm->ref_poc, 3, 0, c->height, map_ahead);
    if (st != XVCB200_OK) return st;
  }
  mark();   // 6: deblocking done
  if (prm->pad) xvcb200_pad_border(c, prm->rec_slot);
  mark();   // 7: padding done
  c->ex.ev_valid = c->ex.profile;
  if (me_results)
    c->check(cudaMemcpyAsync(me_results, c->ex.d_me, sizeof(*me_results) * (size_t)nj, cudaMemcpyDeviceToHost, c->stream), "me results");
  if (tu_results)
    c->check(cudaMemcpyAsync(tu_results, c->ex.d_tu, sizeof(*tu_results) * 3 * (size_t)n, cudaMemcpyDeviceToHost, c->stream), "tu results");
  return c->status;   // asynchronous: xvcb200_sync() completes it
}

}  // extern "C"
