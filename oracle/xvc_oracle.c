/*
 * TEST INFRASTRUCTURE ONLY -- see xvc_oracle.h.  Plain-C restatement of the xvc hot path,
 * written from the reference's behaviour (file:line cited per function), structured for
 * clarity rather than speed: matrix-form transforms instead of partial butterflies,
 * a generic fast Walsh-Hadamard for SATD, table-driven search patterns.
 */
#include "xvc_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef uint16_t Sample;

static inline int clip3(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ===================================================================================
 * Distortion metrics
 * =================================================================================== */

/* ComputeSad_c, sample_metric.cc:671-684.  kind 0: Sample/Sample, 1: Residual/Sample. */
int xo_sad(int kind, int w, int h, const void *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  int sum = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int va = kind == 0 ? ((const uint16_t *)a)[y * sa + x] : ((const int16_t *)a)[y * sa + x];
      sum += iabs(va - (int)b[y * sb + x]);
    }
  return sum;
}

/* ComputeSsd_c, sample_metric.cc:301-314.  `diff * diff` is an int product there; the
 * accumulator is uint64. */
uint64_t xo_ssd(int kind, int w, int h, const void *a, ptrdiff_t sa, const void *b, ptrdiff_t sb) {
  uint64_t ssd = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int va = kind == 0 ? ((const uint16_t *)a)[y * sa + x] : ((const int16_t *)a)[y * sa + x];
      int vb = kind == 2 ? ((const int16_t *)b)[y * sb + x] : ((const uint16_t *)b)[y * sb + x];
      int d = va - vb;
      ssd += (uint64_t)(int64_t)(d * d);
    }
  return ssd;
}

/* In-place fast Walsh-Hadamard over n elements at stride s.  Output order differs from
 * the reference's butterflies, but SATD only sums magnitudes, which is order-free. */
static void wht(int *v, int n, int s) {
  for (int len = 1; len < n; len <<= 1)
    for (int i = 0; i < n; i += len << 1)
      for (int j = i; j < i + len; j++) {
        int p = v[j * s], q = v[(j + len) * s];
        v[j * s] = p + q;
        v[(j + len) * s] = p - q;
      }
}

/* ComputeSatdNxM / ComputeSatd2x2, sample_metric.cc:403-668: Hadamard of the W x H
 * difference tile, sum of magnitudes, then the per-shape normalisation (:633-639). */
static int satd_tile(int W, int H, int first_short, const void *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  int d[16 * 16];
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      int va = first_short ? ((const int16_t *)a)[y * sa + x] : ((const uint16_t *)a)[y * sa + x];
      d[y * W + x] = va - (int)b[y * sb + x];
    }
  for (int y = 0; y < H; y++) wht(d + y * W, W, 1);
  for (int x = 0; x < W; x++) wht(d + x, H, W);
  int sum = 0;
  for (int i = 0; i < W * H; i++) sum += iabs(d[i]);
  if (W == 2 && H == 2) return sum;                       /* :643-668, no scaling */
  if (W == 4 && H == 4) return (sum + 1) >> 1;
  if (W == H) return (sum + 2) >> 2;
  return (int)(2.0 * sum / sqrt((double)(W * H)));
}

/* ComputeSatd<false>, sample_metric.cc:316-389: tile shape by block shape. */
uint64_t xo_satd(int bitdepth, int w, int h, int first_short, const void *a, ptrdiff_t sa, const uint16_t *b,
                 ptrdiff_t sb) {
  int tw, th;
  if (w == 2 || h == 2) { tw = 2; th = 2; }
  else if (w == 4 && h == 4) { tw = 4; th = 4; }
  else if (h == 4 && w > h) { tw = 8; th = 4; }
  else if (w == 4 && h > w) { tw = 4; th = 8; }
  else if (w > h) { tw = 16; th = 8; }
  else if (w < h) { tw = 8; th = 16; }
  else { tw = 8; th = 8; }
  uint64_t sum = 0;
  const size_t esz = 2;
  for (int y = 0; y < h; y += th)
    for (int x = 0; x < w; x += tw)
      sum += (uint64_t)satd_tile(tw, th, first_short, (const char *)a + (y * sa + x) * esz, sa, b + y * sb + x, sb);
  return sum >> (bitdepth - 8);
}

/* SampleMetric::Compare, sample_metric.cc:171-277 (metric dispatch and bit-depth scaling). */
uint64_t xo_compare(int metric, int bitdepth, int w, int h, int first_short, const void *a, ptrdiff_t sa,
                    const uint16_t *b, ptrdiff_t sb) {
  switch (metric) {
    case XVCB200_METRIC_SSD:
      return xo_ssd(first_short ? 1 : 0, w, h, a, sa, b, sb) >> (2 * (bitdepth - 8));
    case XVCB200_METRIC_SATD:
      return xo_satd(bitdepth, w, h, first_short, a, sa, b, sb);
    case XVCB200_METRIC_SAD:
      return (uint64_t)xo_sad(first_short ? 1 : 0, w, h, a, sa, b, sb) >> (bitdepth - 8);
    case XVCB200_METRIC_SAD_FAST: {   /* every second row, doubled: :194-199 */
      uint64_t d = (uint64_t)xo_sad(first_short ? 1 : 0, w, h / 2, a, sa * 2, b, sb * 2);
      return (d * 2) >> (bitdepth - 8);
    }
  }
  return UINT64_MAX;
}

/* `static_cast<Distortion>(dist * weight)`, sample_metric.cc:221-222 */
uint64_t xo_apply_weight(uint64_t dist, double weight) { return (uint64_t)((double)dist * weight); }

/* ===================================================================================
 * Interpolation
 * =================================================================================== */

/* kLumaFilterHighPrec, inter_prediction.cc:55-73 */
static const int16_t k_luma_taps[16][8] = {
    {0, 0, 0, 64, 0, 0, 0, 0},       {0, 1, -3, 63, 4, -2, 1, 0},     {-1, 2, -5, 62, 8, -3, 1, 0},
    {-1, 3, -8, 60, 13, -4, 1, 0},   {-1, 4, -10, 58, 17, -5, 1, 0},  {-1, 4, -11, 52, 26, -8, 3, -1},
    {-1, 3, -9, 47, 31, -10, 4, -1}, {-1, 4, -11, 45, 34, -10, 4, -1}, {-1, 4, -11, 40, 40, -11, 4, -1},
    {-1, 4, -10, 34, 45, -11, 4, -1}, {-1, 4, -10, 31, 47, -9, 3, -1}, {-1, 3, -8, 26, 52, -11, 4, -1},
    {0, 1, -5, 17, 58, -10, 4, -1},  {0, 1, -4, 13, 60, -8, 3, -1},   {0, 1, -3, 8, 62, -5, 2, -1},
    {0, 1, -2, 4, 63, -3, 1, 0}};
/* kChromaFilterHighPrec, inter_prediction.cc:91-126 */
static const int16_t k_chroma_taps[32][4] = {
    {0, 64, 0, 0},   {-1, 63, 2, 0},  {-2, 62, 4, 0},  {-2, 60, 7, -1}, {-2, 58, 10, -2}, {-3, 57, 12, -2},
    {-4, 56, 14, -2}, {-4, 55, 15, -2}, {-4, 54, 16, -2}, {-5, 53, 18, -2}, {-6, 52, 20, -2}, {-6, 49, 24, -3},
    {-6, 46, 28, -4}, {-5, 44, 29, -4}, {-4, 42, 30, -4}, {-4, 39, 33, -4}, {-4, 36, 36, -4}, {-4, 33, 39, -4},
    {-4, 30, 42, -4}, {-4, 29, 44, -5}, {-4, 28, 46, -6}, {-3, 24, 49, -6}, {-2, 20, 52, -6}, {-2, 18, 53, -5},
    {-2, 16, 54, -4}, {-2, 15, 55, -4}, {-2, 14, 56, -4}, {-2, 12, 57, -3}, {-2, 10, 58, -2}, {-1, 7, 60, -2},
    {0, 4, 62, -2},  {0, 2, 63, -1}};
const int16_t *xo_luma_taps(int frac) { return k_luma_taps[frac]; }
const int16_t *xo_chroma_taps(int frac) { return k_chroma_taps[frac]; }

enum { K_INTERNAL_PREC = 14, K_FILTER_PREC = 6, K_INTERNAL_OFFSET = 8192 }; /* inter_prediction.h:58-61 */

/* The six filter kernels, inter_prediction.cc:1207-1385, with the shift/offset rules of
 * inter_prediction.h:218-255.
 * kind: 0 H u16->u16, 1 H u16->i16, 2 V u16->u16, 3 V u16->i16, 4 V i16->u16, 5 V i16->i16. */
void xo_filter(int kind, int chroma, int w, int h, int bitdepth, const int16_t *taps, const void *src, ptrdiff_t ss,
               void *dst, ptrdiff_t ds) {
  const int n = chroma ? 4 : 8;
  const int head = K_INTERNAL_PREC - bitdepth;
  const int src_short = kind >= 4;
  const int dst_sample = kind == 0 || kind == 2 || kind == 4;
  const ptrdiff_t step = (kind <= 1) ? 1 : ss;
  int shift, offset;
  if (!src_short && dst_sample) { shift = K_FILTER_PREC; offset = 1 << (shift - 1); }
  else if (!src_short) { shift = K_FILTER_PREC - head; offset = -(K_INTERNAL_OFFSET << shift); }
  else if (dst_sample) { shift = K_FILTER_PREC + head; offset = (K_INTERNAL_OFFSET << K_FILTER_PREC) + (1 << (shift - 1)); }
  else { shift = K_FILTER_PREC; offset = 0; }
  const int maxv = (1 << bitdepth) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int sum = 0;
      for (int k = 0; k < n; k++) {
        ptrdiff_t idx = y * ss + x + (k - (n / 2 - 1)) * step;
        int s = src_short ? ((const int16_t *)src)[idx] : ((const uint16_t *)src)[idx];
        sum += s * taps[k];
      }
      int val = (sum + offset) >> shift;
      if (dst_sample) {
        /* the vertical variants narrow to int16 before clipping (:1290, :1350); the
         * horizontal u16->u16 variant clips the int directly (:1228-1229) */
        if (kind != 0) val = (int16_t)val;
        ((uint16_t *)dst)[y * ds + x] = (uint16_t)clip3(val, 0, maxv);
      } else {
        ((int16_t *)dst)[y * ds + x] = (int16_t)val;
      }
    }
}

/* SampleBuffer::AddAvg, sample_buffer.h:89-106 */
void xo_add_avg(int w, int h, int offset, int shift, int bitdepth, const int16_t *a, ptrdiff_t sa, const int16_t *b,
                ptrdiff_t sb, uint16_t *dst, ptrdiff_t ds) {
  const int maxv = (1 << bitdepth) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      dst[y * ds + x] = (uint16_t)clip3((a[y * sa + x] + b[y * sb + x] + offset) >> shift, 0, maxv);
}

/* FilterCopyBipred_c, inter_prediction.cc:1462-1473: both steps narrow to int16 */
void xo_filter_copy_bipred(int w, int h, int offset, int shift, const uint16_t *ref, ptrdiff_t rs, int16_t *pred,
                           ptrdiff_t ps) {
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int16_t val = (int16_t)(ref[y * rs + x] << shift);
      pred[y * ps + x] = (int16_t)(val - (int16_t)offset);
    }
}

/* MotionCompUniPred + FilterLuma/FilterChroma (+Bipred variants),
 * inter_prediction.cc:1138-1172, 1387-1448, 1450-1538.  2-D = H pass over h+taps-1 rows
 * into a scratch with stride = w, then V pass. */
void xo_interp(int chroma, int bipred, int w, int h, int bitdepth, int frac_x, int frac_y, const uint16_t *ref,
               ptrdiff_t rs, void *pred, ptrdiff_t ps) {
  const int n = chroma ? 4 : 8;
  const int16_t *th = chroma ? k_chroma_taps[frac_x] : k_luma_taps[frac_x];
  const int16_t *tv = chroma ? k_chroma_taps[frac_y] : k_luma_taps[frac_y];
  if (frac_x == 0 && frac_y == 0) {
    if (bipred) {
      xo_filter_copy_bipred(w, h, K_INTERNAL_OFFSET, K_INTERNAL_PREC - bitdepth, ref, rs, (int16_t *)pred, ps);
    } else {
      for (int y = 0; y < h; y++) memcpy((uint16_t *)pred + y * ps, ref + y * rs, sizeof(uint16_t) * (size_t)w);
    }
  } else if (frac_y == 0) {
    xo_filter(bipred ? 1 : 0, chroma, w, h, bitdepth, th, ref, rs, pred, ps);
  } else if (frac_x == 0) {
    xo_filter(bipred ? 3 : 2, chroma, w, h, bitdepth, tv, ref, rs, pred, ps);
  } else {
    int16_t tmp[64 * (64 + 7)];
    xo_filter(1, chroma, w, h + n - 1, bitdepth, th, ref - (n / 2 - 1) * rs, rs, tmp, w);
    xo_filter(bipred ? 5 : 4, chroma, w, h, bitdepth, tv, tmp + (n / 2 - 1) * w, w, pred, ps);
  }
}

/* ===================================================================================
 * Transform
 * =================================================================================== */

/* The reference's coefficient tables (transform_data.cc, 8-bit "High" precision) are the
 * orthonormal DCT-2 / DCT-5 / DCT-8 / DST-1 / DST-7 bases scaled by 256*sqrt(N) and rounded
 * to nearest; tests/test_oracle_vs_ref.py::test_transform_matrices checks every entry of
 * every table against the reference build.  Row = basis function k, column = sample j. */
static int16_t *g_mat[6][7];
static const double kPi = 3.14159265358979323846;

const int16_t *xo_transform_matrix(int type, int n) {
  if (type == XVCB200_TX_DEFAULT) type = XVCB200_TX_DCT2;
  const int l = ilog2(n);
  if (type < 1 || type > 5 || l < 1 || l > 6 || (1 << l) != n) return NULL;
  if (type != XVCB200_TX_DCT2 && n < 4) return NULL;
  if (g_mat[type][l]) return g_mat[type][l];
  int16_t *m = (int16_t *)malloc(sizeof(int16_t) * (size_t)n * n);
  for (int k = 0; k < n; k++)
    for (int j = 0; j < n; j++) {
      double v = 0;
      switch (type) {
        case XVCB200_TX_DCT2:
          v = k == 0 ? 1.0 : sqrt(2.0) * cos(kPi * (2 * j + 1) * k / (2.0 * n));
          break;
        case XVCB200_TX_DCT5:
          v = sqrt(4.0 / (2 * n - 1)) * cos(2 * kPi * k * j / (2.0 * n - 1)) * sqrt((double)n);
          if (k == 0) v *= 1 / sqrt(2.0);
          if (j == 0) v *= 1 / sqrt(2.0);
          break;
        case XVCB200_TX_DCT8:
          v = sqrt(4.0 / (2 * n + 1)) * cos(kPi * (2 * k + 1) * (2 * j + 1) / (4.0 * n + 2)) * sqrt((double)n);
          break;
        case XVCB200_TX_DST1:
          v = sqrt(2.0 / (n + 1)) * sin(kPi * (k + 1) * (j + 1) / (n + 1.0)) * sqrt((double)n);
          break;
        case XVCB200_TX_DST7:
          v = sqrt(4.0 / (2 * n + 1)) * sin(kPi * (2 * k + 1) * (j + 1) / (2.0 * n + 1)) * sqrt((double)n);
          break;
      }
      m[k * n + j] = (int16_t)lrint(v * 256.0);
    }
  g_mat[type][l] = m;
  return m;
}

/* One forward stage: FwdDct2TransformN / FwdGenericTransformN (transform.cc:1186-1612) in
 * matrix form.  in: `lines` rows of `n` samples; out[k*os + line], k < min(n,32); lines at or
 * beyond the zero-out limit and rows k >= 32 are written as zero (:1568-1577, 1603-1611).
 * Stores narrow to int16 without clipping (:1202). */
static void fwd_stage(const int16_t *m, int n, int shift, int lines, int zero_out, const int16_t *in, ptrdiff_t is,
                      int16_t *out, ptrdiff_t os) {
  const int add = 1 << (shift - 1);
  const int tx_lines = zero_out && lines > 32 ? 32 : lines;
  const int out_rows = n > 32 ? 32 : n;
  for (int k = 0; k < n; k++)
    for (int y = 0; y < lines; y++) {
      int v = 0;
      if (k < out_rows && y < tx_lines) {
        int sum = 0;
        for (int j = 0; j < n; j++) sum += m[k * n + j] * in[y * is + j];
        v = (sum + add) >> shift;
      }
      out[k * os + y] = (int16_t)v;
    }
}

/* FwdPartialDst4 / InvPartialDst4, transform.cc:997-1017, 217-242 (4x4 intra luma DST),
 * as the 4x4 matrix those butterflies factor. */
static const int k_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

static void fwd_dst4(int shift, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  const int add = 1 << (shift - 1);
  for (int y = 0; y < 4; y++)
    for (int k = 0; k < 4; k++) {
      int sum = 0;
      for (int j = 0; j < 4; j++) sum += k_dst4[k][j] * in[y * is + j];
      out[k * os + y] = (int16_t)((sum + add) >> shift);
    }
}
static void inv_dst4(int shift, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  const int add = 1 << (shift - 1);
  for (int y = 0; y < 4; y++)
    for (int j = 0; j < 4; j++) {
      int sum = 0;
      for (int k = 0; k < 4; k++) sum += k_dst4[k][j] * in[k * is + y];
      out[y * os + j] = (int16_t)clip3((sum + add) >> shift, -32768, 32767);
    }
}

/* ForwardTransform::Transform, transform.cc:869-961 (unrestricted: high precision always).
 * Rows first (size = w, type tx_hor) transposed into a 64-stride temp, then columns. */
void xo_fwd_transform(int w, int h, int bitdepth, int tx_hor, int tx_ver, int dst4x4, const int16_t *resi, ptrdiff_t rs,
                      int16_t *coeff, ptrdiff_t cs) {
  int16_t tmp[64 * 64];
  const int shift1 = ilog2(w) + bitdepth - 9 + 2;
  const int shift2 = ilog2(h) + 6 + 2;
  if (dst4x4 && w == 4 && h == 4) {  /* DST has no high-precision variant: shift -= 2 (:1001) */
    fwd_dst4(shift1 - 2, resi, rs, tmp, 64);
    fwd_dst4(shift2 - 2, tmp, 64, coeff, cs);
    return;
  }
  fwd_stage(xo_transform_matrix(tx_hor, w), w, shift1, h, 0, resi, rs, tmp, 64);
  fwd_stage(xo_transform_matrix(tx_ver, h), h, shift2, w, 1, tmp, 64, coeff, cs);
}

/* One inverse stage: InvDct2TransformN / InvGenericTransformN (transform.cc:425-862).
 * in[k*is + line], only the first min(n,32) input rows are read (:699, 725, 843);
 * out[line*os + j] clipped to int16; lines beyond the zero-out limit are zero. */
static void inv_stage(const int16_t *m, int n, int shift, int lines, int zero_out, const int16_t *in, ptrdiff_t is,
                      int16_t *out, ptrdiff_t os) {
  const int add = 1 << (shift - 1);
  const int tx_lines = zero_out && lines > 32 ? 32 : lines;
  const int in_rows = n > 32 ? 32 : n;
  for (int y = 0; y < lines; y++)
    for (int j = 0; j < n; j++) {
      int v = 0;
      if (y < tx_lines) {
        int sum = 0;
        for (int k = 0; k < in_rows; k++) sum += m[k * n + j] * in[k * is + y];
        v = clip3((sum + add) >> shift, -32768, 32767);
      }
      out[y * os + j] = (int16_t)v;
    }
}

/* InverseTransform::Transform, transform.cc:83-182; DC shortcut InvDct2Dc :279-291. */
void xo_inv_transform(int w, int h, int bitdepth, int tx_hor, int tx_ver, int dst4x4, int dc_only, const int16_t *coeff,
                      ptrdiff_t cs, int16_t *resi, ptrdiff_t rs) {
  int16_t tmp[64 * 64];
  const int shift1 = 7 + 2;
  const int shift2 = 20 - bitdepth + 2;
  if (dst4x4 && w == 4 && h == 4) {
    inv_dst4(shift1 - 2, coeff, cs, tmp, 64);
    inv_dst4(shift2 - 2, tmp, 64, resi, rs);
    return;
  }
  if (dc_only && tx_hor <= XVCB200_TX_DCT2 && tx_ver <= XVCB200_TX_DCT2) {
    const int shift = 14 - bitdepth;
    const int16_t c = (int16_t)((((coeff[0] + 1) >> 1) + (1 << (shift - 1))) >> shift);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) resi[y * rs + x] = c;
    return;
  }
  inv_stage(xo_transform_matrix(tx_ver, h), h, shift1, w, 1, coeff, cs, tmp, 64);
  inv_stage(xo_transform_matrix(tx_hor, w), w, shift2, h, 0, tmp, 64, resi, rs);
}

static int transform_shift(int w, int h, int bitdepth) {   /* Quantize::GetTransformShift, quantize.cc:127-131 */
  return 15 - bitdepth - ((ilog2(w) + ilog2(h)) >> 1);
}

/* ForwardTransform::TransformSkip (transform.cc:963-995) / InverseTransform::TransformSkip
 * (:184-215). */
void xo_transform_skip(int forward, int w, int h, int bitdepth, const int16_t *in, ptrdiff_t is, int16_t *out,
                       ptrdiff_t os) {
  const int odd = (ilog2(w) + ilog2(h)) & 1;
  const int scale = odd ? 181 : 1;
  const int ts = transform_shift(w, h, bitdepth);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int v = in[y * is + x] * scale;
      if (forward) {
        int shift = ts + (odd ? -8 : 0);
        v = shift > 0 ? v * (1 << shift) : (v + (1 << (-shift - 1))) >> -shift;
      } else {
        int shift = ts + (odd ? 7 : 0);
        v = shift > 0 ? (v + (1 << (shift - 1))) >> shift : v << -shift;
      }
      out[y * os + x] = (int16_t)v;
    }
}

/* ===================================================================================
 * Quantisation
 * =================================================================================== */

static const uint8_t k_chroma_scale[58] = {  /* Qp::kChromaScale_, quantize.cc:34-38 */
    0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28,
    29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};
static const int k_fwd_scale[6] = {26214, 23302, 20560, 18396, 16384, 14564}; /* quantize.cc:40-42 */
static const int k_inv_scale[6] = {40, 45, 51, 57, 64, 72};                   /* quantize.cc:44-46 */

/* Qp::Qp, ScaleChromaQp, GetChromaDistWeight: quantize.cc:48-92 */
void xo_qp_init(xvcb200_qp *out, int qp, int chroma_format, int bitdepth, double lambda, int table, int off_u,
                int off_v) {
  const int offs[3] = {0, off_u, off_v};
  out->lambda_sqrt = sqrt(lambda);
  for (int c = 0; c < 3; c++) {
    int raw = qp;
    double weight = 1.0;
    if (c > 0) {
      int base = clip3(qp, 0, 57);
      int with_off = clip3(qp + offs[c], 0, 57);
      raw = with_off;
      int delta = with_off - base;
      if (chroma_format == 1 && table == 1) {
        raw = k_chroma_scale[with_off];
        delta = k_chroma_scale[with_off] - base;
      }
      weight = pow(2.0, -delta / 3.0);
    }
    out->qp_raw[c] = raw;
    int qbd = raw + 6 * (bitdepth - 8);
    out->qp_bitdepth[c] = qbd < 0 ? 0 : qbd;
    out->distortion_weight[c] = weight;
    out->lambda[c] = c == 0 ? lambda : lambda / weight;
  }
}

static const uint8_t k_scan4x4[3][16] = {  /* TransformHelper::kScanCoeff4x4, transform.cc:70-76 */
    {0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
    {0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15}};

/* TransformHelper::DeriveSubblockScan, transform.cc:1638-1680 */
static void subblock_scan(int order, int bw, int bh, uint16_t *table) {
  int x = 0, y = 0;
  for (int i = 0; i < bw * bh; i++) {
    table[i] = (uint16_t)(y * bw + x);
    if (order == 0) {          /* up-right diagonal */
      if (x == bw - 1 || y == 0) {
        y += x + 1;
        x = 0;
        if (y >= bh) { x += y - (bh - 1); y = bh - 1; }
      } else { x++; y--; }
    } else if (order == 1) {   /* horizontal */
      if (x == bw - 1) { x = 0; y++; } else x++;
    } else {                   /* vertical */
      if (y == bh - 1) { x++; y = 0; } else y++;
    }
  }
}

/* RdoQuant::CoeffSignHideFast, rdo_quant.cc:448-573.  4x4 sub-blocks visited from the last
 * in scan order to the first; the first one that holds a non-zero level is the "last"
 * sub-block and starts its search at its last non-zero position. */
static int sign_hide_fast(int w, int h, int scan_order, const int16_t *in, ptrdiff_t is, const int16_t *delta,
                          ptrdiff_t dstride, int16_t *out, ptrdiff_t os) {
  const int bw = w >> 2, bh = h >> 2;
  uint16_t sb_scan[256];
  subblock_scan(scan_order, bw, bh, sb_scan);
  const uint8_t *scan = k_scan4x4[scan_order];
  int nnz = 0, seen_last = 0;
  for (int sbi = bw * bh - 1; sbi >= 0; sbi--) {
    const int sx = (sb_scan[sbi] % bw) << 2, sy = (sb_scan[sbi] / bw) << 2;
#define AT(buf, stride, idx) (buf)[(sy + (scan[idx] >> 2)) * (stride) + sx + (scan[idx] & 3)]
    int last = -1, first = 16, sum = 0;
    for (int i = 0; i < 16; i++) {
      int16_t c = AT(out, os, i);
      if (c) { if (i < first) first = i; if (i > last) last = i; sum += c; nnz++; }
    }
    const int is_last_sb = (last >= 0 && !seen_last);
    if (is_last_sb) seen_last = 1;
    if (last - first > 3) {
      const int sign = AT(out, os, first) > 0 ? 0 : 1;
      if (sign != (sum & 1)) {
        int16_t cur_cost = INT16_MAX, cur_change = 0, min_cost = INT16_MAX, min_change = 0;
        int min_index = -1;
        for (int i = is_last_sb ? last : 15; i >= 0; i--) {
          const int16_t lev = AT(out, os, i), dl = AT(delta, dstride, i);
          if (lev != 0) {
            if (dl > 0) { cur_cost = (int16_t)-dl; cur_change = 1; }
            else if (i == first && iabs(lev) == 1) { cur_cost = INT16_MAX; }
            else { cur_cost = dl; cur_change = -1; }
          } else if (i < first && (AT(in, is, i) >= 0 ? 0 : 1) != sign) {
            cur_cost = INT16_MAX;
          } else {
            cur_cost = (int16_t)-dl; cur_change = 1;
          }
          if (cur_cost < min_cost) { min_cost = cur_cost; min_change = cur_change; min_index = i; }
        }
        if (min_index >= 0) {   /* always true here: the last non-zero level has a finite cost */
          int16_t *p = &AT(out, os, min_index);
          if (*p == INT16_MIN || *p == INT16_MAX) min_change = -1;
          if (*p == 0) nnz++;
          *p = (int16_t)(*p + (AT(in, is, min_index) >= 0 ? min_change : -min_change));
          if (*p == 0) nnz--;
        }
      }
    }
#undef AT
  }
  return nnz;
}

/* RdoQuant::QuantFast, rdo_quant.cc:156-201 */
int xo_quant_fast(int w, int h, int bitdepth, int qp_bitdepth, int intra_picture, int sign_hiding, int scan_order,
                  const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  const int odd = (ilog2(w) + ilog2(h)) & 1;
  const int shift = 14 + qp_bitdepth / 6 + transform_shift(w, h, bitdepth) + (odd ? 7 : 0);
  const int scale = k_fwd_scale[qp_bitdepth % 6] * (odd ? 181 : 1);
  const int64_t offset = (int64_t)(intra_picture ? 171 : 85) << (shift - 9);
  int16_t delta[64 * 64];
  int nnz = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int c = in[y * is + x];
      const int64_t mag = (int64_t)iabs(c) * scale;
      const int level = (int)((mag + offset) >> shift);
      nnz += level != 0;
      out[y * os + x] = (int16_t)clip3(c < 0 ? -level : level, -32768, 32767);
      delta[y * 64 + x] = (int16_t)((mag - ((int64_t)level << shift)) >> (shift - 8));
    }
  if (sign_hiding && nnz > 1 && w >= 4 && h >= 4) nnz = sign_hide_fast(w, h, scan_order, in, is, delta, 64, out, os);
  return nnz;
}

/* Quantize::Inverse, quantize.cc:94-125 */
void xo_dequant(int w, int h, int bitdepth, int qp_bitdepth, const int16_t *in, ptrdiff_t is, int16_t *out,
                ptrdiff_t os) {
  const int odd = (ilog2(w) + ilog2(h)) & 1;
  const int shift = 6 - transform_shift(w, h, bitdepth) + (odd ? 8 : 0);
  const int scale = (k_inv_scale[qp_bitdepth % 6] << (qp_bitdepth / 6)) * (odd ? 181 : 1);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int v = in[y * is + x] * scale;
      v = shift > 0 ? (v + (1 << (shift - 1))) >> shift : (int)((unsigned)v << -shift);
      out[y * os + x] = (int16_t)clip3(v, -32768, 32767);
    }
}

/* ===================================================================================
 * Picture-level helpers
 * =================================================================================== */

/* YuvPicture::PadBorder, yuv_pic.cc:118-150 */
void xo_pad_border(xo_picture *pic) {
  for (int c = 0; c < 3; c++) {
    const int w = pic->width[c], h = pic->height[c], s = pic->stride[c], p = pic->pad[c];
    uint16_t *b = pic->base[c];
    for (int y = 0; y < h; y++)
      for (int x = 1; x <= p; x++) {
        b[y * s - x] = b[y * s];
        b[y * s + w - 1 + x] = b[y * s + w - 1];
      }
    for (int y = 1; y <= p; y++) {
      memcpy(b - y * s - p, b - p, sizeof(uint16_t) * (size_t)(w + 2 * p));
      memcpy(b + (h - 1 + y) * s - p, b + (h - 1) * s - p, sizeof(uint16_t) * (size_t)(w + 2 * p));
    }
  }
}

/* InterPrediction::ClipMv, inter_prediction.cc:769-782 (1/16-pel units) */
void xo_clip_mv(int pos_x, int pos_y, int pic_w, int pic_h, int32_t mv[2]) {
  mv[0] = clip3(mv[0], -((64 + 8 + pos_x - 1) << 4), (pic_w + 8 - pos_x - 1) << 4);
  mv[1] = clip3(mv[1], -((64 + 8 + pos_y - 1) << 4), (pic_h + 8 - pos_y - 1) << 4);
}

/* InterPrediction::DetermineMinMaxMv, inter_prediction.cc:801-817; results full-pel */
void xo_min_max_mv(int pos_x, int pos_y, int pic_w, int pic_h, const int32_t center[2], int range, int32_t mv_min[2],
                   int32_t mv_max[2]) {
  int32_t c[2] = {center[0], center[1]};
  xo_clip_mv(pos_x, pos_y, pic_w, pic_h, c);
  int32_t lo[2] = {c[0] - (range << 4), c[1] - (range << 4)};
  int32_t hi[2] = {c[0] + (range << 4), c[1] + (range << 4)};
  xo_clip_mv(pos_x, pos_y, pic_w, pic_h, lo);
  xo_clip_mv(pos_x, pos_y, pic_w, pic_h, hi);
  for (int i = 0; i < 2; i++) { mv_min[i] = lo[i] >> 4; mv_max[i] = hi[i] >> 4; }
}

/* InterSearch::GetNumExpGolombBits, inter_search.cc:1176-1185 */
uint32_t xo_exp_golomb_bits(int v) {
  uint32_t len = 1, u = v <= 0 ? ((uint32_t)(-v) << 1) + 1 : (uint32_t)v << 1;
  while (u != 1) { u >>= 1; len += 2; }
  return len;
}

/* GetMvdBitsFullpel, inter_search.cc:1162-1174 */
static uint32_t mvd_bits_fullpel(const int32_t mvp[2], int x, int y, int down) {
  const int sh = down + 2;
  return xo_exp_golomb_bits((x * 16 - mvp[0]) >> sh) + xo_exp_golomb_bits((y * 16 - mvp[1]) >> sh);
}
/* GetMvdBits, inter_search.cc:1144-1154 */
static uint32_t mvd_bits(const int32_t mvp[2], const int32_t mv[2]) {
  return xo_exp_golomb_bits((mv[0] - mvp[0]) >> 2) + xo_exp_golomb_bits((mv[1] - mvp[1]) >> 2);
}

/* ===================================================================================
 * Integer-pel TZ search
 * =================================================================================== */

typedef struct {
  const uint16_t *org; ptrdiff_t org_stride;
  const uint16_t *ref; ptrdiff_t ref_stride;     /* co-located block in the reference */
  int w, h, bitdepth, metric, down;
  const int32_t *mvp;
  int32_t lo[2], hi[2];
  int32_t best[2];
  uint64_t best_cost;
  int last_pos, last_range;
  uint32_t lambda;
  int evals;
} tz_state;

/* TzSearch::CheckCostBest, inter_tz_search.cc:261-276.  (The reference skips the bit count
 * when dist >= cost_best; cost >= dist makes that a pure shortcut.) */
static int tz_try(tz_state *s, int x, int y) {
  const uint64_t dist = xo_compare(s->metric, s->bitdepth, s->w, s->h, 0, s->org, s->org_stride,
                                   s->ref + y * s->ref_stride + x, s->ref_stride);
  s->evals++;
  if (dist >= s->best_cost) return 0;
  const uint32_t bits = mvd_bits_fullpel(s->mvp, x, y, s->down);
  const uint64_t cost = dist + ((s->lambda * bits) >> 16);   /* uint32 product, as in the reference */
  if (cost >= s->best_cost) return 0;
  s->best_cost = cost; s->best[0] = x; s->best[1] = y;
  return 1;
}

/* One pattern point: direction bits select which window bounds are tested
 * (IsInside<Dir>, inter_tz_search.cc:278-301): up -> y >= min, down -> y <= max,
 * left -> x >= min, right -> x <= max.  `pos` is Dir::index (sum for diagonals):
 * left -1, right +1, up -3, down +3 (inter_tz_search.h). */
static int tz_point(tz_state *s, int x, int y, int pos, int range) {
  const int vert = pos <= -2 ? -1 : (pos >= 2 ? 1 : 0);
  const int horz = pos - 3 * vert;
  if (vert < 0 && y < s->lo[1]) return 0;
  if (vert > 0 && y > s->hi[1]) return 0;
  if (horz < 0 && x < s->lo[0]) return 0;
  if (horz > 0 && x > s->hi[0]) return 0;
  if (!tz_try(s, x, y)) return 0;
  s->last_pos = pos; s->last_range = range;
  return 1;
}

/* FullpelDiamondSearch, inter_tz_search.cc:173-210 */
static int tz_diamond(tz_state *s, int bx, int by, int r) {
  int mod = 0;
  if (r == 1) {
    mod |= tz_point(s, bx, by - 1, -3, 1);
    mod |= tz_point(s, bx - 1, by, -1, 1);
    mod |= tz_point(s, bx + 1, by, 1, 1);
    mod |= tz_point(s, bx, by + 1, 3, 1);
  } else if (r <= 8) {
    const int q = r >> 1;     /* corner points report range r/2 */
    mod |= tz_point(s, bx, by - r, -3, r);
    mod |= tz_point(s, bx - q, by - q, -4, q);
    mod |= tz_point(s, bx + q, by - q, -2, q);
    mod |= tz_point(s, bx - r, by, -1, r);
    mod |= tz_point(s, bx + r, by, 1, r);
    mod |= tz_point(s, bx - q, by + q, 2, q);
    mod |= tz_point(s, bx + q, by + q, 4, q);
    mod |= tz_point(s, bx, by + r, 3, r);
  } else {
    mod |= tz_point(s, bx, by - r, -3, r);
    mod |= tz_point(s, bx - r, by, -1, r);
    mod |= tz_point(s, bx + r, by, 1, r);
    mod |= tz_point(s, bx, by + r, 3, r);
    for (int i = 1; i < 4; i++) {
      const int a = i * (r >> 2), b = r - a;
      mod |= tz_point(s, bx - a, by - b, -4, r);
      mod |= tz_point(s, bx + a, by - b, -2, r);
      mod |= tz_point(s, bx - a, by + b, 2, r);
      mod |= tz_point(s, bx + a, by + b, 4, r);
    }
  }
  return mod;
}

/* FullpelNeighborPointSearch, inter_tz_search.cc:212-259: the two points that complete the
 * square next to the best point, chosen by where it lies relative to its base. */
static void tz_two_point(tz_state *s) {
  const int bx = s->best[0], by = s->best[1];
  switch (s->last_pos) {
    case -4: tz_point(s, bx - 1, by, -1, 1); tz_point(s, bx, by - 1, -3, 1); break;
    case -3: tz_point(s, bx - 1, by - 1, -4, 1); tz_point(s, bx + 1, by - 1, -2, 1); break;
    case -2: tz_point(s, bx, by - 1, -3, 1); tz_point(s, bx + 1, by, 1, 1); break;
    case -1: tz_point(s, bx - 1, by + 1, 2, 1); tz_point(s, bx - 1, by - 1, -4, 1); break;
    case 1: tz_point(s, bx + 1, by - 1, -2, 1); tz_point(s, bx + 1, by + 1, 4, 1); break;
    case 2: tz_point(s, bx - 1, by, -1, 1); tz_point(s, bx, by + 1, 3, 1); break;
    case 3: tz_point(s, bx - 1, by + 1, 2, 1); tz_point(s, bx + 1, by + 1, 4, 1); break;
    case 4: tz_point(s, bx + 1, by, 1, 1); tz_point(s, bx, by + 1, 3, 1); break;
    default: break;
  }
}

/* TzSearch::Search, inter_tz_search.cc:84-171, with the window set-up of
 * InterSearch::MotionEstNormal (inter_search.cc:618-625) and the metric choice of
 * GetFullpelMetric (:1059-1069: kSadFast when height > 8). */
int xo_tz_search(const xo_picture *orig, const xo_picture *ref, int bitdepth, const xvcb200_cu *cu,
                 const xvcb200_me_job *job, uint32_t lambda_me, int32_t mv_out[2], uint32_t *cost_out) {
  const int pw = ref->width[0], ph = ref->height[0];
  const int range = job->search_range;
  tz_state s;
  memset(&s, 0, sizeof(s));
  s.org = orig->base[0] + cu->y * orig->stride[0] + cu->x; s.org_stride = orig->stride[0];
  s.ref = ref->base[0] + cu->y * ref->stride[0] + cu->x;   s.ref_stride = ref->stride[0];
  s.w = cu->w; s.h = cu->h; s.bitdepth = bitdepth;
  s.metric = cu->h > 8 ? XVCB200_METRIC_SAD_FAST : XVCB200_METRIC_SAD;
  s.down = (cu->flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  s.mvp = job->mvp; s.lambda = lambda_me; s.best_cost = UINT64_MAX;
  xo_min_max_mv(cu->x, cu->y, pw, ph, job->mvp, range, s.lo, s.hi);
  int32_t scan_lo[2] = {s.lo[0], s.lo[1]}, scan_hi[2] = {s.hi[0], s.hi[1]};

  int32_t p[2] = {job->mvp[0], job->mvp[1]};
  xo_clip_mv(cu->x, cu->y, pw, ph, p);
  tz_try(&s, p[0] >> 4, p[1] >> 4);
  int moved = 0;
  if (s.best[0] != 0 || s.best[1] != 0) moved = tz_try(&s, 0, 0);
  s.last_range = 0;
  if (cu->depth != 0) {     /* eval_prev_mv_search_result defaults to 1 (encoder_settings.h:81) */
    int32_t q[2] = {job->prev[0] * 16, job->prev[1] * 16};
    xo_clip_mv(cu->x, cu->y, pw, ph, q);
    moved |= tz_try(&s, q[0] >> 4, q[1] >> 4);
    if (moved) {
      const int32_t c[2] = {s.best[0] * 16, s.best[1] * 16};
      xo_min_max_mv(cu->x, cu->y, pw, ph, c, range, scan_lo, scan_hi);
    }
  }

  const int bx = s.best[0], by = s.best[1];
  int misses = 0;
  for (int r = 1; r <= range; r *= 2) {
    if (tz_diamond(&s, bx, by, r)) misses = 0;
    else if (++misses >= 3) break;
  }
  if (s.last_range == 1) { s.last_range = 0; tz_two_point(&s); }

  if (s.last_range > 5) {   /* raster scan of the window on a 5-sample grid */
    s.last_range = 5;
    for (int y = scan_lo[1]; y <= scan_hi[1]; y += 5)
      for (int x = scan_lo[0]; x <= scan_hi[0]; x += 5) tz_try(&s, x, y);
  }

  while (s.last_range > 0) {  /* re-centre until the centre wins */
    const int cx = s.best[0], cy = s.best[1];
    s.last_range = 0;
    for (int r = 1; r <= range; r *= 2) tz_diamond(&s, cx, cy, r);
    if (s.last_range == 1) { s.last_range = 0; tz_two_point(&s); }
  }
  mv_out[0] = s.best[0]; mv_out[1] = s.best[1];
  if (cost_out) *cost_out = (uint32_t)s.best_cost;
  return s.evals;
}

/* ===================================================================================
 * Motion compensation and sub-pel search
 * =================================================================================== */

/* MotionCompensationMv -> ClipMv -> GetFullpelRef -> MotionCompUniPred,
 * inter_prediction.cc:740-758, 1174-1205, 1138-1154 (4:2:0: chroma mv has 1/32 precision). */
static void mc_block(const xo_picture *ref, int comp, int bitdepth, const xvcb200_cu *cu, const int32_t mv_raw[2],
                     int bipred, void *pred, ptrdiff_t ps) {
  int32_t mv[2] = {mv_raw[0], mv_raw[1]};
  xo_clip_mv(cu->x, cu->y, ref->width[0], ref->height[0], mv);
  const int cs = comp ? 1 : 0;
  const int sh = 4 + cs, x = cu->x >> cs, y = cu->y >> cs, w = cu->w >> cs, h = cu->h >> cs;
  const int fx = mv[0] & ((1 << sh) - 1), fy = mv[1] & ((1 << sh) - 1);
  const uint16_t *r = ref->base[comp] + (y + (mv[1] >> sh)) * ref->stride[comp] + x + (mv[0] >> sh);
  xo_interp(comp != 0, bipred, w, h, bitdepth, fx, fy, r, ref->stride[comp], pred, ps);
}

static const int8_t k_half[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
static const int8_t k_qpel[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};

/* InterSearch::SubpelSearch / GetSubpelDist, inter_search.cc:893-964: nine half-pel
 * positions (centre included) then eight quarter-pel positions, SATD + mvd bits. */
void xo_subpel_search(const xo_picture *orig, const xo_picture *ref, int bitdepth, const xvcb200_cu *cu,
                      const int32_t mvp[2], const int32_t mv_fullpel[2], uint32_t lambda_me, int32_t mv_out[2],
                      uint32_t *dist_out, uint32_t *cost_out) {
  uint16_t pred[64 * 64];
  const uint16_t *org = orig->base[0] + cu->y * orig->stride[0] + cu->x;
  uint64_t best_cost = UINT64_MAX, best_dist = UINT64_MAX;
  int32_t best[2] = {mv_fullpel[0] * 16, mv_fullpel[1] * 16};
  if (cu->flags & XVCB200_CU_FULLPEL_MV) {   /* MotionEstNormal, inter_search.cc:650-653 */
    mc_block(ref, 0, bitdepth, cu, best, 0, pred, 64);
    best_dist = xo_satd(bitdepth, cu->w, cu->h, 0, org, orig->stride[0], pred, 64);
    mv_out[0] = best[0]; mv_out[1] = best[1];
    *dist_out = (uint32_t)best_dist; *cost_out = (uint32_t)best_dist;
    return;
  }
  for (int pass = 0; pass < 2; pass++) {
    const int32_t base[2] = {best[0], best[1]};
    const int step = pass == 0 ? 8 : 4;     /* MvDelta(.., prec 1 / 2) in 1/16 units */
    for (int i = pass; i < 9; i++) {
      const int8_t *o = pass == 0 ? k_half[i] : k_qpel[i];
      const int32_t mv[2] = {base[0] + o[0] * step, base[1] + o[1] * step};
      mc_block(ref, 0, bitdepth, cu, mv, 0, pred, 64);
      const uint64_t dist = xo_satd(bitdepth, cu->w, cu->h, 0, org, orig->stride[0], pred, 64);
      if (dist >= best_cost) continue;
      const uint64_t cost = dist + ((lambda_me * mvd_bits(mvp, mv)) >> 16);
      if (cost < best_cost) { best_cost = cost; best_dist = dist; best[0] = mv[0]; best[1] = mv[1]; }
    }
  }
  mv_out[0] = best[0]; mv_out[1] = best[1];
  *dist_out = (uint32_t)best_dist; *cost_out = (uint32_t)best_cost;
}

/* InterSearch::MotionEstNormal, inter_search.cc:606-662, uni-prediction */
void xo_me_search(const xo_picture *orig, const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus,
                  const xvcb200_me_job *jobs, int n, double lambda_sqrt, xvcb200_me_result *results) {
  const uint32_t lambda_me = (uint32_t)floor(65536.0 * lambda_sqrt);
  for (int i = 0; i < n; i++) {
    const xvcb200_me_job *job = &jobs[i];
    const xvcb200_cu *cu = &cus[job->cu];
    const xo_picture *ref = refs[job->list][job->ref_slot];
    xvcb200_me_result *r = &results[i];
    r->num_sad = (uint32_t)xo_tz_search(orig, ref, bitdepth, cu, job, lambda_me, r->mv_fullpel, &r->cost_fullpel);
    xo_subpel_search(orig, ref, bitdepth, cu, job->mvp, r->mv_fullpel, lambda_me, r->mv, &r->dist, &r->cost);
  }
}

/* InterSearch::FullSearch, inter_search.cc:853-891, on the weighted original
 * 2*orig - other_pred (ResidualBuffer::SubtractWeighted, sample_buffer.h:147-161). */
void xo_full_search(const xo_picture *orig, const xo_picture *other_pred, const xo_picture *ref, int bitdepth,
                    const xvcb200_cu *cu, const xvcb200_fullsearch_job *job, uint32_t lambda_me, int32_t mv_out[2],
                    uint32_t *cost_out) {
  int16_t worig[64 * 64];
  for (int y = 0; y < cu->h; y++)
    for (int x = 0; x < cu->w; x++)
      worig[y * 64 + x] = (int16_t)(2 * orig->base[0][(cu->y + y) * orig->stride[0] + cu->x + x] -
                                    other_pred->base[0][(cu->y + y) * other_pred->stride[0] + cu->x + x]);
  int32_t lo[2], hi[2];
  xo_min_max_mv(cu->x, cu->y, ref->width[0], ref->height[0], job->center, job->range, lo, hi);
  const int metric = cu->h > 8 ? XVCB200_METRIC_SAD_FAST : XVCB200_METRIC_SAD;
  const int down = (cu->flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  const uint16_t *r0 = ref->base[0] + cu->y * ref->stride[0] + cu->x;
  uint64_t best = UINT64_MAX;
  mv_out[0] = 0; mv_out[1] = 0;
  for (int y = lo[1]; y <= hi[1]; y++)
    for (int x = lo[0]; x <= hi[0]; x++) {
      const uint64_t dist = xo_compare(metric, bitdepth, cu->w, cu->h, 1, worig, 64, r0 + y * ref->stride[0] + x,
                                       ref->stride[0]);
      if (dist >= best) continue;
      const uint64_t cost = dist + ((lambda_me * mvd_bits_fullpel(job->mvp, x, y, down)) >> 16);
      if (cost < best) { best = cost; mv_out[0] = x; mv_out[1] = y; }
    }
  if (cost_out) *cost_out = (uint32_t)best;
}

/* InterPrediction::MotionCompensation, inter_prediction.cc:710-738 (no LIC, no affine) */
void xo_motion_compensate(const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus, int n,
                          xo_picture *pred) {
  int16_t t0[64 * 64], t1[64 * 64];
  for (int i = 0; i < n; i++) {
    const xvcb200_cu *cu = &cus[i];
    if (cu->flags & XVCB200_CU_INTRA) continue;
    const int l0 = cu->ref_idx[0] >= 0, l1 = cu->ref_idx[1] >= 0;
    for (int c = 0; c < 3; c++) {
      const int cs = c ? 1 : 0, w = cu->w >> cs, h = cu->h >> cs;
      uint16_t *dst = pred->base[c] + (cu->y >> cs) * pred->stride[c] + (cu->x >> cs);
      if (l0 && l1) {
        mc_block(refs[0][cu->ref_idx[0]], c, bitdepth, cu, cu->mv[0], 1, t0, 64);
        mc_block(refs[1][cu->ref_idx[1]], c, bitdepth, cu, cu->mv[1], 1, t1, 64);
        const int head = K_INTERNAL_PREC - bitdepth;
        const int shift = (head > 2 ? head : 2) + 1;   /* AddAvgBi, :1540-1553 */
        xo_add_avg(w, h, (1 << (shift - 1)) + 2 * K_INTERNAL_OFFSET, shift, bitdepth, t0, 64, t1, 64, dst,
                   pred->stride[c]);
      } else {
        const int l = l1 ? 1 : 0;
        mc_block(refs[l][cu->ref_idx[l]], c, bitdepth, cu, cu->mv[l], 0, dst, pred->stride[c]);
      }
    }
  }
}

/* ---- affine motion compensation -------------------------------
 * InterPrediction::MotionCompAffine, inter_prediction.cc:1044-1136: the three control-point
 * MVs (top-left, top-right, bottom-left) are clipped (ClipMv for MotionVector3, :783-799), the
 * block is cut into sub-blocks whose size follows from the MV spread (get_subblock_size,
 * :1071-1086), every sub-block gets the MV of the 4-parameter model at its centre (1/256 of the
 * MV unit, rotation-zoom: ver = (-hor.y, hor.x)), clipped again, and is predicted by the plain
 * MotionCompUniPred. */
static int affine_subblock_size(const int32_t ref[2], const int32_t mv_uni[2], int size, int scale) {
  const int dx = abs(mv_uni[0] - ref[0]), dy = abs(mv_uni[1] - ref[1]);
  const int max_len = dx > dy ? dx : dy;
  if (!max_len) return size;                 /* note: NOT scaled (inter_prediction.cc:1078-1080) */
  int sub = (size >> 2) / max_len;           /* kSizeShift = 6 - kPrecisionShift */
  if (sub < 1) sub = 1;
  while (size % sub) sub--;
  return (sub > 4 ? sub : 4) >> scale;
}

static void mc_affine(const xo_picture *ref, int comp, int bitdepth, const xvcb200_cu *cu, const int32_t mv_raw[3][2],
                      int bipred, void *pred, ptrdiff_t ps) {
  const int cs = comp ? 1 : 0, sh = 4 + cs;
  const int px = cu->x >> cs, py = cu->y >> cs, w = cu->w >> cs, h = cu->h >> cs;
  const int W = ref->width[0], H = ref->height[0];
  int32_t mv[3][2];
  const int min_x = -((64 + 8 + cu->x - 1) * 16), max_x = (W + 8 - cu->x - 1) * 16;
  const int min_y = -((64 + 8 + cu->y - 1) * 16), max_y = (H + 8 - cu->y - 1) * 16;
  for (int i = 0; i < 3; i++) {
    mv[i][0] = mv_raw[i][0] < min_x ? min_x : mv_raw[i][0] > max_x ? max_x : mv_raw[i][0];
    mv[i][1] = mv_raw[i][1] < min_y ? min_y : mv_raw[i][1] > max_y ? max_y : mv_raw[i][1];
  }
  const size_t esz = 2;                       /* both prediction types are 16 bits wide */
  if (mv[0][0] == mv[1][0] && mv[0][1] == mv[1][1]) {      /* :1063-1069: translation, mv[2] ignored */
    const uint16_t *r = ref->base[comp] + (py + (mv[0][1] >> sh)) * ref->stride[comp] + px + (mv[0][0] >> sh);
    xo_interp(comp != 0, bipred, w, h, bitdepth, mv[0][0] & ((1 << sh) - 1), mv[0][1] & ((1 << sh) - 1), r,
              ref->stride[comp], pred, ps);
    return;
  }
  const int sbw = affine_subblock_size(mv[0], mv[1], w, cs), sbh = affine_subblock_size(mv[0], mv[2], h, cs);
  const int dhx = ((mv[1][0] - mv[0][0]) * 256) / w, dhy = ((mv[1][1] - mv[0][1]) * 256) / w;   /* C division: toward zero */
  const int dvx = -dhy, dvy = dhx;
  int hor_x = mv[0][0] * 256, hor_y = mv[0][1] * 256, ver_x = hor_x, ver_y = hor_y;
  for (int sy = 0; sy < h; sy += sbh) {
    for (int sx = 0; sx < w; sx += sbw) {
      int mx = (hor_x + dhx * (sbw >> 1) + dvx * (sbh >> 1)) >> 8;
      int my = (hor_y + dhy * (sbw >> 1) + dvy * (sbh >> 1)) >> 8;
      mx = mx < min_x ? min_x : mx > max_x ? max_x : mx;
      my = my < min_y ? min_y : my > max_y ? max_y : my;
      const uint16_t *r = ref->base[comp] + (py + sy + (my >> sh)) * ref->stride[comp] + px + sx + (mx >> sh);
      xo_interp(comp != 0, bipred, sbw, sbh, bitdepth, mx & ((1 << sh) - 1), my & ((1 << sh) - 1), r, ref->stride[comp],
                (char *)pred + ((size_t)sy * ps + sx) * esz, ps);
      hor_x += dhx * sbw;
      hor_y += dhy * sbw;
    }
    ver_x += dvx * sbh;
    ver_y += dvy * sbh;
    hor_x = ver_x;
    hor_y = ver_y;
  }
}

/* InterPrediction::MotionCompensation (:710-738) for CUs with use_affine: MotionCompRefList takes
 * the GetUseAffine branch (:1021-1023) for each list in use; bi-prediction averages the two 14-bit
 * intermediates with AddAvgBi as for translational CUs. */
void xo_motion_compensate_affine(const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus,
                                 const xvcb200_affine_cu *aff, int n_aff, xo_picture *pred) {
  int16_t t0[64 * 64], t1[64 * 64];
  for (int i = 0; i < n_aff; i++) {
    const xvcb200_cu *cu = &cus[aff[i].cu];
    const int l0 = cu->ref_idx[0] >= 0, l1 = cu->ref_idx[1] >= 0;
    if ((cu->flags & XVCB200_CU_INTRA) || (!l0 && !l1)) continue;
    for (int c = 0; c < 3; c++) {
      const int cs = c ? 1 : 0, w = cu->w >> cs, h = cu->h >> cs;
      uint16_t *dst = pred->base[c] + (cu->y >> cs) * pred->stride[c] + (cu->x >> cs);
      if (l0 && l1) {
        mc_affine(refs[0][cu->ref_idx[0]], c, bitdepth, cu, aff[i].mv[0], 1, t0, 64);
        mc_affine(refs[1][cu->ref_idx[1]], c, bitdepth, cu, aff[i].mv[1], 1, t1, 64);
        const int head = K_INTERNAL_PREC - bitdepth;
        const int shift = (head > 2 ? head : 2) + 1;
        xo_add_avg(w, h, (1 << (shift - 1)) + 2 * K_INTERNAL_OFFSET, shift, bitdepth, t0, 64, t1, 64, dst, pred->stride[c]);
      } else {
        const int l = l1 ? 1 : 0;
        mc_affine(refs[l][cu->ref_idx[l]], c, bitdepth, cu, aff[i].mv[l], 0, dst, pred->stride[c]);
      }
    }
  }
}

/* ---- local illumination compensation ---------------------------
 * InterPrediction::DeriveLicParams, inter_prediction.cc:1578-1673.  `mv` is the CLIPPED 1/16-pel
 * MV of the CU; the neighbour rows are read from the reference picture at the MV rounded to full
 * samples of the component and "clipped" by ClipMv with the NEIGHBOUR CU's position -- ClipMv's
 * bounds are in 1/16 pel while the value is in full samples, as in the reference (:1604, 1618). */
static int size_to_log2(int size) {       /* util::SizeToLog2, utils.cc:29-35 (minimum 1) */
  int l = 1;
  while ((1 << l) < size) l++;
  return l;
}
static int msb_len(unsigned x) { int m = 0; while (x) { m++; x >>= 1; } return m; }

static void lic_params(const xo_picture *ref, const xo_picture *rec, int comp, int bitdepth, const xvcb200_cu *cu,
                       const xvcb200_lic_cu *nb, const int32_t mv[2], int *scale_out, int *offset_out) {
  const int cs = comp ? 1 : 0, sh = 4 + cs;
  const int px = cu->x >> cs, py = cu->y >> cs, w = cu->w >> cs, h = cu->h >> cs;
  const int has_above = nb->above_x >= 0, has_left = nb->left_x >= 0;
  *scale_out = 32; *offset_out = 0;
  if (!has_above && !has_left) return;
  const int fx = (mv[0] + (1 << (sh - 1))) >> sh, fy = (mv[1] + (1 << (sh - 1))) >> sh;
  const int step = (w < h ? w : h) > 8 ? 2 : 1;
  const ptrdiff_t rs = ref->stride[comp], ss = rec->stride[comp];
  const uint16_t *rbase = ref->base[comp] + py * rs + px;
  const uint16_t *sbase = rec->base[comp] + py * ss + px;
  int sum_x = 0, sum_y = 0, sum_xx = 0, sum_xy = 0, nbr = 0;
  if (has_above) {
    int32_t c[2] = {fx, fy};
    xo_clip_mv(nb->above_x, nb->above_y, ref->width[0], ref->height[0], c);
    const uint16_t *r = rbase + c[0] + c[1] * rs - rs, *s = sbase - ss;
    const int dx = step * ((w / h) > 1 ? (w / h) : 1);
    for (int x = 0; x < w; x += dx) {
      sum_x += r[x]; sum_y += s[x]; sum_xx += r[x] * r[x]; sum_xy += r[x] * s[x]; nbr++;
    }
  }
  if (has_left) {
    int32_t c[2] = {fx, fy};
    xo_clip_mv(nb->left_x, nb->left_y, ref->width[0], ref->height[0], c);
    const uint16_t *r = rbase + c[0] + c[1] * rs - 1, *s = sbase - 1;
    const int dy = step * ((h / w) > 1 ? (h / w) : 1);
    for (int y = 0; y < h; y += dy) {
      const int a = r[y * rs], b = s[y * ss];
      sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b; nbr++;
    }
  }
  const int size_shift = size_to_log2(nbr);
  const int base_shift = bitdepth + size_shift - 15 > 0 ? bitdepth + size_shift - 15 : 0;
  const int avg_x = sum_x >> base_shift, avg_y = sum_y >> base_shift;
  const int xx_offset = sum_xx >> 7;
  const int avg_xy = ((sum_xy + xx_offset) >> (2 * base_shift)) << size_shift;
  const int avg_xx = ((sum_xx + xx_offset) >> (2 * base_shift)) << size_shift;
  const int sd_xy = avg_xy - avg_x * avg_y, sd_xx = avg_xx - avg_x * avg_x;
  int shift_xx = msb_len((unsigned)abs(sd_xx)) - 6;
  if (shift_xx < 0) shift_xx = 0;
  const int shift_xy = shift_xx - 12 > 0 ? shift_xx - 12 : 0;
  const int total_shift = 15 - 5 + shift_xx - shift_xy;
  const int sd_xy_s = sd_xy >> shift_xy;
  int sd_xx_s = sd_xx >> shift_xx;
  sd_xx_s = sd_xx_s < 0 ? 0 : sd_xx_s > 63 ? 63 : sd_xx_s;
  if (sd_xx_s == 0) return;
  const int sd_xx_scaled = ((1 << 15) + sd_xx_s / 2) / sd_xx_s;
  int scale = (sd_xy_s * sd_xx_scaled) >> total_shift;
  scale = scale < 0 ? 0 : scale > 128 ? 128 : scale;
  int offset = (sum_y - ((scale * sum_x) >> 5) + (1 << (size_shift - 1))) >> size_shift;
  const int lo = -(1 << (bitdepth - 1)), hi = (1 << (bitdepth - 1)) - 1;
  *scale_out = scale;
  *offset_out = offset < lo ? lo : offset > hi ? hi : offset;
}

/* MotionCompRefList with post_filter (:1024-1040): translational prediction into Samples, then
 * LocalIlluminationComp (:1555-1576) = SampleBuffer::AddLinearModel (sample_buffer.h:108-122). */
static void mc_block_lic(const xo_picture *ref, const xo_picture *rec, int comp, int bitdepth, const xvcb200_cu *cu,
                         const xvcb200_lic_cu *nb, const int32_t mv_raw[2], uint16_t *pred, ptrdiff_t ps) {
  int32_t mv[2] = {mv_raw[0], mv_raw[1]};
  xo_clip_mv(cu->x, cu->y, ref->width[0], ref->height[0], mv);
  mc_block(ref, comp, bitdepth, cu, mv, 0, pred, ps);
  int scale, offset;
  lic_params(ref, rec, comp, bitdepth, cu, nb, mv, &scale, &offset);
  const int cs = comp ? 1 : 0, w = cu->w >> cs, h = cu->h >> cs, maxv = (1 << bitdepth) - 1;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      const int v = ((scale * pred[y * ps + x]) >> 5) + offset;
      pred[y * ps + x] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
    }
}

/* InterPrediction::MotionCompensation (:710-738) for CUs with use_lic */
void xo_motion_compensate_lic(const xo_picture *const refs[2][5], const xo_picture *rec, int bitdepth,
                              const xvcb200_cu *cus, const xvcb200_lic_cu *lic, int n_lic, xo_picture *pred) {
  int16_t t0[64 * 64], t1[64 * 64];
  uint16_t tmp[64 * 64];
  const int head = K_INTERNAL_PREC - bitdepth;
  for (int i = 0; i < n_lic; i++) {
    const xvcb200_cu *cu = &cus[lic[i].cu];
    const int l0 = cu->ref_idx[0] >= 0, l1 = cu->ref_idx[1] >= 0;
    if ((cu->flags & XVCB200_CU_INTRA) || (!l0 && !l1)) continue;
    for (int c = 0; c < 3; c++) {
      const int cs = c ? 1 : 0, w = cu->w >> cs, h = cu->h >> cs;
      uint16_t *dst = pred->base[c] + (cu->y >> cs) * pred->stride[c] + (cu->x >> cs);
      if (l0 && l1) {      /* bi-prediction with intermediate rounding, :724-729 */
        mc_block_lic(refs[0][cu->ref_idx[0]], rec, c, bitdepth, cu, &lic[i], cu->mv[0], tmp, 64);
        xo_filter_copy_bipred(w, h, K_INTERNAL_OFFSET, head, tmp, 64, t0, 64);
        mc_block_lic(refs[1][cu->ref_idx[1]], rec, c, bitdepth, cu, &lic[i], cu->mv[1], tmp, 64);
        xo_filter_copy_bipred(w, h, K_INTERNAL_OFFSET, head, tmp, 64, t1, 64);
        const int shift = (head > 2 ? head : 2) + 1;
        xo_add_avg(w, h, (1 << (shift - 1)) + 2 * K_INTERNAL_OFFSET, shift, bitdepth, t0, 64, t1, 64, dst, pred->stride[c]);
      } else {
        const int l = l1 ? 1 : 0;
        mc_block_lic(refs[l][cu->ref_idx[l]], rec, c, bitdepth, cu, &lic[i], cu->mv[l], dst, pred->stride[c]);
      }
    }
  }
}

/* ===================================================================================
 * Residual coding chain
 * =================================================================================== */

static void cu_qp(const xvcb200_cu *cu, int bitdepth, int table, int off_u, int off_v, xvcb200_qp *q) {
  xo_qp_init(q, cu->qp, 1, bitdepth, 1.0, table, off_u, off_v);
}

/* TransformEncoder::TransformAndReconstruct, transform_encoder.cc:203-285, with
 * RdoQuant::QuantFast as the quantiser, DCT-2 both ways, diagonal scan (inter CU:
 * TransformHelper::DetermineScanOrder, transform.cc:1618-1621), sign hiding on. */
void xo_tq_reconstruct(const xo_picture *orig, const xo_picture *pred, xo_picture *rec, int16_t *const levels[3],
                       int bitdepth, xvcb200_cu *cus, int n, int intra_picture, int table, int off_u, int off_v,
                       xvcb200_tu_result *results) {
  int16_t resi[64 * 64], coef[64 * 64], lev[64 * 64];
  static const int cbf_bit[3] = {XVCB200_CU_CBF_Y, XVCB200_CU_CBF_U, XVCB200_CU_CBF_V};
  const int maxv = (1 << bitdepth) - 1;
  for (int i = 0; i < n; i++) {
    xvcb200_cu *cu = &cus[i];
    xvcb200_qp q;
    cu_qp(cu, bitdepth, table, off_u, off_v, &q);
    for (int c = 0; c < 3; c++) {
      const int cs = c ? 1 : 0, x = cu->x >> cs, y = cu->y >> cs, w = cu->w >> cs, h = cu->h >> cs;
      const uint16_t *o = orig->base[c] + y * orig->stride[c] + x;
      const uint16_t *p = pred->base[c] + y * pred->stride[c] + x;
      uint16_t *r = rec->base[c] + y * rec->stride[c] + x;
      for (int yy = 0; yy < h; yy++)        /* ResidualBuffer::Subtract, sample_buffer.h:130-145 */
        for (int xx = 0; xx < w; xx++) resi[yy * 64 + xx] = (int16_t)(o[yy * orig->stride[c] + xx] - p[yy * pred->stride[c] + xx]);
      xo_fwd_transform(w, h, bitdepth, 0, 0, 0, resi, 64, coef, 64);
      const int nz = xo_quant_fast(w, h, bitdepth, q.qp_bitdepth[c], intra_picture, 1, 0, coef, 64, lev, 64);
      cu->flags = (uint8_t)((cu->flags & ~cbf_bit[c]) | (nz ? cbf_bit[c] : 0));
      int16_t *lp = levels[c] + (size_t)y * orig->width[c] + x;
      for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++) lp[yy * orig->width[c] + xx] = nz ? lev[yy * 64 + xx] : 0;
      if (nz) {
        xo_dequant(w, h, bitdepth, q.qp_bitdepth[c], lev, 64, coef, 64);
        xo_inv_transform(w, h, bitdepth, 0, 0, 0, 0, coef, 64, resi, 64);
        for (int yy = 0; yy < h; yy++)      /* SampleBuffer::AddClip, sample_buffer.h:72-87 */
          for (int xx = 0; xx < w; xx++)
            r[yy * rec->stride[c] + xx] = (uint16_t)clip3(p[yy * pred->stride[c] + xx] + resi[yy * 64 + xx], 0, maxv);
      } else {
        for (int yy = 0; yy < h; yy++) memcpy(r + yy * rec->stride[c], p + yy * pred->stride[c], sizeof(uint16_t) * (size_t)w);
      }
      if (results) {
        results[3 * i + c].num_non_zero = nz;
        results[3 * i + c].ssd = (uint32_t)xo_compare(XVCB200_METRIC_SSD, bitdepth, w, h, 0, o, orig->stride[c], r,
                                                       rec->stride[c]);
      }
    }
  }
}

/* CuDecoder::DecompressComponent, cu_decoder.cc:102-138, from levels and a prediction */
void xo_dequant_reconstruct(const xo_picture *pred, xo_picture *rec, int16_t *const levels[3], int bitdepth,
                            const xvcb200_cu *cus, int n, int table, int off_u, int off_v) {
  int16_t resi[64 * 64], coef[64 * 64];
  static const int cbf_bit[3] = {XVCB200_CU_CBF_Y, XVCB200_CU_CBF_U, XVCB200_CU_CBF_V};
  const int maxv = (1 << bitdepth) - 1;
  for (int i = 0; i < n; i++) {
    const xvcb200_cu *cu = &cus[i];
    xvcb200_qp q;
    cu_qp(cu, bitdepth, table, off_u, off_v, &q);
    for (int c = 0; c < 3; c++) {
      const int cs = c ? 1 : 0, x = cu->x >> cs, y = cu->y >> cs, w = cu->w >> cs, h = cu->h >> cs;
      const uint16_t *p = pred->base[c] + y * pred->stride[c] + x;
      uint16_t *r = rec->base[c] + y * rec->stride[c] + x;
      if (cu->flags & cbf_bit[c]) {
        const int16_t *lp = levels[c] + (size_t)y * rec->width[c] + x;
        xo_dequant(w, h, bitdepth, q.qp_bitdepth[c], lp, rec->width[c], coef, 64);
        xo_inv_transform(w, h, bitdepth, 0, 0, 0, 0, coef, 64, resi, 64);
        for (int yy = 0; yy < h; yy++)
          for (int xx = 0; xx < w; xx++)
            r[yy * rec->stride[c] + xx] = (uint16_t)clip3(p[yy * pred->stride[c] + xx] + resi[yy * 64 + xx], 0, maxv);
      } else {
        for (int yy = 0; yy < h; yy++) memcpy(r + yy * rec->stride[c], p + yy * pred->stride[c], sizeof(uint16_t) * (size_t)w);
      }
    }
  }
}

/* ===================================================================================
 * Deblocking
 * =================================================================================== */

static const uint8_t k_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  0,  0,  0,
                                 1, 1, 1, 1, 1, 1, 1, 1, 1,  2,  2,  2,  2,  3,  3,  3,  3,  4,
                                 4, 4, 5, 5, 6, 6, 7, 8, 9,  10, 11, 13, 14, 16, 18, 20, 22, 24};  /* kTcTable, deblocking_filter.cc:34-38 */
static const uint8_t k_beta[64] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,
                                   6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
                                   26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56,
                                   58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88};  /* kBetaTable, :40-45 */

typedef struct {
  const xvcb200_cu *cus;
  const int32_t *map;       /* CU index per 4x4 block, -1 outside any CU (PictureData::GetCuAt) */
  int map_w, map_h;
  int pic_type;
  int64_t poc[2][5];
  xvcb200_qp *qps;          /* per CU */
} db_ctx;

static int db_cu_at(const db_ctx *d, int x, int y) {
  if (x < 0 || y < 0 || (x >> 2) >= d->map_w || (y >> 2) >= d->map_h) return -1;
  return d->map[(y >> 2) * d->map_w + (x >> 2)];
}

static int64_t db_ref_poc(const db_ctx *d, const xvcb200_cu *cu, int list) {  /* CodingUnit::GetRefPoc, coding_unit.cc:166-172 */
  return cu->ref_idx[list] < 0 ? -1 : d->poc[list][cu->ref_idx[list]];
}

/* DeblockingFilter::GetBoundaryStrength, deblocking_filter.cc:154-241.  Without affine all
 * four corner MVs of a CU are equal (CodingUnit::SetMv fills them, coding_unit.h:248-250),
 * so the corner selection (:166-176) does not change the value read. */
static int db_strength(const db_ctx *d, const xvcb200_cu *p, const xvcb200_cu *q) {
  if ((p->flags | q->flags) & XVCB200_CU_INTRA) return 2;
  if ((p->flags | q->flags) & XVCB200_CU_CBF_Y) return 1;
#define FAR(a, b) (iabs((a)[0] - (b)[0]) >= 16 || iabs((a)[1] - (b)[1]) >= 16)
  if (d->pic_type == 0) {
    const int64_t p0 = db_ref_poc(d, p, 0), p1 = db_ref_poc(d, p, 1), q0 = db_ref_poc(d, q, 0), q1 = db_ref_poc(d, q, 1);
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    const int straight = FAR(p->mv[0], q->mv[0]) || FAR(p->mv[1], q->mv[1]);
    const int crossed = FAR(p->mv[0], q->mv[1]) || FAR(p->mv[1], q->mv[0]);
    if (p0 != p1) return (p0 == q0) ? straight : crossed;
    return straight && crossed;
  }
  if (p->ref_idx[0] != q->ref_idx[0]) return 1;
  return FAR(p->mv[0], q->mv[0]);
#undef FAR
}

/* FilterEdgeLuma + CheckStrongFilter + FilterLumaWeak + FilterLumaStrong,
 * deblocking_filter.cc:243-401, for one 4-sample edge segment.  `across` steps over the
 * edge, `along` steps along it. */
static void db_luma_segment(uint16_t *s, ptrdiff_t across, ptrdiff_t along, int bitdepth, int bs, int qp, int beta_off,
                            int tc_off) {
  const int bd_shift = bitdepth - 8, maxv = (1 << bitdepth) - 1;
  /* the reference clips the beta index to size() = 64, one past the table (:270-271);
   * index 64 is unreachable for qp <= 51 with zero offsets -- defined here as the last entry */
  int ib = clip3(qp + beta_off, 0, 64);
  if (ib > 63) ib = 63;
  const int beta = k_beta[ib] << bd_shift;
#define P(i, l) ((int)s[(l) * along - ((i) + 1) * across])
#define Q(i, l) ((int)s[(l) * along + (i) * across])
  const int dp0 = iabs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dq0 = iabs(Q(0, 0) - 2 * Q(1, 0) + Q(2, 0));
  const int dp3 = iabs(P(2, 3) - 2 * P(1, 3) + P(0, 3)), dq3 = iabs(Q(0, 3) - 2 * Q(1, 3) + Q(2, 3));
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (d0 + d3 >= beta) return;
  const int tc = k_tc[clip3(qp + tc_off + 2 * (bs - 1), 0, 53)] << bd_shift;
  int strong = (d0 << 1) < (beta >> 2) && (d3 << 1) < (beta >> 2);
  for (int l = 0; l < 4 && strong; l += 3)
    strong = (iabs(P(3, l) - P(0, l)) + iabs(Q(0, l) - Q(3, l))) < (beta >> 3) && iabs(P(0, l) - Q(0, l)) < ((tc * 5 + 1) >> 1);
  if (strong) {
    const int t2 = 2 * tc;
    for (int l = 0; l < 4; l++) {
      const int p3 = P(3, l), p2 = P(2, l), p1 = P(1, l), p0 = P(0, l), q0 = Q(0, l), q1 = Q(1, l), q2 = Q(2, l), q3 = Q(3, l);
      /* delta is clipped to +-2tc, narrowed to Sample and added with no final clip (:392-397) */
#define PUT(ptr, old, nv) (ptr) = (uint16_t)((old) + (uint16_t)clip3((nv) - (old), -t2, t2))
      PUT(s[l * along - 3 * across], p2, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      PUT(s[l * along - 2 * across], p1, (p2 + p1 + p0 + q0 + 2) >> 2);
      PUT(s[l * along - 1 * across], p0, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      PUT(s[l * along], q0, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      PUT(s[l * along + across], q1, (p0 + q0 + q1 + q2 + 2) >> 2);
      PUT(s[l * along + 2 * across], q2, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
#undef PUT
    }
    return;
  }
  const int side = (beta + (beta >> 1)) >> 3;
  const int do_p1 = (dp0 + dp3) < side, do_q1 = (dq0 + dq3) < side;
  const int half = tc >> 1;
  for (int l = 0; l < 4; l++) {
    const int p2 = P(2, l), p1 = P(1, l), p0 = P(0, l), q0 = Q(0, l), q1 = Q(1, l), q2 = Q(2, l);
    int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
    if (iabs(delta) >= tc * 10) continue;
    delta = clip3(delta, -tc, tc);
    s[l * along - across] = (uint16_t)clip3(p0 + delta, 0, maxv);
    s[l * along] = (uint16_t)clip3(q0 - delta, 0, maxv);
    if (do_p1) s[l * along - 2 * across] = (uint16_t)clip3(p1 + clip3((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -half, half), 0, maxv);
    if (do_q1) s[l * along + across] = (uint16_t)clip3(q1 + clip3((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -half, half), 0, maxv);
  }
#undef P
#undef Q
}

/* FilterEdgeChroma + FilterChroma<N>, deblocking_filter.cc:403-450 */
static void db_chroma_segment(uint16_t *s, ptrdiff_t across, ptrdiff_t along, int len, int bitdepth, int qp, int tc_off) {
  int it = clip3(qp + tc_off + 2, 0, 54);   /* reference clips to size() = 54 (:407-408); see beta note */
  if (it > 53) it = 53;
  const int tc = k_tc[it] << (bitdepth - 8), maxv = (1 << bitdepth) - 1;
  for (int l = 0; l < len; l++) {
    const int p1 = s[l * along - 2 * across], p0 = s[l * along - across], q0 = s[l * along], q1 = s[l * along + across];
    const int delta = clip3((((q0 - p0) * 4) + p1 - q1 + 4) >> 3, -tc, tc);
    s[l * along - across] = (uint16_t)clip3(p0 + delta, 0, maxv);
    s[l * along] = (uint16_t)clip3(q0 - delta, 0, maxv);
  }
}

/* DeblockingFilter::DeblockPicture / DeblockCtu, deblocking_filter.cc:56-152, for a picture
 * with a single CU tree (inter pictures): all vertical edges in CTU raster order, then all
 * horizontal edges; 4x4 grid; chroma only for bs == 2 on the 8-sample chroma grid. */
void xo_deblock_picture(xo_picture *rec, int bitdepth, const xvcb200_cu *cus, int n, int pic_type, int beta_offset,
                        int tc_offset, int table, int off_u, int off_v, const int64_t ref_poc[2][5]) {
  xo_deblock_band(rec, bitdepth, cus, n, pic_type, beta_offset, tc_offset, table, off_u, off_v, ref_poc, 3, 0,
                  rec->height[0]);
}

/* The same walk restricted to one band of rows and to the selected passes (1 = vertical edges,
 * 2 = horizontal edges): the unit a CTB-row shard executes.  Whole-picture order is recovered
 * when the bands run top to bottom with their halo rows exchanged. */
void xo_deblock_band(xo_picture *rec, int bitdepth, const xvcb200_cu *cus, int n, int pic_type, int beta_offset,
                     int tc_offset, int table, int off_u, int off_v, const int64_t ref_poc[2][5], int pass_mask,
                     int y_begin, int y_end) {
  const int W = rec->width[0], H = rec->height[0];
  db_ctx d;
  d.cus = cus; d.pic_type = pic_type;
  d.map_w = (W + 3) >> 2; d.map_h = (H + 3) >> 2;
  memcpy(d.poc, ref_poc, sizeof(d.poc));
  int32_t *map = (int32_t *)malloc(sizeof(int32_t) * (size_t)d.map_w * d.map_h);
  for (int i = 0; i < d.map_w * d.map_h; i++) map[i] = -1;
  d.qps = (xvcb200_qp *)malloc(sizeof(xvcb200_qp) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    cu_qp(&cus[i], bitdepth, table, off_u, off_v, &d.qps[i]);
    for (int y = cus[i].y >> 2; y < (cus[i].y + cus[i].h) >> 2; y++)
      for (int x = cus[i].x >> 2; x < (cus[i].x + cus[i].w) >> 2; x++) map[y * d.map_w + x] = i;
  }
  d.map = map;
  const int ctus_x = (W + 63) >> 6, ctus_y = (H + 63) >> 6;
  for (int dir = 0; dir < 2; dir++)       /* 0: vertical edges, 1: horizontal edges */
    for (int ctu = 0; ctu < ctus_x * ctus_y; ctu++)
      for (int dy = 0; dy < 64; dy += 4)
        for (int dx = 0; dx < 64; dx += 4) {
          const int x = (ctu % ctus_x) * 64 + dx, y = (ctu / ctus_x) * 64 + dy;
          if (!(pass_mask & (1 << dir)) || y < y_begin || y >= y_end) continue;
          const int iq = db_cu_at(&d, x, y);
          if (iq < 0) continue;
          const int ip = dir == 0 ? db_cu_at(&d, x - 1, y) : db_cu_at(&d, x, y - 1);
          if (ip < 0 || ip == iq) continue;
          const xvcb200_cu *p = &cus[ip], *q = &cus[iq];
          const int bs = db_strength(&d, p, q);
          if (!bs) continue;
          const int qp = (p->qp + q->qp + 1) >> 1;
          uint16_t *s = rec->base[0] + y * rec->stride[0] + x;
          const ptrdiff_t across = dir == 0 ? 1 : rec->stride[0], along = dir == 0 ? rec->stride[0] : 1;
          db_luma_segment(s, across, along, bitdepth, bs, qp, beta_offset, tc_offset);
          if (bs == 2) {
            const int cqp = (d.qps[ip].qp_raw[1] + d.qps[iq].qp_raw[1] + 1) >> 1;   /* cu.GetQp(kU) for both planes (:127) */
            const int cx = x >> 1, cy = y >> 1;
            if ((dir == 0 ? cx : cy) & 7) continue;
            for (int c = 1; c < 3; c++) {
              uint16_t *cs = rec->base[c] + cy * rec->stride[c] + cx;
              db_chroma_segment(cs, dir == 0 ? 1 : rec->stride[c], dir == 0 ? rec->stride[c] : 1, 2, bitdepth, cqp, tc_offset);
            }
          }
        }
  free(map);
  free(d.qps);
}

/* ===================================================================================
 * One inter picture through the whole hot path (the step bench.py times)
 * =================================================================================== */
void xo_encode_picture(const xo_picture *orig, const xo_picture *const refs[2][5], xo_picture *pred, xo_picture *rec,
                       int16_t *const levels[3], int bitdepth, xvcb200_cu *cus, int n,
                       const xvcb200_picture_params *params, xvcb200_me_result *me_results,
                       xvcb200_tu_result *tu_results) {
  const int nl = params->pic_type == 0 ? 2 : 1;
  xvcb200_me_job *jobs = (xvcb200_me_job *)calloc((size_t)n * nl, sizeof(*jobs));
  xvcb200_me_result *res = (xvcb200_me_result *)calloc((size_t)n * nl, sizeof(*res));
  for (int i = 0; i < n; i++)
    for (int l = 0; l < nl; l++) {
      xvcb200_me_job *j = &jobs[i * nl + l];
      j->cu = i; j->ref_slot = 0; j->list = l; j->search_range = params->search_range[l][0];
      j->mvp[0] = cus[i].mv[l][0]; j->mvp[1] = cus[i].mv[l][1];
    }
  xo_me_search(orig, refs, bitdepth, cus, jobs, n * nl, params->lambda_sqrt, res);
  for (int i = 0; i < n; i++) {
    const int best = (nl == 2 && res[2 * i + 1].cost < res[2 * i].cost) ? 1 : 0;
    for (int l = 0; l < 2; l++) {
      cus[i].ref_idx[l] = (int8_t)(l == best ? 0 : -1);
      cus[i].mv[l][0] = l == best ? res[i * nl + l].mv[0] : 0;
      cus[i].mv[l][1] = l == best ? res[i * nl + l].mv[1] : 0;
    }
  }
  if (me_results) memcpy(me_results, res, sizeof(*res) * (size_t)n * nl);
  xo_motion_compensate(refs, bitdepth, cus, n, pred);
  xo_tq_reconstruct(orig, pred, rec, levels, bitdepth, cus, n, 0, params->chroma_offset_table, params->chroma_offset_u,
                    params->chroma_offset_v, tu_results);
  if (params->deblock)
    xo_deblock_picture(rec, bitdepth, cus, n, params->pic_type, params->beta_offset, params->tc_offset,
                       params->chroma_offset_table, params->chroma_offset_u, params->chroma_offset_v, params->ref_poc);
  if (params->pad) xo_pad_border(rec);
  free(jobs);
  free(res);
}
