/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of xvc's intra prediction
 * (xvc_common_lib/intra_prediction.cc) for the SURVEY section 8(f) "next" row: reference
 * sample construction, reference smoothing, DC / planar / angular prediction in the
 * unrestricted (67 mode) configuration, and the SATD scan over all luma modes that opens the
 * encoder's intra mode search (xvc_enc_lib/intra_search.cc:185-216).
 *
 * Parity status: PINNED -- tests/test_oracle_vs_ref.py compares every function with the
 * unmodified reference (oracle/_ref/libxvcref.so), tests/test_oracle_golden.py with the
 * committed vectors generated from it.  Not part of the product (see xvc_oracle.h).
 *
 * Reference sample layout (IntraPrediction::RefState, intra_prediction.h:40-44): an array of
 * 2 x 129 samples; [0] = above-left, [1 .. w+h] = above and above-right, [129 + y] = left and
 * below-left, y < w+h.
 */
#include <stdlib.h>
#include <string.h>

#include "xvc_oracle.h"

#define XO_RS XVCB200_INTRA_REF_STRIDE

/* kAngleTableExt / kInvAngleTableExt, intra_prediction.cc:38-50 */
static const int8_t k_angle[33] = {-32, -29, -26, -23, -21, -19, -17, -15, -13, -11, -9, -7, -5, -3, -2, -1, 0,
                                   1,   2,   3,   5,   7,   9,   11,  13,  15,  17,  19, 21, 23, 26, 29, 32};
static const int16_t k_inv_angle[16] = {8192, 4096, 2731, 1638, 1170, 910, 745, 630, 546, 482, 431, 390, 356, 315, 282, 256};

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }
static int clip_bd(int v, int maxv) { return v < 0 ? 0 : (v > maxv ? maxv : v); }

/* IntraPrediction::ComputeRefSamples, intra_prediction.cc:709-851.  `block` points at the
 * top-left sample of the block inside the reconstructed plane.  above_right / below_left are
 * the numbers of available samples beyond the block (0 = none), as
 * CodingUnit::GetCuSizeAboveRight / GetCuSizeBelowLeft report them. */
void xo_intra_ref_samples(int w, int h, int bitdepth, int has_above_left, int has_above, int above_right, int has_left,
                          int below_left, const uint16_t *block, ptrdiff_t stride, uint16_t *ref) {
  const uint16_t dc = (uint16_t)(1 << (bitdepth - 1));
  const int n = w + h;
  if (!has_above_left && !has_above && !has_left && above_right <= 0 && below_left <= 0) {
    for (int i = 0; i <= n; i++) ref[i] = dc;
    for (int i = 0; i < n; i++) ref[XO_RS + i] = dc;
    return;
  }
  /* One line from the bottom of the left column up to the corner and on to the right end of
   * the row above: line[n-1-y] = left sample y, line[n .. n+w) = the corner (w copies, as the
   * reference keeps them), line[n+w+x] = above sample x. */
  uint16_t line[5 * 64];
  for (int i = 0; i < 2 * n + w; i++) line[i] = dc;
  if (has_above_left)
    for (int i = 0; i < w; i++) line[n + i] = block[-stride - 1];
  if (has_left) {
    for (int y = 0; y < h; y++) line[n - 1 - y] = block[y * stride - 1];
    if (below_left > 0) {
      for (int i = 0; i < below_left; i++) line[n - 1 - h - i] = block[(h + i) * stride - 1];
      for (int i = below_left; i < w; i++) line[n - 1 - h - i] = line[n - h - below_left];   /* beyond the picture */
    }
  }
  if (has_above) {
    for (int x = 0; x < w; x++) line[n + w + x] = block[-stride + x];
    if (above_right > 0) {
      for (int i = 0; i < above_right; i++) line[n + 2 * w + i] = block[-stride + w + i];
      for (int i = above_right; i < h; i++) line[n + 2 * w + i] = line[n + 2 * w + above_right - 1];
    }
  }
  /* padding of what is missing, from the bottom-left end upwards (:806-839) */
  if (below_left <= 0) {
    uint16_t v;
    if (has_left) v = line[w];
    else if (has_above_left) v = line[n];
    else if (has_above) v = line[n + w];
    else v = line[n + 2 * w];
    for (int i = 0; i < w; i++) line[i] = v;
  }
  if (!has_left)
    for (int i = 0; i < h; i++) line[w + i] = line[w - 1];
  if (!has_above_left)
    for (int i = 0; i < w; i++) line[n + i] = line[n - 1];
  if (!has_above)
    for (int i = 0; i < w; i++) line[n + w + i] = line[n + w - 1];
  if (above_right <= 0)
    for (int i = 0; i < h; i++) line[n + 2 * w + i] = line[n + 2 * w - 1];
  for (int x = 0; x <= n; x++) ref[x] = line[n + w - 1 + x];
  for (int y = 0; y < n; y++) ref[XO_RS + y] = line[n - 1 - y];
}

/* IntraPrediction::FilterRefSamples, intra_prediction.cc:853-876: [1 2 1] / 4 along the two
 * reference edges, the far ends unfiltered. */
void xo_intra_filter_ref(int w, int h, const uint16_t *src, uint16_t *dst) {
  const int n = w + h;
  dst[0] = (uint16_t)((2 * src[0] + src[1] + src[XO_RS] + 2) >> 2);
  for (int x = 1; x < n; x++) dst[x] = (uint16_t)((2 * src[x] + src[x - 1] + src[x + 1] + 2) >> 2);
  dst[n] = src[n];
  dst[XO_RS] = (uint16_t)((2 * src[XO_RS] + src[0] + src[XO_RS + 1] + 2) >> 2);
  for (int y = 1; y < n - 1; y++) dst[XO_RS + y] = (uint16_t)((2 * src[XO_RS + y] + src[XO_RS + y - 1] + src[XO_RS + y + 1] + 2) >> 2);
  dst[XO_RS + n - 1] = src[XO_RS + n - 1];
}

/* IntraPrediction::UseFilteredRefSamples, intra_prediction.cc:342-364 (67-mode thresholds);
 * w, h = luma size of the CU. */
int xo_intra_use_filtered_ref(int mode, int w, int h) {
  static const int8_t thr[8] = {0, 20, 20, 14, 2, 0, 20, 0};
  const int size = (ilog2(w) + ilog2(h)) >> 1;
  const int dh = abs(mode - 18), dv = abs(mode - 50);
  return (dh < dv ? dh : dv) > thr[size];
}

/* IntraPrediction::Predict, intra_prediction.cc:81-126, for planar (0), DC (1) and the angular
 * modes 2..66.  luma: component is luma (selects smoothing and the edge filters). */
void xo_intra_predict(int mode, int w, int h, int bitdepth, int luma, const uint16_t *ref_samples,
                      const uint16_t *ref_filtered, uint16_t *out, ptrdiff_t os) {
  const int maxv = (1 << bitdepth) - 1;
  const uint16_t *ref = (luma && ref_filtered && xo_intra_use_filtered_ref(mode, w, h)) ? ref_filtered : ref_samples;
  const int post = luma && w <= 16 && h <= 16;
  if (mode == 0) {                                   /* PlanarPred, :402-424 */
    const int lw = ilog2(w), lh = ilog2(h), shift = lw + lh + 1;
    const uint16_t *above = ref + 1, *left = ref + XO_RS;
    const int tr = above[w], bl = left[h];
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        const int hor = (h - 1 - y) * above[x] + (y + 1) * bl;
        const int ver = (w - 1 - x) * left[y] + (x + 1) * tr;
        out[y * os + x] = (uint16_t)(((hor << lw) + (ver << lh) + (1 << (shift - 1))) >> shift);
      }
    return;
  }
  if (mode == 1) {                                   /* PredIntraDC, :366-400; always the unfiltered samples */
    const uint16_t *above = ref_samples + 1, *left = ref_samples + XO_RS;
    int sum = 0;
    for (int x = 0; x < w; x++) sum += above[x];
    for (int y = 0; y < h; y++) sum += left[y];
    const int dc = (sum + ((w + h) >> 1)) / (w + h);
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) out[y * os + x] = (uint16_t)dc;
    if (post) {
      for (int y = 1; y < h; y++) out[y * os] = (uint16_t)((left[y] + 3 * dc + 2) >> 2);
      for (int x = 1; x < w; x++) out[x] = (uint16_t)((above[x] + 3 * dc + 2) >> 2);
      out[0] = (uint16_t)((above[0] + left[0] + 2 * dc + 2) >> 2);
    }
    return;
  }
  /* AngularPred, :426-558.  Horizontal-class modes (< 34) are the vertical-class prediction of
   * the transposed block with the two edges exchanged. */
  const int horizontal = mode < 34;
  const int angle_offset = horizontal ? 18 - mode : mode - 50;
  const int angle = k_angle[16 + angle_offset];
  const uint16_t *main_edge = horizontal ? ref + XO_RS : ref + 1;     /* the edge the direction starts from */
  const uint16_t *side_edge = horizontal ? ref + 1 : ref + XO_RS;     /* the other edge */
  const uint16_t corner = ref[0];
  const int pw = horizontal ? h : w, ph = horizontal ? w : h;         /* size in the prediction's own orientation */
  for (int py = 0; py < ph; py++)
    for (int px = 0; px < pw; px++) {
      int v;
      if (angle == 0) {
        v = main_edge[px];
        if (post && px == 0) v = clip_bd((int16_t)(v + ((side_edge[py] - corner) >> 1)), maxv);
      } else {
        const int sum = (py + 1) * angle, off = sum >> 5, wgt = sum & 31;
        int s[2];
        for (int t = 0; t < 2; t++) {
          const int j = off + px + t;        /* position on the prediction line, -1 = the corner */
          if (j == -1) s[t] = corner;
          else if (j >= 0) s[t] = main_edge[j];
          else s[t] = side_edge[((128 + (-1 - j) * k_inv_angle[-angle_offset - 1]) >> 8) - 1];   /* projected, :478-487 */
        }
        v = wgt ? ((32 - wgt) * s[0] + wgt * s[1] + 16) >> 5 : s[0];
        if (post && px == 0 && (angle == 1 || angle == -1)) v = clip_bd((int16_t)(v + ((side_edge[py] - corner) >> 2)), maxv);
      }
      if (horizontal) out[px * os + py] = (uint16_t)v;
      else out[py * os + px] = (uint16_t)v;
    }
}

/* First loop of IntraSearch::DetermineSlowIntraModes (intra_search.cc:185-216) without the
 * mode bits: SATD (SampleMetric kSatd, luma) of every mode's prediction against the original
 * block.  satd[67]; the reference itself skips the odd angular modes in this pass. */
void xo_intra_satd_scan(int w, int h, int bitdepth, const uint16_t *orig, ptrdiff_t ostride, const uint16_t *ref_samples,
                        const uint16_t *ref_filtered, uint32_t *satd) {
  uint16_t pred[64 * 64];
  for (int mode = 0; mode < 67; mode++) {
    xo_intra_predict(mode, w, h, bitdepth, 1, ref_samples, ref_filtered, pred, 64);
    satd[mode] = (uint32_t)xo_satd(bitdepth, w, h, 0, orig, ostride, pred, 64);
  }
}

/* ---- chroma from luma (LM chroma) ----------------------------------------------------------
 * IntraPrediction::PredLmChroma / RescaleLuma (4:2:0 branch) / DeriveLmParams,
 * intra_prediction.cc:560-686, 873-913.  The CU's reconstructed luma (and one row above / one
 * column left of it, when the CU is not at the picture border -- the reference tests the POSITION,
 * not the availability) is reduced to chroma resolution with a [1 2 1; 1 2 1]/8 kernel, a linear
 * model chroma ~ luma is fitted on that border against the reconstructed chroma neighbours, and
 * applied to the reduced luma of the block.  Shift counts are masked to five bits where the
 * reference shifts by a computed count (the behaviour of the x86 build this oracle is pinned to). */
static int log2_floor(int x) { int l = 0; while (x > 1) { l++; x >>= 1; } return l; }   /* util::Log2Floor, utils.cc:46-53 */
static int size_to_log2_lm(int size) { int l = 1; while ((1 << l) < size) l++; return l; }

/* sub: (ch+1) x (cw+1) samples, row/column 0 = the border; sub[(y+1) * ss + (x+1)] = reduced luma (x, y) */
static void lm_rescale_luma(const uint16_t *luma, ptrdiff_t ls, int cw, int ch, int has_above, int has_left, uint16_t *sub, int ss) {
  for (int y = has_above ? -1 : 0; y < ch; y++) {
    const uint16_t *s0 = luma + (ptrdiff_t)(2 * y) * ls, *s1 = s0 + ls;
    for (int x = has_left ? -1 : 0; x < cw; x++) {
      int v;
      if (!has_left && x == 0) v = (s0[0] + s1[0] + 1) >> 1;                                            /* :897-903 */
      else v = (s0[2 * x - 1] + 2 * s0[2 * x] + s0[2 * x + 1] + s1[2 * x - 1] + 2 * s1[2 * x] + s1[2 * x + 1] + 4) >> 3;
      sub[(y + 1) * ss + (x + 1)] = (uint16_t)v;
    }
  }
}

static void lm_params(int bitdepth, int cw, int ch, int has_above, int has_left, const uint16_t *chroma, ptrdiff_t cs,
                      const uint16_t *sub, int ss, int *scale_out, int *offset_out, int *shift_out) {
  *scale_out = 0; *offset_out = 1 << (bitdepth - 1); *shift_out = 0;
  if (!has_above && !has_left) return;
  int sum_x = 0, sum_y = 0, sum_xx = 0, sum_xy = 0, nbr = 0;
  if (has_above) {
    const int dx = has_left ? ((cw / ch) > 1 ? cw / ch : 1) : 1;
    for (int x = 0; x < cw; x += dx) {
      const int a = sub[x + 1], b = chroma[-cs + x];
      sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b; nbr++;
    }
  }
  if (has_left) {
    const int dy = has_above ? ((ch / cw) > 1 ? ch / cw : 1) : 1;
    for (int y = 0; y < ch; y += dy) {
      const int a = sub[(y + 1) * ss], b = chroma[y * cs - 1];
      sum_x += a; sum_y += b; sum_xx += a * a; sum_xy += a * b; nbr++;
    }
  }
  int size_shift = size_to_log2_lm(nbr);
  if (size_shift > 15 - bitdepth) {
    const int sh = size_shift + bitdepth - 15;
    sum_x = (sum_x + (1 << (sh - 1))) >> sh; sum_y = (sum_y + (1 << (sh - 1))) >> sh;
    sum_xx = (sum_xx + (1 << (sh - 1))) >> sh; sum_xy = (sum_xy + (1 << (sh - 1))) >> sh;
    size_shift -= sh;
  }
  const int avg_x = sum_x >> size_shift, avg_y = sum_y >> size_shift;
  const int x_frac = sum_x & ((1 << size_shift) - 1), y_frac = sum_y & ((1 << size_shift) - 1);
  const int sd_xy = sum_xy - ((avg_x * avg_y) << size_shift) - avg_x * y_frac - avg_y * x_frac;
  const int sd_xx = sum_xx - ((avg_x * avg_x) << size_shift) - 2 * avg_x * x_frac;
  int shift_xy = 0, shift_xx = 0;
  if (sd_xy != 0) { shift_xy = log2_floor(abs(sd_xy)) - bitdepth + 2; if (shift_xy < 0) shift_xy = 0; }
  if (sd_xx != 0) { shift_xx = log2_floor(abs(sd_xx)) - 5; if (shift_xx < 0) shift_xx = 0; }
  const int sd_xy_s = sd_xy >> shift_xy, sd_xx_s = sd_xx >> shift_xx;
  const int total_shift = bitdepth + shift_xx + 4 + 7 - 13 - shift_xy;
  if (sd_xx_s < 32) { *offset_out = avg_y; return; }
  int scale = (int)((uint32_t)sd_xy_s * (uint32_t)(((1 << (bitdepth + 4)) + sd_xx_s / 2) / sd_xx_s));
  scale = scale >> (total_shift & 31);
  scale = scale < -256 ? -256 : scale > 255 ? 255 : scale;
  scale *= 128;
  const int base_shift = log2_floor(abs(scale) + (scale < 0 ? -1 : 0)) - (scale ? 5 : 0);
  const int shift = 13 - base_shift;
  scale >>= base_shift;
  *scale_out = scale; *shift_out = shift;
  *offset_out = avg_y - ((scale * avg_x) >> (shift & 31));
}

/* One chroma block: luma = the CU's top-left luma sample inside the reconstructed luma plane, chroma = the
 * block's top-left sample inside the reconstructed chroma plane of the component, (x, y) = luma position. */
void xo_intra_lm_chroma(int x, int y, int w, int h, int bitdepth, const uint16_t *luma, ptrdiff_t ls, const uint16_t *chroma,
                        ptrdiff_t cs, uint16_t *pred, ptrdiff_t ps) {
  uint16_t sub[33 * 33];
  const int cw = w >> 1, ch = h >> 1, ss = 33, has_above = y > 0, has_left = x > 0;
  lm_rescale_luma(luma, ls, cw, ch, has_above, has_left, sub, ss);
  int scale, offset, shift;
  lm_params(bitdepth, cw, ch, has_above, has_left, chroma, cs, sub, ss, &scale, &offset, &shift);
  const int maxv = (1 << bitdepth) - 1;
  for (int yy = 0; yy < ch; yy++)
    for (int xx = 0; xx < cw; xx++) {
      const int v = ((scale * sub[(yy + 1) * ss + xx + 1]) >> (shift & 31)) + offset;     /* AddLinearModel, sample_buffer.h:108-122 */
      pred[yy * ps + xx] = (uint16_t)(v < 0 ? 0 : v > maxv ? maxv : v);
    }
}
