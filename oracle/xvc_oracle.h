/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the xvc hot path.
 *
 * This is the parity oracle for xvc_b200's CUDA kernels.  It is NOT part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  Every function cites the reference code it restates (paths relative to the
 * reference tree, divideon/xvc @ e875a2e).
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function here against
 * the unmodified reference compiled by oracle/Makefile (oracle/_ref/libxvcref.so) on seeded
 * inputs, and tests/test_oracle_golden.py checks it against the committed vectors under
 * tests/golden/ that were generated from that same reference build
 * (tests/golden/make_golden.py).
 */
#ifndef XVC_ORACLE_H_
#define XVC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "../include/xvc_b200.h" /* shared plain-C descriptor structs (xvcb200_cu, jobs, results) */

#ifdef __cplusplus
extern "C" {
#endif

/* A picture as three planes; base[c] points at sample (0,0); `pad` addressable samples
 * exist on every side (0 for an unpadded original).  yuv_pic.cc:32-68. */
typedef struct {
  uint16_t *base[3];
  int32_t stride[3];
  int32_t width[3], height[3];
  int32_t pad[3];
} xo_picture;

/* ---- leaf metrics (sample_metric.cc) ---- */
int xo_sad(int kind /*0 u16/u16, 1 i16/u16*/, int w, int h, const void *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb);
uint64_t xo_ssd(int kind /*0 u16/u16, 1 i16/u16, 2 i16/i16*/, int w, int h, const void *a, ptrdiff_t sa,
                const void *b, ptrdiff_t sb);
uint64_t xo_satd(int bitdepth, int w, int h, int first_short, const void *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb);
/* metric: XVCB200_METRIC_*; result before the chroma weight */
uint64_t xo_compare(int metric, int bitdepth, int w, int h, int first_short, const void *a, ptrdiff_t sa,
                    const uint16_t *b, ptrdiff_t sb);
uint64_t xo_apply_weight(uint64_t dist, double weight);

/* ---- leaf interpolation (inter_prediction.cc) ---- */
void xo_filter(int kind, int chroma, int w, int h, int bitdepth, const int16_t *taps, const void *src, ptrdiff_t ss,
               void *dst, ptrdiff_t ds);
void xo_add_avg(int w, int h, int offset, int shift, int bitdepth, const int16_t *a, ptrdiff_t sa, const int16_t *b,
                ptrdiff_t sb, uint16_t *dst, ptrdiff_t ds);
void xo_filter_copy_bipred(int w, int h, int offset, int shift, const uint16_t *ref, ptrdiff_t rs, int16_t *pred,
                           ptrdiff_t ps);
void xo_interp(int chroma, int bipred, int w, int h, int bitdepth, int frac_x, int frac_y, const uint16_t *ref,
               ptrdiff_t rs, void *pred, ptrdiff_t ps);
const int16_t *xo_luma_taps(int frac);   /* kLumaFilterHighPrec[frac], 0..15 */
const int16_t *xo_chroma_taps(int frac); /* kChromaFilterHighPrec[frac], 0..31 */

/* ---- leaf transform / quant (transform.cc, quantize.cc, rdo_quant.cc) ---- */
const int16_t *xo_transform_matrix(int type /*XVCB200_TX_*, DEFAULT==DCT2*/, int n);
void xo_fwd_transform(int w, int h, int bitdepth, int tx_hor, int tx_ver, int dst4x4, const int16_t *resi, ptrdiff_t rs,
                      int16_t *coeff, ptrdiff_t cs);
void xo_inv_transform(int w, int h, int bitdepth, int tx_hor, int tx_ver, int dst4x4, int dc_only, const int16_t *coeff,
                      ptrdiff_t cs, int16_t *resi, ptrdiff_t rs);
void xo_transform_skip(int forward, int w, int h, int bitdepth, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os);
void xo_qp_init(xvcb200_qp *out, int qp, int chroma_format, int bitdepth, double lambda, int table, int off_u, int off_v);
int xo_quant_fast(int w, int h, int bitdepth, int qp_bitdepth, int intra_picture, int sign_hiding, int scan_order,
                  const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os);
void xo_dequant(int w, int h, int bitdepth, int qp_bitdepth, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os);

/* ---- picture level ---- */
void xo_pad_border(xo_picture *pic);
void xo_clip_mv(int pos_x, int pos_y, int pic_w, int pic_h, int32_t mv[2]);
void xo_min_max_mv(int pos_x, int pos_y, int pic_w, int pic_h, const int32_t center[2], int range, int32_t mv_min[2],
                   int32_t mv_max[2]);
uint32_t xo_exp_golomb_bits(int v);

/* TzSearch::Search; lambda_me = floor(65536*lambda_sqrt).  Returns #candidates evaluated. */
int xo_tz_search(const xo_picture *orig, const xo_picture *ref, int bitdepth, const xvcb200_cu *cu,
                 const xvcb200_me_job *job, uint32_t lambda_me, int32_t mv_out[2], uint32_t *cost_out);
/* SubpelSearch (or GetSubpelDist only, when the CU has XVCB200_CU_FULLPEL_MV) */
void xo_subpel_search(const xo_picture *orig, const xo_picture *ref, int bitdepth, const xvcb200_cu *cu,
                      const int32_t mvp[2], const int32_t mv_fullpel[2], uint32_t lambda_me, int32_t mv_out[2],
                      uint32_t *dist_out, uint32_t *cost_out);
/* refs[list][ref_idx]; job.ref_slot is the ref_idx inside job.list */
void xo_me_search(const xo_picture *orig, const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus,
                  const xvcb200_me_job *jobs, int n, double lambda_sqrt, xvcb200_me_result *results);
void xo_full_search(const xo_picture *orig, const xo_picture *other_pred, const xo_picture *ref, int bitdepth,
                    const xvcb200_cu *cu, const xvcb200_fullsearch_job *job, uint32_t lambda_me, int32_t mv_out[2],
                    uint32_t *cost_out);
void xo_motion_compensate(const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus, int n,
                          xo_picture *pred);
/* levels: picture-shaped tight int16 planes (stride = plane width) */
void xo_motion_compensate_affine(const xo_picture *const refs[2][5], int bitdepth, const xvcb200_cu *cus,
                                 const xvcb200_affine_cu *aff, int n_aff, xo_picture *pred);
void xo_motion_compensate_lic(const xo_picture *const refs[2][5], const xo_picture *rec, int bitdepth,
                              const xvcb200_cu *cus, const xvcb200_lic_cu *lic, int n_lic, xo_picture *pred);
void xo_tq_reconstruct(const xo_picture *orig, const xo_picture *pred, xo_picture *rec, int16_t *const levels[3],
                       int bitdepth, xvcb200_cu *cus, int n, int intra_picture, int table, int off_u, int off_v,
                       xvcb200_tu_result *results);
void xo_dequant_reconstruct(const xo_picture *pred, xo_picture *rec, int16_t *const levels[3], int bitdepth,
                            const xvcb200_cu *cus, int n, int table, int off_u, int off_v);
void xo_deblock_picture(xo_picture *rec, int bitdepth, const xvcb200_cu *cus, int n, int pic_type, int beta_offset,
                        int tc_offset, int table, int off_u, int off_v, const int64_t ref_poc[2][5]);
void xo_deblock_band(xo_picture *rec, int bitdepth, const xvcb200_cu *cus, int n, int pic_type, int beta_offset,
                     int tc_offset, int table, int off_u, int off_v, const int64_t ref_poc[2][5], int pass_mask,
                     int y_begin, int y_end);
void xo_encode_picture(const xo_picture *orig, const xo_picture *const refs[2][5], xo_picture *pred, xo_picture *rec,
                       int16_t *const levels[3], int bitdepth, xvcb200_cu *cus, int n,
                       const xvcb200_picture_params *params, xvcb200_me_result *me_results,
                       xvcb200_tu_result *tu_results);

/* ---- intra prediction (intra_prediction.cc, xvc_oracle_intra.c); ref arrays: 2 x XVCB200_INTRA_REF_STRIDE ---- */
void xo_intra_ref_samples(int w, int h, int bitdepth, int has_above_left, int has_above, int above_right, int has_left,
                          int below_left, const uint16_t *block, ptrdiff_t stride, uint16_t *ref);
void xo_intra_filter_ref(int w, int h, const uint16_t *src, uint16_t *dst);
int xo_intra_use_filtered_ref(int mode, int w, int h);
void xo_intra_predict(int mode, int w, int h, int bitdepth, int luma, const uint16_t *ref_samples,
                      const uint16_t *ref_filtered, uint16_t *out, ptrdiff_t os);
void xo_intra_lm_chroma(int x, int y, int w, int h, int bitdepth, const uint16_t *luma, ptrdiff_t ls, const uint16_t *chroma,
                        ptrdiff_t cs, uint16_t *pred, ptrdiff_t ps);
void xo_intra_satd_scan(int w, int h, int bitdepth, const uint16_t *orig, ptrdiff_t ostride, const uint16_t *ref_samples,
                        const uint16_t *ref_filtered, uint32_t *satd);

#ifdef __cplusplus
}
#endif
#endif
