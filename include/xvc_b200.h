/*
 * xvc_b200 -- C ABI of the B200-native xvc hot path (libxvc_b200.so).
 *
 * This is the drop-in boundary.  Everything a host encoder/decoder loop needs from the
 * GPU goes through the plain-C entry points below: no C++ types, no torch types, plain
 * pointers and sizes.  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree, divideon/xvc @ e875a2e).
 *
 * Three flavours:
 *  (A) table-shaped, HOST pointers, one block per call -- the exact signatures of
 *      xvc's SIMD function tables (SampleMetric::SimdFunc, InterPrediction::SimdFunc)
 *      plus the scalar class methods that have no table entry in the reference
 *      (transform, quant, dequant).  Correctness vehicle / drop-in for the table, slow
 *      by construction (one launch + two copies per block).
 *  (B) batched, DEVICE-resident pictures -- the performance surface.  Pictures live in
 *      HBM in "slots"; work is described by arrays of CU descriptors and launched per
 *      picture.  Asynchronous on the context stream; xvcb200_sync() returns the sticky
 *      status.
 *  (C) picture-level: xvcb200_encode_picture() chains ME -> MC -> T/Q/recon -> deblock
 *      -> pad for one inter picture.
 *
 * Sample = uint16_t (reference default build HIGH_BITDEPTH=ON, common.h:34-38),
 * Coeff = Residual = int16_t (common.h:39-40).  Strides are in elements, not bytes.
 * Error convention: table-shaped calls return their value / void like the reference and
 * record failures in a sticky status of the calling thread (xvcb200_last_error); batched calls
 * return an xvcb200_status and never throw.
 */
#ifndef XVC_B200_H_
#define XVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  XVCB200_OK = 0,
  XVCB200_INVALID_ARGUMENT = 1,   /* cf. XVC_ENC_INVALID_ARGUMENT, xvcenc.h:45-61 */
  XVCB200_NO_DEVICE = 2,
  XVCB200_CUDA_ERROR = 3,
  XVCB200_OUT_OF_MEMORY = 4,
  XVCB200_UNSUPPORTED = 5
} xvcb200_status;

/* sticky status of the calling thread for the void/value-returning table calls */
int xvcb200_last_error(void);
const char *xvcb200_last_error_string(void);
void xvcb200_clear_error(void);
/* number of kernels this library has launched since load (bench.py gpu_launches) */
uint64_t xvcb200_launch_count(void);
const char *xvcb200_version(void);
/* CUDA devices visible to the library (0: none -- every compute entry then fails, there is no CPU path) */
int xvcb200_device_count(void);
/* sizeof() of the structs below as compiled (0 cu, 1 me_job, 2 me_result, 3 fullsearch_job,
 * 4 tu_result, 5 picture_params, 6 plane_geom, 7 qp, 8 intra_job, 9 affine_cu, 10 lic_cu) -- lets
 * bindings verify their layout */
int xvcb200_abi_sizeof(int which);

/* ------------------------------------------------------------------------------------
 * (A) table-shaped entry points, host pointers
 * ---------------------------------------------------------------------------------- */

/* SampleMetric::SimdFunc (xvc_enc_lib/sample_metric.h:166-191); C defaults
 * ComputeSad_c / ComputeSsd_c (sample_metric.cc:671-684, 301-314). */
int xvcb200_sad_sample_sample(int width, int height, const uint16_t *src1, ptrdiff_t stride1,
                              const uint16_t *src2, ptrdiff_t stride2);
int xvcb200_sad_short_sample(int width, int height, const int16_t *src1, ptrdiff_t stride1,
                             const uint16_t *src2, ptrdiff_t stride2);
uint64_t xvcb200_ssd_sample_sample(int width, int height, const uint16_t *src1, ptrdiff_t stride1,
                                   const uint16_t *src2, ptrdiff_t stride2);
uint64_t xvcb200_ssd_short_sample(int width, int height, const int16_t *src1, ptrdiff_t stride1,
                                  const uint16_t *src2, ptrdiff_t stride2);
uint64_t xvcb200_ssd_short_short(int width, int height, const int16_t *src1, ptrdiff_t stride1,
                                 const int16_t *src2, ptrdiff_t stride2);

/* SampleMetric::Compare (sample_metric.cc:171-298): metric selected like MetricType
 * (sample_metric.h:34-44), result before the chroma weight (weight applied by caller). */
enum { XVCB200_METRIC_SSD = 0, XVCB200_METRIC_SATD = 1, XVCB200_METRIC_SAD = 2,
       XVCB200_METRIC_SAD_FAST = 3 };
uint64_t xvcb200_compare_sample_sample(int metric, int bitdepth, int width, int height,
                                       const uint16_t *src1, ptrdiff_t stride1,
                                       const uint16_t *src2, ptrdiff_t stride2);
uint64_t xvcb200_compare_short_sample(int metric, int bitdepth, int width, int height,
                                      const int16_t *src1, ptrdiff_t stride1,
                                      const uint16_t *src2, ptrdiff_t stride2);

/* InterPrediction::SimdFunc (xvc_common_lib/inter_prediction.h:176-216); C defaults
 * inter_prediction.cc:1207-1385 (filters), :1462 (copy bipred), :1675 (add_avg).
 * `src` points at the centre sample; the callee backs up taps/2-1 (cc:1215,1275).
 * chroma = 0: 8-tap luma, 1: 4-tap chroma (table index kLC). */
void xvcb200_filter_h_sample_sample(int chroma, int width, int height, int bitdepth,
                                    const int16_t *filter, const uint16_t *src, ptrdiff_t src_stride,
                                    uint16_t *dst, ptrdiff_t dst_stride);
void xvcb200_filter_h_sample_short(int chroma, int width, int height, int bitdepth,
                                   const int16_t *filter, const uint16_t *src, ptrdiff_t src_stride,
                                   int16_t *dst, ptrdiff_t dst_stride);
void xvcb200_filter_v_sample_sample(int chroma, int width, int height, int bitdepth,
                                    const int16_t *filter, const uint16_t *src, ptrdiff_t src_stride,
                                    uint16_t *dst, ptrdiff_t dst_stride);
void xvcb200_filter_v_sample_short(int chroma, int width, int height, int bitdepth,
                                   const int16_t *filter, const uint16_t *src, ptrdiff_t src_stride,
                                   int16_t *dst, ptrdiff_t dst_stride);
void xvcb200_filter_v_short_sample(int chroma, int width, int height, int bitdepth,
                                   const int16_t *filter, const int16_t *src, ptrdiff_t src_stride,
                                   uint16_t *dst, ptrdiff_t dst_stride);
void xvcb200_filter_v_short_short(int chroma, int width, int height, int bitdepth,
                                  const int16_t *filter, const int16_t *src, ptrdiff_t src_stride,
                                  int16_t *dst, ptrdiff_t dst_stride);
void xvcb200_add_avg(int width, int height, int offset, int shift, int bitdepth,
                     const int16_t *src1, intptr_t stride1, const int16_t *src2, intptr_t stride2,
                     uint16_t *dst, intptr_t dst_stride);
void xvcb200_filter_copy_bipred(int width, int height, int16_t offset, int shift,
                                const uint16_t *ref, ptrdiff_t ref_stride,
                                int16_t *pred, ptrdiff_t pred_stride);

/* InterPrediction::FilterLuma / FilterChroma (inter_prediction.cc:1387-1448) and the
 * bi-pred variants (:1475-1538): fractional position in 1/16 (luma) or 1/32 (chroma). */
void xvcb200_interp_block(int chroma, int width, int height, int bitdepth, int frac_x, int frac_y,
                          const uint16_t *ref, ptrdiff_t ref_stride, uint16_t *pred, ptrdiff_t pred_stride);
void xvcb200_interp_block_bipred(int chroma, int width, int height, int bitdepth, int frac_x, int frac_y,
                                 const uint16_t *ref, ptrdiff_t ref_stride, int16_t *pred, ptrdiff_t pred_stride);

/* The reference's table structs, same member order, so a maintainer can memcpy/assign
 * (see INTEGRATION.md).  Indices: [0] width<=2 / luma, [1] width>=4 / chroma. */
typedef struct {
  void (*add_avg[2])(int, int, int, int, int, const int16_t *, intptr_t, const int16_t *, intptr_t,
                     uint16_t *, intptr_t);
  void (*filter_copy_bipred[2])(int, int, int16_t, int, const uint16_t *, ptrdiff_t, int16_t *, ptrdiff_t);
  void (*filter_h_sample_sample[2])(int, int, int, const int16_t *, const uint16_t *, ptrdiff_t, uint16_t *, ptrdiff_t);
  void (*filter_h_sample_short[2])(int, int, int, const int16_t *, const uint16_t *, ptrdiff_t, int16_t *, ptrdiff_t);
  void (*filter_v_sample_sample[2])(int, int, int, const int16_t *, const uint16_t *, ptrdiff_t, uint16_t *, ptrdiff_t);
  void (*filter_v_sample_short[2])(int, int, int, const int16_t *, const uint16_t *, ptrdiff_t, int16_t *, ptrdiff_t);
  void (*filter_v_short_sample[2])(int, int, int, const int16_t *, const int16_t *, ptrdiff_t, uint16_t *, ptrdiff_t);
  void (*filter_v_short_short[2])(int, int, int, const int16_t *, const int16_t *, ptrdiff_t, int16_t *, ptrdiff_t);
} xvcb200_inter_prediction_simd_func;   /* == InterPrediction::SimdFunc, inter_prediction.h:176-216 */

typedef struct {
  int (*sad_sample_sample[7])(int, int, const uint16_t *, ptrdiff_t, const uint16_t *, ptrdiff_t);
  int (*sad_short_sample[7])(int, int, const int16_t *, ptrdiff_t, const uint16_t *, ptrdiff_t);
  uint64_t (*ssd_sample_sample[7])(int, int, const uint16_t *, ptrdiff_t, const uint16_t *, ptrdiff_t);
  uint64_t (*ssd_short_sample[7])(int, int, const int16_t *, ptrdiff_t, const uint16_t *, ptrdiff_t);
  uint64_t (*ssd_short_short[7])(int, int, const int16_t *, ptrdiff_t, const int16_t *, ptrdiff_t);
} xvcb200_sample_metric_simd_func;      /* == SampleMetric::SimdFunc, sample_metric.h:166-191 */

/* simd::InterPredictionSimd::Register / simd::SampleMetricSimd::Register equivalents
 * (simd/inter_prediction_simd.h:36-39, simd/sample_metric_simd.h:36-40). */
void xvcb200_register_inter_prediction(xvcb200_inter_prediction_simd_func *table);
void xvcb200_register_sample_metric(int bitdepth, xvcb200_sample_metric_simd_func *table);

/* Transform types, numbering of TransformType (xvc_common_lib/cu_types.h). */
enum { XVCB200_TX_DEFAULT = 0, XVCB200_TX_DCT2 = 1, XVCB200_TX_DCT5 = 2, XVCB200_TX_DCT8 = 3,
       XVCB200_TX_DST1 = 4, XVCB200_TX_DST7 = 5 };

/* ForwardTransform::Transform (transform.cc:869-961) / TransformSkip (:963-995).
 * tx_hor/tx_ver = cu.GetTransformType(comp,1)/(comp,0); dst4x4 = can_dst_4x4 (:874-876). */
void xvcb200_fwd_transform(int width, int height, int bitdepth, int tx_hor, int tx_ver, int dst4x4,
                           const int16_t *resi, ptrdiff_t resi_stride, int16_t *coeff, ptrdiff_t coeff_stride);
void xvcb200_fwd_transform_skip(int width, int height, int bitdepth,
                                const int16_t *resi, ptrdiff_t resi_stride, int16_t *coeff, ptrdiff_t coeff_stride);
/* InverseTransform::Transform (transform.cc:83-182), TransformSkip (:184-215). */
void xvcb200_inv_transform(int width, int height, int bitdepth, int tx_hor, int tx_ver, int dst4x4, int dc_only,
                           const int16_t *coeff, ptrdiff_t coeff_stride, int16_t *resi, ptrdiff_t resi_stride);
void xvcb200_inv_transform_skip(int width, int height, int bitdepth,
                                const int16_t *coeff, ptrdiff_t coeff_stride, int16_t *resi, ptrdiff_t resi_stride);

/* Qp (quantize.cc:48-92): per-component derived values. */
typedef struct {
  int32_t qp_raw[3];       /* after chroma table/offset */
  int32_t qp_bitdepth[3];  /* qp_raw + 6*(bitdepth-8), >= 0 */
  double distortion_weight[3];
  double lambda[3];
  double lambda_sqrt;
} xvcb200_qp;
void xvcb200_qp_init(xvcb200_qp *out, int qp, int chroma_format /*0:400 1:420 2:422 3:444*/, int bitdepth,
                     double lambda, int chroma_offset_table, int chroma_offset_u, int chroma_offset_v);

/* RdoQuant::QuantFast (rdo_quant.cc:156-201) incl. CoeffSignHideFast (:448-573).
 * qp_bitdepth = Qp::qp_bitdepth_[comp]; scan_order 0 diagonal, 1 horizontal, 2 vertical
 * (TransformHelper::DetermineScanOrder, transform.cc:1614-1636).  Returns num_non_zero. */
int xvcb200_quant_fast(int width, int height, int bitdepth, int qp_bitdepth, int intra_picture,
                       int sign_hiding, int scan_order,
                       const int16_t *in, ptrdiff_t in_stride, int16_t *out, ptrdiff_t out_stride);
/* Quantize::Inverse (quantize.cc:94-125). */
void xvcb200_dequant(int width, int height, int bitdepth, int qp_bitdepth,
                     const int16_t *in, ptrdiff_t in_stride, int16_t *out, ptrdiff_t out_stride);

/* ------------------------------------------------------------------------------------
 * (B) batched, device-resident
 * ---------------------------------------------------------------------------------- */

typedef struct xvcb200_ctx xvcb200_ctx;

/* Geometry of a device picture slot.  Planes are padded like YuvPicture(padding=true)
 * (yuv_pic.cc:32-68: 80 luma / 40 chroma samples on every side are addressable), with a
 * 128-byte aligned pitch and x=0 on a 128-byte boundary. */
typedef struct {
  int32_t width[3], height[3];
  int32_t pitch[3];        /* elements */
  int32_t margin_x[3];     /* allocated columns left of x=0 (>= 80 / 40) */
  int32_t margin_y[3];     /* allocated rows above y=0 (80 / 40) */
} xvcb200_plane_geom;

/* One coding unit (leaf of the CU tree).  What CodingUnit carries for this path
 * (coding_unit.h:62-74, cu_types.h). */
typedef struct {
  int16_t x, y;            /* luma position */
  uint8_t w, h;            /* luma size, 4..64 */
  uint8_t depth;           /* cu.GetDepth() (quad depth); 0 disables prev-MV start (inter_tz_search.cc:115) */
  uint8_t flags;           /* XVCB200_CU_* */
  int8_t qp;               /* raw luma QP of the CU (cu.GetQp(kY)) */
  int8_t ref_idx[2];       /* -1: list unused */
  uint8_t tx_select;       /* reserved (0 = DCT-2 both directions) */
  int32_t mv[2][2];        /* [list][x,y], 1/16 pel */
} xvcb200_cu;
enum { XVCB200_CU_INTRA = 1, XVCB200_CU_FULLPEL_MV = 2, XVCB200_CU_CBF_Y = 4, XVCB200_CU_CBF_U = 8,
       XVCB200_CU_CBF_V = 16, XVCB200_CU_SKIP_ME = 32 };

/* Motion estimation job: InterSearch::MotionEstNormal (inter_search.cc:606-662) for one
 * (CU, reference picture): TzSearch::Search (inter_tz_search.cc:84-171) then SubpelSearch
 * (inter_search.cc:893-949). */
typedef struct {
  int32_t cu;              /* index into the CU array */
  int32_t ref_slot;        /* picture slot of the (padded) reference picture */
  int32_t search_range;    /* InterSearch::GetSearchRangeUniPred (inter_search.cc:1050-1057) */
  int32_t mvp[2];          /* predictor, 1/16 pel */
  int32_t prev[2];         /* previous full-pel search result (previous_fullpel_, inter_search.cc:636) */
  int32_t list;            /* 0/1: which ref list this job belongs to (for the picture pipeline) */
} xvcb200_me_job;

typedef struct {
  int32_t mv_fullpel[2];   /* TzSearch::Search result */
  int32_t mv[2];           /* after SubpelSearch, 1/16 pel */
  uint32_t cost_fullpel;   /* state.cost_best */
  uint32_t dist;           /* *out_dist (SATD of the best sub-pel candidate) */
  uint32_t cost;           /* best_cost of the sub-pel search */
  uint32_t num_sad;        /* candidates evaluated (statistics; not in the reference) */
} xvcb200_me_result;

/* InterSearch::FullSearch (inter_search.cc:853-891): +-range around the clipped window on
 * the weighted original 2*orig - pred_other (sample_buffer.h:147-161). */
typedef struct {
  int32_t cu, ref_slot, other_pred_slot;
  int32_t mvp[2];
  int32_t center[2];       /* mv_bootstrap / mvp used for DetermineMinMaxMv, 1/16 pel */
  int32_t range;
} xvcb200_fullsearch_job;

/* One transform unit = one component of one CU: TransformEncoder::TransformAndReconstruct
 * (transform_encoder.cc:203-285) with QuantFast.  Output levels are written to a
 * picture-shaped int16 plane set ("coefficient slot"). */
typedef struct {
  uint32_t ssd;            /* Sum d^2 >> 2(bd-8), before chroma weight */
  int32_t num_non_zero;
} xvcb200_tu_result;

/* Per-4x4 metadata for deblocking (what GetBoundaryStrength reads through
 * PictureData::GetCuAt, deblocking_filter.cc:154-241).  Built on device from the CU array. */

int xvcb200_ctx_create(xvcb200_ctx **out, int device, int width, int height, int bitdepth,
                       int chroma_format, int num_slots);
void xvcb200_ctx_destroy(xvcb200_ctx *ctx);
/* run on a caller-owned stream (cudaStream_t as void*); 0 = the context's own stream */
int xvcb200_ctx_set_stream(xvcb200_ctx *ctx, void *cuda_stream);
void *xvcb200_stream(xvcb200_ctx *ctx);              /* the cudaStream_t work is enqueued on */
int xvcb200_sync(xvcb200_ctx *ctx);                 /* waits, returns sticky status */
const char *xvcb200_ctx_error_string(xvcb200_ctx *ctx);
int xvcb200_get_geometry(xvcb200_ctx *ctx, xvcb200_plane_geom *geom);
/* device address of sample (0,0) of a plane of a slot (uint16_t*); coefficient slots: int16_t*.
 * Raw pointers (here and xvcb200_slot_region) are ordered with the library's work only through the
 * context stream: use them on xvcb200_stream(ctx), or after xvcb200_sync(ctx). */
int xvcb200_slot_ptr(xvcb200_ctx *ctx, int slot, int comp, void **dev_ptr);

/* the whole device allocation of a slot (three padded planes).  All slots of a context live
 * back to back in one arena, so slots [s, s+n) form one contiguous buffer of n * bytes --
 * e.g. the destination of an NCCL all-gather of reconstructed pictures. */
int xvcb200_slot_region(xvcb200_ctx *ctx, int slot, void **base, uint64_t *bytes);

/* Peer exchange of slots between the contexts of different PROCESSES on one NVLink / NVSwitch
 * node (frame-parallel encoding: a finished, padded reconstruction goes to every GPU that will
 * reference it -- thread_encoder.cc:99-131 has one address space, this is its multi-GPU form).
 * Copy engines move the slot over NVLink; no SM takes part, so the exchange overlaps the next
 * picture's kernels completely (an NCCL all-gather needs SMs, which the persistent search kernel
 * occupies).  handle: XVCB200_IPC_HANDLE_BYTES = the cudaIpcMemHandle_t of the slot arena followed by
 * the arena layout (slot stride, slot count), exchanged out of band; xvcb200_ipc_open_peer returns
 * XVCB200_INVALID_ARGUMENT for a peer whose layout differs from this context's (pushes address the
 * peer by slot * stride).
 * xvcb200_push_slot copies slot `slot` of this context into the same slot of every opened peer,
 * ordered after the work enqueued on the context stream so far; xvcb200_wait_pushes makes the
 * context stream wait for the last push of `slot` (slot < 0: of every slot) -- call it before the
 * slot is overwritten.  Arrival at the consumer is the caller's rendezvous (after
 * xvcb200_wait_pushes + xvcb200_sync on every rank the pushed slots are complete everywhere). */
#define XVCB200_IPC_HANDLE_BYTES 80
int xvcb200_ipc_export(xvcb200_ctx *ctx, void *handle /* XVCB200_IPC_HANDLE_BYTES */);
int xvcb200_ipc_open_peer(xvcb200_ctx *ctx, const void *handle, int *peer_index);
int xvcb200_push_slot(xvcb200_ctx *ctx, int slot);
int xvcb200_wait_pushes(xvcb200_ctx *ctx, int slot);
/* The rendezvous on the device (no host takes part): xvcb200_push_slot_tagged pushes the slot like
 * xvcb200_push_slot and then, on the same copy stream per peer, writes `tag` into the arrival word of that slot in
 * the PEER's arena (one 32-bit word per slot behind the slots, part of the exported allocation);
 * xvcb200_wait_slot_tag makes the consumer's context stream wait (cuStreamWaitValue32, >=) until the arrival word
 * of `slot` in its OWN arena has reached `tag` -- the kernels enqueued afterwards see the pushed content.  Tags of
 * one slot must increase from push to push (a picture counter); the producer of a slot does not wait for its own
 * tag (its content is ordered by its stream).  What ThreadEncoder's "picture done" notification is between
 * threads (thread_encoder.cc:133-159), between GPUs. */
int xvcb200_push_slot_tagged(xvcb200_ctx *ctx, int slot, uint32_t tag);
int xvcb200_wait_slot_tag(xvcb200_ctx *ctx, int slot, uint32_t tag);

/* host <-> device picture transfer; host planes are tight or strided (elements). */
int xvcb200_upload_picture(xvcb200_ctx *ctx, int slot, const uint16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_download_picture(xvcb200_ctx *ctx, int slot, uint16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_download_coeff(xvcb200_ctx *ctx, int slot, int16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_upload_coeff(xvcb200_ctx *ctx, int slot, const int16_t *const planes[3], const ptrdiff_t strides[3]);
/* The same transfers on a second (copy) stream owned by the context, so that they overlap the
 * kernels of another picture.  Ordering is kept by events inside the library: an upload starts
 * once the work enqueued BEFORE the call is done and is awaited by work enqueued after it; a
 * download starts once the work enqueued before the call is done, and later work that writes
 * the same slot (or the CU array) waits for it.  To overlap, alternate between two slots and
 * enqueue the upload of picture n+1 before xvcb200_encode_picture of picture n.  Host buffers
 * should be page-locked.  xvcb200_sync_copies() waits for the copy stream, xvcb200_wait_download()
 * for the transfers of one slot (the copy stream runs them in call order). */
int xvcb200_upload_picture_async(xvcb200_ctx *ctx, int slot, const uint16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_download_picture_async(xvcb200_ctx *ctx, int slot, uint16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_download_coeff_async(xvcb200_ctx *ctx, int slot, int16_t *const planes[3], const ptrdiff_t strides[3]);
int xvcb200_get_cus_async(xvcb200_ctx *ctx, xvcb200_cu *cus, int n);
int xvcb200_wait_download(xvcb200_ctx *ctx, int slot);   /* host waits for the last async download of a slot; slot < 0: the CU array */
int xvcb200_sync_copies(xvcb200_ctx *ctx);
/* one plane incl. its 80 / 40 sample border, tight (h + 2*pad) x (w + 2*pad) */
int xvcb200_download_padded(xvcb200_ctx *ctx, int slot, int comp, uint16_t *dst);
/* YuvPicture::PadBorder (yuv_pic.cc:118-150) */
int xvcb200_pad_border(xvcb200_ctx *ctx, int slot);

/* CU array upload (host -> device list owned by the context). */
int xvcb200_set_cus(xvcb200_ctx *ctx, const xvcb200_cu *cus, int n);
int xvcb200_get_cus(xvcb200_ctx *ctx, xvcb200_cu *cus, int n);

/* lambda_sqrt = Qp::GetLambdaSqrt(); lambda_me = floor(65536*lambda_sqrt) is derived inside. */
int xvcb200_me_search(xvcb200_ctx *ctx, int orig_slot, const xvcb200_me_job *jobs, int n,
                      double lambda_sqrt, xvcb200_me_result *results);
int xvcb200_full_search(xvcb200_ctx *ctx, int orig_slot, const xvcb200_fullsearch_job *jobs, int n,
                        double lambda_sqrt, xvcb200_me_result *results);

/* InterPrediction::MotionCompensation (inter_prediction.cc:710-738) for every CU in the
 * context's CU array, all three components, into pred_slot.  ref_slots[list][ref_idx]. */
int xvcb200_motion_compensate(xvcb200_ctx *ctx, const int32_t ref_slots[2][5], int pred_slot);

/* Affine motion compensation (InterPrediction::MotionCompAffine,
 * inter_prediction.cc:1044-1136, reached from MotionCompensation through MotionCompRefList's
 * GetUseAffine branch :1021-1023 and from the encoder's affine search through
 * MotionCompensationMv(cu, comp, ref_pic, MotionVector3, ...) :761-767).  For every entry: the CU
 * `cu` of the context's CU array is predicted from its reference picture(s) (cu.ref_idx, as in
 * xvcb200_motion_compensate; cu.mv is ignored) with the 4-parameter model given by the control-point
 * MVs mv[list][corner: 0 top-left, 1 top-right, 2 bottom-left][x, y] in 1/16 pel (CodingUnit::
 * GetMvAffine, coding_unit.h:258-264), all three components, into pred_slot. */
typedef struct {
  int32_t cu;
  int32_t mv[2][3][2];
} xvcb200_affine_cu;
int xvcb200_motion_compensate_affine(xvcb200_ctx *ctx, const xvcb200_affine_cu *aff, int n,
                                     const int32_t ref_slots[2][5], int pred_slot);

/* Motion compensation with local illumination compensation (SURVEY 8f rank 4;
 * InterPrediction::MotionCompensation for CUs with GetUseLic(), inter_prediction.cc:710-738 ->
 * MotionCompRefList :1031-1040 -> LocalIlluminationComp / DeriveLicParams :1555-1673).  For every
 * entry the CU `cu` of the context's CU array (cu.ref_idx / cu.mv as in xvcb200_motion_compensate)
 * is predicted per list, a linear model pred' = clip(((scale * pred) >> 5) + offset) is fitted on
 * the row above / column left of the block -- reference picture at the rounded full-pel MV against
 * the CURRENT picture's reconstruction in rec_slot -- and applied; bi-prediction averages the two
 * compensated predictions through FilterCopyBipred + AddAvgBi ("intermediate rounding", :724-729).
 * above_x/above_y, left_x/left_y: luma position of the CU covering (x, y-4) / (x-4, y)
 * (CodingUnit::GetCodingUnitAbove / Left, coding_unit.cc:227-234, 275-282; its position enters the
 * model through ClipMv :1604, 1618), -1 when there is none.  The neighbours' reconstruction must
 * be final in rec_slot: the caller owns coding order (a CTU anti-diagonal at a time, as for intra). */
typedef struct {
  int32_t cu;
  int16_t above_x, above_y, left_x, left_y;
} xvcb200_lic_cu;
int xvcb200_motion_compensate_lic(xvcb200_ctx *ctx, const xvcb200_lic_cu *lic, int n, const int32_t ref_slots[2][5],
                                  int rec_slot, int pred_slot);

/* TransformAndReconstruct for every CU x component: residual = orig - pred, forward
 * transform, QuantFast, dequant, inverse transform, AddClip into rec_slot; levels into
 * coeff_slot; per-TU results[3*n_cus] (order: cu-major, component-minor); cbf flags are
 * written back into the device CU array. */
int xvcb200_tq_reconstruct(xvcb200_ctx *ctx, int orig_slot, int pred_slot, int rec_slot, int coeff_slot,
                           int pic_qp_unused, int intra_picture, int chroma_offset_table,
                           int chroma_offset_u, int chroma_offset_v, xvcb200_tu_result *results);
/* decoder side: CuDecoder::DecompressComponent (cu_decoder.cc:102-138) from levels */
int xvcb200_dequant_reconstruct(xvcb200_ctx *ctx, int pred_slot, int rec_slot, int coeff_slot,
                                int chroma_offset_table, int chroma_offset_u, int chroma_offset_v);

/* DeblockingFilter::DeblockPicture (deblocking_filter.cc:56-77) on rec_slot in place,
 * using the context's CU array.  pic_type: 0 bi, 1 uni (PicturePredictionType).
 * ref_poc[list][ref_idx]: POC of each reference (CodingUnit::GetRefPoc). */
int xvcb200_deblock_picture(xvcb200_ctx *ctx, int rec_slot, int pic_type, int beta_offset, int tc_offset,
                            const int64_t ref_poc[2][5]);
/* same with the segment's chroma QP mapping (SegmentHeader chroma_qp_offset_table / _u / _v); the
 * short form uses table 1, offsets 0 (encoder defaults, encoder_settings.h:93-95) */
int xvcb200_deblock_picture_ex(xvcb200_ctx *ctx, int rec_slot, int pic_type, int beta_offset, int tc_offset,
                               int chroma_offset_table, int chroma_offset_u, int chroma_offset_v,
                               const int64_t ref_poc[2][5]);

/* The two cases the CU array alone does not describe (deblocking_filter.cc:56-77, 88-91, 166-176):
 *   affine / n_affine            CUs coded with affine motion (entries as for xvcb200_motion_compensate_affine):
 *                                the boundary strength compares the vectors at the CU CORNERS nearest to the
 *                                edge segment (up-left / up-right / down-left control points, down-right =
 *                                up-right + down-left - up-left; CodingUnit::SetMv(MotionVector3), coding_unit.h:251-257)
 *   chroma_cus / n_chroma_cus    the SECONDARY CU tree of an intra picture (PictureData::HasSecondaryCuTree): leaf CUs
 *                                of the chroma tree in luma units (x, y, w, h, qp, flags |= XVCB200_CU_INTRA); chroma
 *                                edges are then taken from this tree on an 8-sample luma grid (four chroma lines
 *                                per segment) and the primary tree filters luma only
 * Either pointer may be NULL (count 0).  HOST arrays, copied before the call returns. */
typedef struct {
  const xvcb200_affine_cu *affine;
  int32_t n_affine;
  const xvcb200_cu *chroma_cus;
  int32_t n_chroma_cus;
} xvcb200_deblock_ext;
int xvcb200_deblock_picture_ext(xvcb200_ctx *ctx, int rec_slot, int pic_type, int beta_offset, int tc_offset,
                                int chroma_offset_table, int chroma_offset_u, int chroma_offset_v,
                                const int64_t ref_poc[2][5], const xvcb200_deblock_ext *ext);

/* One band of the picture for CTB-row sharding across GPUs.  pass_mask: 1 = vertical edges of
 * rows [y_begin, y_end), 2 = horizontal edges whose q side lies in [y_begin, y_end) (the edge
 * ON y_begin belongs to this band and reads/writes up to 4/3 rows above it: the caller
 * exchanges those halo rows, see xvc_b200/sharding.py).  The context's CU array must describe
 * the whole picture. */
int xvcb200_deblock_band(xvcb200_ctx *ctx, int rec_slot, int pic_type, int beta_offset, int tc_offset,
                         int chroma_offset_table, int chroma_offset_u, int chroma_offset_v,
                         const int64_t ref_poc[2][5], int pass_mask, int y_begin, int y_end);

/* ------------------------------------------------------------------------------------
 * (B') intra prediction -- SURVEY section 8(f) rank 3, the first component after the hot path.
 * Unrestricted configuration (67 modes: 0 planar, 1 DC, 2..66 angular; 18 horizontal,
 * 50 vertical).  LM chroma and the MPM derivation stay with the caller.
 * ---------------------------------------------------------------------------------- */
#define XVCB200_INTRA_REF_STRIDE 129   /* IntraPrediction::kRefSampleStride_ = 2 * 64 + 1, intra_prediction.h:40 */
#define XVCB200_INTRA_NUM_MODES 67     /* kNbrIntraModesExt, cu_types.h:86 */

/* IntraPrediction::FillReferenceState (intra_prediction.cc:128-147) with the neighbour
 * availability (DetermineNeighbors :688-707) given explicitly: ComputeRefSamples (:709-851)
 * into ref_samples and, if ref_filtered != NULL, FilterRefSamples (:853-876) into ref_filtered.
 * Both arrays: 2 x XVCB200_INTRA_REF_STRIDE samples laid out like IntraPrediction::RefState
 * ([0] above-left, [1 .. w+h] above, [stride + y] left; unused entries are written as 0).
 * `block` = HOST pointer to the block's top-left sample inside the reconstructed plane; only
 * neighbours declared available are read.  above_right / below_left = number of available
 * samples beyond the block (CodingUnit::GetCuSizeAboveRight / GetCuSizeBelowLeft), 0 = none. */
void xvcb200_intra_ref_samples(int width, int height, int bitdepth, int has_above_left, int has_above, int above_right,
                               int has_left, int below_left, const uint16_t *block, ptrdiff_t stride,
                               uint16_t *ref_samples, uint16_t *ref_filtered);
/* IntraPrediction::Predict (intra_prediction.cc:81-126) for planar / DC / angular modes.
 * is_luma selects reference smoothing (UseFilteredRefSamples :342-364; width/height are then
 * the CU's luma size) and the edge filters of blocks up to 16x16.  Host pointers. */
void xvcb200_intra_predict(int mode, int width, int height, int bitdepth, int is_luma, const uint16_t *ref_samples,
                           const uint16_t *ref_filtered, uint16_t *pred, ptrdiff_t stride);

typedef struct {
  int32_t x, y;                      /* luma position of the block */
  uint8_t w, h;
  uint8_t has_above_left, has_above, has_left;
  uint8_t above_right, below_left;   /* available samples beyond the block */
  uint8_t reserved;
} xvcb200_intra_job;
/* First loop of IntraSearch::DetermineSlowIntraModes (xvc_enc_lib/intra_search.cc:185-216) for
 * n luma blocks at once: for every mode the SATD (SampleMetric kSatd incl. the bit-depth shift)
 * of its prediction against the original block in orig_slot.  Reference samples are taken from
 * the luma plane of src_slot (the reconstruction of the blocks coded so far, or any picture for
 * a pre-analysis).  satd: HOST array n x XVCB200_INTRA_NUM_MODES; the mode bits and the sort
 * stay with the caller (they depend on the entropy coder's state). */
int xvcb200_intra_satd_scan(xvcb200_ctx *ctx, int orig_slot, int src_slot, const xvcb200_intra_job *jobs, int n,
                            uint32_t *satd);

/* Chroma from luma: IntraPrediction::Predict(IntraMode::kLmChroma) -> PredLmChroma / RescaleLuma /
 * DeriveLmParams (intra_prediction.cc:113-115, 560-686, 873-913) for n CUs at once, 4:2:0: both chroma
 * blocks of every job (x, y, w, h = the CU's LUMA position and size; the neighbour flags are not used --
 * the reference looks at the position only) predicted from rec_slot -- the CU's reconstructed luma
 * and the luma / chroma samples above and left of it, which must be final -- into the U and V planes
 * of pred_slot.  Asynchronous on the context stream. */
int xvcb200_intra_lm_chroma(xvcb200_ctx *ctx, int rec_slot, const xvcb200_intra_job *jobs, int n, int pred_slot);

/* ------------------------------------------------------------------------------------
 * (C) picture-level hot path
 * ---------------------------------------------------------------------------------- */
typedef struct {
  int32_t orig_slot, pred_slot, rec_slot, coeff_slot;
  int32_t ref_slots[2][5];
  int64_t ref_poc[2][5];
  int32_t num_ref[2];
  int32_t pic_type;                  /* 0 bi, 1 uni */
  int32_t search_range[2][5];
  double lambda_sqrt;
  int32_t chroma_offset_table, chroma_offset_u, chroma_offset_v;
  int32_t beta_offset, tc_offset;
  int32_t deblock, pad;
  int32_t bi_iterations;             /* EncoderSettings::bipred_refinement_iterations (encoder_settings.h:70); 0: no bi-prediction */
  int32_t bits_mode;                 /* 0: the cost of the sub-pel search picks the list (round-1 rule, one picture per list);
                                        1: InterSearch::GetInterPredBits with fast_inter_pred_bits (inter_search.cc:1084-1130) */
} xvcb200_picture_params;

/* InterSearch::SearchMotion (inter_search.cc:199-259) for every CU at once -> MC -> T/Q/recon ->
 * deblock -> pad.  Per list l the num_ref[l] pictures ref_slots[l][0..] are searched (TZ + sub-pel,
 * search_range[l][r]); a list-1 picture whose POC equals a list-0 picture's is not searched again
 * (same_poc_in_l0_mapping_, :536-543).  With bi_iterations > 0 (bi-predicted pictures) the
 * SearchBiIterative passes follow (:392-433: FullSearch +-4 and sub-pel search on the weighted
 * original 2 * orig - pred(other list)).  The predictor of list l is the CU's mv[l] as uploaded
 * (one predictor per list); CUs flagged INTRA or SKIP_ME are left as they are.  Results: the
 * chosen ref_idx / mv written back into the CU array, me_results[n * (num_ref[0] + num_ref[1])]
 * laid out [cu][list 0 pictures, list 1 pictures] (uni searches; may be NULL), tu_results[3*n]
 * (may be NULL).  num_ref[l] <= 0 reads as one picture (list 1: none for pic_type 1). */
/* Optional: transform modes per CU for the residual-coding entries below (what CodingUnit carries in tx_:
 * coding_unit.h:173-191).  Without them every unit is coded with the default transform -- DCT-2, or the 4 x 4
 * DST for the luma block of an intra CU (transform.cc:87-89, 873-875) --, no transform skip, diagonal scan.
 *   tx_ver / tx_hor  luma transform types XVCB200_TX_* (cu.GetTransformType(kY, 0 / 1)): what
 *                    CodingUnit::SetTransformFromSelectIdx (coding_unit.cc:359-424) derived; chroma is DCT-2
 *   tskip            bit c: transform skip of component c (blocks of <= 16 samples, coding_unit.h:202-204)
 *   scan             per component TransformHelper::DetermineScanOrder (transform.cc:1614-1636):
 *                    0 diagonal, 1 horizontal, 2 vertical (sign hiding walks the block in this order)
 * The choice between the candidates (TransformEncoder::CompressAndEvalTransform) compares entropy-coded
 * bit counts and stays with the caller.  modes: HOST array [n_cus]; belongs to the current CU array
 * (xvcb200_set_cus drops it), NULL drops it explicitly. */
typedef struct {
  uint8_t tx_ver, tx_hor;
  uint8_t tskip;
  uint8_t scan[3];
  uint8_t reserved[2];
} xvcb200_tu_mode;
int xvcb200_set_tu_modes(xvcb200_ctx *ctx, const xvcb200_tu_mode *modes);

/* Optional: one predictor per (CU, list, reference picture) instead of the CU's mv[list] -- what
 * InterSearch::GetMvpList (inter_search.cc:489-492: per ref_idx, neighbour vectors scaled by POC distance)
 * gives the reference's search.  mvp: HOST array [n_cus][n_cols][2], 1/16 pel, columns ordered like
 * me_results (list 0 pictures, then list 1 pictures); n_cols must equal num_ref[0] + num_ref[1] of the
 * xvcb200_encode_picture calls that follow.  Belongs to the current CU array: xvcb200_set_cus drops it;
 * n_cols = 0 drops it explicitly. */
int xvcb200_set_mv_predictors(xvcb200_ctx *ctx, const int32_t *mvp, int n_cols);

int xvcb200_encode_picture(xvcb200_ctx *ctx, const xvcb200_picture_params *params,
                           xvcb200_me_result *me_results, xvcb200_tu_result *tu_results);

/* GPU pre-analysis: the CU partition of an inter picture.  The reference decides the
 * partition inside CuEncoder::CompressCu's serial RD recursion (cu_encoder.cc:123-273); this entry decides it
 * from motion-compensated distortion for all CTUs at once: SAD of every 8 x 8 block at every full-pel vector
 * within +-8 of `center` on the luma of ref_slot, summed bottom-up over the quad tree ("SAD tree"), and per
 * node the cheapest of: one CU, two horizontal / vertical halves (each with its own vector), four quadrants
 * -- cost = SAD + ((lambda * bits) >> 16) with exp-Golomb vector bits against `center` plus header_bits_cu per
 * CU and header_bits_split per split decision (0: 8 / 1).  Output: the leaf CUs in coding order (CTU raster,
 * quadrants in z order; depth = quad depth, qp, mv[0] = mv[1] = the winning vector in 1/16 pel: the
 * predictor for the search of xvcb200_encode_picture) and the split flags of all CTUs in raster order,
 * pre-order, one byte per node: 0 leaf, 1 quad, 2 horizontal, 3 vertical (SplitType, cu_types.h:37-42; the
 * two children of a binary split are leaves; splits forced by the picture edge are implicit and not listed).
 * The trees are ones xvc's syntax can carry.  Synchronous; HOST output arrays with capacities (a 64 x 64 CTU
 * yields at most 64 CUs and 128 flags).  tests/partition_model.py states the same rule in numpy. */
typedef struct {
  int32_t orig_slot, ref_slot;
  int32_t center[2];                 /* picture-level predictor, 1/16 pel (rounded down to full samples) */
  double lambda_sqrt;
  int32_t qp;                        /* written into every CU */
  int32_t header_bits_cu, header_bits_split;
} xvcb200_partition_params;
int xvcb200_decide_partition(xvcb200_ctx *ctx, const xvcb200_partition_params *params, xvcb200_cu *cus_out, int cus_cap,
                             int *n_cus, uint8_t *splits_out, int splits_cap, int *n_splits);
/* The same in two halves: _begin enqueues the kernel and the copy of its result on the context stream and returns;
 * _end waits for that copy only and compacts the result.  Work enqueued between the two (the kernels of the picture
 * before) runs while the host processes the partition.  One pre-analysis in flight per context. */
int xvcb200_decide_partition_begin(xvcb200_ctx *ctx, const xvcb200_partition_params *params);
int xvcb200_decide_partition_end(xvcb200_ctx *ctx, xvcb200_cu *cus_out, int cus_cap, int *n_cus, uint8_t *splits_out,
                                 int splits_cap, int *n_splits);

/* Optional per-stage device timing of xvcb200_encode_picture (CUDA events on the context
 * stream): ms[0..6] = job set-up, full-pel TZ search, sub-pel search + list decision, motion
 * compensation, T/Q/recon, deblocking, padding -- of the last call. */
int xvcb200_set_profiling(xvcb200_ctx *ctx, int enable);
int xvcb200_get_stage_times(xvcb200_ctx *ctx, float ms[7]);

#ifdef __cplusplus
}
#endif
#endif  /* XVC_B200_H_ */
