// Internal declarations shared by the kernels and the C-ABI layer of libxvc_b200.so.
#ifndef XVCB_INTERNAL_H_
#define XVCB_INTERNAL_H_

#include <cuda.h>             // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no libcuda link)
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/xvc_b200.h"

namespace xvcb {

typedef uint16_t Sample;

extern std::atomic<uint64_t> g_launch_count;
void set_last_error(int code, const char *what);

// One device picture: three padded planes in a single allocation.
struct DevPicture {
  uint8_t *alloc = nullptr;
  size_t bytes = 0;
  Sample *base[3] = {nullptr, nullptr, nullptr};   // sample (0,0)
};

struct PlaneView {           // what kernels receive
  Sample *base;
  int pitch;                 // elements
  int width, height;
};

struct xvcb_ctx_impl {
  int device = 0;
  int width = 0, height = 0, bitdepth = 10, chroma_format = 1;
  xvcb200_plane_geom geom{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int status = XVCB200_OK;
  std::string error;
  std::vector<DevPicture> slots;
  size_t slot_stride = 0;            // bytes between consecutive slots (one arena)

  // CU array (device) and derived per-picture maps
  xvcb200_cu *d_cus = nullptr;
  int n_cus = 0, cap_cus = 0;
  int32_t *d_cu_map = nullptr;       // CU index per 4x4 block
  uint8_t *d_edge_bs[2] = {nullptr, nullptr};   // boundary strength per 4x4 block: [0] left edge, [1] top edge
  int map_w = 0, map_h = 0;

  // scratch (device), grown on demand
  void *d_scratch = nullptr;  size_t scratch_bytes = 0;
  void *d_scratch2 = nullptr; size_t scratch2_bytes = 0;
  void *h_pinned = nullptr;   size_t pinned_bytes = 0;

  bool fail(int code, const std::string &what) {
    if (status == XVCB200_OK) { status = code; error = what; }
    set_last_error(code, what.c_str());
    return false;
  }
  bool check(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    return fail(e == cudaErrorMemoryAllocation ? XVCB200_OUT_OF_MEMORY : XVCB200_CUDA_ERROR,
                std::string(what) + ": " + cudaGetErrorString(e));
  }
  PlaneView plane(int slot, int comp) const {
    PlaneView v;
    v.base = slots[slot].base[comp];
    v.pitch = geom.pitch[comp];
    v.width = geom.width[comp];
    v.height = geom.height[comp];
    return v;
  }
  void *scratch(size_t bytes);
  void *scratch2(size_t bytes);
  void *pinned(size_t bytes);
};

// Binds the calling thread to the context's device for the duration of an entry point (the CUDA
// current device is per host thread: the reference's ThreadEncoder workers, or a second context
// on another GPU of the same process, call in with whatever device was current) and restores it.
struct DevGuard {
  int prev = -1;
  bool switched = false;
  explicit DevGuard(const xvcb_ctx_impl *c) {
    if (!c) return;
    if (cudaGetDevice(&prev) == cudaSuccess && prev != c->device) switched = cudaSetDevice(c->device) == cudaSuccess;
  }
  ~DevGuard() { if (switched) cudaSetDevice(prev); }
  DevGuard(const DevGuard &) = delete;
  DevGuard &operator=(const DevGuard &) = delete;
};
constexpr int kMaxDevices = 64;      // per-device caches of launch configurations (indexed by cudaGetDevice)

struct Pic3 { PlaneView p[3]; };
inline Pic3 pic3(const xvcb_ctx_impl *c, int slot) {
  Pic3 r; for (int i = 0; i < 3; i++) r.p[i] = c->plane(slot, i); return r;
}

// ---- launchers (each enqueues on `stream`, bumps g_launch_count, returns cudaGetLastError) ----

// metrics.cu: one block, result into *d_out (uint64).  kind: see xo_sad/xo_ssd; metric >= 0 applies
// SampleMetric::Compare scaling (XVCB200_METRIC_*), metric < 0: raw sad (-1) / raw ssd (-2).
cudaError_t launch_block_metric(cudaStream_t s, int metric, int bitdepth, int w, int h, int a_short, int b_short,
                                const void *a, int sa, const void *b, int sb, unsigned long long *d_out);

// interp.cu
cudaError_t launch_block_filter(cudaStream_t s, int kind, int chroma, int w, int h, int bitdepth, const int16_t taps[8],
                                const void *src, int ss, void *dst, int ds);
cudaError_t launch_block_add_avg(cudaStream_t s, int w, int h, int offset, int shift, int bitdepth, const int16_t *a,
                                 int sa, const int16_t *b, int sb, Sample *dst, int ds);
cudaError_t launch_block_copy_bipred(cudaStream_t s, int w, int h, int offset, int shift, const Sample *ref, int rs,
                                     int16_t *pred, int ps);
cudaError_t launch_block_interp(cudaStream_t s, int chroma, int bipred, int w, int h, int bitdepth, int fx, int fy,
                                const Sample *ref, int rs, void *pred, int ps);
cudaError_t launch_motion_compensate(cudaStream_t s, const xvcb200_cu *d_cus, int n, int bitdepth,
                                     const Pic3 refs[2][5], Pic3 pred);
cudaError_t launch_motion_compensate_lic(cudaStream_t s, const xvcb200_cu *d_cus, int n_cus, const xvcb200_lic_cu *d_lic, int n,
                                         int bitdepth, const Pic3 refs[2][5], Pic3 rec, Pic3 pred);
cudaError_t launch_motion_compensate_affine(cudaStream_t s, const xvcb200_cu *d_cus, int n_cus, const xvcb200_affine_cu *d_aff,
                                            int n, int bitdepth, const Pic3 refs[2][5], Pic3 pred);

// transform.cu
cudaError_t launch_block_transform(cudaStream_t s, int forward, int w, int h, int bitdepth, int tx_hor, int tx_ver,
                                   int dst4x4, int dc_only, int skip, const int16_t *in, int is, int16_t *out, int os);
cudaError_t launch_block_quant(cudaStream_t s, int w, int h, int bitdepth, int qp_bd, int intra_pic, int sign_hiding,
                               int scan, const int16_t *in, int is, int16_t *out, int os, int *d_nnz);
cudaError_t launch_block_dequant(cudaStream_t s, int w, int h, int bitdepth, int qp_bd, const int16_t *in, int is,
                                 int16_t *out, int os);
struct TqParams {
  int bitdepth, intra_picture, table, off_u, off_v;
  int decode_only;        // 1: dequant + inverse + reconstruct from levels (CuDecoder::DecompressComponent)
  const xvcb200_tu_mode *modes;   // device, per CU (xvcb200_set_tu_modes), or nullptr
};

// me.cu
cudaError_t launch_tz_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n, int bitdepth,
                             uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes, xvcb200_me_result *d_res,
                             const int *d_job_index, const void *d_groups, int n_groups, void *d_states, int *d_counter,
                             uint32_t *d_pool, int pool_cap, int J = 0, int n_cols = 0, const int *cols = nullptr);
// subpel.cu
cudaError_t launch_subpel_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n,
                                 int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes,
                                 xvcb200_me_result *d_res, int *d_lists /* 15n + 32 ints of scratch */,
                                 cudaStream_t *side, cudaEvent_t *side_ev, int n_side, cudaEvent_t fork_ev);
// (orig may be a plane of int16 weighted-original samples: the kernels read it as signed 16 bit;
//  jobs with search_range == 0 are skipped)

// me_pipe.cu: the motion search of the picture pipeline around the search kernels.  Jobs / results are laid
// out [cu][J]: column j < R[0] = (list 0, ref_idx j), else (list 1, ref_idx j - R[0]).
struct MePipe {
  int n, J, R[2], Rmax;
  int ref_slot[2][5], range[2][5];
  int dup_of[5];            // list-1 ref_idx -> list-0 ref_idx with the same POC (searched once), -1: unique
  int pic_uni;              // 1: uni-predicted picture
  int bits_mode;            // 0: legacy (cost of the sub-pel search decides); 1: fast_inter_pred_bits rule
  int bi_iterations;        // SearchBiIterative passes (0: no bi-prediction)
  int bitdepth;
  uint32_t lambda;
  const int32_t *mvp;       // device, [cu][J][2]: predictor per (CU, list, ref_idx), 1/16 pel; nullptr: the CU's mv[list]
};
struct MeCuState;            // per-CU decision state (me_pipe.cu)
size_t me_cu_state_bytes();
cudaError_t launch_make_me_jobs(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, xvcb200_me_job *d_jobs);
cudaError_t launch_me_uni_decide(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, const xvcb200_me_job *d_jobs,
                                 xvcb200_me_result *d_res, void *d_state);
// Tensor maps (TMA descriptors) of the padded luma planes the bi-prediction full search reads: per job
// column (list 0 pictures, then list 1 pictures) a narrow (kFsBoxNarrow x kFsBoxRows samples) and a wide
// (kFsBoxWide x kFsBoxRows) box over the WHOLE allocation of the plane (margins included), so that box
// coordinates are (margin_x + x, margin_y + y) and rows beyond the allocation read as zero.  The
// innermost box coordinate must sit on a 16-byte boundary (measured: an odd sample offset raises
// "illegal instruction", tools/tma_probe.cu), so a box starts at the window origin rounded down to 8
// samples and is 7 samples wider than the window: 16 + 8 + 7 <= 40, 64 + 8 + 7 <= 80.
constexpr int kFsBoxNarrow = 40, kFsBoxWide = 80, kFsBoxRows = 8;
struct FsTensorMaps { CUtensorMap m[10][2]; };
// One pass of InterSearch::SearchBiIterative up to the full-pel vectors, for every CU still refining: the
// weighted original of the list that is kept -> luma plane of `worig` (int16), InterSearch::FullSearch
// (+-4) of every picture of the list that is searched -> jobs d_bi_jobs / results d_bi_res, [cu][Rmax]
cudaError_t launch_bi_search(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, int iteration, const xvcb200_me_job *d_jobs,
                             const xvcb200_me_result *d_res, void *d_state, PlaneView orig, const PlaneView *d_luma_views,
                             PlaneView worig, const FsTensorMaps &maps, int margin_x, int margin_y, xvcb200_me_job *d_bi_jobs,
                             xvcb200_me_result *d_bi_res, cudaStream_t side, cudaEvent_t side_ev, cudaEvent_t fork_ev);
cudaError_t launch_me_bi_decide(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, const xvcb200_me_job *d_jobs,
                                const xvcb200_me_result *d_bi_res, xvcb200_me_result *d_res, void *d_state);
cudaError_t launch_me_final_decide(cudaStream_t s, xvcb200_cu *d_cus, const MePipe &P, const void *d_state);
// xvcb200_full_search (API): weighted original built per job from `orig` and the luma plane of other_pred_slot
cudaError_t launch_full_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_fullsearch_job *d_jobs, int n,
                               int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_planes,
                               xvcb200_me_result *d_res);

// partition.cu: per CTU 64 CUs / 128 split flags / counts
cudaError_t launch_partition(cudaStream_t s, PlaneView orig, PlaneView ref, int cx16, int cy16, uint32_t lambda_me, int hdr_cu, int hdr_split,
                             int qp, xvcb200_cu *d_cus, int *d_n_cus, uint8_t *d_splits, int *d_n_splits);

// intra.cu
cudaError_t launch_intra_ref(cudaStream_t s, int w, int h, int bitdepth, const int nb[5], const Sample *d_edges, Sample *d_ref,
                             Sample *d_filt);
cudaError_t launch_intra_predict(cudaStream_t s, int mode, int w, int h, int bitdepth, int luma, const Sample *d_ref,
                                 const Sample *d_filt, Sample *d_out, int os);
cudaError_t launch_intra_lm_chroma(cudaStream_t s, const xvcb200_intra_job *d_jobs, int n, int bitdepth, PlaneView luma, PlaneView rec_u,
                                   PlaneView rec_v, PlaneView pred_u, PlaneView pred_v);
cudaError_t launch_intra_satd_scan(cudaStream_t s, const xvcb200_intra_job *d_jobs, int n, int bitdepth, PlaneView orig,
                                   PlaneView src, uint32_t *d_satd);
// deblock.cu
cudaError_t launch_pad_border(cudaStream_t s, Pic3 pic, const int pad[3]);
// slot planes <-> one tight buffer (planes back to back); widths must be multiples of 4 samples, 8-byte aligned rows
cudaError_t launch_plane_pack(cudaStream_t s, Pic3 pic, uint16_t *tight, int to_tight);
struct DeblockParams {
  int bitdepth, pic_type, beta_offset, tc_offset, table, off_u, off_v;
  long long ref_poc[2][5];
  // CUs with affine motion: aff_index[cu] = entry of `aff` or -1 (nullptr: none); their corner vectors enter the boundary strength
  const int *aff_index = nullptr;
  const xvcb200_affine_cu *aff = nullptr;
  // secondary (chroma) CU tree of an intra picture: chroma edges come from it, not from the primary tree
  const xvcb200_cu *chroma_cus = nullptr;
  int n_chroma_cus = 0;
  int32_t *chroma_map = nullptr;
};
cudaError_t launch_deblock(cudaStream_t s, const xvcb200_cu *d_cus, int n, const DeblockParams &p, Pic3 rec,
                           int32_t *d_map, uint8_t *d_bs_v, uint8_t *d_bs_h, int map_w, int map_h, int pass_mask,
                           int y_begin, int y_end, bool map_ready = false);
cudaError_t launch_cu_map(cudaStream_t s, const xvcb200_cu *d_cus, int n, int32_t *d_map, int map_w, int map_h);

}  // namespace xvcb

struct xvcb200_ctx : public xvcb::xvcb_ctx_impl {};

#endif
