// The motion search of the picture pipeline around the search kernels (me.cu, subpel.cu):
// InterSearch::SearchMotion (inter_search.cc:199-259) for every CU of a picture at once.
//
//   uni-prediction   SearchRefIdx (inter_search.cc:456-578) per list: every reference picture of the
//                    list is searched (TZ + sub-pel, MotionEstNormal :606-662), a list-1 picture with
//                    the POC of a list-0 picture reuses that result (same_poc_in_l0_mapping_, :536-543)
//                    and does not count as a "unique" list-1 candidate (:568-571).
//   bi-prediction    SearchBiIterative (:392-433): the list with the higher uni cost is searched
//                    first against the weighted original 2 * orig - pred(other list)
//                    (ResidualBuffer::SubtractWeighted, sample_buffer.h:147-161) with FullSearch
//                    (:853-891, +-inter_search_range_bi = 4 around the list's best uni vector) and the
//                    sub-pel search on the same int16 original; distortion halved (:660); lists
//                    alternate until an iteration brings no gain.
//   decision         bi if its cost is <= both uni costs, else list 0 unless unique list 1 is
//                    cheaper (:245-258).
//
// Rate: the reference's CABAC-independent estimate (GetInterPredBits with fast_inter_pred_bits,
// inter_search.cc:1084-1130): uni = (1 | 3) + ref_idx bits + 1 (mvp flag) + exp-Golomb(mvd), bi = 5 +
// both lists' ref_idx bits + 1 + exp-Golomb(mvd); cost = dist + ((bits * lambda) >> 16).  The
// predictor of a (list, reference picture) is the one given with xvcb200_set_mv_predictors (GetMvpList
// is per ref_idx: neighbour vectors scaled by POC distance), else the vector the CU array carries in
// mv[list] when the picture is handed over; the two-entry mvp list of the reference is collapsed to
// its first entry.
// bits_mode 0 keeps round 1's rule (the cost of the sub-pel search alone picks list 0 or 1).
//
// CUs flagged XVCB200_CU_INTRA or XVCB200_CU_SKIP_ME take no part: no job of theirs is searched and
// their mv / ref_idx are left as the host set them.
#include "xvcb_interp.cuh"

namespace xvcb {

struct MeCuState {
  int32_t uni_mv[2][2];      // best uni-prediction vector per list (list 1: over all its pictures)
  int32_t l1u_mv[2];         // best vector among the unique list-1 pictures
  int32_t bi_mv[2][2];       // bi-prediction state (SearchBiIterative's CU state)
  uint32_t cost_uni[2], cost_l1u, cost_bi;
  int8_t uni_ref[2], l1u_ref, bi_ref[2];
  int8_t search_list, active, bi_done;
};
size_t me_cu_state_bytes() { return sizeof(MeCuState); }

__device__ __forceinline__ bool cu_searched(const xvcb200_cu &cu) {
  return !(cu.flags & (XVCB200_CU_INTRA | XVCB200_CU_SKIP_ME));
}
__device__ __forceinline__ uint32_t ref_idx_bits(int r, int R) { return R <= 1 ? 0u : (uint32_t)(r + 1 - (r == R - 1)); }
__device__ __forceinline__ uint32_t mvd_bits_down(const int32_t mvp[2], const int32_t mv[2], int down) {
  return exp_golomb_bits((mv[0] - mvp[0]) >> (2 + down)) + exp_golomb_bits((mv[1] - mvp[1]) >> (2 + down));
}

// jobs[cu * J + j]; search_range = 0 marks a job that is not searched
__global__ void make_me_jobs_kernel(const xvcb200_cu *__restrict__ cus, const __grid_constant__ MePipe P,
                                    xvcb200_me_job *__restrict__ jobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n * P.J) return;
  const int c = i / P.J, jc = i - c * P.J;
  const int l = jc >= P.R[0] ? 1 : 0, r = l ? jc - P.R[0] : jc;
  const xvcb200_cu cu = cus[c];
  xvcb200_me_job j;
  j.cu = c; j.ref_slot = P.ref_slot[l][r];
  j.search_range = (!cu_searched(cu) || (l == 1 && P.dup_of[r] >= 0)) ? 0 : P.range[l][r];
  if (P.mvp) { j.mvp[0] = P.mvp[2 * (size_t)i]; j.mvp[1] = P.mvp[2 * (size_t)i + 1]; }      // GetMvpList is per (list, ref_idx)
  else { j.mvp[0] = cu.mv[l][0]; j.mvp[1] = cu.mv[l][1]; }
  j.prev[0] = 0; j.prev[1] = 0; j.list = l;
  jobs[i] = j;
}

cudaError_t launch_make_me_jobs(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, xvcb200_me_job *d_jobs) {
  if (P.n <= 0) return cudaSuccess;
  g_launch_count++;
  make_me_jobs_kernel<<<(P.n * P.J + 255) / 256, 256, 0, s>>>(d_cus, P, d_jobs);
  return cudaGetLastError();
}

// SearchRefIdx per list on the finished uni searches (a thread per CU)
__global__ void me_uni_decide_kernel(const xvcb200_cu *__restrict__ cus, const __grid_constant__ MePipe P,
                                     const xvcb200_me_job *__restrict__ jobs, xvcb200_me_result *__restrict__ res,
                                     MeCuState *__restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const xvcb200_cu cu = cus[i];
  MeCuState st;
  memset(&st, 0, sizeof(st));
  st.active = cu_searched(cu) ? 1 : 0;
  st.uni_ref[0] = st.uni_ref[1] = st.l1u_ref = st.bi_ref[0] = st.bi_ref[1] = -1;
  st.cost_uni[0] = st.cost_uni[1] = st.cost_l1u = st.cost_bi = 0xffffffffu;
  if (st.active) {
    const int down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
    for (int l = 0; l < 2; l++)
      for (int r = 0; r < P.R[l]; r++) {
        const int j = l ? P.R[0] + r : r;
        const bool dup = l == 1 && P.dup_of[r] >= 0;
        xvcb200_me_result q = res[(size_t)i * P.J + (dup ? P.dup_of[r] : j)];
        if (dup) res[(size_t)i * P.J + j] = q;          // unipred_best_mv_[L1][r] = the list-0 vector (:536-549)
        uint32_t cost = q.cost;
        if (P.bits_mode) {
          const uint32_t bits = (P.pic_uni ? 1u : 3u) + ref_idx_bits(r, P.R[l]) + 1u + mvd_bits_down(jobs[(size_t)i * P.J + j].mvp, q.mv, down);
          cost = q.dist + ((bits * P.lambda) >> 16);
        }
        if (cost < st.cost_uni[l]) { st.cost_uni[l] = cost; st.uni_ref[l] = (int8_t)r; st.uni_mv[l][0] = q.mv[0]; st.uni_mv[l][1] = q.mv[1]; }
        if (l == 1 && !dup && cost < st.cost_l1u) { st.cost_l1u = cost; st.l1u_ref = (int8_t)r; st.l1u_mv[0] = q.mv[0]; st.l1u_mv[1] = q.mv[1]; }
      }
  }
  state[i] = st;
}

cudaError_t launch_me_uni_decide(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, const xvcb200_me_job *d_jobs,
                                 xvcb200_me_result *d_res, void *d_state) {
  if (P.n <= 0) return cudaSuccess;
  g_launch_count++;
  me_uni_decide_kernel<<<(P.n + 127) / 128, 128, 0, s>>>(d_cus, P, d_jobs, d_res, static_cast<MeCuState *>(d_state));
  return cudaGetLastError();
}

// ---------------------------------------------------------------- InterSearch::FullSearch
// Every full-pel position of the clipped +-range window around the bootstrap vector, row-major, on
// the int16 weighted original (SampleMetric on Residual vs Sample: SAD, or SAD over every second
// row x 2 for blocks higher than 8, inter_search.cc:1059-1069).  One CTA of 128 threads per job:
// the window (w + 2 range) x (h + 2 range) and the weighted original are staged in shared memory
// as packed 16-bit pairs, both biased by 2^bitdepth so that the signed difference becomes an
// unsigned one (|a - b| = max - min per 16-bit lane, VIMNMX.U16x2); a thread per candidate.
constexpr int kFsRange = 8;                                   // largest range served (the reference searches +-4)
constexpr int kFsWinPitch = (64 + 2 * kFsRange) / 2 + 3;      // words per window row incl. the word the funnel shift reads ahead; odd: rows land in different banks
constexpr int kFsOrgPitch = 33;

__device__ __forceinline__ uint32_t fs_absdiff2(uint32_t a, uint32_t b) { return __vmaxu2(a, b) - __vminu2(a, b); }

template <class GetWorig>
__device__ __forceinline__ void full_search_cta(const xvcb200_cu &cu, int ref_w, int ref_h, const Sample *ref_base, int ref_pitch,
                                                int mvpx, int mvpy, int cx16, int cy16, int range, int bitdepth, uint32_t lambda,
                                                GetWorig get_worig, uint32_t *s_win, uint32_t *s_org, unsigned long long *s_best,
                                                xvcb200_me_result *out) {
  const int tid = threadIdx.x;
  const int w = cu.w, h = cu.h;
  int lo[2], hi[2];
  min_max_mv(cu.x, cu.y, ref_w, ref_h, cx16, cy16, range, lo, hi);
  const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, total = nx * ny;
  const uint32_t bias = (uint32_t)1 << bitdepth, bias2 = bias | (bias << 16);
  const int fast = h > 8, rstep = fast ? 2 : 1, rows = fast ? h >> 1 : h;
  // window rows that the metric visits: candidate row offsets 0 .. ny-1 plus block rows 0, rstep, ...
  const int wx0 = (cu.x + lo[0]) & ~1;                       // even start: aligned 32-bit loads
  const int wpairs = ((cu.x + hi[0] + w + 1) >> 1) - (wx0 >> 1);
  const int wrows = h + ny - 1;
  if (tid == 0) *s_best = ~0ull;
  for (int k = tid; k < wrows * wpairs; k += 128) {
    const int y = k / wpairs, x = k - y * wpairs;
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(ref_base + (cu.y + lo[1] + y) * ref_pitch + wx0) + x);
    s_win[y * kFsWinPitch + x] = v + bias2;
  }
  const int lw2 = 30 - __clz(w);
  for (int k = tid; k < (rows << lw2); k += 128) {
    const int y = (k >> lw2) * rstep, x = (k & ((w >> 1) - 1)) * 2;
    const uint32_t a = (uint32_t)(get_worig(x, y) + (int)bias) & 0xffffu, b = (uint32_t)(get_worig(x + 1, y) + (int)bias) & 0xffffu;
    s_org[(k >> lw2) * kFsOrgPitch + (x >> 1)] = a | (b << 16);
  }
  __syncthreads();
  const int down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  for (int t = tid; t < total; t += 128) {
    const int cyi = t / nx, cxi = t - cyi * nx;
    const int ox = cu.x + lo[0] + cxi - wx0;                 // sample offset of the candidate inside the window row
    const int sh = (ox & 1) << 4;
    const uint32_t *wp = s_win + cyi * kFsWinPitch + (ox >> 1);
    uint32_t sad = 0;
    for (int r = 0; r < rows; r++) {
      const uint32_t *wr = wp + r * rstep * kFsWinPitch, *op = s_org + r * kFsOrgPitch;
      uint32_t prev = wr[0];
      for (int c = 0; c < (w >> 1); c += 4) {                // <= 4 differences of <= 3 * 2^bitdepth per 16-bit lane
        uint32_t acc = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (c + u < (w >> 1)) {
            const uint32_t nxt = wr[c + u + 1];
            acc += fs_absdiff2(op[c + u], __funnelshift_r(prev, nxt, sh));
            prev = nxt;
          }
        }
        sad += (acc & 0xffffu) + (acc >> 16);
      }
    }
    const uint32_t dist = fast ? (sad * 2) >> (bitdepth - 8) : sad >> (bitdepth - 8);
    const int cx = lo[0] + cxi, cy = lo[1] + cyi;
    const uint32_t cost = dist + ((lambda * mvd_bits_fullpel(mvpx, mvpy, cx, cy, down)) >> 16);
    atomicMin(s_best, ((unsigned long long)cost << 32) | (unsigned)t);     // first minimum in scan order (:868-886)
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long b = *s_best;
    const int t = (int)(b & 0xffffffffu);
    const int bx = lo[0] + t % nx, by = lo[1] + t / nx;
    out->mv_fullpel[0] = bx; out->mv_fullpel[1] = by;
    out->mv[0] = bx * 16; out->mv[1] = by * 16;
    out->cost_fullpel = (uint32_t)(b >> 32); out->dist = 0; out->cost = (uint32_t)(b >> 32); out->num_sad = (uint32_t)total;
  }
}

// ---- one SearchBiIterative pass up to the full-pel vector, one CTA per CU -------------------------------
//   1. thread 0 picks the list to search (:402-403, :428), writes the pass's jobs, and -- the search windows
//      depend on nothing the CTA computes -- issues cp.async.bulk.tensor.2d loads (TMA) of the windows of
//      the first two reference pictures of that list: (w + 8) x (h + 8) samples of the padded plane as
//      boxes of kFsBoxRows rows, narrow or wide by block width, starting at the window origin rounded
//      down to 8 samples (the copy engine wants 16-byte aligned box origins), arrival on an mbarrier each;
//   2. while they fly, all threads build the prediction of the list that is kept (luma, uni-prediction:
//      MotionCompensation with InterDir = that list, :415-418) and the weighted original 2 * orig - pred
//      (ResidualBuffer::SubtractWeighted) -- to the `worig` plane for the sub-pel search that follows, and
//      packed (rows the metric visits, biased by 2^bitdepth) into shared memory;
//   3. InterSearch::FullSearch per reference picture (:853-891): the window is copied once more shifted
//      by one sample, so that candidates at odd and at even offsets both read aligned 32-bit words; a
//      thread per (candidate, half of the block's rows): SAD = VIMNMX.U16x2 max - min on pairs biased by
//      2^bitdepth (the original is a broadcast read, neighbouring candidates share window words), halves
//      joined by a shared-memory atomicAdd, the first minimum in scan order through a 64-bit
//      (cost << 32 | position) atomicMin.  A third and later picture of the list re-uses a window buffer
//      as soon as its search is done.  (Measured: a warp per candidate with a REDUX.SUM per candidate costs
//      6x the instructions on the small blocks that make up most of a picture.)
// job.prev of the pass's jobs carries the bootstrap vector (GetBestUniPredMv, :497) in 1/16 pel.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int kFsWinRows = 72;                         // (64 + 8) rows at most
constexpr int kFsBiRange = 4;                          // encoder_settings.h:65 inter_search_range_bi
constexpr int kFsOrgPitch16 = 2 * kFsOrgPitch;         // the packed original addressed as uint16

struct FsGeom { int nx, ny, lox, loy, ox0, wp, wrows, wide, mvpx, mvpy, ref_slot; };

constexpr int kFsCands = (2 * kFsBiRange + 1) * (2 * kFsBiRange + 1);

// Two size classes, each a launch over all CUs (a CTA whose CU belongs to the other class leaves at once): the
// CTA's life is a chain of dependent global loads and barriers, so what counts is how many CTAs an SM holds.
// Blocks up to 16 x 16 (most CUs of a picture) need 6 KB of shared memory and 96 threads (a thread per candidate),
// larger ones 46 KB and 352 threads (a thread per candidate and quarter of the rows); the two launches run side by
// side on two streams (the small class alone is issue bound, the large one latency bound).
template <int MAXD, int kBiThreads>
__global__ void __launch_bounds__(kBiThreads) bi_search_kernel(const xvcb200_cu *__restrict__ cus, const __grid_constant__ MePipe P, int iteration,
                                                        const xvcb200_me_job *__restrict__ jobs, const xvcb200_me_result *__restrict__ res,
                                                        MeCuState *__restrict__ state, PlaneView orig, const PlaneView *__restrict__ luma,
                                                        PlaneView worig, const __grid_constant__ FsTensorMaps maps, int margin_x,
                                                        int margin_y, xvcb200_me_job *__restrict__ bi_jobs,
                                                        xvcb200_me_result *__restrict__ bi_res) {
  // interpolation scratch (MAXD x (MAXD + 7) int16) + prediction (MAXD x MAXD); the shifted window re-uses it afterwards
  constexpr int kBox = MAXD <= 16 ? kFsBoxNarrow : kFsBoxWide, kWinRows = MAXD + 2 * kFsBiRange;
  constexpr int kScratch = MAXD * (MAXD + 7) + MAXD * MAXD > kWinRows * kBox ? MAXD * (MAXD + 7) + MAXD * MAXD : kWinRows * kBox;
  __shared__ __align__(16) uint16_t s_scratch[kScratch];
  __shared__ __align__(128) uint16_t s_win[2][kWinRows * kBox];
  __shared__ uint32_t s_org[(MAXD <= 16 ? 8 : MAXD / 2) * kFsOrgPitch];
  __shared__ unsigned long long s_best;
  __shared__ uint2 s_cand[kFsCands];
  __shared__ uint32_t s_sad[kFsCands];
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ int s_sl;
  int16_t *tmp = reinterpret_cast<int16_t *>(s_scratch);
  Sample *pred = s_scratch + MAXD * (MAXD + 7);
  uint16_t *s_shift = s_scratch;
  const int i = blockIdx.x, tid = threadIdx.x;
  const xvcb200_cu cu = cus[i];
  if ((cu.w <= 16 && cu.h <= 16) != (MAXD <= 16)) return;            // the other class's CU
  MeCuState *st = &state[i];
  const bool run = st->active && !st->bi_done && st->uni_ref[0] >= 0 && st->uni_ref[1] >= 0;
  const int w = cu.w, h = cu.h;
  const int fast = h > 8, rstep = fast ? 2 : 1, rows = fast ? h >> 1 : h;

  // window geometry of the pass's job r (list sl)
  auto geom = [&](int sl, int r) {
    FsGeom g;
    const size_t col = (size_t)i * P.J + (sl ? P.R[0] + r : r);
    g.mvpx = jobs[col].mvp[0]; g.mvpy = jobs[col].mvp[1];
    g.ref_slot = P.ref_slot[sl][r];
    const PlaneView ref = luma[g.ref_slot];
    int lo[2], hi[2];
    min_max_mv(cu.x, cu.y, ref.width, ref.height, res[col].mv[0], res[col].mv[1], kFsBiRange, lo, hi);
    g.nx = hi[0] - lo[0] + 1; g.ny = hi[1] - lo[1] + 1; g.lox = lo[0]; g.loy = lo[1];
    g.ox0 = (cu.x + lo[0]) & 7;                                     // window origin inside the 16-byte aligned box
    g.wide = w + 2 * kFsBiRange + 7 > kFsBoxNarrow;
    g.wp = (g.wide ? kFsBoxWide : kFsBoxNarrow) >> 1;
    g.wrows = h + g.ny - 1;
    return g;
  };
  auto issue_window = [&](int sl, int r, int stage) {               // one thread
    const FsGeom g = geom(sl, r);
    const int box_w = 2 * g.wp, nbox = (g.wrows + kFsBoxRows - 1) / kFsBoxRows;
    const CUtensorMap *map = &maps.m[sl ? P.R[0] + r : r][g.wide];
    mbar_arrive_expect_tx(&s_full[stage], (uint32_t)(nbox * box_w * kFsBoxRows * 2));
    const int x0 = margin_x + cu.x + g.lox - g.ox0, y0 = margin_y + cu.y + g.loy;
    for (int k = 0; k < nbox; k++) tma_load_2d(&s_win[stage][k * kFsBoxRows * box_w], map, x0, y0 + k * kFsBoxRows, &s_full[stage]);
  };

  if (tid == 0) {
    int sl = st->search_list;
    if (run && iteration == 0) {
      // the list with the higher uni cost first (:402-403); CU state = best of both lists (:231-233)
      sl = st->cost_uni[0] <= st->cost_uni[1] ? 1 : 0;
      st->search_list = (int8_t)sl;
      for (int l = 0; l < 2; l++) { st->bi_ref[l] = st->uni_ref[l]; st->bi_mv[l][0] = st->uni_mv[l][0]; st->bi_mv[l][1] = st->uni_mv[l][1]; }
    }
    if (!run && !st->bi_done) st->bi_done = 1;
    s_sl = sl;
    for (int r = 0; r < P.Rmax; r++) {
      xvcb200_me_job j;
      j.cu = i; j.list = sl; j.ref_slot = 0; j.search_range = 0;
      j.mvp[0] = j.mvp[1] = 0; j.prev[0] = j.prev[1] = 0;
      if (run && r < P.R[sl]) {
        const size_t col = (size_t)i * P.J + (sl ? P.R[0] + r : r);
        const xvcb200_me_result q = res[col];
        j.mvp[0] = jobs[col].mvp[0]; j.mvp[1] = jobs[col].mvp[1];
        j.ref_slot = P.ref_slot[sl][r];
        j.search_range = kFsBiRange;
        j.prev[0] = q.mv[0]; j.prev[1] = q.mv[1];
      }
      bi_jobs[(size_t)i * P.Rmax + r] = j;
    }
    if (run) {
      mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
      mbar_init_fence();
      for (int r = 0; r < min(P.R[sl], 2); r++) issue_window(sl, r, r);
    }
  }
  __syncthreads();
  if (!run) return;
  const int sl = s_sl, other = 1 - sl;
  {
    const PlaneView rp = luma[P.ref_slot[other][st->bi_ref[other]]];
    int mx = st->bi_mv[other][0], my = st->bi_mv[other][1];
    clip_mv(cu.x, cu.y, rp.width, rp.height, mx, my);
    const Sample *r = rp.base + (cu.y + (my >> 4)) * rp.pitch + cu.x + (mx >> 4);
    interp_cta<false, 8>(w, h, P.bitdepth, mx & 15, my & 15, r, rp.pitch, pred, MAXD, tmp, tid, kBiThreads);
  }
  __syncthreads();
  const int bias = 1 << P.bitdepth;
  const uint32_t bias2 = (uint32_t)bias | ((uint32_t)bias << 16);
  const int lw = 31 - __clz(w), lpw = lw - 1, nelem = rows << lpw;
  {
    int16_t *dst = reinterpret_cast<int16_t *>(worig.base);
    uint16_t *org16 = reinterpret_cast<uint16_t *>(s_org);
    for (int k = tid; k < w * h; k += kBiThreads) {
      const int y = k >> lw, x = k & (w - 1);
      const int v = 2 * (int)orig.base[(cu.y + y) * orig.pitch + cu.x + x] - (int)pred[y * MAXD + x];
      dst[(cu.y + y) * worig.pitch + cu.x + x] = (int16_t)v;
      if (!(y & (rstep - 1))) org16[(y >> (rstep - 1)) * kFsOrgPitch16 + x] = (uint16_t)(v + bias);
    }
  }
  __syncthreads();                                                   // the scratch is free from here on: s_shift
  const int down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  for (int r = 0; r < P.R[sl]; r++) {
    const int stage = r & 1;
    const FsGeom g = geom(sl, r);
    mbar_wait(&s_full[stage], (uint32_t)(r >> 1) & 1);
    {   // the window once more, one sample to the left (the word after a row's last one is never read by a candidate)
      const uint32_t *src = reinterpret_cast<const uint32_t *>(s_win[stage]);
      uint32_t *dstw = reinterpret_cast<uint32_t *>(s_shift);
      for (int k = tid; k < g.wrows * g.wp - 1; k += kBiThreads) dstw[k] = __funnelshift_r(src[k], src[k + 1], 16);
    }
    // per candidate, once: where its window starts (word offset, bit 31 = odd sample offset -> shifted copy) and its rate
    const int total = g.nx * g.ny;
    static_assert(kBiThreads >= kFsCands, "a thread per candidate");
    if (tid < total) {
      const int cyi = tid / g.nx, cxi = tid - cyi * g.nx, ox = g.ox0 + cxi;
      s_cand[tid] = make_uint2((uint32_t)(cyi * g.wp + (ox >> 1)) | ((uint32_t)(ox & 1) << 31),
                               (P.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, g.lox + cxi, g.loy + cyi, down)) >> 16);
    }
    if (tid < kFsCands) s_sad[tid] = 0;
    if (tid == 0) s_best = ~0ull;
    __syncthreads();
    // a thread per (candidate, part of the rows): the original is a broadcast read, neighbouring candidates share words
    constexpr int kParts = kBiThreads >= 4 * kFsCands ? 4 : (kBiThreads >= 2 * kFsCands ? 2 : 1);
    if (tid < kParts * kFsCands) {
      const int half = tid / kFsCands, t = tid - half * kFsCands;
      if (t < total) {
        const uint2 cd = s_cand[t];
        const int hrows = rows / kParts, pairs = 1 << lpw;
        const uint32_t *pw = reinterpret_cast<const uint32_t *>((cd.x >> 31) ? s_shift : s_win[stage]) + (cd.x & 0x7fffffffu) + half * hrows * rstep * g.wp;
        const uint32_t *po = s_org + half * hrows * kFsOrgPitch;
        uint32_t sad = 0;
        for (int r = 0; r < hrows; r++) {
          if (pairs == 2) {
            const uint32_t acc = fs_absdiff2(po[0], pw[0] + bias2) + fs_absdiff2(po[1], pw[1] + bias2);
            sad += (acc & 0xffffu) + (acc >> 16);
          } else {
            for (int c = 0; c < pairs; c += 4) {          // 4 differences of <= 3 * 2^bitdepth per 16-bit lane
              const uint32_t acc = fs_absdiff2(po[c], pw[c] + bias2) + fs_absdiff2(po[c + 1], pw[c + 1] + bias2) +
                                   fs_absdiff2(po[c + 2], pw[c + 2] + bias2) + fs_absdiff2(po[c + 3], pw[c + 3] + bias2);
              sad += (acc & 0xffffu) + (acc >> 16);
            }
          }
          pw += rstep * g.wp; po += kFsOrgPitch;
        }
        atomicAdd(&s_sad[t], sad);
      }
    }
    __syncthreads();
    if (tid < total) {
      const uint32_t sad = s_sad[tid];
      const uint32_t dist = fast ? (sad * 2) >> (P.bitdepth - 8) : sad >> (P.bitdepth - 8);
      atomicMin(&s_best, ((unsigned long long)(dist + s_cand[tid].y) << 32) | (unsigned)tid);
    }
    __syncthreads();                                                 // every read of the window and of s_shift is done
    if (tid == 0) {
      const unsigned long long b = s_best;
      const int t = (int)(b & 0xffffffffu);
      const int bx = g.lox + t % g.nx, by = g.loy + t / g.nx;
      xvcb200_me_result *out = &bi_res[(size_t)i * P.Rmax + r];
      out->mv_fullpel[0] = bx; out->mv_fullpel[1] = by;
      out->mv[0] = bx * 16; out->mv[1] = by * 16;
      out->cost_fullpel = (uint32_t)(b >> 32); out->dist = 0; out->cost = (uint32_t)(b >> 32); out->num_sad = (uint32_t)total;
      if (r + 2 < P.R[sl]) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_window(sl, r + 2, stage);
      }
    }
  }
}

cudaError_t launch_bi_search(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, int iteration, const xvcb200_me_job *d_jobs,
                             const xvcb200_me_result *d_res, void *d_state, PlaneView orig, const PlaneView *d_luma_views,
                             PlaneView worig, const FsTensorMaps &maps, int margin_x, int margin_y, xvcb200_me_job *d_bi_jobs,
                             xvcb200_me_result *d_bi_res, cudaStream_t side, cudaEvent_t side_ev, cudaEvent_t fork_ev) {
  if (P.n <= 0) return cudaSuccess;
  g_launch_count += 2;
  cudaStream_t s1 = side ? side : s;
  if (side) {
    cudaEventRecord(fork_ev, s);
    cudaStreamWaitEvent(s1, fork_ev, 0);
  }
  bi_search_kernel<64, 352><<<P.n, 352, 0, s>>>(d_cus, P, iteration, d_jobs, d_res, static_cast<MeCuState *>(d_state), orig, d_luma_views, worig,
                                             maps, margin_x, margin_y, d_bi_jobs, d_bi_res);
  bi_search_kernel<16, 96><<<P.n, 96, 0, s1>>>(d_cus, P, iteration, d_jobs, d_res, static_cast<MeCuState *>(d_state), orig, d_luma_views, worig,
                                            maps, margin_x, margin_y, d_bi_jobs, d_bi_res);
  if (side) {
    cudaEventRecord(side_ev, s1);
    cudaStreamWaitEvent(s, side_ev, 0);
  }
  return cudaGetLastError();
}

// xvcb200_full_search: the same search for explicit jobs, weighted original = 2 * orig - luma of other_pred_slot
__global__ void __launch_bounds__(128) full_search_kernel(const xvcb200_cu *__restrict__ cus, const xvcb200_fullsearch_job *__restrict__ jobs,
                                                          int bitdepth, uint32_t lambda, PlaneView orig,
                                                          const PlaneView *__restrict__ planes, xvcb200_me_result *__restrict__ res) {
  __shared__ uint32_t s_win[(64 + 2 * kFsRange) * kFsWinPitch];
  __shared__ uint32_t s_org[64 * kFsOrgPitch];
  __shared__ unsigned long long s_best;
  const xvcb200_fullsearch_job job = jobs[blockIdx.x];
  const xvcb200_cu cu = cus[job.cu];
  const PlaneView ref = planes[job.ref_slot], other = planes[job.other_pred_slot];
  const Sample *po = orig.base + cu.y * orig.pitch + cu.x, *pp = other.base + cu.y * other.pitch + cu.x;
  const int opitch = orig.pitch, ppitch = other.pitch;
  full_search_cta(cu, ref.width, ref.height, ref.base, ref.pitch, job.mvp[0], job.mvp[1], job.center[0], job.center[1], job.range,
                  bitdepth, lambda, [&](int x, int y) { return 2 * (int)po[y * opitch + x] - (int)pp[y * ppitch + x]; }, s_win, s_org,
                  &s_best, &res[blockIdx.x]);
}

cudaError_t launch_full_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_fullsearch_job *d_jobs, int n,
                               int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_planes,
                               xvcb200_me_result *d_res) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  full_search_kernel<<<n, 128, 0, s>>>(d_cus, d_jobs, bitdepth, lambda_me, orig, d_planes, d_res);
  return cudaGetLastError();
}

// SearchRefIdx of one SearchBiIterative pass on the finished bi searches (a thread per CU)
__global__ void me_bi_decide_kernel(const xvcb200_cu *__restrict__ cus, const __grid_constant__ MePipe P,
                                    const xvcb200_me_job *__restrict__ jobs, const xvcb200_me_result *__restrict__ bi_res, xvcb200_me_result *__restrict__ res,
                                    MeCuState *__restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  MeCuState st = state[i];
  if (!st.active || st.bi_done) return;
  const xvcb200_cu cu = cus[i];
  const int down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  const int sl = st.search_list, other = 1 - sl;
  const uint32_t prev_best = st.cost_bi;
  const size_t col_other = (size_t)i * P.J + (other ? P.R[0] + st.bi_ref[other] : st.bi_ref[other]);
  const uint32_t bits_other = ref_idx_bits(st.bi_ref[other], P.R[other]) + 1u + mvd_bits_down(jobs[col_other].mvp, st.bi_mv[other], down);
  for (int r = 0; r < P.R[sl]; r++) {
    const xvcb200_me_result q = bi_res[(size_t)i * P.Rmax + r];
    if (P.bi_iterations > 1) {                      // SetBestUniPredMv also after a bi search (:549-553)
      xvcb200_me_result *u = &res[(size_t)i * P.J + (sl ? P.R[0] + r : r)];
      u->mv[0] = q.mv[0]; u->mv[1] = q.mv[1];
    }
    const uint32_t bits = 5u + bits_other + ref_idx_bits(r, P.R[sl]) + 1u + mvd_bits_down(jobs[(size_t)i * P.J + (sl ? P.R[0] + r : r)].mvp, q.mv, down);
    const uint32_t cost = (q.dist >> 1) + ((bits * P.lambda) >> 16);       // MotionEstNormal halves the bi distortion (:660)
    if (cost < st.cost_bi) { st.cost_bi = cost; st.bi_ref[sl] = (int8_t)r; st.bi_mv[sl][0] = q.mv[0]; st.bi_mv[sl][1] = q.mv[1]; }
  }
  if (st.cost_bi == prev_best) st.bi_done = 1;      // :425-427
  st.search_list = (int8_t)other;                   // :428
  state[i] = st;
}

cudaError_t launch_me_bi_decide(cudaStream_t s, const xvcb200_cu *d_cus, const MePipe &P, const xvcb200_me_job *d_jobs,
                                const xvcb200_me_result *d_bi_res, xvcb200_me_result *d_res, void *d_state) {
  if (P.n <= 0) return cudaSuccess;
  g_launch_count++;
  me_bi_decide_kernel<<<(P.n + 127) / 128, 128, 0, s>>>(d_cus, P, d_jobs, d_bi_res, d_res, static_cast<MeCuState *>(d_state));
  return cudaGetLastError();
}

// SearchMotion's final choice (:245-258) written into the CU array
__global__ void me_final_decide_kernel(xvcb200_cu *__restrict__ cus, const __grid_constant__ MePipe P,
                                       const MeCuState *__restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const MeCuState st = state[i];
  if (!st.active) return;
  xvcb200_cu *cu = &cus[i];
  int8_t ref[2] = {-1, -1};
  int32_t mv[2][2] = {{0, 0}, {0, 0}};
  if (P.bi_iterations > 0 && st.cost_bi != 0xffffffffu && st.cost_bi <= st.cost_uni[0] && st.cost_bi <= st.cost_l1u) {
    for (int l = 0; l < 2; l++) { ref[l] = st.bi_ref[l]; mv[l][0] = st.bi_mv[l][0]; mv[l][1] = st.bi_mv[l][1]; }
  } else if (st.cost_uni[0] <= st.cost_l1u) {
    ref[0] = st.uni_ref[0]; mv[0][0] = st.uni_mv[0][0]; mv[0][1] = st.uni_mv[0][1];
  } else {
    ref[1] = st.l1u_ref; mv[1][0] = st.l1u_mv[0]; mv[1][1] = st.l1u_mv[1];
  }
  for (int l = 0; l < 2; l++) { cu->ref_idx[l] = ref[l]; cu->mv[l][0] = mv[l][0]; cu->mv[l][1] = mv[l][1]; }
}

cudaError_t launch_me_final_decide(cudaStream_t s, xvcb200_cu *d_cus, const MePipe &P, const void *d_state) {
  if (P.n <= 0) return cudaSuccess;
  g_launch_count++;
  me_final_decide_kernel<<<(P.n + 127) / 128, 128, 0, s>>>(d_cus, P, static_cast<const MeCuState *>(d_state));
  return cudaGetLastError();
}

}  // namespace xvcb
