// Sub-pel motion search: InterSearch::SubpelSearch / GetSubpelDist (inter_search.cc:893-964) for a
// batch of (CU, reference picture) jobs, after the full-pel search wrote mv_fullpel.
//
// The reference evaluates 1 + 8 half-pel and 8 quarter-pel candidates one after the other, each
// by a full motion compensation (FilterLuma, inter_prediction.cc:1387-1430) and a SATD.  The
// candidates of one pass share almost all of their filter work:
//
//   half-pel pass   the 8 neighbours use one horizontal half-pel plane of (w+1) x (h+8) raw
//                   filter sums (integer columns X0-1 .. X0+w-1): its rows give the two
//                   horizontal-only candidates, one vertical half-pel pass over it gives a
//                   (w+1) x (h+1) plane holding all four diagonal candidates, one vertical pass
//                   over the reference gives the two vertical-only ones.  3 filter passes
//                   instead of 12.
//   quarter-pel     the 8 neighbours have three x coordinates; per x one horizontal pass over
//                   h+8 rows, then one vertical pass per candidate.  <= 3 + 8 instead of 16.
//
// Each filter pass is a register sliding window: a thread produces a run of 8 (9) outputs along
// the filter direction from 15 (16) inputs, all operands in shared memory.  SATD runs on up to
// four candidates of one plane at once (warp-cooperative Hadamard, xvcb_satd.cuh).  The arithmetic
// (shifts, offsets, int16 narrowing, clipping) is that of xvcb_interp.cuh, i.e. bit-exact to the
// reference; candidates are then replayed in the reference's order with its strict `<` tests.
//
// Teams: blocks of <= 256 samples are searched by one warp (four jobs per CTA), larger ones by
// a CTA of 128 or 256 threads; each class is one persistent launch that strides over a
// device-built job list.  Jobs whose candidates would be moved by InterPrediction::ClipMv
// (inter_prediction.cc:769-782) or with a 4-sample side go to the generic kernel, which
// interpolates every candidate on its own exactly as the reference does.
#include <mutex>

#include "xvcb_interp.cuh"
#include "xvcb_satd.cuh"

namespace xvcb {

// ---------------------------------------------------------------- generic path (one candidate at a time)
__global__ void __launch_bounds__(128) subpel_generic_kernel(const xvcb200_cu *__restrict__ cus,
                                                             const xvcb200_me_job *__restrict__ jobs,
                                                             const int *__restrict__ list, const int *__restrict__ count,
                                                             int bitdepth, uint32_t lambda, PlaneView orig,
                                                             const PlaneView *__restrict__ ref_planes,
                                                             xvcb200_me_result *__restrict__ res) {
  __shared__ int16_t tmp[64 * 71];
  __shared__ Sample pred[64 * 64];
  __shared__ int16_t org[64 * 64];       // original block, or the int16 weighted original of a bi-prediction pass
  __shared__ unsigned part[4];
  const int tid = threadIdx.x;
  const int n_list = *count;
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
    const int ji = list[li];
    const xvcb200_me_job job = jobs[ji];
    const xvcb200_cu cu = cus[job.cu];
    const PlaneView ref = ref_planes[job.ref_slot];
    const int w = cu.w, h = cu.h;
    const int lw = 31 - __clz(w);
    __syncthreads();
    for (int i = tid; i < w * h; i += 128) {
      const int y = i >> lw, x = i & (w - 1);
      org[y * 64 + x] = (int16_t)orig.base[(cu.y + y) * orig.pitch + cu.x + x];
    }
    const int fx0 = res[ji].mv_fullpel[0] * 16, fy0 = res[ji].mv_fullpel[1] * 16;
    const bool fullpel_only = (cu.flags & XVCB200_CU_FULLPEL_MV) != 0;
    uint32_t best_cost = 0xffffffffu, best_dist = 0xffffffffu;
    int best_x = fx0, best_y = fy0;
    const int8_t half[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {1, -1}, {-1, 1}, {1, 1}};
    const int8_t qpel[9][2] = {{0, 0}, {0, -1}, {0, 1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {1, 1}};
    auto diff = [&](int x, int y) { return (int)org[y * 64 + x] - (int)pred[y * 64 + x]; };
    for (int pass = 0; pass < 2; pass++) {
      const int bx = best_x, by = best_y;
      const int step = pass == 0 ? 8 : 4;      // MvDelta(.., 1) / MvDelta(.., 2) in 1/16 units
      for (int i = pass; i < 9; i++) {
        const int mvx = bx + (pass == 0 ? half[i][0] : qpel[i][0]) * step;
        const int mvy = by + (pass == 0 ? half[i][1] : qpel[i][1]) * step;
        int cx = mvx, cy = mvy;                // MotionCompensationMv clips a copy (inter_prediction.cc:747-748)
        clip_mv(cu.x, cu.y, ref.width, ref.height, cx, cy);
        const Sample *r = ref.base + (cu.y + (cy >> 4)) * ref.pitch + cu.x + (cx >> 4);
        __syncthreads();                       // previous candidate's SATD reads are done
        interp_cta<false, 8>(w, h, bitdepth, cx & 15, cy & 15, r, ref.pitch, pred, 64, tmp, tid, 128);
        __syncthreads();
        unsigned s = satd_block_partial(diff, w, h, tid, 128);
        s = warp_sum(s);
        if ((tid & 31) == 0) part[tid >> 5] = s;
        __syncthreads();
        const uint32_t dist = (part[0] + part[1] + part[2] + part[3]) >> (bitdepth - 8);
        if (fullpel_only) { best_dist = dist; best_cost = dist; break; }
        if (dist < best_cost) {
          const uint32_t cost = dist + ((lambda * mvd_bits(job.mvp[0], job.mvp[1], mvx, mvy)) >> 16);
          if (cost < best_cost) { best_cost = cost; best_dist = dist; best_x = mvx; best_y = mvy; }
        }
      }
      if (fullpel_only) break;
    }
    if (tid == 0) {
      res[ji].mv[0] = best_x; res[ji].mv[1] = best_y;
      res[ji].dist = best_dist; res[ji].cost = best_cost;
    }
  }
}

// ---------------------------------------------------------------- job classes
// lists: 15 segments of n ints + counts[16].  The classify kernel fills, by block area / shape:
//   0: 1024 samples   1: 4096   2: generic kernel   3..9: blocks of <= 256 samples by shape
//   11: 2048 samples  12: 512
// and the concat kernel builds the three lists the team kernels walk:
//   10 = 3..9 (one warp per job; shapes adjacent, so that the four warps of a CTA and neighbouring
//        CTAs run the same instructions at about the same time: the kernel is instruction-fetch
//        bound when every warp is somewhere else in the code)
//   13 = 1, 11 (CTA of 256)   14 = 0, 12 (CTA of 128) -- larger blocks first: jobs are fetched
//        one at a time, so the short ones fill the tail.
constexpr int kSubpelLists = 15;
__global__ void subpel_classify_kernel(const xvcb200_cu *__restrict__ cus, const xvcb200_me_job *__restrict__ jobs, int n,
                                       const PlaneView *__restrict__ ref_planes, const xvcb200_me_result *__restrict__ res,
                                       int *__restrict__ lists, int *__restrict__ counts) {
  const int ji = blockIdx.x * blockDim.x + threadIdx.x;
  int seg = -1;
  if (ji < n && jobs[ji].search_range != 0) {                // range 0: a column that is not searched (see MePipe)
    const xvcb200_me_job job = jobs[ji];
    const xvcb200_cu cu = cus[job.cu];
    const int area = (int)cu.w * cu.h;
    // every candidate lies within +-12/16 of the full-pel vector: where ClipMv would move one of them the job goes
    // to the generic kernel (decided here, so that kernel runs beside the team kernels instead of after them)
    const PlaneView ref = ref_planes[job.ref_slot];
    const int fx0 = res[ji].mv_fullpel[0] * 16, fy0 = res[ji].mv_fullpel[1] * 16;
    int ax = fx0 - 12, ay = fy0 - 12, bx = fx0 + 12, by = fy0 + 12;
    clip_mv(cu.x, cu.y, ref.width, ref.height, ax, ay);
    clip_mv(cu.x, cu.y, ref.width, ref.height, bx, by);
    const bool clipped = ax != fx0 - 12 || ay != fy0 - 12 || bx != fx0 + 12 || by != fy0 + 12;
    if (cu.w < 8 || cu.h < 8 || clipped) seg = 2;
    else if (area > 256) seg = area == 4096 ? 1 : (area == 2048 ? 11 : (area == 1024 ? 0 : 12));
    else seg = 3 + (28 - __clz((int)cu.w)) * 3 + (28 - __clz((int)cu.h));     // (log2 w - 3) * 3 + (log2 h - 3): 0,1,2,3,4,6
  }
  // one atomic per (warp, class): most jobs of a picture fall into two or three classes
  const unsigned peers = __match_any_sync(XVCB_FULL, seg);
  if (seg >= 0) {
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&counts[seg], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    lists[(size_t)seg * n + base + __popc(peers & ((1u << lane) - 1))] = ji;
  }
}

__global__ void subpel_concat_kernel(int n, int *__restrict__ lists, int *__restrict__ counts) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  {   // the seven shape lists of the small blocks -> list 10
    int end[7], total = 0;
#pragma unroll
    for (int q = 0; q < 7; q++) { total += counts[9 - q]; end[q] = total; }      // 32x8 / 16x16 ... first, 8x8 last
    if (li == 0) counts[10] = total;
    if (li < total) {
      int q = 0, first = 0;
#pragma unroll
      for (int t = 0; t < 6; t++)
        if (li >= end[t]) { q = t + 1; first = end[t]; }
      lists[(size_t)10 * n + li] = lists[(size_t)(9 - q) * n + (li - first)];
    }
  }
#pragma unroll
  for (int o = 0; o < 2; o++) {   // 4096 then 2048 -> list 13; 1024 then 512 -> list 14
    const int a = o == 0 ? 1 : 0, b = o == 0 ? 11 : 12, out = 13 + o;
    const int na = counts[a], nb = counts[b];
    if (li == 0) counts[out] = na + nb;
    if (li < na) lists[(size_t)out * n + li] = lists[(size_t)a * n + li];
    else if (li < na + nb) lists[(size_t)out * n + li] = lists[(size_t)b * n + (li - na)];
  }
}

// shared memory of one team (bytes): reference window, horizontal plane, prediction plane, original.
// Row pitches are chosen odd in 32-bit words so that threads working on consecutive rows hit
// distinct banks.
// `fused` (the one-warp teams): three intermediate planes and eight prediction planes, see the quarter-pel pass.
__host__ __device__ inline int subpel_team_bytes(int w, int h, bool fused) {
  const int nt = fused ? 3 : 1, np = fused ? 8 : 1;
  const int samples = (h + 9) * (w + 10) + nt * (h + 9) * (w + 2) + np * (h + 1) * (w + 2) + h * (w + 2);
  return (2 * samples + 15) & ~15;
}

// The eight taps of a phase as four pairs of int8 (every tap of kLumaFilterHighPrec lies in -11 .. 63), the second
// operand of IDP.2A: t[j] = tap 2j | tap 2j+1 << 8.
struct Taps8 { int t[4]; };
__device__ __forceinline__ int pack_taps(int a, int b) { return (a & 0xff) | ((b & 0xff) << 8); }
__device__ __forceinline__ Taps8 luma_taps(int frac) {
  Taps8 r;
#pragma unroll
  for (int k = 0; k < 4; k++) r.t[k] = pack_taps((int)c_luma_taps[frac][2 * k], (int)c_luma_taps[frac][2 * k + 1]);
  return r;
}

// Up to 9 FIR outputs spaced `step` elements apart along the filter direction: out(r) =
// sum_k taps[k] * in[(r + k) * step].  Reads 16 inputs (the last one only matters for count 9).
template <typename ST, class Emit>
__device__ __forceinline__ void fir_run(const ST *in, int step, int count, const Taps8 &taps, Emit emit) {
  // inputs are 16-bit (samples, or the signed 14-bit intermediate): neighbours packed in pairs (one PRMT each), then
  // two multiply-adds per instruction (IDP.2A: a signed 16-bit pair times a signed 8-bit pair) -- 4 instead of 8 per output
  int win[16];
#pragma unroll
  for (int k = 0; k < 16; k++) win[k] = (int)in[k * step];
  int pr[15];
#pragma unroll
  for (int k = 0; k < 15; k++) pr[k] = (int)__byte_perm((unsigned)win[k], (unsigned)win[k + 1], 0x5410);
#pragma unroll
  for (int r = 0; r < 9; r++) {
    if (r < count) {
      int sum = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) sum = __dp2a_lo(pr[r + 2 * k], taps.t[k], sum);
      emit(r, sum);
    }
  }
}

template <int T> __device__ __forceinline__ void team_sync() {
  if (T == 32) __syncwarp(); else __syncthreads();
}

// SATD of up to four candidates that live in one prediction plane at offsets (ox, oy) in {0,1}^2
// (packed two bits per candidate in `offs`).  The candidates are stacked into one list of tile
// rows so that small blocks still fill the team.
template <int TW, int TH>
__device__ __forceinline__ void satd_stacked(const int16_t *org, int op, const Sample *pred, int pp, unsigned offs, int nc,
                                             int w, int h, int tid, int nthreads, unsigned (&acc)[4], int cstride) {
  const int ltx = 31 - __clz(w / TW);
  const int lper = ltx + (31 - __clz(h));      // log2(tile rows per candidate)
  const int total = nc << lper;
  const int lane = tid & 31;
  for (int g0 = 0; g0 < total; g0 += nthreads) {
    const int g = g0 + tid;
    const bool active = g < total;
    const int c = active ? g >> lper : 0, gi = g & ((1 << lper) - 1);
    const int tile = gi / TH, r = gi % TH;
    const int tx = (tile & ((1 << ltx) - 1)) * TW, ty = (tile >> ltx) * TH + r;
    const int ox = (offs >> (2 * c)) & 1, oy = (offs >> (2 * c + 1)) & 1;
    const int16_t *po = org + ty * op + tx;
    const Sample *pq = pred + c * cstride + (ty + oy) * pp + tx + ox;      // cstride: candidates in planes of their own
    int v[TW];
#pragma unroll
    for (int i = 0; i < TW; i++) v[i] = active ? (int)po[i] - (int)pq[i] : 0;
#pragma unroll
    for (int len = 1; len < TW; len <<= 1)
#pragma unroll
      for (int i = 0; i < TW; i += 2 * len)
#pragma unroll
        for (int j = i; j < i + len; j++) {
          const int p = v[j], q = v[j + len];
          v[j] = p + q;
          v[j + len] = p - q;
        }
#pragma unroll
    for (int o = 1; o < TH; o <<= 1) {
      const bool upper = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < TW; i++) {
        const int pv = __shfl_xor_sync(XVCB_FULL, v[i], o);
        v[i] = upper ? pv - v[i] : v[i] + pv;
      }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < TW; i++) s = (int)__sad(v[i], 0, (unsigned)s);      // |v| + s in one instruction
#pragma unroll
    for (int o = 1; o < TH; o <<= 1) s += __shfl_xor_sync(XVCB_FULL, s, o);
    if (active && r == 0) {
      const unsigned t = (unsigned)satd_norm<TW, TH>(s);
#pragma unroll
      for (int cc = 0; cc < 4; cc++) acc[cc] += c == cc ? t : 0u;
    }
  }
}

// Tile choice by block shape (sample_metric.cc:322-387) for sides >= 8, then the team-wide sums.
template <int T>
__device__ __forceinline__ void satd_candidates(const int16_t *org, int op, const Sample *pred, int pp, unsigned offs, int nc,
                                                int w, int h, int tid, unsigned *s_part, unsigned (&sum)[4], int cstride = 0) {
  unsigned acc[4] = {0, 0, 0, 0};
  if (w > h) satd_stacked<16, 8>(org, op, pred, pp, offs, nc, w, h, tid, T, acc, cstride);
  else if (w < h) satd_stacked<8, 16>(org, op, pred, pp, offs, nc, w, h, tid, T, acc, cstride);
  else satd_stacked<8, 8>(org, op, pred, pp, offs, nc, w, h, tid, T, acc, cstride);
#pragma unroll
  for (int c = 0; c < 4; c++) acc[c] = warp_sum(acc[c]);
  if (T == 32) {
#pragma unroll
    for (int c = 0; c < 4; c++) sum[c] = acc[c];
    __syncwarp();
  } else {
    if ((tid & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 4; c++) s_part[(tid >> 5) * 4 + c] = acc[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; c++) {
      unsigned t = 0;
      for (int wi = 0; wi < T / 32; wi++) t += s_part[wi * 4 + c];
      sum[c] = t;
    }
    __syncthreads();
  }
}

template <int T>
__global__ void __launch_bounds__(T == 32 ? 128 : T, T == 32 ? 6 : 1) subpel_team_kernel(
    const xvcb200_cu *__restrict__ cus, const xvcb200_me_job *__restrict__ jobs, const int *__restrict__ list,
    const int *__restrict__ count, int *__restrict__ fetch, int team_bytes, int bitdepth,
    uint32_t lambda, PlaneView orig, const PlaneView *__restrict__ ref_planes, xvcb200_me_result *__restrict__ res) {
  extern __shared__ __align__(16) unsigned char subpel_smem[];
  __shared__ unsigned s_part[T == 32 ? 4 : (T / 32) * 4];
  __shared__ unsigned s_sd[T == 32 ? 4 : 1][20];
  unsigned *sd = s_sd[T == 32 ? (threadIdx.x >> 5) : 0];
  const int tid = T == 32 ? (threadIdx.x & 31) : threadIdx.x;
  __shared__ int s_li;
  __shared__ int s_taps[16][4];          // the luma filters as IDP.2A operands: tasks of one fused pass use different phases per lane
  if (T == 32) {
    if (threadIdx.x < 64)
      s_taps[threadIdx.x >> 2][threadIdx.x & 3] = pack_taps((int)c_luma_taps[threadIdx.x >> 2][2 * (threadIdx.x & 3)],
                                                            (int)c_luma_taps[threadIdx.x >> 2][2 * (threadIdx.x & 3) + 1]);
    __syncthreads();
  }
  unsigned char *base = subpel_smem + (T == 32 ? (threadIdx.x >> 5) * team_bytes : 0);
  const int n_list = *count;
  const int maxv = (1 << bitdepth) - 1;
  int sh1, off1, sh2, off2;
  filter_shift_offset(false, false, bitdepth, sh1, off1);     // horizontal stage of the 2-D filter
  filter_shift_offset(true, true, bitdepth, sh2, off2);       // vertical stage on the intermediate
  for (;;) {
    // jobs are taken one at a time from the class list (a static stride leaves 2 vs 3 jobs per team)
    int li = 0;
    if (T == 32) {
      if (tid == 0) li = atomicAdd(fetch, 1);
      li = __shfl_sync(XVCB_FULL, li, 0);
    } else {
      if (tid == 0) s_li = atomicAdd(fetch, 1);
      __syncthreads();
      li = s_li;
      __syncthreads();
    }
    if (li >= n_list) break;
    const int ji = list[li];
    const xvcb200_me_job job = jobs[ji];
    const xvcb200_cu cu = cus[job.cu];
    const PlaneView ref = ref_planes[job.ref_slot];
    const int w = cu.w, h = cu.h;
    const int mfx = res[ji].mv_fullpel[0], mfy = res[ji].mv_fullpel[1];
    const int fx0 = mfx * 16, fy0 = mfy * 16;
    // (jobs whose candidates ClipMv would move were handed to the generic kernel by the classification)
    const bool fullpel_only = (cu.flags & XVCB200_CU_FULLPEL_MV) != 0;
    const int RP = w + 10, TP = w + 2, PP = w + 2, OP = w + 2;
    Sample *sref = reinterpret_cast<Sample *>(base);
    int16_t *tmp = reinterpret_cast<int16_t *>(sref + (h + 9) * RP);
    const int tstride = (h + 9) * TP, pstride = (h + 1) * PP;          // one intermediate / prediction plane
    Sample *pred = reinterpret_cast<Sample *>(tmp + (T == 32 ? 3 : 1) * tstride);
    int16_t *org = reinterpret_cast<int16_t *>(pred + (T == 32 ? 8 : 1) * pstride);    // signed: also holds a weighted original
    team_sync<T>();                      // the previous job of this team is done with the buffers
    // reference window: rows Y0-4 .. Y0+h+3, columns from the even sample at or left of X0-4
    // (aligned 32-bit loads; `co` = 0/1 is where X0-4 sits in the staged row), original block.
    const int co = (cu.x + mfx - 4) & 1;
    {
      const Sample *src = ref.base + (cu.y + mfy - 4) * ref.pitch + (cu.x + mfx - 4 - co);
      const int wr = RP >> 1, total = (h + 8) * wr;
      uint32_t *dst = reinterpret_cast<uint32_t *>(sref);
      for (int i0 = tid; i0 < total; i0 += 8 * T) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {          // eight independent loads in flight per thread
          const int i = i0 + u * T;
          if (i < total) {
            const int y = i / wr, x = i - y * wr;
            v[u] = __ldg(reinterpret_cast<const uint32_t *>(src + y * ref.pitch) + x);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int i = i0 + u * T;
          if (i < total) dst[i] = v[u];
        }
      }
      const int lw2 = 30 - __clz(w);          // log2(w / 2)
      const Sample *so = orig.base + cu.y * orig.pitch + cu.x;
      uint32_t *od = reinterpret_cast<uint32_t *>(org);
      for (int i0 = tid; i0 < (w * h) >> 1; i0 += 4 * T) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * T;
          if (i < (w * h) >> 1) v[u] = __ldg(reinterpret_cast<const uint32_t *>(so + (i >> lw2) * orig.pitch) + (i & ((w >> 1) - 1)));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int i = i0 + u * T;
          if (i < (w * h) >> 1) od[(i >> lw2) * (OP >> 1) + (i & ((w >> 1) - 1))] = v[u];
        }
      }
    }
    team_sync<T>();

    // H pass: tmp[row][c] for all h+8 rows, c in [0, ncols); raw sums of rows [prow0, prow0+h)
    // also give the horizontal-only prediction.
    auto h_product = [&](int frac, int col0, int ncols, bool want_pred, int prow0) {
      const Taps8 taps = luma_taps(frac);
      const int rows_n = h + 8, runs = w >> 3;
      for (int task = tid; task < rows_n * runs; task += T) {
        const int run = task / rows_n, row = task - run * rows_n;
        const int c0 = run * 8;
        const int cnt = run == runs - 1 ? ncols - c0 : 8;
        const bool to_pred = want_pred && row >= prow0 && row < prow0 + h;
        int16_t *tp = tmp + row * TP + c0;
        Sample *pp = pred + (row - prow0) * PP + c0;
        fir_run(sref + row * RP + co + col0 + c0, 1, cnt, taps, [&](int r, int sum) {
          tp[r] = (int16_t)((sum + off1) >> sh1);
          if (to_pred) pp[r] = (Sample)clip3i((sum + 32) >> 6, 0, maxv);
        });
      }
    };
    // V pass over the intermediate (FROM_TMP) or over the reference window: pred[r][c], r in
    // [0, nrows), c in [0, ncols); the taps of output row 0 start at source row row0.
    auto v_product = [&](bool from_tmp, int frac, int src_col0, int row0, int nrows, int ncols) {
      const Taps8 taps = luma_taps(frac);
      const int runs = h >> 3;
      for (int task = tid; task < ncols * runs; task += T) {
        const int run = task / ncols, col = task - run * ncols;
        const int r0 = run * 8;
        const int cnt = run == runs - 1 ? nrows - r0 : 8;
        Sample *pp = pred + r0 * PP + col;
        if (from_tmp) {
          fir_run(tmp + (row0 + r0) * TP + col, TP, cnt, taps, [&](int r, int sum) {
            pp[r * PP] = (Sample)clip3i((int)(int16_t)((sum + off2) >> sh2), 0, maxv);
          });
        } else {
          fir_run(sref + (row0 + r0) * RP + co + src_col0 + col, RP, cnt, taps, [&](int r, int sum) {
            pp[r * PP] = (Sample)clip3i((int)(int16_t)((sum + 32) >> 6), 0, maxv);
          });
        }
      }
    };

    uint32_t best_cost = 0xffffffffu, best_dist = 0xffffffffu;
    int best_x = fx0, best_y = fy0;
    auto consider = [&](uint32_t satd, int mvx, int mvy) {      // inter_search.cc:925-934
      const uint32_t dist = satd >> (bitdepth - 8);
      if (dist < best_cost) {
        const uint32_t cost = dist + ((lambda * mvd_bits(job.mvp[0], job.mvp[1], mvx, mvy)) >> 16);
        if (cost < best_cost) { best_cost = cost; best_dist = dist; best_x = mvx; best_y = mvy; }
      }
    };
    // The search as a list of steps, each = [one filter pass] + SATD of the 1..4 candidates its
    // plane holds, executed by ONE loop so that every routine exists once in the code (the
    // kernel has to stay inside the instruction cache: warps of different teams run different
    // steps at the same time).  SATD sums land in sd[]: 0 = full-pel, 1..8 = half-pel list,
    // 9 + 3j + k = quarter-pel candidate (x index j, y index k).
    int bx = fx0, by = fy0;
#pragma unroll 1
    for (int s = 0; s < (T == 32 ? 7 : 16); s++) {
      int kind = 0;            // 0 none, 1 horizontal, 2 vertical from the reference, 3 vertical from the intermediate
      int frac = 8, col0 = 0, row0 = 0, nrows = h, ncols = w, want_pred = 0, nc = 1, dst = 0;
      unsigned offs = 0;
      const Sample *pq = pred;
      int ppitch = PP, cstride = 0;
      if (s == 0) { pq = sref + 4 * RP + 4 + co; ppitch = RP; }
      else if (s == 1) { kind = 2; col0 = 4; nrows = h + 1; nc = 2; offs = 2u << 2; dst = 1; }             // (0,-1) (0,1)
      else if (s == 2) { kind = 1; ncols = w + 1; want_pred = 1; row0 = 4; nc = 2; offs = 1u << 2; dst = 3; }  // (-1,0) (1,0)
      else if (s == 3) { kind = 3; nrows = h + 1; ncols = w + 1; nc = 4; offs = (1u << 2) | (2u << 4) | (3u << 6); dst = 5; }
      else if (T == 32 && s > 4) {      // one-warp teams: SATD of the quarter-pel candidates, four planes at a time
        nc = 4; cstride = pstride; pq = pred + (s - 5) * 4 * pstride; dst = s == 5 ? 9 : 14;
      } else {
        if (s == 4) {          // half-pel decisions in list order, then the quarter-pel pass around the winner
          team_sync<T>();
          consider(sd[0], fx0, fy0);
          const int hx[8] = {0, 0, -1, 1, -1, 1, -1, 1}, hy[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
#pragma unroll
          for (int i = 0; i < 8; i++) consider(sd[1 + i], fx0 + hx[i] * 8, fy0 + hy[i] * 8);
          bx = best_x; by = best_y;
          if (T == 32) {
            // One-warp teams (blocks of <= 256 samples, most of them 8 x 8): a step of the list above keeps 8 or 16
            // lanes busy on such a block.  Here the quarter-pel pass is three fused steps over planes of their own:
            // all horizontal planes at once (x indices with a fractional column; on an integer row they also give
            // candidate (j, 1) directly), all vertical passes at once (a task = candidate, column, run of 8 rows, each
            // with its own filter phase), then (steps 5, 6) the SATD of the eight candidates four at a time.
            const bool xint = (bx & 15) == 0, yint = (by & 15) == 0;
            {
              const int rows_n = h + 8, per = rows_n * (w >> 3), prow0 = 4 + ((by >> 4) - mfy);
              for (int task = tid; task < (xint ? 2 : 3) * per; task += T) {
                const int p = task / per, rem = task - p * per;
                const int j = xint ? 2 * p : p;
                const int xv = bx + (j - 1) * 4, ixr = (xv >> 4) - mfx;
                const int run = rem / rows_n, row = rem - run * rows_n, c0 = run * 8;
                const bool to_pred = yint && j != 1 && row >= prow0 && row < prow0 + h;
                const int idx = 3 * j + 1, slot = idx < 4 ? idx : idx - 1;
                int16_t *tp = tmp + p * tstride + row * TP + c0;
                Sample *pp = pred + slot * pstride + (row - prow0) * PP + c0;
                Taps8 taps;
#pragma unroll
                for (int k = 0; k < 4; k++) taps.t[k] = s_taps[xv & 15][k];
                fir_run(sref + row * RP + co + ixr + 1 + c0, 1, 8, taps, [&](int r, int sum) {
                  tp[r] = (int16_t)((sum + off1) >> sh1);
                  if (to_pred) pp[r] = (Sample)clip3i((sum + 32) >> 6, 0, maxv);
                });
              }
            }
            team_sync<T>();
            {
              const int per = w * (h >> 3);
              for (int task = tid; task < 8 * per; task += T) {
                const int slot = task / per, rem = task - slot * per;
                const int idx = slot < 4 ? slot : slot + 1, j = idx / 3, k = idx - 3 * j;
                const int yv = by + (k - 1) * 4, iyr = (yv >> 4) - mfy;
                if ((yv & 15) == 0) continue;                 // came out of the horizontal pass
                const int xv = bx + (j - 1) * 4, ixr = (xv >> 4) - mfx;
                const int run = rem / w, col = rem - run * w, r0 = run * 8;
                Sample *pp = pred + slot * pstride + r0 * PP + col;
                Taps8 taps;
#pragma unroll
                for (int t = 0; t < 4; t++) taps.t[t] = s_taps[yv & 15][t];
                if ((xv & 15) != 0) {
                  fir_run(tmp + (xint ? j >> 1 : j) * tstride + (iyr + 1 + r0) * TP + col, TP, 8, taps, [&](int r, int sum) {
                    pp[r * PP] = (Sample)clip3i((int)(int16_t)((sum + off2) >> sh2), 0, maxv);
                  });
                } else {
                  fir_run(sref + (iyr + 1 + r0) * RP + co + 4 + ixr + col, RP, 8, taps, [&](int r, int sum) {
                    pp[r * PP] = (Sample)clip3i((int)(int16_t)((sum + 32) >> 6), 0, maxv);
                  });
                }
              }
            }
            nc = 0;                                          // the planes are complete after the step's barrier; steps 5, 6: SATD
          }
        }
        if (T != 32) {
        const int q = s - 4, j = q >> 2, kk = q & 3;
        const int xv = bx + (j - 1) * 4;
        const int ixr = (xv >> 4) - mfx, fxj = xv & 15;
        if (kk == 0) {         // horizontal plane of x index j (+ the candidate on an integer row, if any)
          if (fxj == 0) continue;
          int k0 = -1;
          for (int k = 0; k < 3; k++)
            if (((by + (k - 1) * 4) & 15) == 0 && !(j == 1 && k == 1)) k0 = k;
          kind = 1; frac = fxj; col0 = ixr + 1; want_pred = k0 >= 0;
          row0 = 4 + (k0 >= 0 ? ((by + (k0 - 1) * 4) >> 4) - mfy : 0);
          nc = k0 >= 0 ? 1 : 0; dst = 9 + 3 * j + (k0 >= 0 ? k0 : 0);
        } else {
          const int k = kk - 1;
          const int yv = by + (k - 1) * 4;
          const int iyr = (yv >> 4) - mfy, fyk = yv & 15;
          if (fyk == 0 || (j == 1 && k == 1)) continue;
          kind = fxj != 0 ? 3 : 2; frac = fyk; col0 = 4 + ixr; row0 = iyr + 1; dst = 9 + 3 * j + k;
        }
        }
      }
      if (kind == 1) h_product(frac, col0, ncols, want_pred != 0, row0);
      else if (kind != 0) v_product(kind == 3, frac, col0, row0, nrows, ncols);
      team_sync<T>();
      if (nc > 0) {
        unsigned sum[4];
        satd_candidates<T>(org, OP, pq, ppitch, offs, nc, w, h, tid, s_part, sum, cstride);
        if (tid == 0) {
#pragma unroll
          for (int c = 0; c < 4; c++)
            if (c < nc) sd[dst + c] = sum[c];
        }
      }
      if (fullpel_only) break;
    }
    team_sync<T>();
    if (fullpel_only) {
      best_dist = best_cost = sd[0] >> (bitdepth - 8);
    } else {
      // order: (0,-1) (0,1) (-1,-1) (1,-1) (-1,0) (1,0) (-1,1) (1,1)
      const int qx[8] = {0, 0, -1, 1, -1, 1, -1, 1}, qy[8] = {-1, 1, -1, -1, 0, 0, 1, 1};
#pragma unroll
      for (int i = 0; i < 8; i++) consider(sd[9 + 3 * (qx[i] + 1) + qy[i] + 1], bx + qx[i] * 4, by + qy[i] * 4);
    }
    if (tid == 0) {
      res[ji].mv[0] = best_x; res[ji].mv[1] = best_y;
      res[ji].dist = best_dist; res[ji].cost = best_cost;
    }
  }
}

static int max_team_bytes(int max_area, int min_area, bool fused) {
  int best = 0;
  for (int w = 8; w <= 64; w <<= 1)
    for (int h = 8; h <= 64; h <<= 1)
      if (w * h <= max_area && w * h > min_area && subpel_team_bytes(w, h, fused) > best) best = subpel_team_bytes(w, h, fused);
  return best;
}

cudaError_t launch_subpel_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n,
                                 int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes,
                                 xvcb200_me_result *d_res, int *d_lists, cudaStream_t *side, cudaEvent_t *side_ev,
                                 int n_side, cudaEvent_t fork_ev) {
  if (n <= 0) return cudaSuccess;
  // launch configuration per device (see launch_tz_search)
  struct Cfg { int num_sms = 0, bytes0 = 0, bytes1 = 0, bytes2 = 0, occ0 = 1, occ1 = 1, occ2 = 1; };
  static std::mutex cfg_mutex;
  static Cfg cfg_by_dev[kMaxDevices];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  Cfg cfg;
  {
    std::lock_guard<std::mutex> lock(cfg_mutex);
    Cfg &c = cfg_by_dev[dev];
    if (!c.num_sms) {
      int sms = 0;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      c.bytes0 = max_team_bytes(256, 0, true);
      c.bytes1 = max_team_bytes(1024, 256, false);
      c.bytes2 = max_team_bytes(4096, 1024, false);
      cudaError_t e = cudaFuncSetAttribute(subpel_team_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * c.bytes0);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(subpel_team_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.bytes1);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(subpel_team_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, c.bytes2);
      if (e != cudaSuccess) return e;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.occ0, subpel_team_kernel<32>, 128, 4 * c.bytes0);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.occ1, subpel_team_kernel<128>, 128, c.bytes1);
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c.occ2, subpel_team_kernel<256>, 256, c.bytes2);
      if (c.occ0 < 1) c.occ0 = 1;
      if (c.occ1 < 1) c.occ1 = 1;
      if (c.occ2 < 1) c.occ2 = 1;
      c.num_sms = sms;
    }
    cfg = c;
  }
  const int num_sms = cfg.num_sms, bytes0 = cfg.bytes0, bytes1 = cfg.bytes1, bytes2 = cfg.bytes2, occ0 = cfg.occ0, occ1 = cfg.occ1,
            occ2 = cfg.occ2;
  int *counts = d_lists + kSubpelLists * (size_t)n;
  cudaError_t e = cudaMemsetAsync(counts, 0, 20 * sizeof(int), s);      // list lengths [0..14] + the three fetch counters [16..18]
  if (e != cudaSuccess) return e;
  g_launch_count += 6;
  subpel_classify_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_cus, d_jobs, n, d_ref_planes, d_res, d_lists, counts);
  subpel_concat_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, d_lists, counts);
  int *slow = d_lists + 2 * (size_t)n;
  // The three size classes and the generic kernel (4-wide blocks, vectors at the picture border) are independent
  // persistent grids, each sized to fill the register file on its own; on separate streams the next class's CTAs
  // move in as soon as CTAs of the previous one retire (its tail) instead of after its last CTA.
  const bool fork = n_side >= 2;
  cudaStream_t s1 = fork ? side[0] : s, s2 = fork ? side[1] : s, s3 = n_side >= 3 ? side[2] : s;
  if (fork) {
    cudaEventRecord(fork_ev, s);
    cudaStreamWaitEvent(s1, fork_ev, 0);
    cudaStreamWaitEvent(s2, fork_ev, 0);
    if (n_side >= 3) cudaStreamWaitEvent(s3, fork_ev, 0);
  }
  // persistent grids: as many CTAs as fit, each strides over its class list (largest blocks first)
  subpel_team_kernel<256><<<num_sms * occ2, 256, bytes2, s>>>(d_cus, d_jobs, d_lists + 13 * (size_t)n, counts + 13, counts + 16,
                                                         bytes2, bitdepth, lambda_me, orig, d_ref_planes, d_res);
  subpel_team_kernel<128><<<num_sms * occ1, 128, bytes1, s1>>>(d_cus, d_jobs, d_lists + 14 * (size_t)n, counts + 14, counts + 17, bytes1, bitdepth,
                                                           lambda_me, orig, d_ref_planes, d_res);
  subpel_team_kernel<32><<<num_sms * occ0, 128, 4 * bytes0, s2>>>(d_cus, d_jobs, d_lists + 10 * (size_t)n, counts + 10, counts + 18,
                                                              bytes0, bitdepth, lambda_me, orig, d_ref_planes, d_res);
  subpel_generic_kernel<<<num_sms * 2, 128, 0, s3>>>(d_cus, d_jobs, slow, counts + 2, bitdepth, lambda_me, orig, d_ref_planes, d_res);
  if (fork) {
    cudaEventRecord(side_ev[0], s1);
    cudaEventRecord(side_ev[1], s2);
    cudaStreamWaitEvent(s, side_ev[0], 0);
    cudaStreamWaitEvent(s, side_ev[1], 0);
    if (n_side >= 3) {
      cudaEventRecord(side_ev[2], s3);
      cudaStreamWaitEvent(s, side_ev[2], 0);
    }
  }
  return cudaGetLastError();
}

}  // namespace xvcb
