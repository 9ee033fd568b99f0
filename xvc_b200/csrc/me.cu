// Motion estimation: InterSearch::MotionEstNormal (inter_search.cc:606-662) for a batch of
// (CU, reference picture) jobs.
//
//   tz_search_kernel   TzSearch::Search (inter_tz_search.cc:84-171), one WARP per job.
//   subpel_kernel      InterSearch::SubpelSearch / GetSubpelDist (inter_search.cc:893-964), one CTA per job.
//   full_search_kernel InterSearch::FullSearch (inter_search.cc:853-891), one warp per job.
//
// Exactness of the parallel search.  The reference walks its candidate list in order and
// keeps a candidate only when `cost < cost_best` (strict), where cost = dist + rate >= dist;
// its `dist >= cost_best` test is therefore a pure shortcut.  For any ordered list the final
// state equals: best = first candidate attaining the minimum cost over the list, kept only if
// that minimum is below the incoming best; last_position / last_range are those of that
// candidate; "changed" = that minimum is below the incoming best.  The kernels evaluate a
// whole list at once (one candidate per lane, lane index = list position) and take the
// minimum of (cost << 5 | lane) -- the same winner, found in parallel.
#include "xvcb_interp.cuh"
#include "xvcb_satd.cuh"

namespace xvcb {

struct MeGeom {            // per job, warp-uniform
  int w, h, x, y;          // luma block
  int rows, rstep;         // rows visited by the metric (kSadFast: every second row, sample_metric.cc:194-199)
  int lpw;                 // log2(pairs per row)
  int G, lG;               // lanes per candidate, log2
  int bd_shift, fast;
  int mvpx, mvpy, down;
  uint32_t lambda;
};

struct TzBest {
  int x, y;
  uint32_t cost;
  int last_pos, last_range;
};

__device__ __forceinline__ uint32_t ld_pair(const Sample *p) {
  return (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 16);
}

// Sum of |a - b| over two packed 16-bit lanes: max - min per lane, native VIMNMX.U16x2.
__device__ __forceinline__ uint32_t absdiff2(uint32_t a, uint32_t b) { return __vmaxu2(a, b) - __vminu2(a, b); }

// Where reference samples come from: the padded reference plane in global memory and, when the
// search windows of a job group fit, a copy of their bounding box staged in shared memory
// (rows of `spw` 32-bit words, two samples per word; spw is odd so that lanes reading rows
// 5*k apart -- the raster grid -- fall into distinct banks).
struct RefSrc {
  const Sample *plane; int gpitch;      // sample (0,0) of the reference luma plane
  const uint32_t *sm; int spw;          // staged box, null when not staged
  int rx0, ry0, rx1, ry1;               // staged box in picture coordinates, [rx0,rx1) x [ry0,ry1), rx0 even
};

__device__ __forceinline__ bool block_staged(const RefSrc &src, const MeGeom &g, int cx, int cy) {
  const int X = g.x + cx, Y = g.y + cy;
  return src.sm != nullptr && X >= src.rx0 && X + g.w <= src.rx1 && Y >= src.ry0 && Y + g.h <= src.ry1;
}

// Lane j holds candidate j (cx, cy, valid) of a list of K <= 32 candidates; returns in lane j
// the metric value (SampleMetric::Compare kSad / kSadFast incl. the bit-depth shift) of candidate j.
// G lanes share one candidate (G = min(32, pairs)), 32/G candidates per pass.
template <int P>
__device__ __forceinline__ uint32_t eval_candidates(const MeGeom &g, const uint32_t (&o)[P], const RefSrc &src,
                                                    int cx, int cy, bool valid, int K, int lane) {
  const int NG = 32 >> g.lG;                 // candidates per pass
  const int gl = lane & (g.G - 1), grp = lane >> g.lG;
  uint32_t mine = 0xffffffffu;
  for (int c0 = 0; c0 < K; c0 += NG) {
    const int c = c0 + grp;
    const int sx = __shfl_sync(XVCB_FULL, cx, c & 31);
    const int sy = __shfl_sync(XVCB_FULL, cy, c & 31);
    const bool sv = __shfl_sync(XVCB_FULL, (int)valid, c & 31) && c < K;
    uint32_t acc = 0;
    if (sv) {
      if (block_staged(src, g, sx, sy)) {
        const int ox = g.x + sx - src.rx0, oy = g.y + sy - src.ry0;
#pragma unroll
        for (int k0 = 0; k0 < P; k0 += 8) {
          uint32_t packed = 0;   // up to 8 x 4095 per 16-bit lane: no carry between the lanes
#pragma unroll
          for (int k = k0; k < (k0 + 8 < P ? k0 + 8 : P); k++) {
            const int q = gl + (k << g.lG);
            const int row = q >> g.lpw, col = q & ((1 << g.lpw) - 1);
            const int sxo = ox + col * 2;
            const uint32_t *wp = src.sm + (oy + row * g.rstep) * src.spw + (sxo >> 1);
            packed += absdiff2(o[k], __funnelshift_r(wp[0], wp[1], (sxo & 1) << 4));
          }
          acc += (packed & 0xffff) + (packed >> 16);
        }
      } else {
        const Sample *r = src.plane + (g.y + sy) * src.gpitch + g.x + sx;
#pragma unroll
        for (int k0 = 0; k0 < P; k0 += 8) {
          uint32_t packed = 0;
#pragma unroll
          for (int k = k0; k < (k0 + 8 < P ? k0 + 8 : P); k++) {
            const int q = gl + (k << g.lG);
            const int row = q >> g.lpw, col = q & ((1 << g.lpw) - 1);
            packed += absdiff2(o[k], ld_pair(r + row * g.rstep * src.gpitch + col * 2));
          }
          acc += (packed & 0xffff) + (packed >> 16);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1)
      if (off < g.G) acc += __shfl_xor_sync(XVCB_FULL, acc, off);
    // candidate j was computed in pass j / NG by group j % NG
    const uint32_t got = __shfl_sync(XVCB_FULL, acc, (lane & (NG - 1)) << g.lG);
    if ((lane & ~(NG - 1)) == c0) mine = got;
  }
  if (!valid || lane >= K) return 0xffffffffu;
  return g.fast ? (mine * 2) >> g.bd_shift : mine >> g.bd_shift;
}

// Applies an evaluated candidate list to the running best (see the file comment).
__device__ __forceinline__ bool apply_candidates(TzBest &b, uint32_t dist, int cx, int cy, int pos, int range,
                                                 const MeGeom &g, int lane) {
  uint32_t key = 0xffffffffu;
  if (dist != 0xffffffffu) {
    const uint32_t cost = dist + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
    key = (cost << 5) | (uint32_t)lane;
  }
  const uint32_t win = __reduce_min_sync(XVCB_FULL, key);
  if (win == 0xffffffffu || (win >> 5) >= b.cost) return false;
  const int wl = win & 31;
  b.cost = win >> 5;
  b.x = __shfl_sync(XVCB_FULL, cx, wl);
  b.y = __shfl_sync(XVCB_FULL, cy, wl);
  b.last_pos = __shfl_sync(XVCB_FULL, pos, wl);
  b.last_range = __shfl_sync(XVCB_FULL, range, wl);
  return true;
}

// IsInside<Dir> (inter_tz_search.cc:278-301): the direction of a pattern point selects which
// window bounds are tested.  pos = Dir::index sum: left -1, right +1, up -3, down +3.
__device__ __forceinline__ bool inside(int x, int y, int pos, const int lo[2], const int hi[2]) {
  const int vert = pos <= -2 ? -1 : (pos >= 2 ? 1 : 0);
  const int horz = pos - 3 * vert;
  if (vert < 0 && y < lo[1]) return false;
  if (vert > 0 && y > hi[1]) return false;
  if (horz < 0 && x < lo[0]) return false;
  if (horz > 0 && x > hi[0]) return false;
  return true;
}

// Lane -> point of FullpelDiamondSearch (inter_tz_search.cc:173-210), in the reference's order.
__device__ __forceinline__ int diamond_point(int r, int lane, int &dx, int &dy, int &pos, int &rep) {
  dx = dy = pos = 0; rep = r;
  if (r == 1) {
    if (lane < 4) {
      dx = lane == 1 ? -1 : (lane == 2 ? 1 : 0);
      dy = lane == 0 ? -1 : (lane == 3 ? 1 : 0);
      pos = lane == 0 ? -3 : (lane == 1 ? -1 : (lane == 2 ? 1 : 3));
    }
    return 4;
  }
  if (r <= 8) {
    const int q = r >> 1;
    switch (lane) {
      case 0: dy = -r; pos = -3; break;
      case 1: dx = -q; dy = -q; pos = -4; rep = q; break;
      case 2: dx = q; dy = -q; pos = -2; rep = q; break;
      case 3: dx = -r; pos = -1; break;
      case 4: dx = r; pos = 1; break;
      case 5: dx = -q; dy = q; pos = 2; rep = q; break;
      case 6: dx = q; dy = q; pos = 4; rep = q; break;
      case 7: dy = r; pos = 3; break;
      default: break;
    }
    return 8;
  }
  if (lane < 4) {
    switch (lane) {
      case 0: dy = -r; pos = -3; break;
      case 1: dx = -r; pos = -1; break;
      case 2: dx = r; pos = 1; break;
      default: dy = r; pos = 3; break;
    }
  } else if (lane < 16) {
    const int i = ((lane - 4) >> 2) + 1, a = i * (r >> 2), bb = r - a;
    switch ((lane - 4) & 3) {
      case 0: dx = -a; dy = -bb; pos = -4; break;
      case 1: dx = a; dy = -bb; pos = -2; break;
      case 2: dx = -a; dy = bb; pos = 2; break;
      default: dx = a; dy = bb; pos = 4; break;
    }
  }
  return 16;
}

// Lane -> point of FullpelNeighborPointSearch (inter_tz_search.cc:212-259).
__device__ __forceinline__ int two_point(int last_pos, int lane, int &dx, int &dy, int &pos) {
  // {dx0,dy0,pos0, dx1,dy1,pos1} per last_position -4..4
  int t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
  switch (last_pos) {
    case -4: t0 = -1; t2 = -1; t4 = -1; t5 = -3; break;
    case -3: t0 = -1; t1 = -1; t2 = -4; t3 = 1; t4 = -1; t5 = -2; break;
    case -2: t1 = -1; t2 = -3; t3 = 1; t5 = 1; break;
    case -1: t0 = -1; t1 = 1; t2 = 2; t3 = -1; t4 = -1; t5 = -4; break;
    case 1: t0 = 1; t1 = -1; t2 = -2; t3 = 1; t4 = 1; t5 = 4; break;
    case 2: t0 = -1; t2 = -1; t4 = 1; t5 = 3; break;
    case 3: t0 = -1; t1 = 1; t2 = 2; t3 = 1; t4 = 1; t5 = 4; break;
    case 4: t0 = 1; t2 = 1; t4 = 1; t5 = 3; break;
    default: return 0;
  }
  dx = lane == 1 ? t3 : t0; dy = lane == 1 ? t4 : t1; pos = lane == 1 ? t5 : t2;
  return 2;
}

// State of one search between its phases (kept in global scratch, L2 resident).
struct TzJobState {
  int bx, by; uint32_t cost; int last_pos, last_range;
  int lo[2], hi[2], slo[2], shi[2];
  uint32_t evals;
  int need_raster;
};

template <int P>
__device__ __forceinline__ void load_orig_regs(const MeGeom &g, PlaneView orig, int lane, uint32_t (&o)[P]) {
  const int gl = lane & (g.G - 1);
  const Sample *ob = orig.base + g.y * orig.pitch + g.x;
#pragma unroll
  for (int k = 0; k < P; k++) {
    const int q = gl + (k << g.lG);
    const int row = q >> g.lpw, col = q & ((1 << g.lpw) - 1);
    o[k] = ld_pair(ob + row * g.rstep * orig.pitch + col * 2);
  }
}

template <int P>
__device__ __forceinline__ void neighbour_points(const MeGeom &g, const uint32_t (&o)[P], const RefSrc &src, TzBest &b,
                                                 const int lo[2], const int hi[2], uint32_t &evals, int lane) {
  if (b.last_range != 1) return;
  b.last_range = 0;
  int dx = 0, dy = 0, pos = 0;
  const int K = two_point(b.last_pos, lane, dx, dy, pos);
  if (K == 0) return;
  const int cx = b.x + dx, cy = b.y + dy;
  const bool valid = lane < K && inside(cx, cy, pos, lo, hi);
  evals += __popc(__ballot_sync(XVCB_FULL, valid));
  const uint32_t d = eval_candidates<P>(g, o, src, cx, cy, valid, K, lane);
  apply_candidates(b, d, cx, cy, pos, 1, g, lane);
}

// Phase 1 of TzSearch::Search: start points, first diamond pass, 2-point refinement
// (inter_tz_search.cc:102-144).
template <int P>
__device__ void tz_phase1(const MeGeom &g, const xvcb200_cu &cu, const xvcb200_me_job &job, int pic_w, int pic_h,
                          const uint32_t (&o)[P], const RefSrc &src, int lane, TzJobState &st) {
  const int range = job.search_range;
  int lo[2], hi[2], slo[2], shi[2];
  min_max_mv(g.x, g.y, pic_w, pic_h, g.mvpx, g.mvpy, range, lo, hi);
  slo[0] = lo[0]; slo[1] = lo[1]; shi[0] = hi[0]; shi[1] = hi[1];
  TzBest b;
  b.x = 0; b.y = 0; b.cost = 0xffffffffu; b.last_pos = 0; b.last_range = 0;
  uint32_t evals = 0;
  {   // predictor, zero, previous search result (:102-131)
    int px = g.mvpx, py = g.mvpy;
    clip_mv(g.x, g.y, pic_w, pic_h, px, py);
    px >>= 4; py >>= 4;
    int qx = job.prev[0] * 16, qy = job.prev[1] * 16;
    clip_mv(g.x, g.y, pic_w, pic_h, qx, qy);
    qx >>= 4; qy >>= 4;
    const bool use_zero = (px != 0 || py != 0);
    const bool use_prev = cu.depth != 0;
    const int cx = lane == 0 ? px : (lane == 1 ? 0 : qx);
    const int cy = lane == 0 ? py : (lane == 1 ? 0 : qy);
    const bool valid = lane == 0 || (lane == 1 && use_zero) || (lane == 2 && use_prev);
    const uint32_t d = eval_candidates<P>(g, o, src, cx, cy, valid, 3, lane);
    evals += 1 + use_zero + use_prev;
    uint32_t cost = 0xffffffffu;
    if (d != 0xffffffffu) cost = d + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
    const uint32_t c0 = __shfl_sync(XVCB_FULL, cost, 0), c1 = __shfl_sync(XVCB_FULL, cost, 1),
                   c2 = __shfl_sync(XVCB_FULL, cost, 2);
    b.cost = c0; b.x = px; b.y = py;
    bool moved = false;
    if (use_zero && c1 < b.cost) { b.cost = c1; b.x = 0; b.y = 0; moved = true; }
    if (use_prev) {
      if (c2 < b.cost) { b.cost = c2; b.x = qx; b.y = qy; moved = true; }
      if (moved) min_max_mv(g.x, g.y, pic_w, pic_h, b.x * 16, b.y * 16, range, slo, shi);
    }
    b.last_range = 0;
  }
  {   // first diamond pass around the start point, stops after three rounds without a hit (:133-143)
    const int bx = b.x, by = b.y;
    int misses = 0;
    for (int r = 1; r <= range; r *= 2) {
      int dx, dy, pos, rep;
      const int K = diamond_point(r, lane, dx, dy, pos, rep);
      const int cx = bx + dx, cy = by + dy;
      const bool valid = lane < K && inside(cx, cy, pos, lo, hi);
      evals += __popc(__ballot_sync(XVCB_FULL, valid));
      const uint32_t d = eval_candidates<P>(g, o, src, cx, cy, valid, K, lane);
      if (apply_candidates(b, d, cx, cy, pos, rep, g, lane)) misses = 0;
      else if (++misses >= 3) break;
    }
  }
  neighbour_points<P>(g, o, src, b, lo, hi, evals, lane);
  st.bx = b.x; st.by = b.y; st.cost = b.cost; st.last_pos = b.last_pos; st.last_range = b.last_range;
  st.lo[0] = lo[0]; st.lo[1] = lo[1]; st.hi[0] = hi[0]; st.hi[1] = hi[1];
  st.slo[0] = slo[0]; st.slo[1] = slo[1]; st.shi[0] = shi[0]; st.shi[1] = shi[1];
  st.evals = evals;
  st.need_raster = b.last_range > 5;       // kFullSearchGranularity (:91, :146)
}

// Raster scan of the window on a 5-sample grid by ONE warp (:145-155); used when the window
// is not staged in shared memory.
template <int P>
__device__ void tz_raster_warp(const MeGeom &g, const uint32_t (&o)[P], const RefSrc &src, int lane, TzJobState &st) {
  TzBest b;
  b.x = st.bx; b.y = st.by; b.cost = st.cost; b.last_pos = st.last_pos; b.last_range = 5;
  const int nx = (st.shi[0] - st.slo[0]) / 5 + 1, ny = (st.shi[1] - st.slo[1]) / 5 + 1;
  if (st.shi[0] >= st.slo[0] && st.shi[1] >= st.slo[1]) {
    const int total = nx * ny;
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      const int j = t / nx, i = t - j * nx;
      const int cx = st.slo[0] + 5 * i, cy = st.slo[1] + 5 * j;
      const uint32_t d = eval_candidates<P>(g, o, src, cx, cy, t < total, min(32, total - t0), lane);
      apply_candidates(b, d, cx, cy, b.last_pos, b.last_range, g, lane);   // CheckCostBest alone keeps last_*
    }
    st.evals += total;
  }
  st.bx = b.x; st.by = b.y; st.cost = b.cost; st.last_range = 5;
  st.need_raster = 0;
}

// Phase 3: re-centre until the centre wins (:157-168), then the result.
template <int P>
__device__ void tz_phase3(const MeGeom &g, int range, const uint32_t (&o)[P], const RefSrc &src, int lane,
                          const TzJobState &st, xvcb200_me_result *out) {
  TzBest b;
  b.x = st.bx; b.y = st.by; b.cost = st.cost; b.last_pos = st.last_pos; b.last_range = st.last_range;
  uint32_t evals = st.evals;
  while (b.last_range > 0) {
    const int bx = b.x, by = b.y;
    b.last_range = 0;
    for (int r = 1; r <= range; r *= 2) {
      int dx, dy, pos, rep;
      const int K = diamond_point(r, lane, dx, dy, pos, rep);
      const int cx = bx + dx, cy = by + dy;
      const bool valid = lane < K && inside(cx, cy, pos, st.lo, st.hi);
      evals += __popc(__ballot_sync(XVCB_FULL, valid));
      const uint32_t d = eval_candidates<P>(g, o, src, cx, cy, valid, K, lane);
      apply_candidates(b, d, cx, cy, pos, rep, g, lane);
    }
    neighbour_points<P>(g, o, src, b, st.lo, st.hi, evals, lane);
  }
  if (lane == 0) {
    out->mv_fullpel[0] = b.x; out->mv_fullpel[1] = b.y;
    out->cost_fullpel = b.cost;
    out->num_sad = evals;
  }
}

__device__ __forceinline__ MeGeom me_geom(const xvcb200_cu &cu, int bitdepth, uint32_t lambda, int mvpx, int mvpy) {
  MeGeom g;
  g.w = cu.w; g.h = cu.h; g.x = cu.x; g.y = cu.y;
  g.fast = cu.h > 8;                       // InterSearch::GetFullpelMetric, inter_search.cc:1059-1069
  g.rows = g.fast ? cu.h >> 1 : cu.h;
  g.rstep = g.fast ? 2 : 1;
  g.lpw = ilog2i(cu.w) - 1;
  const int pairs = g.rows << g.lpw;
  g.G = pairs < 32 ? pairs : 32;
  g.lG = ilog2i(g.G);
  g.bd_shift = bitdepth - 8;
  g.mvpx = mvpx; g.mvpy = mvpy;
  g.down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  g.lambda = lambda;
  return g;
}

// which: 1 = phase 1 (+ raster and phase 3 when `through`), 3 = phase 3
template <int P>
__device__ void tz_job_phase(int which, bool through, const MeGeom &g, const xvcb200_cu &cu, const xvcb200_me_job &job,
                             PlaneView orig, int pic_w, int pic_h, const RefSrc &src, int lane, TzJobState *st_g,
                             xvcb200_me_result *out) {
  uint32_t o[P];
  load_orig_regs<P>(g, orig, lane, o);
  TzJobState st;
  if (which == 1) {
    tz_phase1<P>(g, cu, job, pic_w, pic_h, o, src, lane, st);
    if (through) {
      if (st.need_raster) tz_raster_warp<P>(g, o, src, lane, st);
      tz_phase3<P>(g, job.search_range, o, src, lane, st, out);
    } else if (lane == 0) {
      *st_g = st;
    }
  } else {
    st = *st_g;
    tz_phase3<P>(g, job.search_range, o, src, lane, st, out);
  }
}

__device__ void tz_job_dispatch(int which, bool through, const MeGeom &g, const xvcb200_cu &cu, const xvcb200_me_job &job,
                                PlaneView orig, int pic_w, int pic_h, const RefSrc &src, int lane, TzJobState *st_g,
                                xvcb200_me_result *out) {
  const int pairs = g.rows << g.lpw;
  switch (pairs >> 5) {
    case 0: case 1: tz_job_phase<1>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
    case 2: tz_job_phase<2>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
    case 4: tz_job_phase<4>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
    case 8: tz_job_phase<8>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
    case 16: tz_job_phase<16>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
    default: tz_job_phase<32>(which, through, g, cu, job, orig, pic_w, pic_h, src, lane, st_g, out); break;
  }
}

// One raster candidate column for 32 candidate rows (lane = row): sum of |orig - ref| over the
// block rows the metric visits.  `rp` = this lane's first reference word, `so` = the original
// block as packed pairs in shared memory (broadcast reads), shift = 0 / 16 for even / odd x.
template <int LPW>
__device__ __forceinline__ uint32_t raster_sad(const uint32_t *rp, int row_words, const uint32_t *so, int rows,
                                               int shift) {
  constexpr int PW = 1 << LPW;
  uint32_t total = 0;
  for (int r = 0; r < rows; r++) {
    uint32_t prev = rp[0];
#pragma unroll
    for (int c0 = 0; c0 < PW; c0 += 16) {
      uint32_t acc = 0;       // <= 16 x 4095 per 16-bit lane
#pragma unroll
      for (int c = c0; c < (c0 + 16 < PW ? c0 + 16 : PW); c++) {
        const uint32_t nxt = rp[c + 1];
        acc += absdiff2(so[c], __funnelshift_r(prev, nxt, shift));
        prev = nxt;
      }
      total += (acc & 0xffff) + (acc >> 16);
    }
    rp += row_words;
    so += PW;
  }
  return total;
}

constexpr int kTzThreads = 256;
constexpr int kTzWarps = kTzThreads / 32;

struct TzGroup { int first, count; };    // run of entries in job_index: jobs sharing a reference picture and a CTU

// TzSearch::Search for job groups.  One persistent CTA per SM; per group: bounding box of the
// search windows -> shared memory, phase 1 per warp, raster scans CTA-wide (one candidate per
// lane, no reductions in the inner loop), phase 3 per warp.
__global__ void __launch_bounds__(kTzThreads, 1)
tz_search_kernel(const xvcb200_cu *__restrict__ cus, const xvcb200_me_job *__restrict__ jobs,
                 const int *__restrict__ job_index, const TzGroup *__restrict__ groups, int n_groups,
                 int *__restrict__ counter, int bitdepth, uint32_t lambda, PlaneView orig,
                 const PlaneView *__restrict__ ref_planes, xvcb200_me_result *__restrict__ res,
                 TzJobState *__restrict__ states, int region_budget_words) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *s_orig = smem;                       // 1024 words: original block of the job being raster-scanned
  uint32_t *s_region = smem + 1024;
  __shared__ int s_group, s_next, s_box[4];
  __shared__ unsigned long long s_red[kTzWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (;;) {
    __syncthreads();                             // previous group is completely done with shared memory
    if (tid == 0) {
      s_group = atomicAdd(counter, 1);
      s_next = 0;
      s_box[0] = s_box[1] = 1 << 30; s_box[2] = s_box[3] = -(1 << 30);
    }
    __syncthreads();
    const int grp = s_group;
    if (grp >= n_groups) break;
    const TzGroup G = groups[grp];
    const PlaneView ref = ref_planes[jobs[job_index[G.first]].ref_slot];

    // bounding box of the jobs' search windows (block extent included)
    for (int k = tid; k < G.count; k += kTzThreads) {
      const xvcb200_me_job job = jobs[job_index[G.first + k]];
      const xvcb200_cu cu = cus[job.cu];
      int lo[2], hi[2];
      min_max_mv(cu.x, cu.y, ref.width, ref.height, job.mvp[0], job.mvp[1], job.search_range, lo, hi);
      atomicMin(&s_box[0], cu.x + lo[0]); atomicMin(&s_box[1], cu.y + lo[1]);
      atomicMax(&s_box[2], cu.x + hi[0] + cu.w); atomicMax(&s_box[3], cu.y + hi[1] + cu.h);
    }
    __syncthreads();
    RefSrc src;
    src.plane = ref.base; src.gpitch = ref.pitch;
    src.rx0 = s_box[0] & ~7; src.ry0 = s_box[1]; src.rx1 = s_box[2]; src.ry1 = s_box[3];
    const int bw = src.rx1 - src.rx0, bh = src.ry1 - src.ry0;
    src.spw = ((bw + 1) / 2 + 1) | 1;
    const bool staged = (long long)src.spw * bh <= region_budget_words;
    src.sm = staged ? s_region : nullptr;
    if (staged) {     // 16-byte global loads (rx0 is a multiple of 8 samples), 4-byte shared stores
      const int cpr = (bw + 7) >> 3;
      for (int idx = tid; idx < bh * cpr; idx += kTzThreads) {
        const int row = idx / cpr, ch = idx - row * cpr;
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(ref.base + (src.ry0 + row) * ref.pitch + src.rx0) + ch);
        uint32_t *d = s_region + row * src.spw + ch * 4;
        const int left = src.spw - ch * 4;
        d[0] = v.x;
        if (left > 1) d[1] = v.y;
        if (left > 2) d[2] = v.z;
        if (left > 3) d[3] = v.w;
      }
    }
    __syncthreads();

    // phase 1 (and everything else when the windows are not staged): one warp per job
    for (;;) {
      int k = 0;
      if (lane == 0) k = atomicAdd(&s_next, 1);
      k = __shfl_sync(XVCB_FULL, k, 0);
      if (k >= G.count) break;
      const int ji = job_index[G.first + k];
      const xvcb200_me_job job = jobs[ji];
      const xvcb200_cu cu = cus[job.cu];
      const MeGeom g = me_geom(cu, bitdepth, lambda, job.mvp[0], job.mvp[1]);
      tz_job_dispatch(1, !staged, g, cu, job, orig, ref.width, ref.height, src, lane, &states[ji], &res[ji]);
    }
    if (!staged) continue;
    __threadfence_block();
    __syncthreads();

    // raster scans, CTA-wide, one job after the other
    for (int k = 0; k < G.count; k++) {
      const int ji = job_index[G.first + k];
      TzJobState *stp = &states[ji];
      if (!stp->need_raster) continue;           // uniform: every thread reads the same word
      const xvcb200_me_job job = jobs[ji];
      const xvcb200_cu cu = cus[job.cu];
      const MeGeom g = me_geom(cu, bitdepth, lambda, job.mvp[0], job.mvp[1]);
      const int slox = stp->slo[0], sloy = stp->slo[1], shix = stp->shi[0], shiy = stp->shi[1];
      const uint32_t cost_in = stp->cost;
      const int nx = (shix - slox) / 5 + 1, ny = (shiy - sloy) / 5 + 1;
      const bool nonempty = shix >= slox && shiy >= sloy;
      const bool fits = nonempty && g.x + slox >= src.rx0 && g.x + shix + g.w <= src.rx1 && g.y + sloy >= src.ry0 &&
                        g.y + shiy + g.h <= src.ry1;
      if (!fits) {       // window moved by the start points: one warp scans it from global memory
        if (warp == 0) {
          TzJobState st = *stp;
          // (registers for the largest block class; the scan is the rare path)
          const int pairs = g.rows << g.lpw;
#define XVCB_RW(PP) { uint32_t o[PP]; load_orig_regs<PP>(g, orig, lane, o); tz_raster_warp<PP>(g, o, src, lane, st); }
          switch (pairs >> 5) {
            case 0: case 1: XVCB_RW(1) break;
            case 2: XVCB_RW(2) break;
            case 4: XVCB_RW(4) break;
            case 8: XVCB_RW(8) break;
            case 16: XVCB_RW(16) break;
            default: XVCB_RW(32) break;
          }
#undef XVCB_RW
          if (lane == 0) *stp = st;
        }
        __threadfence_block();
        __syncthreads();
        continue;
      }
      // original block -> shared memory as packed pairs [row][pair]
      const int pw = 1 << g.lpw;
      for (int q = tid; q < (g.rows << g.lpw); q += kTzThreads) {
        const int row = q >> g.lpw, col = q & (pw - 1);
        s_orig[q] = ld_pair(orig.base + (g.y + row * g.rstep) * orig.pitch + g.x + col * 2);
      }
      __syncthreads();
      const int passes = (ny + 31) >> 5;
      uint32_t best_cost = 0xffffffffu, best_t = 0;
      for (int task = warp; task < nx * passes; task += kTzWarps) {
        const int i = task % nx, j = (task / nx) * 32 + lane;
        const int jj = min(j, ny - 1);             // idle lanes recompute the last row (no stray reads)
        const int cx = slox + 5 * i, cy = sloy + 5 * jj;
        const int ox = g.x + cx - src.rx0, oy = g.y + cy - src.ry0;
        const uint32_t *rp = s_region + oy * src.spw + (ox >> 1);
        const int shift = (ox & 1) << 4, rw = g.rstep * src.spw;
        uint32_t sad;
        switch (g.lpw) {
          case 1: sad = raster_sad<1>(rp, rw, s_orig, g.rows, shift); break;
          case 2: sad = raster_sad<2>(rp, rw, s_orig, g.rows, shift); break;
          case 3: sad = raster_sad<3>(rp, rw, s_orig, g.rows, shift); break;
          case 4: sad = raster_sad<4>(rp, rw, s_orig, g.rows, shift); break;
          default: sad = raster_sad<5>(rp, rw, s_orig, g.rows, shift); break;
        }
        if (j < ny) {
          const uint32_t dist = g.fast ? (sad * 2) >> g.bd_shift : sad >> g.bd_shift;
          const uint32_t cost = dist + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
          const uint32_t t = (uint32_t)(j * nx + i);                      // position in the reference's scan order
          if (cost < best_cost || (cost == best_cost && t < best_t)) { best_cost = cost; best_t = t; }
        }
      }
      unsigned long long key = ((unsigned long long)best_cost << 32) | best_t;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long other = __shfl_xor_sync(XVCB_FULL, key, off);
        key = other < key ? other : key;
      }
      if (lane == 0) s_red[warp] = key;
      __syncthreads();
      if (tid == 0) {
        unsigned long long m = s_red[0];
        for (int w2 = 1; w2 < kTzWarps; w2++) m = s_red[w2] < m ? s_red[w2] : m;
        const uint32_t c = (uint32_t)(m >> 32), t = (uint32_t)m;
        if (c < cost_in) {                         // strict: ties keep the earlier best (:266-268)
          stp->cost = c;
          stp->bx = slox + 5 * (int)(t % nx);
          stp->by = sloy + 5 * (int)(t / nx);
        }
        stp->last_range = 5;
        stp->evals += nx * ny;
        stp->need_raster = 0;
      }
      __threadfence_block();
      __syncthreads();
    }

    // phase 3: one warp per job
    if (tid == 0) s_next = 0;
    __syncthreads();
    for (;;) {
      int k = 0;
      if (lane == 0) k = atomicAdd(&s_next, 1);
      k = __shfl_sync(XVCB_FULL, k, 0);
      if (k >= G.count) break;
      const int ji = job_index[G.first + k];
      const xvcb200_me_job job = jobs[ji];
      const xvcb200_cu cu = cus[job.cu];
      const MeGeom g = me_geom(cu, bitdepth, lambda, job.mvp[0], job.mvp[1]);
      tz_job_dispatch(3, false, g, cu, job, orig, ref.width, ref.height, src, lane, &states[ji], &res[ji]);
    }
  }
}

// ---------------------------------------------------------------- sub-pel search
__global__ void __launch_bounds__(128) subpel_kernel(const xvcb200_cu *__restrict__ cus,
                                                     const xvcb200_me_job *__restrict__ jobs, int n, int bitdepth,
                                                     uint32_t lambda, PlaneView orig,
                                                     const PlaneView *__restrict__ ref_planes,
                                                     xvcb200_me_result *__restrict__ res) {
  __shared__ int16_t tmp[64 * 71];
  __shared__ Sample pred[64 * 64];
  __shared__ Sample org[64 * 64];
  __shared__ unsigned part[4];
  const int tid = threadIdx.x;
  const int ji = blockIdx.x;
  const xvcb200_me_job job = jobs[ji];
  const xvcb200_cu cu = cus[job.cu];
  const PlaneView ref = ref_planes[job.ref_slot];
  const int w = cu.w, h = cu.h;
  for (int i = tid; i < w * h; i += 128) {
    const int y = i / w, x = i - y * w;
    org[y * 64 + x] = orig.base[(cu.y + y) * orig.pitch + cu.x + x];
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

cudaError_t launch_tz_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n, int bitdepth,
                             uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes, xvcb200_me_result *d_res,
                             const int *d_job_index, const void *d_groups, int n_groups, void *d_states, int *d_counter) {
  if (n <= 0 || n_groups <= 0) return cudaSuccess;
  static int smem_bytes = 0, num_sms = 0;
  if (!smem_bytes) {
    int dev = 0, max_optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, tz_search_kernel);
    smem_bytes = max_optin - (int)fa.sharedSizeBytes - 1024;
    cudaError_t e = cudaFuncSetAttribute(tz_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) { smem_bytes = 0; return e; }
  }
  cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  g_launch_count++;
  const int grid = n_groups < num_sms ? n_groups : num_sms;
  tz_search_kernel<<<grid, kTzThreads, smem_bytes, s>>>(d_cus, d_jobs, d_job_index, static_cast<const TzGroup *>(d_groups),
                                                        n_groups, d_counter, bitdepth, lambda_me, orig, d_ref_planes, d_res,
                                                        static_cast<TzJobState *>(d_states), smem_bytes / 4 - 1024);
  return cudaGetLastError();
}
size_t tz_state_bytes() { return sizeof(TzJobState); }

cudaError_t launch_subpel_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n,
                                 int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes,
                                 xvcb200_me_result *d_res) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  subpel_kernel<<<n, 128, 0, s>>>(d_cus, d_jobs, n, bitdepth, lambda_me, orig, d_ref_planes, d_res);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- bi-prediction full search
// InterSearch::FullSearch (inter_search.cc:853-891): every full-pel position of the clipped
// +-range window, row-major, on the weighted original 2*orig - other_pred
// (ResidualBuffer::SubtractWeighted, sample_buffer.h:147-161).  One warp per job, one
// candidate per lane pass; metric on int16 vs Sample.
__global__ void __launch_bounds__(128) full_search_kernel(const xvcb200_cu *__restrict__ cus,
                                                          const xvcb200_fullsearch_job *__restrict__ jobs, int n,
                                                          int bitdepth, uint32_t lambda, PlaneView orig,
                                                          const PlaneView *__restrict__ planes,
                                                          xvcb200_me_result *__restrict__ res) {
  __shared__ int16_t worig[4][64 * 64];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int ji = blockIdx.x * 4 + wi;
  if (ji >= n) return;
  const xvcb200_fullsearch_job job = jobs[ji];
  const xvcb200_cu cu = cus[job.cu];
  const PlaneView ref = planes[job.ref_slot], other = planes[job.other_pred_slot];
  const int w = cu.w, h = cu.h;
  int16_t *wo = worig[wi];
  for (int i = lane; i < w * h; i += 32) {
    const int y = i / w, x = i - y * w;
    wo[y * 64 + x] = (int16_t)(2 * (int)orig.base[(cu.y + y) * orig.pitch + cu.x + x] -
                               (int)other.base[(cu.y + y) * other.pitch + cu.x + x]);
  }
  __syncwarp();
  int lo[2], hi[2];
  min_max_mv(cu.x, cu.y, ref.width, ref.height, job.center[0], job.center[1], job.range, lo, hi);
  const bool fast = h > 8;
  const int rows = fast ? h >> 1 : h, rstep = fast ? 2 : 1;
  const int down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  const int nx = hi[0] - lo[0] + 1, ny = hi[1] - lo[1] + 1, total = nx * ny;
  const Sample *ref0 = ref.base + cu.y * ref.pitch + cu.x;
  uint32_t best = 0xffffffffu;
  int bx = 0, by = 0;
  for (int t0 = 0; t0 < total; t0 += 32) {
    const int t = t0 + lane;
    uint32_t key = 0xffffffffu;
    int cx = 0, cy = 0;
    if (t < total) {
      cy = lo[1] + t / nx; cx = lo[0] + t % nx;
      const Sample *r = ref0 + cy * ref.pitch + cx;
      uint32_t sad = 0;
      for (int y = 0; y < rows; y++)
        for (int x = 0; x < w; x++) sad += abs((int)wo[y * rstep * 64 + x] - (int)__ldg(r + y * rstep * ref.pitch + x));
      const uint32_t dist = fast ? (sad * 2) >> (bitdepth - 8) : sad >> (bitdepth - 8);
      const uint32_t cost = dist + ((lambda * mvd_bits_fullpel(job.mvp[0], job.mvp[1], cx, cy, down)) >> 16);
      key = (cost << 5) | lane;
    }
    const uint32_t win = __reduce_min_sync(XVCB_FULL, key);
    if (win != 0xffffffffu && (win >> 5) < best) {
      best = win >> 5;
      bx = __shfl_sync(XVCB_FULL, cx, win & 31);
      by = __shfl_sync(XVCB_FULL, cy, win & 31);
    }
  }
  if (lane == 0) {
    res[ji].mv_fullpel[0] = bx; res[ji].mv_fullpel[1] = by;
    res[ji].mv[0] = bx * 16; res[ji].mv[1] = by * 16;
    res[ji].cost_fullpel = best; res[ji].dist = 0; res[ji].cost = best; res[ji].num_sad = total;
  }
}

cudaError_t launch_full_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_fullsearch_job *d_jobs, int n,
                               int bitdepth, uint32_t lambda_me, PlaneView orig, const PlaneView *d_planes,
                               xvcb200_me_result *d_res) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  full_search_kernel<<<(n + 3) / 4, 128, 0, s>>>(d_cus, d_jobs, n, bitdepth, lambda_me, orig, d_planes, d_res);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- picture pipeline glue
// jobs for (CU i, list l, ref_idx 0) with the CU's mv[l] as predictor
__global__ void make_me_jobs_kernel(const xvcb200_cu *__restrict__ cus, int n, int nl, int slot0, int slot1, int range0,
                                    int range1, xvcb200_me_job *__restrict__ jobs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * nl) return;
  const int c = i / nl, l = i - c * nl;
  xvcb200_me_job j;
  j.cu = c; j.ref_slot = l ? slot1 : slot0; j.search_range = l ? range1 : range0;
  j.mvp[0] = cus[c].mv[l][0]; j.mvp[1] = cus[c].mv[l][1];
  j.prev[0] = 0; j.prev[1] = 0; j.list = l;
  jobs[i] = j;
}

cudaError_t launch_make_me_jobs(cudaStream_t s, const xvcb200_cu *d_cus, int n, int nl, const int ref_slot[2],
                                const int range[2], xvcb200_me_job *d_jobs) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  make_me_jobs_kernel<<<(n * nl + 255) / 256, 256, 0, s>>>(d_cus, n, nl, ref_slot[0], ref_slot[1], range[0], range[1], d_jobs);
  return cudaGetLastError();
}

// best list by sub-pel cost (ties -> L0); the loser is cleared like SearchRefIdx does for
// uni-prediction (inter_search.cc:476-480)
__global__ void me_decide_kernel(xvcb200_cu *__restrict__ cus, int n, int nl, const xvcb200_me_result *__restrict__ res) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int best = (nl == 2 && res[2 * i + 1].cost < res[2 * i].cost) ? 1 : 0;
  for (int l = 0; l < 2; l++) {
    cus[i].ref_idx[l] = (int8_t)(l == best ? 0 : -1);
    cus[i].mv[l][0] = l == best ? res[i * nl + l].mv[0] : 0;
    cus[i].mv[l][1] = l == best ? res[i * nl + l].mv[1] : 0;
  }
}

cudaError_t launch_me_decide(cudaStream_t s, xvcb200_cu *d_cus, int n, int nl, const xvcb200_me_result *d_res) {
  if (n <= 0) return cudaSuccess;
  g_launch_count++;
  me_decide_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_cus, n, nl, d_res);
  return cudaGetLastError();
}

}  // namespace xvcb
