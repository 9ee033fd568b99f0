// Motion estimation: InterSearch::MotionEstNormal (inter_search.cc:606-662) for a batch of
// (CU, reference picture) jobs.
//
//   tz_search_kernel   TzSearch::Search (inter_tz_search.cc:84-171): persistent CTAs, a CTU's jobs
//                      per step, searched by warp teams (16 / 4 / 1 warps by block size).
//   (sub-pel search: subpel.cu)
//   full_search_kernel InterSearch::FullSearch (inter_search.cc:853-891), one warp per job.
//
// Exactness of the parallel search.  The reference walks its candidate list in order and
// keeps a candidate only when `cost < cost_best` (strict), where cost = dist + rate >= dist;
// its `dist >= cost_best` test is therefore a pure shortcut.  For any ordered list the final
// state equals: best = first candidate attaining the minimum cost over the list, kept only if
// that minimum is below the incoming best; last_position / last_range are those of that
// candidate; "changed" = that minimum is below the incoming best.  The kernels evaluate a
// whole list at once and take the minimum of (cost << 5 | list position) -- the same winner,
// found in parallel.  The rounds of one diamond pass share their centre, so their lists do not
// depend on each other either: a whole pass is evaluated at once and the reference's
// round-by-round decisions are replayed on the stored costs (see SearchEval).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

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

// Sum of |orig - ref| over `nrows` rows of PW = 2^LPW sample pairs.  rp: first (4-byte aligned)
// reference word of the first row; a row of pairs starting at an odd sample is assembled from
// two neighbouring words with a funnel shift (shift = 16), at an even sample shift = 0.
// so: the original block as packed pairs.  GLOBAL: rp is global memory (read-only path).
template <int LPW, bool GLOBAL>
__device__ __forceinline__ uint32_t sad_rows(const uint32_t *rp, int ref_row_words, const uint32_t *so, int so_row_words,
                                             int nrows, int shift) {
  constexpr int PW = 1 << LPW;
  uint32_t total = 0;
#pragma unroll 4
  for (int r = 0; r < nrows; r++) {
    uint32_t prev = GLOBAL ? __ldg(rp) : rp[0];
#pragma unroll
    for (int c0 = 0; c0 < PW; c0 += 16) {
      uint32_t acc = 0;       // <= 16 x 4095 per 16-bit lane: no carry between the lanes
#pragma unroll
      for (int c = c0; c < (c0 + 16 < PW ? c0 + 16 : PW); c++) {
        const uint32_t nxt = GLOBAL ? __ldg(rp + c + 1) : rp[c + 1];
        acc += absdiff2(so[c], __funnelshift_r(prev, nxt, shift));
        prev = nxt;
      }
      total += (acc & 0xffff) + (acc >> 16);
    }
    rp += ref_row_words;
    so += so_row_words;
  }
  return total;
}

// The same from global memory (read-only path) for candidates outside the staged box: rare, so
// one compact out-of-line routine for all shapes.
__device__ __noinline__ uint32_t sad_rows_global(const uint32_t *rp, int ref_row_words, const uint32_t *so,
                                                 int so_row_words, int nrows, int shift, int pw) {
  uint32_t total = 0;
  for (int r = 0; r < nrows; r++) {
    uint32_t prev = __ldg(rp), acc = 0;
    for (int c = 0; c < pw; c++) {
      const uint32_t nxt = __ldg(rp + c + 1);
      acc += absdiff2(so[c], __funnelshift_r(prev, nxt, shift));
      prev = nxt;
      if ((c & 15) == 15) { total += (acc & 0xffff) + (acc >> 16); acc = 0; }
    }
    total += (acc & 0xffff) + (acc >> 16);
    rp += ref_row_words;
    so += so_row_words;
  }
  return total;
}

__device__ __forceinline__ uint32_t sad_rows_lpw(int lpw, const uint32_t *rp, int ref_row_words, const uint32_t *so,
                                                 int so_row_words, int nrows, int shift) {
  switch (lpw) {
    case 1: return sad_rows<1, false>(rp, ref_row_words, so, so_row_words, nrows, shift);
    case 2: return sad_rows<2, false>(rp, ref_row_words, so, so_row_words, nrows, shift);
    case 3: return sad_rows<3, false>(rp, ref_row_words, so, so_row_words, nrows, shift);
    case 4: return sad_rows<4, false>(rp, ref_row_words, so, so_row_words, nrows, shift);
    default: return sad_rows<5, false>(rp, ref_row_words, so, so_row_words, nrows, shift);
  }
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

// State of one search between its phases (global scratch, L2 resident).
struct TzJobState {
  int bx, by; uint32_t cost; int last_pos, last_range;
  int lo[2], hi[2], slo[2], shi[2];
  uint32_t evals;
  int need_raster;
};

// Search pattern of FullpelDiamondSearch as a table: for the three pattern classes (radius 1:
// 4 points, 2..8: 8 points, >= 16: 16 points) point k in units of 1, r/2, r/4, its direction
// index and whether its recorded range is r/2 (the diagonal points of the 8-point pattern) --
// packed (ux + 8) | (uy + 8) << 4 | (pos + 8) << 8 | half << 12.
__device__ __forceinline__ int pack_pattern_point(int cls, int k) {
  const int r = cls == 0 ? 1 : (cls == 1 ? 2 : 16), unit = cls == 2 ? 4 : 1;
  int dx, dy, pos, rep;
  diamond_point(r, k, dx, dy, pos, rep);
  return (dx / unit + 8) | ((dy / unit + 8) << 4) | ((pos + 8) << 8) | ((rep != r ? 1 : 0) << 12);
}
// point k of the round with radius 2^ri around (ax, ay)
__device__ __forceinline__ void pattern_point(const int *s_pat, int ri, int k, int ax, int ay, int &cx, int &cy, int &pos,
                                              int &rep) {
  const int cls = ri == 0 ? 0 : (ri <= 3 ? 1 : 2);
  const int pk = s_pat[cls * 16 + k];
  const int r = 1 << ri, unit = cls == 0 ? 1 : (cls == 1 ? r >> 1 : r >> 2);
  cx = ax + ((pk & 15) - 8) * unit;
  cy = ay + (((pk >> 4) & 15) - 8) * unit;
  pos = ((pk >> 8) & 15) - 8;
  rep = (pk >> 12) ? r >> 1 : r;
}

// The candidate evaluator of the search.  All rounds of a diamond pass share their centre, so
// their candidates do not depend on each other: the whole pass (4 + 8 + 8 + 8 + 16 + ... points,
// "slots" in the reference's order) is evaluated at once and the reference's sequential
// decisions (strict compares, three-miss rule, evaluation counts) are then replayed on the stored
// costs -- one long dependent chain per pass instead of one per round.
// A job is searched by a TEAM of 2^lt warps that run the control flow redundantly (identical
// state) and split the evaluation passes; a pass = 32 >> lg slots, 2^lg lanes per candidate
// (each lane sums whole block rows r = sub, sub + 2^lg, ..., the partial sums meet in lg
// shuffles).  Costs return to "holder" lanes: slot s lives in lane s & 31, register s >> 5 --
// through one shuffle per pass for a single warp, through the team's scratch (128 words of
// shared memory) for a team.
enum { kEvalDiamond, kEvalList };
struct SearchEval {
  const MeGeom &g;
  const uint32_t *so;        // original block as packed pairs; rows visited by the metric are so_row_words apart
  int so_row_words;
  const Sample *plane;       // sample (0,0) of the reference luma plane (4-byte aligned, even pitch)
  int gpitch;
  const uint32_t *sm;        // reference box staged in shared memory (rows of spw words), or null
  int spw, rx0, ry0, rx1, ry1;
  const int *s_pat;
  int lane;
  int tw, lt;                // rank of this warp in the team, log2(team size)
  uint32_t *t_cost;
  int bar_id;

  // kEvalDiamond: slots = the points of rounds 0..nrounds-1 around (ax, ay), valid if inside the
  // window [lo, hi].  kEvalList: slot k < nslots = the candidate held by lane k (ex, ey, evalid).
  // cost[q] of lane l = cost of slot 32 q + l, 0xffffffff for invalid / non-existent slots.
  // Only the slots [s_lo, nslots) are evaluated (s_lo > 0: the rest of a pass whose first rounds were evaluated
  // before; cost[] of the earlier slots is then unspecified).
  __device__ __forceinline__ void run(int mode, int lg, int s_lo, int nslots, int ax, int ay, const int lo[2], const int hi[2], int ex,
                                      int ey, bool evalid, uint32_t (&cost)[4]) {
    const int spp = 32 >> lg;
    const int npass = (nslots + spp - 1) >> (5 - lg);
    if (s_lo == 0) {
#pragma unroll
      for (int q = 0; q < 4; q++) cost[q] = 0xffffffffu;
    }
    for (int p = (s_lo >> (5 - lg)) + tw; p < npass; p += 1 << lt) {
      const int s0 = p << (5 - lg);
      const int s = s0 + (lane >> lg), sub = lane & ((1 << lg) - 1);
      int cx, cy;
      bool valid;
      if (mode == kEvalDiamond) {
        int ri, k, pos, rep;
        if (s < 4) { ri = 0; k = s; }
        else if (s < 28) { ri = 1 + ((s - 4) >> 3); k = (s - 4) & 7; }
        else { ri = 4 + ((s - 28) >> 4); k = (s - 28) & 15; }
        pattern_point(s_pat, ri, k, ax, ay, cx, cy, pos, rep);
        valid = s >= s_lo && s < nslots && inside(cx, cy, pos, lo, hi);
      } else {
        cx = __shfl_sync(XVCB_FULL, ex, s & 31);
        cy = __shfl_sync(XVCB_FULL, ey, s & 31);
        valid = __shfl_sync(XVCB_FULL, (int)evalid, s & 31) && s < nslots;
      }
      uint32_t c = 0xffffffffu;
      if (__ballot_sync(XVCB_FULL, valid)) {
        uint32_t acc = 0;
        if (valid) {
          const int X = g.x + cx, Y = g.y + cy;
          const uint32_t *op = so + sub * so_row_words;
          if (sm != nullptr && X >= rx0 && X + g.w <= rx1 && Y >= ry0 && Y + g.h <= ry1) {
            const int ox = X - rx0, oy = Y - ry0 + sub * g.rstep;
            acc = sad_rows_lpw(g.lpw, sm + oy * spw + (ox >> 1), (g.rstep * spw) << lg, op, so_row_words << lg,
                               g.rows >> lg, (ox & 1) << 4);
          } else {
            const Sample *row0 = plane + (Y + sub * g.rstep) * gpitch + (X & ~1);
            acc = sad_rows_global(reinterpret_cast<const uint32_t *>(row0), (g.rstep * gpitch << lg) >> 1, op,
                                  so_row_words << lg, g.rows >> lg, (X & 1) << 4, 1 << g.lpw);
          }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          if (off < (1 << lg)) acc += __shfl_xor_sync(XVCB_FULL, acc, off);
        if (valid) {
          const uint32_t dist = g.fast ? (acc * 2) >> g.bd_shift : acc >> g.bd_shift;
          c = dist + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
        }
      }
      if (lt == 0) {
        const int base = s0 & 31, q = s0 >> 5;
        const uint32_t v = __shfl_sync(XVCB_FULL, c, ((lane - base) << lg) & 31);
        if (lane >= base && lane < base + spp) {
          if (q == 0) cost[0] = v;
          else if (q == 1) cost[1] = v;
          else if (q == 2) cost[2] = v;
          else cost[3] = v;
        }
      } else if (sub == 0) {
        t_cost[s] = c;
      }
    }
    if (lt > 0) {
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 << lt) : "memory");
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (32 * q + lane < (npass << (5 - lg))) cost[q] = t_cost[32 * q + lane];
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(32 << lt) : "memory");
    }
  }
};

__device__ __forceinline__ MeGeom me_geom(const xvcb200_cu &cu, int bitdepth, uint32_t lambda, int mvpx, int mvpy) {
  MeGeom g;
  g.w = cu.w; g.h = cu.h; g.x = cu.x; g.y = cu.y;
  g.fast = cu.h > 8;                       // InterSearch::GetFullpelMetric, inter_search.cc:1059-1069
  g.rows = g.fast ? cu.h >> 1 : cu.h;
  g.rstep = g.fast ? 2 : 1;
  g.lpw = 30 - __clz((int)cu.w);
  const int pairs = g.rows << g.lpw;
  g.G = pairs < 32 ? pairs : 32;
  g.lG = 31 - __clz(g.G);
  g.bd_shift = bitdepth - 8;
  g.mvpx = mvpx; g.mvpy = mvpy;
  g.down = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 2 : 0;
  g.lambda = lambda;
  return g;
}

constexpr int kStageDepth = 4;
__device__ __forceinline__ void stage_box(const Sample *plane00, int pitch, int rx0, int ry0, int bh, int cpr, int spw,
                                          uint32_t *s_region, int tid, int nthreads) {
  const int total = bh * cpr;
  const uint32_t magic = 0xffffffffu / (uint32_t)cpr + 1u;    // idx / cpr == umulhi(idx, magic) for idx * cpr < 2^32
  for (int idx0 = tid; idx0 < total; idx0 += kStageDepth * nthreads) {
    uint4 v[kStageDepth];
#pragma unroll
    for (int u = 0; u < kStageDepth; u++) {            // independent 16-byte loads in flight per thread
      const int idx = idx0 + u * nthreads;
      if (idx < total) {
        const int row = (int)__umulhi((uint32_t)idx, magic), ch = idx - row * cpr;
        v[u] = __ldg(reinterpret_cast<const uint4 *>(plane00 + (ry0 + row) * pitch + rx0) + ch);
      }
    }
#pragma unroll
    for (int u = 0; u < kStageDepth; u++) {
      const int idx = idx0 + u * nthreads;
      if (idx < total) {
        const int row = (int)__umulhi((uint32_t)idx, magic), ch = idx - row * cpr;
        uint32_t *d = s_region + row * spw + ch * 4;
        const int left = spw - ch * 4;
        d[0] = v[u].x;
        if (left > 1) d[1] = v[u].y;
        if (left > 2) d[2] = v[u].z;
        if (left > 3) d[3] = v[u].w;
      }
    }
  }
}

// The staged sample box <-> its 8-sample segment sums, IN PLACE, a row per thread.  Forward:
// word i of a row (samples 2i, 2i+1) becomes (S8[2i], S8[2i+1]), S8[x] = s[x] + ... + s[x+7], by a
// sliding window over a four-word delay line; the last four words of the row stay raw samples
// (no candidate reads S8 there: x <= box width - 8).  Inverse, right to left:
// s[x] = S8[x] - S8[x+1] + s[x+8], the eight samples to the right being already restored (or
// the raw tail).  All arithmetic mod 2^16, so the round trip is exact whatever the tail holds.
// Replaces a second copy of every reference plane (its segment sums) and two stagings per group.
__device__ __forceinline__ void box_to_s8(uint32_t *s_region, int spw, int bh, int tid, int nthreads) {
  const int nconv = spw - 4;
  for (int r = tid; r < bh; r += nthreads) {
    uint32_t *row = s_region + r * spw;
    uint32_t w0 = row[0], w1 = row[1], w2 = row[2], w3 = row[3];
    uint32_t S = (w0 & 0xffffu) + (w0 >> 16) + (w1 & 0xffffu) + (w1 >> 16) + (w2 & 0xffffu) + (w2 >> 16) + (w3 & 0xffffu) + (w3 >> 16);
#pragma unroll 4
    for (int i = 0; i < nconv; i++) {
      const uint32_t w4 = row[i + 4];
      const uint32_t S1 = S - (w0 & 0xffffu) + (w4 & 0xffffu);
      row[i] = (S & 0xffffu) | (S1 << 16);
      S = S1 - (w0 >> 16) + (w4 >> 16);
      w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
  }
}
__device__ __forceinline__ void s8_to_box(uint32_t *s_region, int spw, int bh, int tid, int nthreads) {
  const int nconv = spw - 4;
  for (int r = tid; r < bh; r += nthreads) {
    uint32_t *row = s_region + r * spw;
    uint32_t w1 = row[nconv], w2 = row[nconv + 1], w3 = row[nconv + 2], w4 = row[nconv + 3];
    uint32_t Snext = (w1 & 0xffffu) + (w1 >> 16) + (w2 & 0xffffu) + (w2 >> 16) + (w3 & 0xffffu) + (w3 >> 16) + (w4 & 0xffffu) + (w4 >> 16);
#pragma unroll 4
    for (int i = nconv - 1; i >= 0; i--) {
      const uint32_t v = row[i];
      const uint32_t a = v & 0xffffu, b = v >> 16;
      const uint32_t hi = (b - Snext + (w4 >> 16)) & 0xffffu;
      const uint32_t lo = (a - b + (w4 & 0xffffu)) & 0xffffu;
      const uint32_t w0 = lo | (hi << 16);
      row[i] = w0;
      Snext = a;
      w4 = w3; w3 = w2; w2 = w1; w1 = w0;
    }
  }
}

// Lower bound of the SAD of a candidate from 8-sample segment sums (successive elimination):
// sum_rows sum_k |A8[r][k] - S8[y+r][x+8k]| <= SAD (triangle inequality per segment), with S8 =
// segment sums of the reference at every position (segment_sum_kernel) and A8 = segment sums of
// the original block, two per word.  VABSDIFF.U32 does |a - b| + c in one instruction.
// Evaluated for CH neighbouring grid columns (candidates 5 samples apart in x) of one grid row per
// lane at once: a segment sum of the original block is fetched once and used for CH candidates,
// whose S8 operands sit at compile-time offsets from one row pointer -- two instructions
// (LDS.U16 + VABSDIFF) per segment difference.
constexpr int kBoundCols = 8;
template <int NSEG>
__device__ __forceinline__ void seg_bound_cols(const uint16_t *rp, int row_stride, const uint32_t *seg, int rows,
                                               uint32_t (&lb)[kBoundCols]) {
  if (NSEG == 1) {
    for (int r = 0; r < rows; r += 2) {          // rows is a multiple of 4; a word of seg = two rows
      const uint32_t a2 = seg[r >> 1];
      const uint32_t s0 = a2 & 0xffffu, s1 = a2 >> 16;
      const uint16_t *rq = rp + row_stride;
#pragma unroll
      for (int u = 0; u < kBoundCols; u++) lb[u] = __usad((unsigned)rp[5 * u], s0, lb[u]);
#pragma unroll
      for (int u = 0; u < kBoundCols; u++) lb[u] = __usad((unsigned)rq[5 * u], s1, lb[u]);
      rp = rq + row_stride;
    }
  } else {
    for (int r = 0; r < rows; r++) {
#pragma unroll
      for (int k = 0; k < NSEG; k += 2) {
        const uint32_t a2 = seg[(r * NSEG + k) >> 1];
        const uint32_t s0 = a2 & 0xffffu, s1 = a2 >> 16;
#pragma unroll
        for (int u = 0; u < kBoundCols; u++) lb[u] = __usad((unsigned)rp[k * 8 + 5 * u], s0, lb[u]);
#pragma unroll
        for (int u = 0; u < kBoundCols; u++) lb[u] = __usad((unsigned)rp[k * 8 + 8 + 5 * u], s1, lb[u]);
      }
      rp += row_stride;
    }
  }
}

constexpr int kProfGroups = 4096;        // XVCB_TZ_PROF: per-group records kept
constexpr int kTzThreads = 512;
constexpr int kTzWarps = kTzThreads / 32;
constexpr int kMaxGroupJobs = 64;        // jobs handled per pass over a group
constexpr int kTileWords = 64 * 33;      // original CTU as packed pairs, rows padded to 33 words
constexpr int kSegWords = 256;           // segment sums of the blocks being bounded: 512 x uint16 (a tiled CTU needs <= 512)

struct TzGroup { int first, count; };    // run of entries in job_index: jobs sharing a reference picture and a CTU

// Picture pipeline: jobs are laid out [cu][J] (J = reference pictures searched per CU); `groups` then
// describes runs of the CU list `job_index` (the CUs of one CTU) and launch group G stands for run
// G / n of it searched on job column j[G % n].  J == 0: `job_index` holds job indices, one run per group.
struct TzJobSel { int J, n, j[10]; };

struct SJob {                            // one job of the current group, in shared memory
  int ji;
  short x, y; unsigned char w, h, depth, fullpel;
  int mvpx, mvpy, prevx, prevy, range;
  int slox, sloy, nx, ny;                // raster grid (valid when need != 0)
  uint32_t cost_in;
  int need;                              // 0: no raster; 1: raster, box staged; 2: raster, window outside the staged box
  int list_off, list_cnt;                // survivors in the pool (shared by the jobs of one bound batch); list_off < 0: dense scan
  int task0;                             // first bound task of the job in its batch, -1: not in the batch
  int seg_off;                           // its segment sums in s_seg
  unsigned long long key;                // best (cost << 32 | scan position) of the exact pass
};

__device__ __forceinline__ MeGeom sjob_geom(const SJob &j, int bitdepth, uint32_t lambda) {
  xvcb200_cu cu;
  cu.x = j.x; cu.y = j.y; cu.w = j.w; cu.h = j.h; cu.depth = j.depth;
  cu.flags = j.fullpel ? XVCB200_CU_FULLPEL_MV : 0;
  return me_geom(cu, bitdepth, lambda, j.mvpx, j.mvpy);
}

// TzSearch::Search (inter_tz_search.cc:84-171) for job groups.  One persistent CTA per SM takes
// groups from a counter.  A group = a run of jobs of one CTU on one reference picture, searched in
// chunks: as many consecutive jobs as have the bounding box of their search windows inside the
// shared-memory region; that box and the original CTU are staged once per chunk and shared by all
// phases of all its jobs.
//
//   first loop  per job (warp teams of 16 / 4 / 1 by block size, fetched largest first): start points,
//               the first diamond pass -- evaluated in two instalments, radii 1..8 and the rest: the
//               three-miss rule (:133-143) usually ends it inside the first -- and the 2-point step;
//               a job that needs no raster scan goes straight on to the refinement (re-centre until
//               the centre wins, :157-168) and is finished.  A job whose first pass ended far out
//               (last_range > 5) parks its state (SJob::need = 1: scan window inside the staged box,
//               2: outside -- a start point other than the predictor re-centres the window, :121-125).
//   raster      two rounds, CTA-wide: the jobs with need == 1 on the staged box, then the bounding box
//               of the OTHER windows is staged (as many as fit) and scanned the same way; what still
//               does not fit is scanned densely from global memory.  A scan (:145-155) visits the
//               window on the 5-sample grid, ONE CANDIDATE PER LANE: a warp takes grid columns (fixed
//               x: warp-uniform alignment) and 32 grid rows (5 picture rows apart; the odd row pitch
//               of the box puts them in 32 distinct banks).
//               Successive elimination (exact): first a lower bound of every candidate's SAD from
//               8-sample segment sums,  sum_rows sum_k |A8[r][k] - S8[y+r][x+8k]| <= SAD  (triangle
//               inequality per segment), hence bound_cost <= cost.  A candidate can only replace
//               the incoming best if cost < cost_in (strict compare of CheckCostBest, :266), so
//               candidates with bound_cost >= cost_in are dropped without changing the result.  The
//               box is converted to its segment sums in place for the bound pass and back; survivors
//               go to a per-CTA pool and are evaluated exactly, each spread over 2^lgw lanes by block
//               size and survivor count (per-job winners through a 64-bit shared-memory atomicMin).
//   second loop the refinement of the scanned jobs (need == 3), on the box staged last.
__global__ void __launch_bounds__(kTzThreads, 1)
tz_search_kernel(const xvcb200_cu *__restrict__ cus, const xvcb200_me_job *__restrict__ jobs,
                 const int *__restrict__ job_index, const TzGroup *__restrict__ groups, int n_groups,
                 const __grid_constant__ TzJobSel sel, int *__restrict__ counter, int bitdepth, uint32_t lambda, PlaneView orig,
                 const PlaneView *__restrict__ ref_planes,
                 xvcb200_me_result *__restrict__ res, TzJobState *__restrict__ states, int region_budget_words,
                 uint32_t *__restrict__ pool_all, int pool_cap, unsigned long long *__restrict__ prof) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t *s_tile = smem;                                   // original CTU, packed pairs, 33 words per row
  uint16_t *s_seg = reinterpret_cast<uint16_t *>(smem + kTileWords);
  SJob *s_job = reinterpret_cast<SJob *>(smem + kTileWords + kSegWords);
  uint32_t *s_region = smem + kTileWords + kSegWords + kMaxGroupJobs * (sizeof(SJob) / 4);
  __shared__ int s_group, s_box[4], s_next, s_count, s_pool_used, s_any_raster, s_batch_end, s_batch_tasks;
  __shared__ int s_cls[10], s_ncoop;
  __shared__ unsigned char s_order[kMaxGroupJobs];    // jobs of the group, largest block first
  __shared__ uint32_t s_tcost[kTzWarps / 4][128];     // per warp team: the costs of one evaluation
  __shared__ int s_pat[48];
  __shared__ int s_fetch[kTzWarps];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *pool = pool_all + (size_t)blockIdx.x * pool_cap;
  long long t_mark = prof ? clock64() : 0;
  const long long t_kernel = t_mark;
  int n_staged = 0;
  auto lap = [&](int slot) {     // optional phase timing (XVCB_TZ_PROF=1): cycles of thread 0, summed over CTAs
    if (prof && threadIdx.x == 0) { const long long now = clock64(); atomicAdd(&prof[slot], (unsigned long long)(now - t_mark)); t_mark = now; }
  };

  if (tid < 48) s_pat[tid] = pack_pattern_point(tid >> 4, tid & 15);
  __syncthreads();

  for (;;) {
    __syncthreads();                             // previous group is completely done with shared memory
    if (tid == 0) s_group = atomicAdd(counter, 1);
    __syncthreads();
    const int grp = s_group;
    if (grp >= n_groups) break;
    const long long t_group = prof ? clock64() : 0;
    int g_raster = 0, g_chunks = 0;
    const TzGroup G = groups[sel.J ? grp / sel.n : grp];
    const int jcol = sel.J ? sel.j[grp % sel.n] : 0;
    auto job_of = [&](int k) { const int e = job_index[G.first + k]; return sel.J ? e * sel.J + jcol : e; };
    const int ref_slot = jobs[job_of(0)].ref_slot;
    const PlaneView ref = ref_planes[ref_slot];

    // The jobs of a group are searched in chunks: as many consecutive jobs (<= kMaxGroupJobs) as have the
    // bounding box of their search windows inside the shared-memory region -- predictors that differ between
    // the CUs of a CTU widen the box (with one predictor for all of them a +-128 window just fits).
    int kn_adv = 0;
    for (int k0 = 0; k0 < G.count; k0 += kn_adv) {
      const int kmax = min(kMaxGroupJobs, G.count - k0);
      __syncthreads();
      // job descriptors -> shared memory; the search window of each (block extent included) parked in the raster fields
      for (int k = tid; k < kmax; k += kTzThreads) {
        const int ji = job_of(k0 + k);
        const xvcb200_me_job job = jobs[ji];
        const xvcb200_cu cu = cus[job.cu];
        SJob &sj = s_job[k];
        sj.ji = ji; sj.x = cu.x; sj.y = cu.y; sj.w = cu.w; sj.h = cu.h; sj.depth = cu.depth;
        sj.fullpel = (cu.flags & XVCB200_CU_FULLPEL_MV) ? 1 : 0;
        sj.mvpx = job.mvp[0]; sj.mvpy = job.mvp[1]; sj.prevx = job.prev[0]; sj.prevy = job.prev[1];
        sj.range = job.search_range; sj.need = 0; sj.list_off = -1; sj.list_cnt = 0;
        sj.key = ~0ull;
        int lo[2], hi[2];
        min_max_mv(cu.x, cu.y, ref.width, ref.height, job.mvp[0], job.mvp[1], job.search_range, lo, hi);
        sj.slox = cu.x + lo[0]; sj.sloy = cu.y + lo[1]; sj.nx = cu.x + hi[0] + cu.w; sj.ny = cu.y + hi[1] + cu.h;
      }
      __syncthreads();
      if (tid == 0) {
        int b0 = s_job[0].slox, b1 = s_job[0].sloy, b2 = s_job[0].nx, b3 = s_job[0].ny, kfit = 1;
        for (; kfit < kmax; kfit++) {
          const SJob &sj = s_job[kfit];
          const int n0 = min(b0, sj.slox), n1 = min(b1, sj.sloy), n2 = max(b2, sj.nx), n3 = max(b3, sj.ny);
          const int nbw = n2 - (n0 & ~7);
          if ((long long)((((nbw + 1) / 2 + 1) | 1)) * (n3 - n1) > region_budget_words) break;
          b0 = n0; b1 = n1; b2 = n2; b3 = n3;
        }
        s_box[0] = b0; s_box[1] = b1; s_box[2] = b2; s_box[3] = b3;
        s_count = kfit;
        s_pool_used = 0; s_next = 0; s_any_raster = 0;
      }
      if (tid < 10) s_cls[tid] = 0;
      __syncthreads();
      const int kn = s_count;
      kn_adv = kn;
      g_chunks++;
      for (int k = tid; k < kn; k += kTzThreads) atomicAdd(&s_cls[__clz((int)s_job[k].w * s_job[k].h) - 19], 1);   // area 4096 -> class 0 ... 16 -> class 8
      __syncthreads();
      if (tid == 0) {          // class counts -> first slot of each class; blocks of >= 2048 samples are searched CTA-wide
        s_ncoop = s_cls[0] + s_cls[1];
        int off = 0;
        for (int c = 0; c < 10; c++) { const int n = s_cls[c]; s_cls[c] = off; off += n; }
        s_next = 0;
      }
      __syncthreads();
      for (int k = tid; k < kn; k += kTzThreads)
        s_order[atomicAdd(&s_cls[__clz((int)s_job[k].w * s_job[k].h) - 19], 1)] = (unsigned char)k;
      __syncthreads();
      int rx0 = s_box[0] & ~7, ry0 = s_box[1], rx1 = s_box[2], ry1 = s_box[3];
      int bw = rx1 - rx0, bh = ry1 - ry0;
      int spw = ((bw + 1) / 2 + 1) | 1;
      int cpr = (bw + 7) >> 3;
      bool staged = (long long)spw * bh <= region_budget_words;
      const int ctu_x = s_job[0].x & ~63, ctu_y = s_job[0].y & ~63;
      lap(0);
      // original CTU -> shared memory (all jobs of a group lie in one CTU)
      for (int q = tid; q < 64 * 32; q += kTzThreads) {
        const int row = q >> 5, col = q & 31;
        s_tile[row * 33 + col] = __ldg(reinterpret_cast<const uint32_t *>(orig.base + (ctu_y + row) * orig.pitch + ctu_x) + col);
      }
      if (staged) stage_box(ref.base, ref.pitch, rx0, ry0, bh, cpr, spw, s_region, tid, kTzThreads);
      __syncthreads();
      if (prof && tid == 0) {      // first staging by the ordinal of the group in this CTA, and its bytes
        atomicAdd(&prof[18 + min(n_staged, 3)], (unsigned long long)(clock64() - t_mark));
        atomicAdd(&prof[22], (unsigned long long)(staged ? bh * cpr * 16 : 0));
        n_staged++;
      }
      lap(1);

      // ---------------- phase 1 (ph = 0: start points, first diamond pass, 2-point step), the raster
      // scan of the jobs that ended far out, phase 3 (ph = 1: re-centre until the centre wins).
      // Both phases are ONE loop with ONE evaluation site driven by a small state machine, so the
      // SAD routines exist once in the code: the search is issue-bound and every warp of the CTA
      // is somewhere else in it, the kernel has to stay inside the instruction cache.
      // Jobs are fetched largest first.  Blocks of >= 2048 samples (groups of one or two jobs) are
      // searched by all 16 warps as one team; at the first smaller block the team splits into
      // single warps, one job each.  The team synchronises on named barrier 1.
#pragma unroll 1
      for (int ph = 0; ph < 2; ph++) {
        if (ph == 1) {
          if (!s_any_raster) break;                // every job of the chunk finished in the first loop
          if (prof && tid == 0)
            for (int k = 0; k < kn; k++) g_raster += s_job[k].need != 0;
          // Two rounds (SJob::need: 1 = scan window inside the staged box, 2 = outside, 3 = scanned).  Round 0 scans
          // the jobs whose window lies in the staged box.  A start point other than the predictor re-centres the
          // scan window (DetermineMinMaxMv around the best start, :121-125), typically for several CUs of the CTU
          // alike: round 1 stages the bounding box of THOSE windows (as many as fit) and scans them the same way;
          // only what still does not fit is scanned from global memory.  The refinement then runs on the box
          // staged last (candidates outside it are read from global memory).
#pragma unroll 1
          for (int rr = 0; rr < 2; rr++) {
            if (rr == 1) {
              if (tid == 0) {
                int b0 = 0, b1 = 0, b2 = 0, b3 = 0, taken = 0;
                for (int k = 0; k < kn; k++) {
                  SJob &sj = s_job[k];
                  if (sj.need != 2) continue;
                  const int X0 = sj.x + sj.slox, Y0 = sj.y + sj.sloy, X1 = X0 + 5 * (sj.nx - 1) + sj.w, Y1 = Y0 + 5 * (sj.ny - 1) + sj.h;
                  const int n0 = taken ? min(b0, X0) : X0, n1 = taken ? min(b1, Y0) : Y0, n2 = taken ? max(b2, X1) : X1, n3 = taken ? max(b3, Y1) : Y1;
                  const int nbw = n2 - (n0 & ~7);
                  if ((long long)((((nbw + 1) / 2 + 1) | 1)) * (n3 - n1) > region_budget_words) continue;
                  b0 = n0; b1 = n1; b2 = n2; b3 = n3; taken++;
                  sj.need = 1; sj.list_off = -1; sj.list_cnt = 0;
                }
                int left = 0;
                for (int k = 0; k < kn; k++) left += s_job[k].need == 2;
                s_box[0] = b0; s_box[1] = b1; s_box[2] = b2; s_box[3] = b3;
                s_batch_tasks = taken; s_batch_end = left;
                s_pool_used = 0;
              }
              __syncthreads();
              if (s_batch_tasks == 0 && s_batch_end == 0) break;      // no scan window outside the box (uniform)
              if (s_batch_tasks > 0) {
                rx0 = s_box[0] & ~7; ry0 = s_box[1]; rx1 = s_box[2]; ry1 = s_box[3];
                bw = rx1 - rx0; bh = ry1 - ry0;
                spw = ((bw + 1) / 2 + 1) | 1;
                cpr = (bw + 7) >> 3;
                staged = true;
                stage_box(ref.base, ref.pitch, rx0, ry0, bh, cpr, spw, s_region, tid, kTzThreads);
              }
              __syncthreads();
            } else {
              if (tid == 0) {
                int n1 = 0;
                for (int k = 0; k < kn; k++) n1 += s_job[k].need == 1;
                s_batch_tasks = n1;
              }
              __syncthreads();
            }
            const bool any_in_box = s_batch_tasks > 0;
            __syncthreads();                                   // s_batch_tasks is written again below
            // ---------------- raster pass 1: segment-sum bound, survivors -> pool
            if (staged && any_in_box) {
              box_to_s8(s_region, spw, bh, tid, kTzThreads);
              __syncthreads();
              lap(3);
              const uint16_t *s8reg = reinterpret_cast<const uint16_t *>(s_region);
              const uint32_t *seg32 = reinterpret_cast<const uint32_t *>(s_seg);
              // Jobs are bounded in batches (as many as have room for their segment sums); inside a
              // batch the (kBoundCols grid columns, 32 grid rows) tasks of ALL jobs form one list that
              // is dealt to the warps.  The last column chunk of a job is shifted left to stay inside
              // the window; the columns it shares with its neighbour are not reported twice.
              for (int kb = 0; kb < kn;) {                         // uniform
                if (tid == 0) {
                  int used = 0, tasks = 0, k = kb;
                  for (; k < kn; k++) {
                    SJob &sj = s_job[k];
                    sj.task0 = -1;
                    if (sj.need != 1 || sj.w < 8 || sj.nx < kBoundCols || sj.nx * sj.ny > 65535) continue;
                    const int segs = (sj.h > 8 ? sj.h >> 1 : sj.h) * (sj.w >> 3);
                    if (used + segs > 2 * kSegWords) break;
                    sj.seg_off = used; sj.task0 = tasks;
                    used += (segs + 1) & ~1;
                    tasks += ((sj.nx + kBoundCols - 1) / kBoundCols) * ((sj.ny + 31) >> 5);
                  }
                  s_batch_end = k; s_batch_tasks = tasks; s_next = 0; s_count = 0;
                }
                __syncthreads();
                const int ke = s_batch_end, ntasks = s_batch_tasks;
                for (int k = kb + warp; k < ke; k += kTzWarps) {    // segment sums of the original blocks, a warp per job
                  const SJob &sj = s_job[k];
                  if (sj.task0 < 0) continue;
                  const int rstep = sj.h > 8 ? 2 : 1, rows = sj.h > 8 ? sj.h >> 1 : sj.h;
                  const int lsg = 28 - __clz((int)sj.w);            // log2(w / 8)
                  const uint32_t *tp = s_tile + (sj.y - ctu_y) * 33 + ((sj.x - ctu_x) >> 1);
                  for (int q = lane; q < (rows << lsg); q += 32) {
                    const int row = q >> lsg, sg = q & ((1 << lsg) - 1);
                    const uint32_t *p = tp + row * rstep * 33 + sg * 4;
                    const uint32_t s2 = p[0] + p[1] + p[2] + p[3];            // two 16-bit partial sums, no carry (<= 4 x 4095)
                    s_seg[sj.seg_off + q] = (uint16_t)((s2 & 0xffff) + (s2 >> 16));
                  }
                }
                __syncthreads();
                const int base = s_pool_used;
                const int room = pool_cap - base;
                {
                  int kcur = kb - 1, tend = 0, tbeg = 0;           // the job the warp's current task belongs to
                  int nx = 1, ny = 1, nch = 1, slox = 0, sloy = 0, lsg = 0, rstride = 0, rows = 0, gx = 0, gy = 0, fast = 0, sh = 2;
                  int mvpx = 0, mvpy = 0;
                  uint32_t cost_in = 0;
                  const uint32_t *segp = seg32;
                  for (int task = warp; task < ntasks; task += kTzWarps) {
                    while (task >= tend) {                           // tasks arrive in increasing order
                      kcur++;
                      const SJob &sj = s_job[kcur];
                      if (sj.task0 < 0) continue;
                      nx = sj.nx; ny = sj.ny; slox = sj.slox; sloy = sj.sloy; cost_in = sj.cost_in;
                      nch = (nx + kBoundCols - 1) / kBoundCols;
                      tbeg = sj.task0; tend = tbeg + nch * ((ny + 31) >> 5);
                      fast = sj.h > 8; rows = fast ? sj.h >> 1 : sj.h;
                      lsg = 28 - __clz((int)sj.w);
                      rstride = (fast ? 2 : 1) * 2 * spw;
                      gx = sj.x - rx0; gy = sj.y - ry0;
                      mvpx = sj.mvpx; mvpy = sj.mvpy; sh = sj.fullpel ? 4 : 2;
                      segp = seg32 + (sj.seg_off >> 1);
                    }
                    const int local = task - tbeg;
                    const int pass = local / nch, ch = local - pass * nch;
                    const int skip = max(0, (ch + 1) * kBoundCols - nx);       // leading columns owned by the previous chunk
                    const int i0 = ch * kBoundCols - skip;
                    const int j = pass * 32 + lane, jj = min(j, ny - 1);      // idle lanes recompute the last row (no stray reads)
                    const int cy = sloy + 5 * jj;
                    const uint32_t bits_y = exp_golomb_bits((cy * 16 - mvpy) >> sh);
                    const uint16_t *rp = s8reg + (gy + cy) * (2 * spw) + gx + slox + 5 * i0;
                    uint32_t lb[kBoundCols];
    #pragma unroll
                    for (int u = 0; u < kBoundCols; u++) lb[u] = 0;
                    switch (lsg) {
                      case 0: seg_bound_cols<1>(rp, rstride, segp, rows, lb); break;
                      case 1: seg_bound_cols<2>(rp, rstride, segp, rows, lb); break;
                      case 2: seg_bound_cols<4>(rp, rstride, segp, rows, lb); break;
                      default: seg_bound_cols<8>(rp, rstride, segp, rows, lb); break;
                    }
    #pragma unroll
                    for (int u = 0; u < kBoundCols; u++) {
                      const int i = i0 + u, cx = slox + 5 * i;
                      const uint32_t rate = (lambda * (exp_golomb_bits((cx * 16 - mvpx) >> sh) + bits_y)) >> 16;
                      // dropped iff dist_bound + rate >= cost_in, in raw SAD units: lb >= thr
                      uint32_t thr = 0;
                      if (j < ny && u >= skip && rate < cost_in) {
                        const uint32_t need = (cost_in - rate) << (bitdepth - 8);
                        thr = fast ? (need + 1) >> 1 : need;
                      }
                      const bool keep = lb[u] < thr;
                      const unsigned mask = __ballot_sync(XVCB_FULL, keep);
                      if (mask) {
                        int wbase = 0;
                        if (lane == 0) wbase = atomicAdd(&s_count, __popc(mask));
                        wbase = __shfl_sync(XVCB_FULL, wbase, 0);
                        const int slot = wbase + __popc(mask & ((1u << lane) - 1));
                        if (keep && slot < room) pool[base + slot] = ((uint32_t)kcur << 16) | (uint32_t)(j * nx + i);
                      }
                    }
                  }
                }
                __syncthreads();
                if (tid == 0) {
                  unsigned long long cands = 0;
                  for (int k = kb; k < ke; k++) {
                    SJob &sj = s_job[k];
                    if (sj.task0 < 0) continue;
                    cands += (unsigned long long)(sj.nx * sj.ny);
                    if (s_count <= room) { sj.list_off = base; sj.list_cnt = s_count; }
                  }
                  if (s_count <= room) s_pool_used = base + s_count;
                  if (prof) { atomicAdd(&prof[8], cands); atomicAdd(&prof[9], (unsigned long long)s_count); }
                }
                __syncthreads();
                kb = ke;
              }
              lap(4);
              s8_to_box(s_region, spw, bh, tid, kTzThreads);
              __syncthreads();
              lap(5);
            }

            // ---------------- raster pass 2a: survivors of all jobs, one flat loop, one candidate per lane
            // A group holds a few hundred survivors -- fewer than the CTA has threads -- so a survivor per lane
            // leaves the pass waiting for the one lane that walks a 64x64 block alone.  A warp takes 32
            // consecutive entries and spreads each over 2^lgw lanes (rows sub, sub + 2^lgw, ... of the block;
            // lgw from the largest block among the 32 so that a lane sums about 16 sample pairs), the partial
            // sums meet in lgw shuffles.
            {
              const int total = s_pool_used;
              // ... as long as there are fewer lane tasks than about eight per thread: with thousands of survivors (poor
              // predictors: most jobs of the group are scanned) a survivor per lane keeps every lane busy anyway
              const int lg_cap = total >= 8 * kTzThreads ? 0 : 31 - __clz((8 * kTzThreads) / max(total, 1));
              for (int e0 = warp * 32; e0 < total; e0 += kTzThreads) {
                const int e = e0 + lane;
                uint32_t ent = 0;
                int lgj = 0;
                bool live = false;
                if (e < total) {
                  ent = pool[e];
                  const SJob &sj = s_job[ent >> 16];
                  live = !(sj.list_off < 0 || e < sj.list_off || e >= sj.list_off + sj.list_cnt);   // not the list of an overflowed job
                  const int rows = sj.h > 8 ? sj.h >> 1 : sj.h;
                  const int lpairs = (31 - __clz(rows)) + (30 - __clz((int)sj.w));
                  lgj = live ? min(31 - __clz(rows), max(0, lpairs - 4)) : 0;
                }
                const int lgw = min(__reduce_max_sync(XVCB_FULL, lgj), lg_cap);
                const unsigned live_mask = __ballot_sync(XVCB_FULL, live);
                for (int q = 0; q < (1 << lgw); q++) {
                  const int src = ((q << 5) + lane) >> lgw, sub = lane & ((1 << lgw) - 1);
                  const uint32_t en = __shfl_sync(XVCB_FULL, ent, src);
                  const bool on = (live_mask >> src) & 1u;
                  if (!__ballot_sync(XVCB_FULL, on)) continue;
                  SJob &sj = s_job[en >> 16];
                  const uint32_t t = en & 0xffffu;
                  const MeGeom g = sjob_geom(sj, bitdepth, lambda);
                  const int j = (int)t / sj.nx, i = (int)t - j * sj.nx;
                  const int cx = sj.slox + 5 * i, cy = sj.sloy + 5 * j;
                  uint32_t sad = 0;
                  if (on && sub < g.rows) {
                    const int ox = g.x + cx - rx0, oy = g.y + cy - ry0 + sub * g.rstep;
                    sad = sad_rows_lpw(g.lpw, s_region + oy * spw + (ox >> 1), (g.rstep * spw) << lgw,
                                       s_tile + (sj.y - ctu_y + sub * g.rstep) * 33 + ((sj.x - ctu_x) >> 1), (33 * g.rstep) << lgw,
                                       (g.rows - sub + (1 << lgw) - 1) >> lgw, (ox & 1) << 4);
                  }
#pragma unroll
                  for (int off = 16; off > 0; off >>= 1)
                    if (off < (1 << lgw)) sad += __shfl_xor_sync(XVCB_FULL, sad, off);
                  if (on && sub == 0) {
                    const uint32_t dist = g.fast ? (sad * 2) >> g.bd_shift : sad >> g.bd_shift;
                    const uint32_t cost = dist + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
                    atomicMin(&sj.key, ((unsigned long long)cost << 32) | t);
                  }
                }
              }
            }
            // ---------------- raster pass 2b: dense scans (no survivor list), CTA-wide per job
            for (int k = 0; k < kn; k++) {
              SJob &sj = s_job[k];
              if (!((sj.need == 1 && sj.list_off < 0) || (rr == 1 && sj.need == 2))) continue;      // uniform
              const MeGeom g = sjob_geom(sj, bitdepth, lambda);
              const int slox = sj.slox, sloy = sj.sloy, nx = sj.nx, ny = sj.ny;
              const bool in_box = sj.need == 1;
              const uint32_t *tp = s_tile + (sj.y - ctu_y) * 33 + ((sj.x - ctu_x) >> 1);
              uint32_t best_cost = 0xffffffffu, best_t = 0;
              for (int j0 = 0; j0 < ny; j0 += 32) {
                const int j = j0 + lane, jj = min(j, ny - 1);     // idle lanes recompute the last row (no stray reads)
                const int cy = sloy + 5 * jj;
                for (int i = warp; i < nx; i += kTzWarps) {
                  const int cx = slox + 5 * i;
                  uint32_t sad;
                  if (in_box) {
                    const int ox = g.x + cx - rx0, oy = g.y + cy - ry0;
                    sad = sad_rows_lpw(g.lpw, s_region + oy * spw + (ox >> 1), g.rstep * spw, tp, 33 * g.rstep, g.rows, (ox & 1) << 4);
                  } else {                                  // window outside / larger than the staged box: same walk from global memory
                    const int X = g.x + cx;
                    const Sample *row0 = ref.base + (g.y + cy) * ref.pitch + (X & ~1);
                    sad = sad_rows_global(reinterpret_cast<const uint32_t *>(row0), (g.rstep * ref.pitch) >> 1, tp, 33 * g.rstep,
                                          g.rows, (X & 1) << 4, 1 << g.lpw);
                  }
                  if (j < ny) {
                    const uint32_t dist = g.fast ? (sad * 2) >> g.bd_shift : sad >> g.bd_shift;
                    const uint32_t cost = dist + ((g.lambda * mvd_bits_fullpel(g.mvpx, g.mvpy, cx, cy, g.down)) >> 16);
                    const uint32_t t = (uint32_t)(j * nx + i);                      // position in the reference's scan order
                    if (cost < best_cost || (cost == best_cost && t < best_t)) { best_cost = cost; best_t = t; }
                  }
                }
              }
              if (best_cost != 0xffffffffu) atomicMin(&sj.key, ((unsigned long long)best_cost << 32) | best_t);
            }
            __syncthreads();
            // winners -> job state (strict compare: ties keep the earlier best, :266-268)
            for (int k = tid; k < kn; k += kTzThreads) {
              SJob &sj = s_job[k];
              if (!(sj.need == 1 || (rr == 1 && sj.need == 2))) continue;
              sj.need = 3;
              TzJobState *stp = &states[sj.ji];
              const uint32_t c = (uint32_t)(sj.key >> 32), t = (uint32_t)sj.key;
              if (sj.key != ~0ull && c < sj.cost_in) {
                stp->cost = c;
                stp->bx = sj.slox + 5 * (int)(t % sj.nx);
                stp->by = sj.sloy + 5 * (int)(t / sj.nx);
              }
              stp->last_range = 5;
              stp->evals += sj.nx * sj.ny;
              stp->need_raster = 0;
            }
            __threadfence_block();
            __syncthreads();
          }
          if (tid == 0) s_next = 0;
          __syncthreads();
          lap(6);
        }
        int tsize = kTzWarps;
        long long t_w = prof ? clock64() : 0;
        auto tick = [&](int slot) { if (prof) { const long long now = clock64(); if (lane == 0) atomicAdd(&prof[slot], (unsigned long long)(now - t_w)); t_w = now; } };
        for (;;) {
          tick(16);                                   // end-of-job / idle
          const int tfirst = warp & ~(tsize - 1);
          int o = 0;
          if (tsize == 1) {
            if (lane == 0) o = atomicAdd(&s_next, 1);
            o = __shfl_sync(XVCB_FULL, o, 0);
          } else {
            const int bar = 1 + (tfirst >> 2);
            if (warp == tfirst && lane == 0) s_fetch[tfirst] = atomicAdd(&s_next, 1);
            asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * tsize) : "memory");
            o = s_fetch[tfirst];
            asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(32 * tsize) : "memory");
          }
          if (o >= kn) break;
          SJob &sj = s_job[s_order[o]];
          if (ph == 1 && sj.need == 0) continue;      // finished in the first loop
          const int area = (int)sj.w * sj.h;
          const int want = area >= 2048 ? kTzWarps : (area >= 512 ? 4 : 1);
          if (want < tsize) {
            const bool first_sub = (warp & (tsize - 1)) < want;    // the sub-team that keeps this job
            tsize = want;
            if (!first_sub) continue;
          }
          const int tf = warp & ~(tsize - 1), tw = warp - tf;
          tick(10);                                   // fetch
          const MeGeom g = sjob_geom(sj, bitdepth, lambda);
          SearchEval ev{g, s_tile + (sj.y - ctu_y) * 33 + ((sj.x - ctu_x) >> 1), 33 * g.rstep, ref.base, ref.pitch,
                        staged ? s_region : nullptr, spw, rx0, ry0, rx1, ry1, s_pat, lane, tw, 31 - __clz(tsize),
                        s_tcost[tf >> 2], 1 + (tf >> 2)};
          const int lrows = 31 - __clz(g.rows);
          // lanes per candidate: diamond passes / the 2-3 point lists
          const int lg_d = min(lrows, tsize == kTzWarps ? 4 : (tsize == 4 ? 2 : ((g.rows << g.lpw) >= 64 ? 1 : 0)));
          const int lg_l = min(lrows, tsize == kTzWarps ? 4 : 3);
          const int range = sj.range;
          const int nrounds = 32 - __clz(range);                       // radii 1, 2, 4, ... <= range
          const int nslots = nrounds <= 1 ? 4 * nrounds : (nrounds <= 4 ? 8 * nrounds - 4 : 16 * nrounds - 36);
          TzJobState *stp = &states[sj.ji];
          int lo[2], hi[2], slo[2], shi[2];
          TzBest b;
          uint32_t evals;
          uint32_t cost[4];
          if (ph == 0) {      // inter_tz_search.cc:92-131: predictor, zero, previous search result
            min_max_mv(g.x, g.y, ref.width, ref.height, g.mvpx, g.mvpy, range, lo, hi);
            slo[0] = lo[0]; slo[1] = lo[1]; shi[0] = hi[0]; shi[1] = hi[1];
            int px = g.mvpx, py = g.mvpy;
            clip_mv(g.x, g.y, ref.width, ref.height, px, py);
            px >>= 4; py >>= 4;
            int qx = sj.prevx * 16, qy = sj.prevy * 16;
            clip_mv(g.x, g.y, ref.width, ref.height, qx, qy);
            qx >>= 4; qy >>= 4;
            const bool use_zero = px != 0 || py != 0, use_prev = sj.depth != 0;
            const int ex = lane == 0 ? px : (lane == 1 ? 0 : qx), ey = lane == 0 ? py : (lane == 1 ? 0 : qy);
            tick(11);                                 // setup
            ev.run(kEvalList, lg_l, 0, 3, 0, 0, lo, hi, ex, ey, lane == 0 || (lane == 1 && use_zero) || (lane == 2 && use_prev), cost);
            tick(12);                                 // start eval
            evals = 1 + use_zero + use_prev;
            const uint32_t c0 = __shfl_sync(XVCB_FULL, cost[0], 0), c1 = __shfl_sync(XVCB_FULL, cost[0], 1),
                           c2 = __shfl_sync(XVCB_FULL, cost[0], 2);
            b.cost = c0; b.x = px; b.y = py;
            bool moved = false;
            if (use_zero && c1 < b.cost) { b.cost = c1; b.x = 0; b.y = 0; moved = true; }
            if (use_prev) {
              if (c2 < b.cost) { b.cost = c2; b.x = qx; b.y = qy; moved = true; }
              if (moved) min_max_mv(g.x, g.y, ref.width, ref.height, b.x * 16, b.y * 16, range, slo, shi);
            }
            b.last_pos = 0; b.last_range = 1 << 30;     // enters the loop below exactly once
          } else {
            b.x = stp->bx; b.y = stp->by; b.cost = stp->cost; b.last_pos = stp->last_pos; b.last_range = stp->last_range;
            lo[0] = stp->lo[0]; lo[1] = stp->lo[1]; hi[0] = stp->hi[0]; hi[1] = stp->hi[1];
            slo[0] = slo[1] = shi[0] = shi[1] = 0;
            evals = stp->evals;
            tick(11);
          }
          // The first diamond pass around the start point stops after three rounds without a hit (:133-143),
          // then the 2-point step; a job that needs no raster scan goes straight on to the refinement
          // (re-centre until the centre wins, every round of every pass, :157-168) -- no CTA-wide barrier
          // between the two for the majority of the jobs.  A job that needs the scan parks its state and
          // is picked up again (ph 1) after the CTA-wide raster passes.
          // The rounds of the first pass are evaluated in two instalments, radii 1..8 (28 slots) and the
          // rest: with a good predictor the three-miss rule ends the pass inside the first one.
          bool first = ph == 0, parked = false;
          while (b.last_range > 0) {
            const int ax = b.x, ay = b.y;
            b.last_range = 0;
            uint32_t win[10];
            int misses = 0, wri = -1, wk = 0;
            bool stopped = false;
            const int split = first && nrounds > 4 ? 4 : nrounds;      // rounds in the first instalment
#pragma unroll 1
            for (int inst = 0; inst < 2; inst++) {
              const int r_lo = inst == 0 ? 0 : split, r_hi = inst == 0 ? split : nrounds;
              if (r_lo >= r_hi || stopped) break;
              tick(11);
              ev.run(kEvalDiamond, lg_d, r_lo == 0 ? 0 : 28, r_hi == nrounds ? nslots : 28, ax, ay, lo, hi, 0, 0, false, cost);
              tick(13);                               // diamond eval
              if (prof && lane == 0) atomicAdd(&prof[17], 1ull);
              // FullpelDiamondSearch replayed (:173-210).  The winner of every round (first minimum in
              // the reference's candidate order) and its number of evaluated points do not depend on
              // the running best: independent warp reductions, then the sequential decisions
              // (strict compare, three-miss rule of the first pass) on uniform values.
              uint32_t nv_lo = 0, nv_hi = 0;                   // evaluated points per round, 5 bits each
#pragma unroll
              for (int ri = 0; ri < 10; ri++) {
                constexpr int kBeg[10] = {0, 4, 12, 20, 28, 44, 60, 76, 92, 108};
                const int K = ri == 0 ? 4 : (ri <= 3 ? 8 : 16);
                const int sbeg = kBeg[ri], q0 = sbeg >> 5, q1 = (sbeg + K - 1) >> 5;
                win[ri] = 0xffffffffu;
                if (ri >= r_lo && ri < r_hi) {
                  uint32_t key = 0xffffffffu;
#pragma unroll
                  for (int q = 0; q < 4; q++) {
                    if (q != q0 && q != q1) continue;
                    const int rel = 32 * q + lane - sbeg;
                    if ((unsigned)rel < (unsigned)K && cost[q] != 0xffffffffu) key = min(key, (cost[q] << 5) | (uint32_t)rel);
                  }
                  win[ri] = __reduce_min_sync(XVCB_FULL, key);
                  const uint32_t nv = __popc(__ballot_sync(XVCB_FULL, key != 0xffffffffu));
                  if (ri < 6) nv_lo |= nv << (5 * ri); else nv_hi |= nv << (5 * (ri - 6));
                }
              }
#pragma unroll
              for (int ri = 0; ri < 10; ri++) {
                if (ri >= r_lo && ri < r_hi && !stopped) {
                  evals += ((ri < 6 ? nv_lo >> (5 * ri) : nv_hi >> (5 * (ri - 6))) & 31u);
                  if (win[ri] != 0xffffffffu && (win[ri] >> 5) < b.cost) {
                    b.cost = win[ri] >> 5; wri = ri; wk = win[ri] & 31; misses = 0;
                  } else if (first && ++misses >= 3) {
                    stopped = true;
                  }
                }
              }
            }
            if (wri >= 0) pattern_point(s_pat, wri, wk, ax, ay, b.x, b.y, b.last_pos, b.last_range);
            tick(14);                                 // replay
            if (b.last_range == 1) {                         // FullpelNeighborPointSearch, :212-259
              b.last_range = 0;
              int dx = 0, dy = 0, pos = 0;
              const int K = two_point(b.last_pos, lane, dx, dy, pos);
              const int cx = b.x + dx, cy = b.y + dy;
              const bool valid = lane < K && inside(cx, cy, pos, lo, hi);
              const unsigned vm = __ballot_sync(XVCB_FULL, valid);
              if (vm) {
                evals += __popc(vm);
                ev.run(kEvalList, lg_l, 0, 2, 0, 0, lo, hi, cx, cy, valid, cost);
                const uint32_t key = lane < 2 && cost[0] != 0xffffffffu ? (cost[0] << 5) | (uint32_t)lane : 0xffffffffu;
                const uint32_t win2 = __reduce_min_sync(XVCB_FULL, key);
                if (win2 != 0xffffffffu && (win2 >> 5) < b.cost) {
                  b.cost = win2 >> 5;
                  int wx = 0, wy = 0, wpos = 0;
                  two_point(b.last_pos, win2 & 31, wx, wy, wpos);
                  b.x += wx; b.y += wy; b.last_pos = wpos; b.last_range = 1;
                }
              }
            }
            tick(15);                                 // neighbour
            if (first) {
              first = false;
              if (b.last_range > 5) {                        // kFullSearchGranularity (:91, :146)
                if (shi[0] >= slo[0] && shi[1] >= slo[1]) { parked = true; break; }
                b.last_range = 5;                            // empty scan window (:146-147)
              }
            }
          }
          if (tw == 0 && lane == 0) {
            if (parked) {
              stp->bx = b.x; stp->by = b.y; stp->cost = b.cost; stp->last_pos = b.last_pos; stp->last_range = b.last_range;
              stp->lo[0] = lo[0]; stp->lo[1] = lo[1]; stp->hi[0] = hi[0]; stp->hi[1] = hi[1];
              stp->evals = evals;
              sj.slox = slo[0]; sj.sloy = slo[1];
              sj.nx = (shi[0] - slo[0]) / 5 + 1; sj.ny = (shi[1] - slo[1]) / 5 + 1;
              sj.cost_in = b.cost;
              const bool fits = staged && g.x + slo[0] >= rx0 && g.x + shi[0] + g.w <= rx1 && g.y + slo[1] >= ry0 &&
                                g.y + shi[1] + g.h <= ry1;
              sj.need = fits ? 1 : 2;
              s_any_raster = 1;
            } else {
              xvcb200_me_result *out = &res[sj.ji];
              out->mv_fullpel[0] = b.x; out->mv_fullpel[1] = b.y;
              out->cost_fullpel = b.cost;
              out->num_sad = evals;
            }
          }
        }
        __threadfence_block();
        __syncthreads();
        lap(ph == 0 ? 2 : 7);
      }
    }
    if (prof && tid == 0 && grp < kProfGroups) {
      unsigned long long *pg = prof + 24 + 4 * grp;
      pg[0] = (unsigned long long)(clock64() - t_group); pg[1] = (unsigned long long)G.count; pg[3] = (unsigned long long)g_raster | ((unsigned long long)g_chunks << 32);
      unsigned long long area = 0;
      for (int k = 0; k < G.count; k++) { const xvcb200_cu cu = cus[jobs[job_of(k)].cu]; area += (unsigned long long)cu.w * cu.h; }
      pg[2] = area;
    }
  }
  if (prof && tid == 0) prof[24 + 4 * kProfGroups + blockIdx.x] = (unsigned long long)(clock64() - t_kernel);
}

cudaError_t launch_tz_search(cudaStream_t s, const xvcb200_cu *d_cus, const xvcb200_me_job *d_jobs, int n, int bitdepth,
                             uint32_t lambda_me, PlaneView orig, const PlaneView *d_ref_planes, xvcb200_me_result *d_res,
                             const int *d_job_index, const void *d_groups, int n_groups, void *d_states, int *d_counter,
                             uint32_t *d_pool, int pool_cap, int J, int n_cols, const int *cols) {
  if (n <= 0 || n_groups <= 0) return cudaSuccess;
  TzJobSel sel;
  sel.J = J; sel.n = J ? n_cols : 1;
  for (int k = 0; k < 10; k++) sel.j[k] = (J && k < n_cols) ? cols[k] : 0;
  if (J) n_groups *= n_cols;
  if (n_groups <= 0) return cudaSuccess;
  // launch configuration per device (a process may hold contexts on several GPUs; the dynamic
  // shared-memory opt-in is a per-device function attribute)
  static std::mutex cfg_mutex;
  static int smem_by_dev[kMaxDevices] = {0}, sms_by_dev[kMaxDevices] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  int smem_bytes, num_sms;
  {
    std::lock_guard<std::mutex> lock(cfg_mutex);
    if (!smem_by_dev[dev]) {
      int max_optin = 0, sms = 0;
      cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      cudaFuncAttributes fa;
      cudaFuncGetAttributes(&fa, tz_search_kernel);
      const int bytes = max_optin - (int)fa.sharedSizeBytes - 512;
      cudaError_t e = cudaFuncSetAttribute(tz_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (e != cudaSuccess) return e;
      smem_by_dev[dev] = bytes; sms_by_dev[dev] = sms;
    }
    smem_bytes = smem_by_dev[dev]; num_sms = sms_by_dev[dev];
  }
  cudaError_t e = cudaMemsetAsync(d_counter, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
    static const bool want_prof = getenv("XVCB_TZ_PROF") != nullptr;
  // device memory, not managed: the first touch of a managed page from the kernel faults and stalls the
  // SM for ~0.1 ms, which the laps then attribute to whatever phase comes first
  // (debugging aid, XVCB_TZ_PROF=1: one counter block per device, single-threaded use)
  static unsigned long long *d_prof_by_dev[kMaxDevices] = {nullptr};
  static unsigned long long prof[24 + 4 * kProfGroups + 256];
  unsigned long long *&d_prof = d_prof_by_dev[dev];
  if (want_prof && !d_prof) cudaMalloc(&d_prof, sizeof(prof));
  if (want_prof) cudaMemsetAsync(d_prof, 0, sizeof(prof), s);
  const int grid = n_groups < num_sms ? n_groups : num_sms;
  const int fixed_words = kTileWords + kSegWords + kMaxGroupJobs * (int)(sizeof(SJob) / 4);
  g_launch_count++;
  tz_search_kernel<<<grid, kTzThreads, smem_bytes, s>>>(
      d_cus, d_jobs, d_job_index, static_cast<const TzGroup *>(d_groups), n_groups, sel, d_counter, bitdepth, lambda_me, orig,
      d_ref_planes, d_res, static_cast<TzJobState *>(d_states), smem_bytes / 4 - fixed_words, d_pool, pool_cap,
      want_prof ? d_prof : nullptr);
  if (want_prof) {
    cudaMemcpyAsync(prof, d_prof, sizeof(prof), cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    fprintf(stderr, "[tz prof] cycles/CTA: box %.0f stage %.0f phase1 %.0f stageS8 %.0f bound %.0f stageRef %.0f exact %.0f phase3 %.0f | raster candidates %llu survivors %llu (%.2f%%)\n",
            (double)prof[0] / grid, (double)prof[1] / grid, (double)prof[2] / grid, (double)prof[3] / grid, (double)prof[4] / grid,
            (double)prof[5] / grid, (double)prof[6] / grid, (double)prof[7] / grid, prof[8], prof[9],
            100.0 * (double)prof[9] / (double)(prof[8] ? prof[8] : 1));
    fprintf(stderr, "[tz prof2] warp-cycles/CTA: fetch %.0f setup %.0f start %.0f diamond %.0f replay %.0f neighbour %.0f tail/idle %.0f | diamond passes (warp level) %llu\n",
            (double)prof[10] / grid, (double)prof[11] / grid, (double)prof[12] / grid, (double)prof[13] / grid, (double)prof[14] / grid,
            (double)prof[15] / grid, (double)prof[16] / grid, prof[17]);
    if (const char *path = getenv("XVCB_TZ_PROF_GROUPS")) {      // per-group records: cycles, jobs, samples, raster jobs
      if (FILE *f = fopen(path, "w")) {
        for (int g = 0; g < n_groups && g < kProfGroups; g++)
          fprintf(f, "g %d %llu %llu %llu %llu\n", g, prof[24 + 4 * g], prof[24 + 4 * g + 1], prof[24 + 4 * g + 2], prof[24 + 4 * g + 3]);
        for (int b = 0; b < grid && b < 256; b++) fprintf(f, "cta %d %llu\n", b, prof[24 + 4 * kProfGroups + b]);
        fclose(f);
      }
    }
    fprintf(stderr, "[tz prof3] first staging, cycles/CTA by ordinal of the group in its CTA: 1st %.0f 2nd %.0f 3rd %.0f later %.0f | staged bytes/CTA %.0f groups %d\n",
            (double)prof[18] / grid, (double)prof[19] / grid, (double)prof[20] / grid, (double)prof[21] / grid, (double)prof[22] / grid, n_groups);
  }
  return cudaGetLastError();
}
size_t tz_state_bytes() { return sizeof(TzJobState); }
int tz_max_ctas() { int dev = 0, n = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }

}  // namespace xvcb
