// In-loop deblocking (DeblockingFilter, deblocking_filter.cc) and border padding
// (YuvPicture::PadBorder, yuv_pic.cc:118-150) for a whole device-resident picture.
//
// Parallel decomposition.  The reference visits every 4x4 grid position of every CTU in
// raster order, all vertical edges first, then all horizontal edges (:56-77).  A vertical
// edge segment (4 rows) reads 4 and writes up to 3 samples on each side, so two edges 4
// samples apart in the same 4-row group are order dependent (left one first); edges in
// different 4-row groups, and edges 8 or more apart, touch disjoint samples.  The unit of
// parallel work is therefore a CHAIN: a maximal run of consecutive active edges (boundary
// strength > 0) 4 samples apart inside one 4-row group (vertical pass) or one 4-column group
// (horizontal pass).  The thread whose edge starts a chain filters the whole chain in the
// reference's order; every other thread of that chain does nothing.  Chroma edges sit on an
// 8-sample chroma grid and only touch one sample each side, so they are independent.
#include "xvcb_device.cuh"

namespace xvcb {

__constant__ uint8_t c_tc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,  0,  0,  0,  0,  0,
                                 1, 1, 1, 1, 1, 1, 1, 1, 1, 2,  2,  2,  2,  3,  3,  3,  3,  4,
                                 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};   // kTcTable, :34-38
__constant__ uint8_t c_beta[64] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,
                                   6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24,
                                   26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56,
                                   58, 60, 62, 64, 66, 68, 70, 72, 74, 76, 78, 80, 82, 84, 86, 88};  // kBetaTable, :40-45
__constant__ uint8_t c_db_chroma_scale[58] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
                                              20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35,
                                              36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};

// ---------------------------------------------------------------- CU map (PictureData::MarkUsedInPic, picture_data.cc:196-210)
__global__ void cu_map_kernel(const xvcb200_cu *__restrict__ cus, int n, int32_t *__restrict__ map, int map_w, int map_h) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const xvcb200_cu cu = cus[i];
  const int bw = cu.w >> 2, bh = cu.h >> 2, bx = cu.x >> 2, by = cu.y >> 2;
  for (int e = threadIdx.x; e < bw * bh; e += blockDim.x) {
    const int x = bx + e % bw, y = by + e / bw;
    if (x < map_w && y < map_h) map[y * map_w + x] = i;
  }
}

struct DbRefPoc { long long poc[2][5]; };

__device__ __forceinline__ long long ref_poc_of(const xvcb200_cu &cu, int list, const DbRefPoc &rp) {
  return cu.ref_idx[list] < 0 ? -1 : rp.poc[list][cu.ref_idx[list]];   // CodingUnit::GetRefPoc, coding_unit.cc:166-172
}
__device__ __forceinline__ bool mv_far(const int32_t a[2], const int32_t b[2]) {
  return abs(a[0] - b[0]) >= 16 || abs(a[1] - b[1]) >= 16;               // one integer sample at 1/16 pel
}

// The vector of a CU at one of its corners (CodingUnit::GetMv(list, corner), coding_unit.h:245-257): the control
// points of an affine CU (up-left, up-right, down-left; down-right = up-right + down-left - up-left), else the
// CU's one vector.  corner: 0 up-left, 1 up-right, 2 down-left, 3 down-right (MvCorner, cu_types.h:214-220).
struct DbAffine { const int *index; const xvcb200_affine_cu *cus; };
__device__ __forceinline__ void corner_mv(const xvcb200_cu &cu, int ci, int list, int corner, const DbAffine &af, int32_t out[2]) {
  const int a = af.index ? af.index[ci] : -1;
  if (a < 0) { out[0] = cu.mv[list][0]; out[1] = cu.mv[list][1]; return; }
  const int32_t (*m)[2] = af.cus[a].mv[list];
  if (corner < 3) { out[0] = m[corner][0]; out[1] = m[corner][1]; }
  else { out[0] = m[1][0] + m[2][0] - m[0][0]; out[1] = m[1][1] + m[2][1] - m[0][1]; }
}

// DeblockingFilter::GetBoundaryStrength, deblocking_filter.cc:154-241, for the edge segment whose q side starts at
// luma position (pos_x, pos_y); dir 0: vertical edge (p left of q), 1: horizontal edge (p above q).  The corner
// whose vector is compared depends on which half of the CU the segment lies in (:166-176).
__device__ int boundary_strength(const xvcb200_cu &p, const xvcb200_cu &q, int ip, int iq, int pos_x, int pos_y, int dir,
                                 int pic_type, const DbRefPoc &rp, const DbAffine &af) {
  if ((p.flags | q.flags) & XVCB200_CU_INTRA) return 2;
  if ((p.flags | q.flags) & XVCB200_CU_CBF_Y) return 1;
  int corner_p, corner_q;
  if (dir == 0) {
    corner_p = (pos_y - p.y) < (p.h >> 1) ? 1 : 3;
    corner_q = (pos_y - q.y) < (q.h >> 1) ? 0 : 2;
  } else {
    corner_p = (pos_x - p.x) < (p.w >> 1) ? 2 : 3;
    corner_q = (pos_x - q.x) < (q.w >> 1) ? 0 : 1;
  }
  int32_t mp0[2], mp1[2], mq0[2], mq1[2];
  corner_mv(p, ip, 0, corner_p, af, mp0); corner_mv(q, iq, 0, corner_q, af, mq0);
  if (pic_type == 0) {
    corner_mv(p, ip, 1, corner_p, af, mp1); corner_mv(q, iq, 1, corner_q, af, mq1);
    const long long p0 = ref_poc_of(p, 0, rp), p1 = ref_poc_of(p, 1, rp), q0 = ref_poc_of(q, 0, rp), q1 = ref_poc_of(q, 1, rp);
    if (!((p0 == q0 && p1 == q1) || (p0 == q1 && p1 == q0))) return 1;
    const bool straight = mv_far(mp0, mq0) || mv_far(mp1, mq1);
    const bool crossed = mv_far(mp0, mq1) || mv_far(mp1, mq0);
    if (p0 != p1) return (p0 == q0) ? straight : crossed;
    return straight && crossed;
  }
  if (p.ref_idx[0] != q.ref_idx[0]) return 1;
  return mv_far(mp0, mq0);
}

// boundary strength of the left (bs_v) and top (bs_h) edge of every 4x4 block; 0 = no edge
__global__ void edge_bs_kernel(const xvcb200_cu *__restrict__ cus, const int32_t *__restrict__ map, int map_w, int map_h,
                               int pic_type, const __grid_constant__ DbRefPoc rp, DbAffine af, uint8_t *__restrict__ bs_v,
                               uint8_t *__restrict__ bs_h) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= map_w * map_h) return;
  const int cx = cell % map_w, cy = cell / map_w;
  const int iq = map[cell];
  uint8_t v = 0, h = 0;
  if (iq >= 0) {
    const xvcb200_cu q = cus[iq];
    if (cx > 0) {
      const int ip = map[cell - 1];
      if (ip >= 0 && ip != iq) v = (uint8_t)boundary_strength(cus[ip], q, ip, iq, cx * 4, cy * 4, 0, pic_type, rp, af);
    }
    if (cy > 0) {
      const int ip = map[cell - map_w];
      if (ip >= 0 && ip != iq) h = (uint8_t)boundary_strength(cus[ip], q, ip, iq, cx * 4, cy * 4, 1, pic_type, rp, af);
    }
  }
  bs_v[cell] = v;
  bs_h[cell] = h;
}

// FilterEdgeLuma + CheckStrongFilter + FilterLumaWeak + FilterLumaStrong (:243-401) for one
// 4-line edge segment.  `across` steps over the edge, `along` steps along it.
__device__ void luma_segment(Sample *s, int across, int along, int bitdepth, int bs, int qp, int beta_off, int tc_off) {
  const int bd_shift = bitdepth - 8, maxv = (1 << bitdepth) - 1;
  // the reference clips the beta index to size() = 64, one past the table (:270-271); unreachable
  // for qp <= 51 with zero offsets -- defined here as the last entry
  const int beta = c_beta[min(clip3i(qp + beta_off, 0, 64), 63)] << bd_shift;
  int px[4][4], qx[4][4];   // [line][distance from the edge]
#pragma unroll
  for (int l = 0; l < 4; l++)
#pragma unroll
    for (int i = 0; i < 4; i++) {
      px[l][i] = s[l * along - (i + 1) * across];
      qx[l][i] = s[l * along + i * across];
    }
  const int dp0 = abs(px[0][2] - 2 * px[0][1] + px[0][0]), dq0 = abs(qx[0][0] - 2 * qx[0][1] + qx[0][2]);
  const int dp3 = abs(px[3][2] - 2 * px[3][1] + px[3][0]), dq3 = abs(qx[3][0] - 2 * qx[3][1] + qx[3][2]);
  const int d0 = dp0 + dq0, d3 = dp3 + dq3;
  if (d0 + d3 >= beta) return;
  const int tc = c_tc[clip3i(qp + tc_off + 2 * (bs - 1), 0, 53)] << bd_shift;
  bool strong = (d0 << 1) < (beta >> 2) && (d3 << 1) < (beta >> 2);
#pragma unroll
  for (int l = 0; l < 4; l += 3)
    strong = strong && (abs(px[l][3] - px[l][0]) + abs(qx[l][0] - qx[l][3])) < (beta >> 3) &&
             abs(px[l][0] - qx[l][0]) < ((tc * 5 + 1) >> 1);
  if (strong) {
    const int t2 = 2 * tc;
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int p3 = px[l][3], p2 = px[l][2], p1 = px[l][1], p0 = px[l][0];
      const int q0 = qx[l][0], q1 = qx[l][1], q2 = qx[l][2], q3 = qx[l][3];
      // delta clipped to +-2tc, narrowed to Sample, added with no final clip (:392-397)
#define XVCB_PUT(off, old, nv) s[l * along + (off) * across] = (Sample)((old) + (Sample)clip3i((nv) - (old), -t2, t2))
      XVCB_PUT(-3, p2, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      XVCB_PUT(-2, p1, (p2 + p1 + p0 + q0 + 2) >> 2);
      XVCB_PUT(-1, p0, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      XVCB_PUT(0, q0, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      XVCB_PUT(1, q1, (p0 + q0 + q1 + q2 + 2) >> 2);
      XVCB_PUT(2, q2, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
#undef XVCB_PUT
    }
    return;
  }
  const int side = (beta + (beta >> 1)) >> 3;
  const bool do_p1 = (dp0 + dp3) < side, do_q1 = (dq0 + dq3) < side;
  const int half = tc >> 1;
#pragma unroll
  for (int l = 0; l < 4; l++) {
    const int p2 = px[l][2], p1 = px[l][1], p0 = px[l][0], q0 = qx[l][0], q1 = qx[l][1], q2 = qx[l][2];
    int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
    if (abs(delta) >= tc * 10) continue;
    delta = clip3i(delta, -tc, tc);
    s[l * along - across] = (Sample)clip3i(p0 + delta, 0, maxv);
    s[l * along] = (Sample)clip3i(q0 - delta, 0, maxv);
    if (do_p1) s[l * along - 2 * across] = (Sample)clip3i(p1 + clip3i((((p2 + p0 + 1) >> 1) - p1 + delta) >> 1, -half, half), 0, maxv);
    if (do_q1) s[l * along + across] = (Sample)clip3i(q1 + clip3i((((q2 + q0 + 1) >> 1) - q1 - delta) >> 1, -half, half), 0, maxv);
  }
}

// FilterEdgeChroma + FilterChroma<N> (:403-450): N = 2 chroma lines per 4-sample luma segment in 4:2:0, 4 per
// 8-sample segment of the secondary tree
template <int NLINES>
__device__ void chroma_segment(Sample *s, int across, int along, int bitdepth, int qp, int tc_off) {
  const int tc = c_tc[min(clip3i(qp + tc_off + 2, 0, 54), 53)] << (bitdepth - 8);   // index clip as in luma
  const int maxv = (1 << bitdepth) - 1;
#pragma unroll
  for (int l = 0; l < NLINES; l++) {
    const int p1 = s[l * along - 2 * across], p0 = s[l * along - across], q0 = s[l * along], q1 = s[l * along + across];
    const int delta = clip3i((((q0 - p0) * 4) + p1 - q1 + 4) >> 3, -tc, tc);
    s[l * along - across] = (Sample)clip3i(p0 + delta, 0, maxv);
    s[l * along] = (Sample)clip3i(q0 - delta, 0, maxv);
  }
}

template <int DIR>   // 0: vertical edges, 1: horizontal edges
__global__ void __launch_bounds__(128) deblock_kernel(const xvcb200_cu *__restrict__ cus, const int32_t *__restrict__ map,
                                                      const uint8_t *__restrict__ bs_arr, int map_w, int map_h,
                                                      DeblockParams prm, Pic3 rec, int cy_begin, int cy_end) {
  // thread -> 4x4 block; for DIR 1 consecutive threads still walk along x so loads coalesce.
  // [cy_begin, cy_end): band of 4-row groups this launch owns (CTB-row sharding); an edge row
  // belongs to the band that holds its q side.
  const int cell = blockIdx.x * blockDim.x + threadIdx.x + cy_begin * map_w;
  if (cell >= map_w * cy_end) return;
  const int cx = cell % map_w, cy = cell / map_w;
  const int bs0 = bs_arr[cell];
  if (bs0 == 0) return;
  const int step = DIR == 0 ? 1 : map_w;                 // next edge of the chain
  const int pos = DIR == 0 ? cx : cy, lim = DIR == 0 ? map_w : cy_end;
  const int first = DIR == 0 ? 0 : cy_begin;             // chains restart at the top of a band
  const PlaneView ly = rec.p[0];
  const int across = DIR == 0 ? 1 : ly.pitch, along = DIR == 0 ? ly.pitch : 1;

  // chroma: bs == 2 edges whose chroma coordinate is a multiple of 8 (:131-149); with a secondary CU tree the
  // chroma edges are that tree's (:88-91)
  if (bs0 == 2 && ((DIR == 0 ? cx : cy) & 3) == 0 && !prm.chroma_cus) {
    const int iq = map[cell], ip = map[cell - step];
    const int qpp = chroma_qp_raw(cus[ip].qp, prm.off_u, prm.table, c_db_chroma_scale);   // cu.GetQp(kU) for both planes (:127)
    const int qpq = chroma_qp_raw(cus[iq].qp, prm.off_u, prm.table, c_db_chroma_scale);
    const int cqp = (qpp + qpq + 1) >> 1;
    for (int c = 1; c < 3; c++) {
      const PlaneView pc = rec.p[c];
      Sample *s = pc.base + (cy * 2) * pc.pitch + cx * 2;
      chroma_segment<2>(s, DIR == 0 ? 1 : pc.pitch, DIR == 0 ? pc.pitch : 1, prm.bitdepth, cqp, prm.tc_offset);
    }
  }

  // luma: only the head of a chain works
  if (pos > first && bs_arr[cell - step] != 0) return;
  int e = cell;
  for (int k = pos; k < lim; k++, e += step) {
    const int bs = bs_arr[e];
    if (bs == 0) break;
    const int iq = map[e], ip = map[e - step];
    const int qp = (cus[ip].qp + cus[iq].qp + 1) >> 1;
    const int ex = DIR == 0 ? k : cx, ey = DIR == 0 ? cy : k;
    Sample *s = ly.base + (ey * 4) * ly.pitch + ex * 4;
    luma_segment(s, across, along, prm.bitdepth, bs, qp, prm.beta_offset, prm.tc_offset);
  }
}

// Chroma edges of the secondary CU tree (intra pictures, deblocking_filter.cc:65-68, 88-91): DeblockCtu walks that
// tree on an 8-sample luma grid; an edge between two of its CUs with boundary strength 2 is filtered where its
// chroma coordinate is a multiple of 8, four chroma lines per segment.  A thread per 8 x 8 luma cell.
template <int DIR>
__global__ void __launch_bounds__(128) deblock_chroma_tree_kernel(const xvcb200_cu *__restrict__ cus2, const int32_t *__restrict__ map2,
                                                                  int map_w, int map_h, DeblockParams prm, Pic3 rec) {
  const int gw = map_w >> 1, gh = map_h >> 1;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= gw * gh) return;
  const int cx = (g % gw) * 2, cy = (g / gw) * 2;            // 4x4 cell of the luma position (8 gx, 8 gy)
  if (((DIR == 0 ? cx : cy) & 3) != 0) return;                // chroma coordinate (4 cx / 2) multiple of 8
  if ((DIR == 0 ? cx : cy) == 0) return;
  const int cell = cy * map_w + cx, step = DIR == 0 ? 1 : map_w;
  const int iq = map2[cell], ip = map2[cell - step];
  if (iq < 0 || ip < 0 || ip == iq) return;
  DbRefPoc rp;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) rp.poc[l][i] = prm.ref_poc[l][i];
  const DbAffine none = {nullptr, nullptr};
  if (boundary_strength(cus2[ip], cus2[iq], ip, iq, cx * 4, cy * 4, DIR, prm.pic_type, rp, none) != 2) return;
  const int qpp = chroma_qp_raw(cus2[ip].qp, prm.off_u, prm.table, c_db_chroma_scale);
  const int qpq = chroma_qp_raw(cus2[iq].qp, prm.off_u, prm.table, c_db_chroma_scale);
  const int cqp = (qpp + qpq + 1) >> 1;
  for (int c = 1; c < 3; c++) {
    const PlaneView pc = rec.p[c];
    Sample *s = pc.base + (cy * 2) * pc.pitch + cx * 2;
    chroma_segment<4>(s, DIR == 0 ? 1 : pc.pitch, DIR == 0 ? pc.pitch : 1, prm.bitdepth, cqp, prm.tc_offset);
  }
}

// CU index of every 4x4 block: depends on the CU geometry only, so the picture pipeline builds it
// on a side stream while the search runs.
cudaError_t launch_cu_map(cudaStream_t s, const xvcb200_cu *d_cus, int n, int32_t *d_map, int map_w, int map_h) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(d_map, 0xff, sizeof(int32_t) * map_w * map_h, s);
  if (e != cudaSuccess) return e;
  g_launch_count++;
  cu_map_kernel<<<n, 64, 0, s>>>(d_cus, n, d_map, map_w, map_h);
  return cudaGetLastError();
}

cudaError_t launch_deblock(cudaStream_t s, const xvcb200_cu *d_cus, int n, const DeblockParams &p, Pic3 rec,
                           int32_t *d_map, uint8_t *d_bs_v, uint8_t *d_bs_h, int map_w, int map_h, int pass_mask,
                           int y_begin, int y_end, bool map_ready) {
  if (n <= 0) return cudaSuccess;
  const int cells = map_w * map_h;
  const int cy0 = y_begin >> 2, cy1 = y_end >> 2;
  const int band_cells = map_w * (cy1 - cy0);
  if (band_cells <= 0) return cudaSuccess;
  if (!map_ready) {
    cudaError_t e = launch_cu_map(s, d_cus, n, d_map, map_w, map_h);
    if (e != cudaSuccess) return e;
  }
  DbRefPoc rp;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) rp.poc[l][i] = p.ref_poc[l][i];
  g_launch_count++;
  const DbAffine af = {p.aff_index, p.aff};
  edge_bs_kernel<<<(cells + 127) / 128, 128, 0, s>>>(d_cus, d_map, map_w, map_h, p.pic_type, rp, af, d_bs_v, d_bs_h);
  const bool tree2 = p.chroma_cus && p.n_chroma_cus > 0 && p.chroma_map;
  const int cells2 = (map_w >> 1) * (map_h >> 1);
  if (tree2) {
    cudaError_t e = launch_cu_map(s, p.chroma_cus, p.n_chroma_cus, p.chroma_map, map_w, map_h);
    if (e != cudaSuccess) return e;
  }
  // luma and chroma touch different planes; within chroma the vertical pass precedes the horizontal one (:62-76)
  if (pass_mask & 1) {
    g_launch_count++;
    deblock_kernel<0><<<(band_cells + 127) / 128, 128, 0, s>>>(d_cus, d_map, d_bs_v, map_w, map_h, p, rec, cy0, cy1);
    if (tree2) {
      g_launch_count++;
      deblock_chroma_tree_kernel<0><<<(cells2 + 127) / 128, 128, 0, s>>>(p.chroma_cus, p.chroma_map, map_w, map_h, p, rec);
    }
  }
  if (pass_mask & 2) {
    g_launch_count++;
    deblock_kernel<1><<<(band_cells + 127) / 128, 128, 0, s>>>(d_cus, d_map, d_bs_h, map_w, map_h, p, rec, cy0, cy1);
    if (tree2) {
      g_launch_count++;
      deblock_chroma_tree_kernel<1><<<(cells2 + 127) / 128, 128, 0, s>>>(p.chroma_cus, p.chroma_map, map_w, map_h, p, rec);
    }
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------- PadBorder
// Every sample outside the picture takes the value of the nearest picture sample (rows are
// replicated first, then columns over all rows incl. the new ones: corners = corner samples).
// One thread per BORDER sample of the three planes (blockIdx.y = plane): the two bands of `pad`
// rows above / below over the full padded width, then the two side bands of the picture rows.
struct PadPlanes { PlaneView p[3]; int pad[3]; };
__global__ void pad_border_kernel(const __grid_constant__ PadPlanes pp) {
  const PlaneView pl = pp.p[blockIdx.y];
  const int pad = pp.pad[blockIdx.y];
  const int fw = pl.width + 2 * pad, band = pad * fw;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int x, y;
  if (idx < 2 * band) {
    const int i = idx < band ? idx : idx - band;
    const int r = i / fw;
    x = i - r * fw - pad;
    y = idx < band ? r - pad : pl.height + r;
  } else {
    const int i = idx - 2 * band;
    if (i >= 2 * pad * pl.height) return;
    y = i / (2 * pad);
    const int c = i - y * 2 * pad;
    x = c < pad ? c - pad : pl.width + c - pad;
  }
  const int sx = clip3i(x, 0, pl.width - 1), sy = clip3i(y, 0, pl.height - 1);
  pl.base[y * pl.pitch + x] = pl.base[sy * pl.pitch + sx];
}

cudaError_t launch_pad_border(cudaStream_t s, Pic3 pic, const int pad[3]) {
  PadPlanes pp;
  int most = 0;
  for (int c = 0; c < 3; c++) {
    pp.p[c] = pic.p[c];
    pp.pad[c] = pad[c];
    const int total = 2 * pad[c] * (pic.p[c].width + 2 * pad[c]) + 2 * pad[c] * pic.p[c].height;
    most = total > most ? total : most;
  }
  g_launch_count++;
  pad_border_kernel<<<dim3((most + 255) / 256, 3), 256, 0, s>>>(pp);
  return cudaGetLastError();
}

// ---------------------------------------------------------------- picture <-> tight staging
// The three planes of a picture between the padded slot layout and one tight buffer (planes
// back to back, row pitch = width), so that the PCIe transfer is one contiguous copy per plane
// instead of one DMA descriptor per row.  Widths are multiples of 4 samples (8-byte vectors).
__global__ void plane_pack_kernel(Pic3 pic, uint16_t *__restrict__ tight, int to_tight) {
  const int comp = blockIdx.y;
  const PlaneView pl = pic.p[comp];
  size_t off = 0;
  for (int c = 0; c < comp; c++) off += (size_t)pic.p[c].width * pic.p[c].height;
  const int vw = pl.width >> 2;                       // 8-byte vectors per row
  const int total = vw * pl.height;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int y = i / vw, x = (i - y * vw) * 4;
    uint2 *t = reinterpret_cast<uint2 *>(tight + off + (size_t)y * pl.width + x);
    uint2 *s = reinterpret_cast<uint2 *>(pl.base + y * pl.pitch + x);
    if (to_tight) *t = *s; else *s = *t;
  }
}

cudaError_t launch_plane_pack(cudaStream_t s, Pic3 pic, uint16_t *tight, int to_tight) {
  g_launch_count++;
  plane_pack_kernel<<<dim3(296, 3), 256, 0, s>>>(pic, tight, to_tight);
  return cudaGetLastError();
}

}  // namespace xvcb
