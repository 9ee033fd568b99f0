// GPU pre-analysis: the CU partition of an inter picture.
//
// The reference decides the partition inside CuEncoder::CompressCu's RD recursion (cu_encoder.cc:123-273:
// every split shape x every mode x every transform candidate, each through the whole T/Q chain and a CABAC
// bit count) -- serial, and the reason its encoder spends ~90 core-seconds on a 1080p picture.  There is no
// reference behaviour to be exact against here; the contract is (a) a partition xvc's syntax can carry
// (quad splits down to 8 x 8, a quad-tree leaf may be split once more horizontally or vertically: the trees
// workload.make_partition_tree produces and oracle/ref_shim.cc's writer signals), and (b) decided from
// motion-compensated distortion, exactly reproducible on a CPU (tests/partition_model.py is the numpy
// statement of the same rule; the kernel is compared with it bit for bit).
//
// One CTA per CTU:
//   1. the reference window (64 + 2R)^2 around the CTU displaced by the picture-level predictor and the
//      original CTU are staged in shared memory (window coordinates clamped to the padded plane, which
//      continues the picture by replication);
//   2. SAD of every 8 x 8 block at every full-pel vector of the +-R window: T8[block][vector], the
//      "SAD tree" leaves (R = 8: 64 x 289 values in shared memory);
//   3. bottom-up over 16 x 16, 32 x 32, 64 x 64 nodes: a node's table is the sum of its four children's, so
//      the distortion of a CU of any shape at any vector of the window is a sum of table entries.  Per node,
//      cost = min over vectors (SAD + ((lambda * mvd bits) >> 16)) + header, for: not split, split
//      horizontally (two halves, each its own vector), vertically, or into four (the children's best
//      costs).  A warp per node, lanes over the vectors, min through REDUX on (cost << 9 | vector);
//   4. top-down emission of the CUs in coding order with the winning vector as the CU's mv[0] (the
//      predictor of the search that follows) and of the split flags, one thread.
// Nodes that cross the picture edge are split; parts outside are dropped.
#include "xvcb_device.cuh"

namespace xvcb {

constexpr int kPaR = 8;                          // vectors of the window: (2R + 1)^2
constexpr int kPaSide = 2 * kPaR + 1;
constexpr int kPaVec = kPaSide * kPaSide;        // 289
constexpr int kPaWin = 64 + 2 * kPaR;            // 80 samples
constexpr int kPaWinPitch = kPaWin / 2 + 1;      // words per window row (+1: the word the funnel shift reads ahead)
constexpr int kPaThreads = 256;
constexpr uint32_t kPaInf = 0x3fffffffu;

struct PaNode {                                  // decision of one quad-tree node
  uint32_t best;                                 // cost of the best alternative
  uint8_t split;                                 // 0 none, 1 quad, 2 horizontal, 3 vertical, 255: outside the picture
  uint8_t inside;                                // 1: wholly inside the picture
  uint16_t mv_none, mv_a, mv_b;                  // vector index (my * side + mx): unsplit / first half / second half
};

__device__ __forceinline__ uint32_t pa_rate(int m, uint32_t lambda) {
  const int dy = m / kPaSide - kPaR, dx = m % kPaSide - kPaR;
  return (lambda * (exp_golomb_bits(dx * 4) + exp_golomb_bits(dy * 4))) >> 16;      // mvd in quarter samples (cu_types.h:124-144)
}

__global__ void __launch_bounds__(kPaThreads) partition_kernel(PlaneView orig, PlaneView ref, int cx16, int cy16, uint32_t lambda,
                                                               int hdr_cu, int hdr_split, int qp, xvcb200_cu *__restrict__ cus_out,
                                                               int *__restrict__ n_cus_out, uint8_t *__restrict__ splits_out,
                                                               int *__restrict__ n_splits_out) {
  extern __shared__ __align__(16) uint32_t pa_smem[];
  uint32_t *s_win = pa_smem;                                   // reference window, packed pairs
  uint32_t *s_org = s_win + kPaWin * kPaWinPitch;              // original CTU, packed pairs
  uint32_t *s_t16 = s_org + 64 * 32;                           // tables of the 16 x 16 and 32 x 32 nodes
  uint32_t *s_t32 = s_t16 + 16 * kPaVec;
  uint32_t *s_rate = s_t32 + 4 * kPaVec;
  uint16_t *s_t8 = reinterpret_cast<uint16_t *>(s_rate + kPaVec + 1);   // SAD of 8 x 8 block b = by * 8 + bx at vector m, saturated at 65 535
  __shared__ PaNode s_n8[64], s_n16[16], s_n32[4], s_n64[1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ctus_x = (orig.width + 63) >> 6;
  const int ctu = blockIdx.x, x0 = (ctu % ctus_x) * 64, y0 = (ctu / ctus_x) * 64;
  const int cx = cx16 >> 4, cy = cy16 >> 4;                    // window centre, full samples

  // 1. staging
  for (int i = tid; i < kPaWin * (kPaWin / 2); i += kPaThreads) {
    const int r = i / (kPaWin / 2), c = (i - r * (kPaWin / 2)) * 2;
    const int gy = clip3i(y0 + cy - kPaR + r, -80, ref.height + 79);
    const int gx0 = clip3i(x0 + cx - kPaR + c, -80, ref.width + 79), gx1 = clip3i(x0 + cx - kPaR + c + 1, -80, ref.width + 79);
    s_win[r * kPaWinPitch + (c >> 1)] = (uint32_t)ref.base[gy * ref.pitch + gx0] | ((uint32_t)ref.base[gy * ref.pitch + gx1] << 16);
  }
  for (int i = tid; i < kPaWin; i += kPaThreads) s_win[i * kPaWinPitch + kPaWin / 2] = 0;
  for (int i = tid; i < 64 * 32; i += kPaThreads) {
    const int r = i >> 5, c = (i & 31) * 2;
    const int gy = min(y0 + r, orig.height - 1), gx = min(x0 + c, orig.width - 2);      // outside the picture: never used
    s_org[i] = *reinterpret_cast<const uint32_t *>(orig.base + gy * orig.pitch + gx);
  }
  for (int m = tid; m < kPaVec; m += kPaThreads) s_rate[m] = pa_rate(m, lambda);
  __syncthreads();

  // 2. T8 (64 * 1023 fits 16 bits; 12-bit content saturates at 65 535 where a block is off by > 1023 on average).
  // A thread takes one block and one row of the vector window (17 vectors, my fixed): per block row the 12 window
  // words and the 4 original words are loaded once and serve all 17 vectors -- even mx read the words as they
  // are, odd mx the 11 words shifted by one sample (one funnel shift each, shared by the 8 odd vectors).  The
  // per-lane 16-bit sums are folded every 4 rows (4 rows x 4 pairs x 4095 = 65 520 fits).
  for (int p = tid; p < 64 * kPaSide; p += kPaThreads) {
    const int b = p / kPaSide, my = p - b * kPaSide;
    const int by = b >> 3, bx = b & 7;
    const uint32_t *w = s_win + (by * 8 + my) * kPaWinPitch + bx * 4;
    const uint32_t *o = s_org + by * 8 * 32 + bx * 4;
    uint32_t sad[kPaSide], acc[kPaSide];
#pragma unroll
    for (int m = 0; m < kPaSide; m++) sad[m] = acc[m] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++) {
      uint32_t ww[12], sw[11], oo[4];
#pragma unroll
      for (int c = 0; c < 12; c++) ww[c] = w[c];
#pragma unroll
      for (int c = 0; c < 4; c++) oo[c] = o[c];
#pragma unroll
      for (int c = 0; c < 11; c++) sw[c] = __funnelshift_r(ww[c], ww[c + 1], 16);
#pragma unroll
      for (int m = 0; m < kPaSide; m++) {
        const int k = m >> 1;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const uint32_t v = (m & 1) ? sw[k + c] : ww[k + c];
          acc[m] += __vmaxu2(oo[c], v) - __vminu2(oo[c], v);
        }
      }
      if ((r & 3) == 3) {
#pragma unroll
        for (int m = 0; m < kPaSide; m++) { sad[m] += (acc[m] & 0xffffu) + (acc[m] >> 16); acc[m] = 0; }
      }
      w += kPaWinPitch; o += 32;
    }
    uint16_t *t = s_t8 + b * kPaVec + my * kPaSide;
#pragma unroll
    for (int m = 0; m < kPaSide; m++) t[m] = (uint16_t)min(sad[m], 65535u);
  }
  __syncthreads();

  // 3a. 8 x 8 nodes: only "not split"
  for (int b = warp; b < 64; b += kPaThreads / 32) {
    const int by = b >> 3, bx = b & 7;
    const bool inside = x0 + bx * 8 + 8 <= orig.width && y0 + by * 8 + 8 <= orig.height;
    uint32_t key = 0xffffffffu;
    for (int m = lane; m < kPaVec; m += 32) key = min(key, (((uint32_t)s_t8[b * kPaVec + m] + s_rate[m]) << 9) | (uint32_t)m);
    key = __reduce_min_sync(XVCB_FULL, key);
    if (lane == 0) {
      PaNode n;
      n.inside = inside; n.split = inside ? 0 : 255;
      n.best = inside ? (key >> 9) + hdr_cu : 0;          // a block outside the picture costs nothing and is dropped
      n.mv_none = (uint16_t)(key & 511); n.mv_a = n.mv_b = 0;
      s_n8[b] = n;
    }
  }
  // tables of the 16 x 16 nodes
  for (int p = tid; p < 16 * kPaVec; p += kPaThreads) {
    const int q = p / kPaVec, m = p - q * kPaVec, qy = q >> 2, qx = q & 3;
    const int b0 = (qy * 2) * 8 + qx * 2;
    s_t16[p] = (uint32_t)s_t8[b0 * kPaVec + m] + s_t8[(b0 + 1) * kPaVec + m] + s_t8[(b0 + 8) * kPaVec + m] + s_t8[(b0 + 9) * kPaVec + m];
  }
  __syncthreads();

  // One node of size S from its four children: tables c0 (top-left), c1 (top-right), c2 (bottom-left), c3
  // (bottom-right) and the children's decisions.  Executed by one warp.
  auto decide = [&](auto tab, int i0, int i1, int i2, int i3, const PaNode &k0, const PaNode &k1, const PaNode &k2, const PaNode &k3,
                    int px, int py, int size, PaNode *out) {
    const bool inside = px + size <= orig.width && py + size <= orig.height;
    const bool any = px < orig.width && py < orig.height;
    uint32_t kn = 0xffffffffu, kt = 0xffffffffu, kb = 0xffffffffu, kl = 0xffffffffu, kr = 0xffffffffu;
    if (inside)
      for (int m = lane; m < kPaVec; m += 32) {
        const uint32_t a = tab[i0 * kPaVec + m], b = tab[i1 * kPaVec + m], c = tab[i2 * kPaVec + m], d = tab[i3 * kPaVec + m], r = s_rate[m];
        kn = min(kn, ((a + b + c + d + r) << 9) | (uint32_t)m);
        kt = min(kt, ((a + b + r) << 9) | (uint32_t)m);
        kb = min(kb, ((c + d + r) << 9) | (uint32_t)m);
        kl = min(kl, ((a + c + r) << 9) | (uint32_t)m);
        kr = min(kr, ((b + d + r) << 9) | (uint32_t)m);
      }
    kn = __reduce_min_sync(XVCB_FULL, kn); kt = __reduce_min_sync(XVCB_FULL, kt); kb = __reduce_min_sync(XVCB_FULL, kb);
    kl = __reduce_min_sync(XVCB_FULL, kl); kr = __reduce_min_sync(XVCB_FULL, kr);
    if (lane == 0) {
      PaNode n;
      n.inside = inside; n.mv_none = n.mv_a = n.mv_b = 0;
      const uint32_t quad = k0.best + k1.best + k2.best + k3.best + hdr_split;
      if (!any) { n.split = 255; n.best = 0; }
      else if (!inside) { n.split = 1; n.best = quad; }                       // crosses the picture edge: split
      else {
        const uint32_t none = (kn >> 9) + hdr_cu + hdr_split;
        const uint32_t hor = (kt >> 9) + (kb >> 9) + 2 * hdr_cu + 2 * hdr_split;
        const uint32_t ver = (kl >> 9) + (kr >> 9) + 2 * hdr_cu + 2 * hdr_split;
        // ties: the coarser alternative (fewer CUs) wins, in the order none, horizontal, vertical, quad
        n.split = 0; n.best = none; n.mv_none = (uint16_t)(kn & 511);
        if (hor < n.best) { n.split = 2; n.best = hor; }
        if (ver < n.best) { n.split = 3; n.best = ver; }
        if (quad < n.best) { n.split = 1; n.best = quad; }
        if (n.split == 2) { n.mv_a = (uint16_t)(kt & 511); n.mv_b = (uint16_t)(kb & 511); }
        if (n.split == 3) { n.mv_a = (uint16_t)(kl & 511); n.mv_b = (uint16_t)(kr & 511); }
      }
      *out = n;
    }
  };

  // 3b. 16 x 16 nodes (children: 8 x 8 blocks, tables in s_t8)
  for (int q = warp; q < 16; q += kPaThreads / 32) {
    const int qy = q >> 2, qx = q & 3, b0 = (qy * 2) * 8 + qx * 2;
    decide(s_t8, b0, b0 + 1, b0 + 8, b0 + 9, s_n8[b0], s_n8[b0 + 1], s_n8[b0 + 8], s_n8[b0 + 9], x0 + qx * 16, y0 + qy * 16, 16, &s_n16[q]);
  }
  for (int p = tid; p < 4 * kPaVec; p += kPaThreads) {
    const int q = p / kPaVec, m = p - q * kPaVec, q0 = ((q >> 1) * 2) * 4 + (q & 1) * 2;
    s_t32[p] = s_t16[q0 * kPaVec + m] + s_t16[(q0 + 1) * kPaVec + m] + s_t16[(q0 + 4) * kPaVec + m] + s_t16[(q0 + 5) * kPaVec + m];
  }
  __syncthreads();
  // 3c. 32 x 32 nodes
  for (int q = warp; q < 4; q += kPaThreads / 32) {
    const int q0 = ((q >> 1) * 2) * 4 + (q & 1) * 2;
    decide(s_t16, q0, q0 + 1, q0 + 4, q0 + 5, s_n16[q0], s_n16[q0 + 1], s_n16[q0 + 4], s_n16[q0 + 5], x0 + (q & 1) * 32, y0 + (q >> 1) * 32, 32,
           &s_n32[q]);
  }
  __syncthreads();
  // 3d. the CTU
  if (warp == 0) decide(s_t32, 0, 1, 2, 3, s_n32[0], s_n32[1], s_n32[2], s_n32[3], x0, y0, 64, &s_n64[0]);
  __syncthreads();

  // 4. emission in coding order (pre-order; quadrants top-left, top-right, bottom-left, bottom-right)
  if (tid == 0) {
    xvcb200_cu *cus = cus_out + (size_t)ctu * 64;
    uint8_t *splits = splits_out + (size_t)ctu * 128;
    int nc = 0, ns = 0;
    auto leaf = [&](int x, int y, int w, int h, int depth, int m) {
      xvcb200_cu u;
      memset(&u, 0, sizeof(u));
      u.x = (int16_t)x; u.y = (int16_t)y; u.w = (uint8_t)w; u.h = (uint8_t)h; u.depth = (uint8_t)depth; u.qp = (int8_t)qp;
      u.ref_idx[0] = u.ref_idx[1] = -1;
      u.mv[0][0] = (cx + m % kPaSide - kPaR) * 16; u.mv[0][1] = (cy + m / kPaSide - kPaR) * 16;
      u.mv[1][0] = u.mv[0][0]; u.mv[1][1] = u.mv[0][1];
      cus[nc++] = u;
    };
    auto node = [&](const PaNode &n, int x, int y, int size, int depth, bool &quad) {      // everything but the recursion
      quad = false;
      if (n.split == 255) return;
      if (n.inside || n.split != 1) splits[ns++] = n.split;                 // a split forced by the picture edge is implicit
      if (n.split == 0) leaf(x, y, size, size, depth, n.mv_none);
      else if (n.split == 2) { splits[ns++] = 0; leaf(x, y, size, size / 2, depth, n.mv_a); splits[ns++] = 0; leaf(x, y + size / 2, size, size / 2, depth, n.mv_b); }
      else if (n.split == 3) { splits[ns++] = 0; leaf(x, y, size / 2, size, depth, n.mv_a); splits[ns++] = 0; leaf(x + size / 2, y, size / 2, size, depth, n.mv_b); }
      else quad = true;
    };
    bool q64, q32, q16, q8;
    node(s_n64[0], x0, y0, 64, 0, q64);
    if (q64)
      for (int a = 0; a < 4; a++) {
        const int ax = x0 + (a & 1) * 32, ay = y0 + (a >> 1) * 32;
        node(s_n32[a], ax, ay, 32, 1, q32);
        if (!q32) continue;
        for (int b = 0; b < 4; b++) {
          const int i16 = ((a >> 1) * 2 + (b >> 1)) * 4 + (a & 1) * 2 + (b & 1);
          const int bx = ax + (b & 1) * 16, by = ay + (b >> 1) * 16;
          node(s_n16[i16], bx, by, 16, 2, q16);
          if (!q16) continue;
          for (int c = 0; c < 4; c++) {
            const int i8 = ((i16 >> 2) * 2 + (c >> 1)) * 8 + (i16 & 3) * 2 + (c & 1);
            node(s_n8[i8], bx + (c & 1) * 8, by + (c >> 1) * 8, 8, 3, q8);
          }
        }
      }
    n_cus_out[ctu] = nc;
    n_splits_out[ctu] = ns;
  }
}

// cus_out: [ctu][64], splits_out: [ctu][128], counts per CTU
cudaError_t launch_partition(cudaStream_t s, PlaneView orig, PlaneView ref, int cx16, int cy16, uint32_t lambda_me, int hdr_cu, int hdr_split,
                             int qp, xvcb200_cu *d_cus, int *d_n_cus, uint8_t *d_splits, int *d_n_splits) {
  const int n_ctus = ((orig.width + 63) >> 6) * ((orig.height + 63) >> 6);
  const int smem = (kPaWin * kPaWinPitch + 64 * 32 + 16 * kPaVec + 4 * kPaVec + kPaVec + 1) * 4 + 64 * kPaVec * 2;
  cudaError_t e = cudaFuncSetAttribute(partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per device; cheap
  if (e != cudaSuccess) return e;
  g_launch_count++;
  partition_kernel<<<n_ctus, kPaThreads, smem, s>>>(orig, ref, cx16, cy16, lambda_me, hdr_cu, hdr_split, qp, d_cus, d_n_cus, d_splits, d_n_splits);
  return cudaGetLastError();
}

}  // namespace xvcb
