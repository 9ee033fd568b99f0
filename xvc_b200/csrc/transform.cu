// Integer transforms, quantisation and the fused residual-coding chain.
//
//  * single-block kernels behind the class-method-shaped C ABI (ForwardTransform::Transform,
//    InverseTransform::Transform, RdoQuant::QuantFast, Quantize::Inverse) for every
//    transform type and shape;
//  * tq_kernel<LW,LH>: TransformEncoder::TransformAndReconstruct (transform_encoder.cc:203-285)
//    for every transform unit of a picture, one launch per block shape.
//
// Arithmetic contract (unrestricted mode, transform.cc:83-182, 869-961): 8-bit "High"
// matrices (DC = 256) for every size; forward = rows (shift log2w + bd - 9 + 2) then columns
// (shift log2h + 6 + 2), stored to int16 WITHOUT clipping; inverse = columns (shift 9) then
// rows (shift 20 - bd + 2), every output clipped to int16; for 64-point dimensions only the
// first 32 coefficients are produced / consumed (kTransformZeroOutMinSize).
#include "xvcb_device.cuh"
#include "xvcb_tables.inc"

namespace xvcb {

// all matrices, int16, global memory (L1/L2 resident): used by the any-type single-block path
__device__ const int16_t g_mats[XVCB_MAT_TOTAL] = {XVCB_MAT_VALUES};
__constant__ int c_mat_off[6][7] = XVCB_MAT_OFFSETS;
static const int16_t h_mats[XVCB_MAT_TOTAL] = {XVCB_MAT_VALUES};
// DCT-2 matrices as int32 in constant memory: warp-uniform operand of the fused kernel's IMADs
__constant__ int c_dct2[XVCB_MAT_DCT2_TOTAL];
static bool g_dct2_loaded[16] = {false};

// The same matrices for the fused kernel, int32 in global memory (L1 resident; the threads of a
// transform unit read the same 16 bytes).  Per transform type and size N (offset = c_mat_off[type][log2 N],
// the offset of the N x N matrix above; the 4 x 4 DST of intra luma blocks follows at XVCB_MAT_TOTAL),
// with R = min(N, 32) the number of coefficients a 64-point transform keeps (kTransformZeroOutMinSize):
//   g_tq_fwdT[j * R + k] = m[k][j]   forward: the R outputs of one input sample are contiguous
//   g_tq_inv [k * N + j] = m[k][j]   inverse: the N outputs of one coefficient are contiguous
constexpr int kTqDst4Off = XVCB_MAT_TOTAL;
constexpr int kTqTableInts = XVCB_MAT_TOTAL + 16;
__device__ __align__(16) int g_tq_fwdT[kTqTableInts];
__device__ __align__(16) int g_tq_inv[kTqTableInts];
static const int h_mat_off[6][7] = XVCB_MAT_OFFSETS;
static const int h_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

static cudaError_t ensure_dct2_constant() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 16 && g_dct2_loaded[dev]) return cudaSuccess;
  static int h_dct2[XVCB_MAT_DCT2_TOTAL], h_fwdT[kTqTableInts], h_inv[kTqTableInts];
  for (int i = 0; i < XVCB_MAT_DCT2_TOTAL; i++) h_dct2[i] = h_mats[i];
  for (int type = XVCB200_TX_DCT2; type <= XVCB200_TX_DST7; type++)
    for (int lg = (type == XVCB200_TX_DCT2 ? 1 : 2); lg <= 6; lg++) {
      const int n = 1 << lg, r = n > 32 ? 32 : n, off = h_mat_off[type][lg];
      for (int k = 0; k < r; k++)
        for (int j = 0; j < n; j++) {
          h_fwdT[off + j * r + k] = h_mats[off + k * n + j];
          h_inv[off + k * n + j] = h_mats[off + k * n + j];
        }
    }
  for (int k = 0; k < 4; k++)
    for (int j = 0; j < 4; j++) {
      h_fwdT[kTqDst4Off + j * 4 + k] = h_dst4[k][j];
      h_inv[kTqDst4Off + k * 4 + j] = h_dst4[k][j];
    }
  cudaError_t e = cudaMemcpyToSymbol(c_dct2, h_dct2, sizeof(h_dct2));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tq_fwdT, h_fwdT, sizeof(h_fwdT));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_tq_inv, h_inv, sizeof(h_inv));
  if (e == cudaSuccess && dev < 16) g_dct2_loaded[dev] = true;
  return e;
}

// Qp::kChromaScale_, kFwdQuantScales_, kInvQuantScales_ (quantize.cc:34-46)
__constant__ uint8_t c_chroma_scale[58] = {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19,
                                           20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 29, 30, 31, 32, 33, 33, 34, 34, 35, 35,
                                           36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};
__constant__ int c_fwd_scale[6] = {26214, 23302, 20560, 18396, 16384, 14564};
__constant__ int c_inv_scale[6] = {40, 45, 51, 57, 64, 72};
// TransformHelper::kScanCoeff4x4 (transform.cc:70-76): diagonal, horizontal, vertical
__constant__ uint8_t c_scan4x4[3][16] = {{0, 4, 1, 8, 5, 2, 12, 9, 6, 3, 13, 10, 7, 14, 11, 15},
                                         {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15},
                                         {0, 4, 8, 12, 1, 5, 9, 13, 2, 6, 10, 14, 3, 7, 11, 15}};

__device__ __forceinline__ int clip16(int v) { return clip3i(v, -32768, 32767); }

// ================================================================ shared device pieces

struct QuantParams { int shift, scale; long long offset; };
// RdoQuant::QuantFast set-up, rdo_quant.cc:160-170
__device__ __forceinline__ QuantParams quant_params(int lw, int lh, int bitdepth, int qp_bd, int intra_pic) {
  QuantParams q;
  const int odd = (lw + lh) & 1;
  q.shift = 14 + qp_bd / 6 + transform_shift(lw, lh, bitdepth) + (odd ? 7 : 0);
  q.scale = c_fwd_scale[qp_bd % 6] * (odd ? 181 : 1);
  q.offset = (long long)(intra_pic ? 171 : 85) << (q.shift - 9);
  return q;
}
// one coefficient of QuantFast (rdo_quant.cc:181-192): level (sign restored, clipped) and delta
__device__ __forceinline__ void quant_one(int c, const QuantParams &q, int16_t &level_out, int16_t &delta_out) {
  const long long mag = (long long)abs(c) * q.scale;
  const int level = (int)((mag + q.offset) >> q.shift);
  level_out = (int16_t)clip16(c < 0 ? -level : level);
  delta_out = (int16_t)((mag - ((long long)level << q.shift)) >> (q.shift - 8));
}
// Quantize::Inverse set-up and one coefficient, quantize.cc:94-125
struct DequantParams { int shift, scale; };
__device__ __forceinline__ DequantParams dequant_params(int lw, int lh, int bitdepth, int qp_bd) {
  DequantParams d;
  const int odd = (lw + lh) & 1;
  d.shift = 6 - transform_shift(lw, lh, bitdepth) + (odd ? 8 : 0);
  d.scale = (c_inv_scale[qp_bd % 6] << (qp_bd / 6)) * (odd ? 181 : 1);
  return d;
}
__device__ __forceinline__ int16_t dequant_one(int level, const DequantParams &d) {
  int v = level * d.scale;
  v = d.shift > 0 ? (v + (1 << (d.shift - 1))) >> d.shift : (int)((unsigned)v << -d.shift);
  return (int16_t)clip16(v);
}

// Index of sub-block (sx, sy) in the sub-block scan (TransformHelper::DeriveSubblockScan,
// transform.cc:1638-1680) of a bw x bh grid, in closed form.  Diagonal: anti-diagonals in
// increasing x+y, each walked from its bottom-left cell towards the top-right.
__device__ __forceinline__ int subblock_scan_index(int order, int bw, int bh, int sx, int sy) {
  if (order == 1) return sy * bw + sx;
  if (order == 2) return sx * bh + sy;
  const int d = sx + sy;
  int before = 0;
  for (int t = 0; t < d; t++) before += min(t, bw - 1) - max(0, t - (bh - 1)) + 1;
  return before + sx - max(0, d - (bh - 1));
}

// RdoQuant::CoeffSignHideFast for ONE 4x4 sub-block (rdo_quant.cc:461-567).  `in`, `delta`,
// `out` point at the sub-block's top-left coefficient.  is_last: this is the last sub-block
// in scan order that holds a non-zero level (search starts at its last non-zero position).
__device__ __forceinline__ void sign_hide_subblock(int scan_order, bool is_last, const int16_t *in, int is,
                                                   const int16_t *delta, int dstride, int16_t *out, int os) {
  const uint8_t *scan = c_scan4x4[scan_order];
#define XVCB_AT(buf, stride, i) (buf)[(scan[i] >> 2) * (stride) + (scan[i] & 3)]
  int last = -1, first = 16, sum = 0;
  for (int i = 0; i < 16; i++) {
    const int c = XVCB_AT(out, os, i);
    if (c) { first = min(first, i); last = max(last, i); sum += c; }
  }
  if (last - first <= 3) return;          // kSignHidingThreshold
  const int sign = XVCB_AT(out, os, first) > 0 ? 0 : 1;
  if (sign == (sum & 1)) return;
  // costs and changes are Coeff (int16) in the reference, including the wrap of -delta
  int16_t cur_cost = 32767, cur_change = 0, min_cost = 32767, min_change = 0;
  int min_index = -1;
  for (int i = is_last ? last : 15; i >= 0; i--) {
    const int16_t lev = XVCB_AT(out, os, i), dl = XVCB_AT(delta, dstride, i);
    if (lev != 0) {
      if (dl > 0) { cur_cost = (int16_t)-dl; cur_change = 1; }
      else if (i == first && abs((int)lev) == 1) { cur_cost = 32767; }
      else { cur_cost = dl; cur_change = -1; }
    } else if (i < first && (XVCB_AT(in, is, i) >= 0 ? 0 : 1) != sign) {
      cur_cost = 32767;
    } else {
      cur_cost = (int16_t)-dl; cur_change = 1;
    }
    if (cur_cost < min_cost) { min_cost = cur_cost; min_change = cur_change; min_index = i; }
  }
  if (min_index < 0) return;   // cannot happen: the last non-zero level always has a finite cost
  int16_t *p = &XVCB_AT(out, os, min_index);
  if (*p == -32768 || *p == 32767) min_change = -1;
  *p = (int16_t)(*p + (XVCB_AT(in, is, min_index) >= 0 ? min_change : -min_change));
#undef XVCB_AT
}

// ================================================================ single block, any type

// DST 4x4 (FwdPartialDst4 / InvPartialDst4, transform.cc:997-1017, 217-242) as the matrix
// its butterflies factor.
__constant__ int c_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

__device__ __forceinline__ const int16_t *mat_ptr(int type, int n) {
  if (type == XVCB200_TX_DEFAULT) type = XVCB200_TX_DCT2;
  return g_mats + c_mat_off[type][ilog2i(n)];
}

// forward stage: out[k*os + line] = (sum_j m[k][j] * in[line*is + j] + add) >> shift (int16 wrap)
__device__ void fwd_stage_generic(const int16_t *m, bool dst4, int n, int shift, int lines, bool zero_out,
                                  const int16_t *in, int is, int16_t *out, int os, int tid, int nthreads) {
  const int add = 1 << (shift - 1);
  const int tx_lines = (zero_out && lines > 32) ? 32 : lines;
  const int out_rows = n > 32 ? 32 : n;
  for (int e = tid; e < n * lines; e += nthreads) {
    const int k = e / lines, y = e - k * lines;
    int v = 0;
    if (k < out_rows && y < tx_lines) {
      int sum = 0;
      for (int j = 0; j < n; j++) sum += (dst4 ? c_dst4[k][j] : (int)m[k * n + j]) * in[y * is + j];
      v = (sum + add) >> shift;
    }
    out[k * os + y] = (int16_t)v;
  }
}
// inverse stage: out[line*os + j] = clip16((sum_k m[k][j] * in[k*is + line] + add) >> shift)
__device__ void inv_stage_generic(const int16_t *m, bool dst4, int n, int shift, int lines, bool zero_out,
                                  const int16_t *in, int is, int16_t *out, int os, int tid, int nthreads) {
  const int add = 1 << (shift - 1);
  const int tx_lines = (zero_out && lines > 32) ? 32 : lines;
  const int in_rows = n > 32 ? 32 : n;
  for (int e = tid; e < n * lines; e += nthreads) {
    const int y = e / n, j = e - y * n;
    int v = 0;
    if (y < tx_lines) {
      int sum = 0;
      for (int k = 0; k < in_rows; k++) sum += (dst4 ? c_dst4[k][j] : (int)m[k * n + j]) * in[k * is + y];
      v = clip16((sum + add) >> shift);
    }
    out[y * os + j] = (int16_t)v;
  }
}

__global__ void __launch_bounds__(256) block_transform_kernel(int forward, int w, int h, int bitdepth, int tx_hor,
                                                              int tx_ver, int dst4x4, int dc_only, int skip,
                                                              const int16_t *in, int is, int16_t *out, int os) {
  __shared__ int16_t tmp[64 * 64];
  const int tid = threadIdx.x;
  const int lw = ilog2i(w), lh = ilog2i(h);
  if (skip) {   // TransformSkip, transform.cc:963-995 (forward) / 184-215 (inverse)
    const int odd = (lw + lh) & 1, scale = odd ? 181 : 1, ts = transform_shift(lw, lh, bitdepth);
    for (int e = tid; e < w * h; e += 256) {
      const int y = e / w, x = e - y * w;
      int v = in[y * is + x] * scale;
      if (forward) {
        const int shift = ts + (odd ? -8 : 0);
        v = shift > 0 ? v * (1 << shift) : (v + (1 << (-shift - 1))) >> -shift;
      } else {
        const int shift = ts + (odd ? 7 : 0);
        v = shift > 0 ? (v + (1 << (shift - 1))) >> shift : (int)((unsigned)v << -shift);
      }
      out[y * os + x] = (int16_t)v;
    }
    return;
  }
  const bool dst = dst4x4 && w == 4 && h == 4;
  const int hp = dst ? 0 : 2;   // DST 4x4 has no high-precision variant (transform.cc:220, 1001)
  if (forward) {
    fwd_stage_generic(mat_ptr(tx_hor, w), dst, w, lw + bitdepth - 9 + hp, h, false, in, is, tmp, 64, tid, 256);
    __syncthreads();
    fwd_stage_generic(mat_ptr(tx_ver, h), dst, h, lh + 6 + hp, w, true, tmp, 64, out, os, tid, 256);
  } else {
    if (!dst && dc_only && tx_hor <= XVCB200_TX_DCT2 && tx_ver <= XVCB200_TX_DCT2) {  // InvDct2Dc, :279-291
      const int shift = 14 - bitdepth;
      const int16_t c = (int16_t)((((in[0] + 1) >> 1) + (1 << (shift - 1))) >> shift);
      for (int e = tid; e < w * h; e += 256) out[(e / w) * os + e % w] = c;
      return;
    }
    inv_stage_generic(mat_ptr(tx_ver, h), dst, h, 7 + hp, w, true, in, is, tmp, 64, tid, 256);
    __syncthreads();
    inv_stage_generic(mat_ptr(tx_hor, w), dst, w, 20 - bitdepth + hp, h, false, tmp, 64, out, os, tid, 256);
  }
}

cudaError_t launch_block_transform(cudaStream_t s, int forward, int w, int h, int bitdepth, int tx_hor, int tx_ver,
                                   int dst4x4, int dc_only, int skip, const int16_t *in, int is, int16_t *out, int os) {
  g_launch_count++;
  block_transform_kernel<<<1, 256, 0, s>>>(forward, w, h, bitdepth, tx_hor, tx_ver, dst4x4, dc_only, skip, in, is, out, os);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) block_quant_kernel(int w, int h, int bitdepth, int qp_bd, int intra_pic,
                                                          int sign_hiding, int scan, const int16_t *in, int is,
                                                          int16_t *out, int os, int *nnz_out) {
  __shared__ int16_t delta[64 * 64];
  __shared__ int nnz, last_sb;
  const int tid = threadIdx.x;
  const int lw = ilog2i(w), lh = ilog2i(h);
  if (tid == 0) { nnz = 0; last_sb = -1; }
  __syncthreads();
  const QuantParams q = quant_params(lw, lh, bitdepth, qp_bd, intra_pic);
  int mine = 0;
  for (int e = tid; e < w * h; e += 256) {
    const int y = e / w, x = e - y * w;
    int16_t lev, dl;
    quant_one(in[y * is + x], q, lev, dl);
    out[y * os + x] = lev;
    delta[y * 64 + x] = dl;
    mine += lev != 0;
  }
  atomicAdd(&nnz, mine);
  __syncthreads();
  if (sign_hiding && nnz > 1 && w >= 4 && h >= 4) {
    const int bw = w >> 2, bh = h >> 2;
    for (int sb = tid; sb < bw * bh; sb += 256) {
      const int sx = sb % bw, sy = sb / bw;
      bool any = false;
      for (int i = 0; i < 16; i++) any |= out[(sy * 4 + (i >> 2)) * os + sx * 4 + (i & 3)] != 0;
      if (any) atomicMax(&last_sb, subblock_scan_index(scan, bw, bh, sx, sy));
    }
    __syncthreads();
    for (int sb = tid; sb < bw * bh; sb += 256) {
      const int sx = sb % bw, sy = sb / bw;
      sign_hide_subblock(scan, subblock_scan_index(scan, bw, bh, sx, sy) == last_sb, in + sy * 4 * is + sx * 4, is,
                         delta + sy * 4 * 64 + sx * 4, 64, out + sy * 4 * os + sx * 4, os);
    }
    __syncthreads();
    if (tid == 0) nnz = 0;
    __syncthreads();
    mine = 0;
    for (int e = tid; e < w * h; e += 256) mine += out[(e / w) * os + e % w] != 0;
    atomicAdd(&nnz, mine);
    __syncthreads();
  }
  if (tid == 0) *nnz_out = nnz;
}

cudaError_t launch_block_quant(cudaStream_t s, int w, int h, int bitdepth, int qp_bd, int intra_pic, int sign_hiding,
                               int scan, const int16_t *in, int is, int16_t *out, int os, int *d_nnz) {
  g_launch_count++;
  block_quant_kernel<<<1, 256, 0, s>>>(w, h, bitdepth, qp_bd, intra_pic, sign_hiding, scan, in, is, out, os, d_nnz);
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) block_dequant_kernel(int w, int h, int bitdepth, int qp_bd, const int16_t *in,
                                                            int is, int16_t *out, int os) {
  const DequantParams d = dequant_params(ilog2i(w), ilog2i(h), bitdepth, qp_bd);
  for (int e = threadIdx.x; e < w * h; e += 256) {
    const int y = e / w, x = e - y * w;
    out[y * os + x] = dequant_one(in[y * is + x], d);
  }
}

cudaError_t launch_block_dequant(cudaStream_t s, int w, int h, int bitdepth, int qp_bd, const int16_t *in, int is,
                                 int16_t *out, int os) {
  g_launch_count++;
  block_dequant_kernel<<<1, 256, 0, s>>>(w, h, bitdepth, qp_bd, in, is, out, os);
  return cudaGetLastError();
}

// ================================================================ fused T/Q/recon, DCT-2
//
// A transform stage is a set of LINES (rows or columns of the block); a line of a long transform
// is shared by SPLIT threads, each producing a contiguous run of its outputs.  A thread loads its
// line of N int16 into registers once and accumulates all of its outputs together: for every
// input sample the matrix coefficients of its outputs are contiguous in the int32 tables above,
// fetched with 16-byte loads whose address is uniform across the lines of a warp (one L1
// transaction), so the inner loop is IMADs on independent accumulators.  Small blocks share a
// CTA.  Shared-memory rows are padded by one 32-bit word so that the per-thread line reads
// (stride = pitch) are bank-conflict free.

template <int N> struct Dct2 { static constexpr int kOff = Dct2<N / 2>::kOff + (N / 2) * (N / 2); };
template <> struct Dct2<2> { static constexpr int kOff = 0; };

// threads sharing one line of an N-point transform
template <int N> struct Split { static constexpr int kS = N >= 64 ? 4 : (N >= 32 ? 2 : 1); };

// forward line: in = N contiguous int16 (4-byte aligned); out[k*os] for k < N.  This thread: outputs
// [split * KPT, (split + 1) * KPT) of the R = min(N, 32) computed ones (+ its share of the zeros).
template <int N>
__device__ __forceinline__ void fwd_line(const int16_t *in, int split, int16_t *out, int os, int shift, bool zero_line, int mat_off) {
  constexpr int R = N > 32 ? 32 : N, S = Split<N>::kS, KPT = R / S;
  const int k0 = split * KPT;
  if (!zero_line) {
    int v[N];
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const uint32_t p = *reinterpret_cast<const uint32_t *>(in + j);
      v[j] = (int)(int16_t)(p & 0xffff);
      v[j + 1] = (int)(int16_t)(p >> 16);
    }
    int acc[KPT];
#pragma unroll
    for (int i = 0; i < KPT; i++) acc[i] = 1 << (shift - 1);
    const int *mt = g_tq_fwdT + mat_off + k0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      if (KPT >= 4) {
#pragma unroll
        for (int i = 0; i < KPT; i += 4) {
          const int4 c = __ldg(reinterpret_cast<const int4 *>(mt + j * R + i));
          acc[i] += c.x * v[j]; acc[i + 1] += c.y * v[j]; acc[i + 2] += c.z * v[j]; acc[i + 3] += c.w * v[j];
        }
      } else {
#pragma unroll
        for (int i = 0; i < KPT; i++) acc[i] += __ldg(mt + j * R + i) * v[j];
      }
    }
#pragma unroll
    for (int i = 0; i < KPT; i++) out[(k0 + i) * os] = (int16_t)(acc[i] >> shift);
  } else {
#pragma unroll 4
    for (int i = 0; i < KPT; i++) out[(k0 + i) * os] = 0;
  }
  if (N > R) {      // coefficients 32..63 of a 64-point transform are zero
#pragma unroll 4
    for (int i = 0; i < KPT; i++) out[(R + k0 + i) * os] = 0;
  }
}

// inverse line: in[k*is] for k < min(N,32); out = N contiguous int16, clipped.  This thread:
// outputs [split * JPT, (split + 1) * JPT).
template <int N>
__device__ __forceinline__ void inv_line(const int16_t *in, int is, int split, int16_t *out, int shift, bool zero_line, int mat_off) {
  constexpr int R = N > 32 ? 32 : N, S = Split<N>::kS, JPT = N / S;
  const int j0 = split * JPT;
  if (zero_line) {
#pragma unroll 4
    for (int i = 0; i < JPT; i++) out[j0 + i] = 0;
    return;
  }
  int v[R];
#pragma unroll
  for (int k = 0; k < R; k++) v[k] = in[k * is];
  int acc[JPT];
#pragma unroll
  for (int i = 0; i < JPT; i++) acc[i] = 1 << (shift - 1);
  const int *m = g_tq_inv + mat_off + j0;
#pragma unroll
  for (int k = 0; k < R; k++) {
    if (JPT >= 4) {
#pragma unroll
      for (int i = 0; i < JPT; i += 4) {
        const int4 c = __ldg(reinterpret_cast<const int4 *>(m + k * N + i));
        acc[i] += c.x * v[k]; acc[i + 1] += c.y * v[k]; acc[i + 2] += c.z * v[k]; acc[i + 3] += c.w * v[k];
      }
    } else {
#pragma unroll
      for (int i = 0; i < JPT; i++) acc[i] += __ldg(m + k * N + i) * v[k];
    }
  }
#pragma unroll
  for (int i = 0; i < JPT; i++) out[j0 + i] = (int16_t)clip16(acc[i] >> shift);
}

struct TqPlanes {
  PlaneView orig[3], pred[3], rec[3];
  int16_t *lev[3];
  int lev_pitch[3];
};

template <int LW, int LH> struct TqShape {
  static constexpr int W = 1 << LW, H = 1 << LH;
  static constexpr int T1 = H * Split<W>::kS;            // threads of a stage along the rows (N = W, H lines)
  static constexpr int T2 = W * Split<H>::kS;            // ... along the columns (N = H, W lines)
  static constexpr int TT = T1 > T2 ? T1 : T2;           // threads per transform unit
  static constexpr int NT = TT < 32 ? 32 : TT;           // threads per CTA
  static constexpr int TPB = NT / TT;                    // transform units per CTA
};

template <int LW, int LH>
__global__ void __launch_bounds__(TqShape<LW, LH>::NT)
tq_kernel(xvcb200_cu *__restrict__ cus, const int *__restrict__ tu_list, int n_tu, TqParams prm,
          const __grid_constant__ TqPlanes pl, xvcb200_tu_result *__restrict__ results) {
  using Sh = TqShape<LW, LH>;
  constexpr int W = Sh::W, H = Sh::H, TT = Sh::TT, TPB = Sh::TPB, T1 = Sh::T1, T2 = Sh::T2;
  constexpr int PA = W + 2, PB = H + 2;          // row pitches (int16) of the [H][W] and [W][H] buffers
  constexpr int SZB = W * PB > H * PA ? W * PB : H * PA;
  __shared__ __align__(16) int16_t s_a[TPB][H * PA];   // residual -> delta -> reconstructed residual
  __shared__ __align__(16) int16_t s_b[TPB][SZB];      // transposed intermediate of both transforms; the levels in between
  __shared__ __align__(16) int16_t s_c[TPB][H * PA];   // coefficients / dequantised coefficients
  __shared__ int s_nnz[TPB], s_last[TPB];
  __shared__ unsigned long long s_ssd[TPB];

  const int tid = threadIdx.x;
  const int t = tid / TT, ti = tid % TT;            // TU slot in this CTA, thread in the TU
  const int tu = blockIdx.x * TPB + t;
  const bool valid = tu < n_tu;
  const int id = valid ? tu_list[tu] : 0;
  const int ci = id / 3, comp = id - ci * 3;
  const xvcb200_cu cu = cus[ci];
  const int cs = comp ? 1 : 0;
  const int x0 = cu.x >> cs, y0 = cu.y >> cs;
  const int bd = prm.bitdepth;
  int16_t *A = s_a[t], *B = s_b[t], *C = s_c[t], *D = s_b[t];   // D (levels, [H][PA]) lives in B between the transforms

  // transform types, transform skip and coefficient scan of this unit (xvcb200_set_tu_modes; none: DCT-2 -- or the
  // 4 x 4 DST of an intra CU's luma block, transform.cc:87-89, 873-875 --, no skip, diagonal scan)
  int ty_ver = XVCB200_TX_DEFAULT, ty_hor = XVCB200_TX_DEFAULT, scan = 0;
  bool tskip = false;
  if (prm.modes) {
    const xvcb200_tu_mode md = prm.modes[ci];
    if (comp == 0) { ty_ver = md.tx_ver; ty_hor = md.tx_hor; }
    tskip = W * H <= 16 && ((md.tskip >> comp) & 1);
    scan = md.scan[comp];
  }
  const bool dst4 = W == 4 && H == 4 && comp == 0 && (cu.flags & XVCB200_CU_INTRA) && ty_ver == XVCB200_TX_DEFAULT && ty_hor == XVCB200_TX_DEFAULT;
  const int hp = dst4 ? 0 : 2;            // the 4 x 4 DST has no high-precision variant (transform.cc:220, 1001)
  const int off_hor = dst4 ? kTqDst4Off : c_mat_off[ty_hor == XVCB200_TX_DEFAULT ? XVCB200_TX_DCT2 : ty_hor][LW];
  const int off_ver = dst4 ? kTqDst4Off : c_mat_off[ty_ver == XVCB200_TX_DEFAULT ? XVCB200_TX_DCT2 : ty_ver][LH];

  int qp_raw = cu.qp;
  if (comp) qp_raw = chroma_qp_raw(cu.qp, comp == 1 ? prm.off_u : prm.off_v, prm.table, c_chroma_scale);
  const int qp_bd = max(0, qp_raw + 6 * (bd - 8));

  const PlaneView po = pl.orig[comp], pp = pl.pred[comp], pr = pl.rec[comp];
  const Sample *orig = po.base + y0 * po.pitch + x0;
  const Sample *pred = pp.base + y0 * pp.pitch + x0;
  Sample *rec = pr.base + y0 * pr.pitch + x0;
  int16_t *lev = pl.lev[comp] + y0 * pl.lev_pitch[comp] + x0;

  if (ti == 0) { s_nnz[t] = 0; s_last[t] = -1; s_ssd[t] = 0; }
  __syncthreads();

  bool cbf;
  if (!prm.decode_only) {
    // residual = orig - pred (ResidualBuffer::Subtract, sample_buffer.h:130-145)
    if (valid)
      for (int e = ti; e < W * H; e += TT) {
        const int y = e / W, x = e % W;
        A[y * PA + x] = (int16_t)((int)orig[y * po.pitch + x] - (int)pred[y * pp.pitch + x]);
      }
    __syncthreads();
    // (barriers are unconditional: the units sharing a CTA may differ in their modes)
    if (W * H <= 16 && tskip) {
      // ForwardTransform::TransformSkip (transform.cc:963-995)
      const int odd = (LW + LH) & 1, sh = transform_shift(LW, LH, bd) + (odd ? -8 : 0), scale = odd ? 181 : 1;
      if (valid)
        for (int e = ti; e < W * H; e += TT) {
          const int y = e / W, x = e % W, v = (int)A[y * PA + x] * scale;
          C[y * PA + x] = (int16_t)(sh > 0 ? v * (1 << sh) : (v + (1 << (-sh - 1))) >> -sh);
        }
    } else if (valid && ti < T1) {
      // forward: rows (N = W, lines = H) into B[k][y]; columns (N = H, lines = W) into C[x][y']
      fwd_line<W>(A + (ti % H) * PA, ti / H, B + (ti % H), PB, LW + bd - 9 + hp, false, off_hor);
    }
    __syncthreads();
    if (!tskip && valid && ti < T2) fwd_line<H>(B + (ti % W) * PB, ti / W, C + (ti % W), PA, LH + 6 + hp, (ti % W) >= 32, off_ver);
    __syncthreads();
    // QuantFast (rdo_quant.cc:156-201)
    const QuantParams q = quant_params(LW, LH, bd, qp_bd, prm.intra_picture);
    int mine = 0;
    if (valid)
      for (int e = ti; e < W * H; e += TT) {
        const int y = e / W, x = e % W;
        int16_t lv, dl;
        quant_one(C[y * PA + x], q, lv, dl);
        D[y * PA + x] = lv;
        A[y * PA + x] = dl;
        mine += lv != 0;
      }
    if (mine) atomicAdd(&s_nnz[t], mine);
    __syncthreads();
    if (W >= 4 && H >= 4) {   // sign hiding in the unit's scan order (TransformHelper::DetermineScanOrder, transform.cc:1614-1636)
      constexpr int BW = W >= 4 ? W / 4 : 1, BH = H >= 4 ? H / 4 : 1;
      const bool run = valid && s_nnz[t] > 1;
      if (run)
        for (int sb = ti; sb < BW * BH; sb += TT) {
          const int sx = sb % BW, sy = sb / BW;
          bool any = false;
#pragma unroll
          for (int i = 0; i < 16; i++) any |= D[(sy * 4 + (i >> 2)) * PA + sx * 4 + (i & 3)] != 0;
          if (any) atomicMax(&s_last[t], subblock_scan_index(scan, BW, BH, sx, sy));
        }
      __syncthreads();
      if (run)
        for (int sb = ti; sb < BW * BH; sb += TT) {
          const int sx = sb % BW, sy = sb / BW;
          sign_hide_subblock(scan, subblock_scan_index(scan, BW, BH, sx, sy) == s_last[t], C + sy * 4 * PA + sx * 4, PA,
                             A + sy * 4 * PA + sx * 4, PA, D + sy * 4 * PA + sx * 4, PA);
        }
      __syncthreads();
      if (run && ti == 0) s_nnz[t] = 0;
      __syncthreads();
      if (run) {
        mine = 0;
        for (int e = ti; e < W * H; e += TT) mine += D[(e / W) * PA + e % W] != 0;
        if (mine) atomicAdd(&s_nnz[t], mine);
      }
      __syncthreads();
    }
    cbf = s_nnz[t] != 0;
    if (valid)   // levels out (zero block when cbf == 0)
      for (int e = ti; e < W * H; e += TT) {
        const int y = e / W, x = e % W;
        lev[y * pl.lev_pitch[comp] + x] = cbf ? D[y * PA + x] : (int16_t)0;
      }
  } else {
    const int bit = comp == 0 ? XVCB200_CU_CBF_Y : (comp == 1 ? XVCB200_CU_CBF_U : XVCB200_CU_CBF_V);
    cbf = (cu.flags & bit) != 0;
    if (valid && cbf)
      for (int e = ti; e < W * H; e += TT) {
        const int y = e / W, x = e % W;
        D[y * PA + x] = lev[y * pl.lev_pitch[comp] + x];
      }
    __syncthreads();
  }

  // dequant (Quantize::Inverse) -> C
  const DequantParams dq = dequant_params(LW, LH, bd, qp_bd);
  if (valid && cbf)
    for (int e = ti; e < W * H; e += TT) {
      const int y = e / W, x = e % W;
      C[y * PA + x] = dequant_one(D[y * PA + x], dq);
    }
  __syncthreads();
  if (W * H <= 16 && tskip) {
    // InverseTransform::TransformSkip (transform.cc:184-215)
    const int odd = (LW + LH) & 1, sh = transform_shift(LW, LH, bd) + (odd ? 7 : 0), scale = odd ? 181 : 1;
    if (valid && cbf)
      for (int e = ti; e < W * H; e += TT) {
        const int y = e / W, x = e % W, v = (int)C[y * PA + x] * scale;
        A[y * PA + x] = (int16_t)(sh > 0 ? (v + (1 << (sh - 1))) >> sh : (int)((unsigned)v << -sh));
      }
  } else if (valid && cbf && ti < T2) {
    // inverse: columns (N = H, lines = W) into B[x][j]; rows (N = W, lines = H) into A[y][x]
    inv_line<H>(C + (ti % W), PA, ti / W, B + (ti % W) * PB, 7 + hp, (ti % W) >= 32, off_ver);
  }
  __syncthreads();
  if (!tskip && valid && cbf && ti < T1) inv_line<W>(B + (ti % H), PB, ti / H, A + (ti % H) * PA, 20 - bd + hp, false, off_hor);
  __syncthreads();

  // reconstruct (SampleBuffer::AddClip, sample_buffer.h:72-87; cbf == 0: copy of the prediction)
  const int maxv = (1 << bd) - 1;
  unsigned long long ssd = 0;
  if (valid)
    for (int e = ti; e < W * H; e += TT) {
      const int y = e / W, x = e % W;
      const int p = pred[y * pp.pitch + x];
      const int r = cbf ? clip3i(p + A[y * PA + x], 0, maxv) : p;
      rec[y * pr.pitch + x] = (Sample)r;
      if (!prm.decode_only) {
        const int d = (int)orig[y * po.pitch + x] - r;
        ssd += (unsigned long long)(d * d);
      }
    }
  if (!prm.decode_only) {
    if (ssd) atomicAdd(&s_ssd[t], ssd);
    __syncthreads();
    if (valid && ti == 0) {
      if (results) {
        results[id].ssd = (uint32_t)(s_ssd[t] >> (2 * (bd - 8)));
        results[id].num_non_zero = s_nnz[t];
      }
      // cbf flag of this component back into the CU (three TUs of a CU may run concurrently)
      const unsigned bit = comp == 0 ? XVCB200_CU_CBF_Y : (comp == 1 ? XVCB200_CU_CBF_U : XVCB200_CU_CBF_V);
      unsigned *word = reinterpret_cast<unsigned *>(&cus[ci]) + 1;   // bytes 4..7: w, h, depth, flags
      if (cbf) atomicOr(word, bit << 24); else atomicAnd(word, ~(bit << 24));
    }
  }
}

template <int LW, int LH>
static void launch_tq_class(cudaStream_t s, xvcb200_cu *d_cus, const int *d_list, int count, const TqParams &p,
                            const TqPlanes &pl, xvcb200_tu_result *d_res) {
  constexpr int NT = TqShape<LW, LH>::NT, TPB = TqShape<LW, LH>::TPB;
  g_launch_count++;
  tq_kernel<LW, LH><<<(count + TPB - 1) / TPB, NT, 0, s>>>(d_cus, d_list, count, p, pl, d_res);
}

// h_class_count / h_class_offset: per shape class [lw][lh] (1..6) counts and offsets into
// d_tu_list, prepared on the host when the CU array is set.
// The shape classes are independent of each other: their launches are spread round-robin over
// `n_side` side streams that fork from / join into `s` through events, so the many small grids
// overlap on the GPU instead of running back to back.
cudaError_t launch_tq_reconstruct_classes(cudaStream_t s, xvcb200_cu *d_cus, const int *d_tu_list,
                                          const int class_count[7][7], const int class_offset[7][7],
                                          const TqParams &p, Pic3 orig, Pic3 pred, Pic3 rec, int16_t *const lev[3],
                                          const int lev_pitch[3], xvcb200_tu_result *d_res, cudaStream_t *side,
                                          cudaEvent_t *side_ev, int n_side, cudaEvent_t fork_ev) {
  cudaError_t e = ensure_dct2_constant();
  if (e != cudaSuccess) return e;
  if (n_side > 0) {
    cudaEventRecord(fork_ev, s);
    for (int i = 0; i < n_side; i++) cudaStreamWaitEvent(side[i], fork_ev, 0);
  }
  int rr = 0;
  TqPlanes pl;
  for (int c = 0; c < 3; c++) {
    pl.orig[c] = orig.p[c]; pl.pred[c] = pred.p[c]; pl.rec[c] = rec.p[c];
    pl.lev[c] = lev[c]; pl.lev_pitch[c] = lev_pitch[c];
  }
#define XVCB_TQ(LW, LH)                                                                                              \
  if (class_count[LW][LH] > 0)                                                                                     \
    launch_tq_class<LW, LH>(n_side > 0 ? side[rr++ % n_side] : s, d_cus, d_tu_list + class_offset[LW][LH],          \
                            class_count[LW][LH], p, pl, d_res);
#define XVCB_TQ_ROW(LW) XVCB_TQ(LW, 1) XVCB_TQ(LW, 2) XVCB_TQ(LW, 3) XVCB_TQ(LW, 4) XVCB_TQ(LW, 5) XVCB_TQ(LW, 6)
  // largest transforms first: they have the longest critical path
  XVCB_TQ_ROW(6) XVCB_TQ_ROW(5) XVCB_TQ_ROW(4) XVCB_TQ_ROW(3) XVCB_TQ_ROW(2) XVCB_TQ_ROW(1)
#undef XVCB_TQ_ROW
#undef XVCB_TQ
  for (int i = 0; i < n_side; i++) {
    cudaEventRecord(side_ev[i], side[i]);
    cudaStreamWaitEvent(s, side_ev[i], 0);
  }
  return cudaGetLastError();
}

}  // namespace xvcb
