// Single-block distortion metrics behind the table-shaped C ABI
// (SampleMetric::SimdFunc entries and SampleMetric::Compare, sample_metric.cc:171-314).
// These serve drop-in parity; the throughput path computes the same metrics inside the
// batched ME / TQ kernels (me.cu, transform.cu).
#include "xvcb_satd.cuh"

namespace xvcb {

std::atomic<uint64_t> g_launch_count{0};

template <typename AT, typename BT>
__global__ void __launch_bounds__(128) block_metric_kernel(int metric, int bitdepth, int w, int h, const AT *a, int sa,
                                                           const BT *b, int sb, unsigned long long *out) {
  const int tid = threadIdx.x;
  unsigned long long acc = 0;
  if (metric == XVCB200_METRIC_SATD) {
    auto diff = [&](int x, int y) { return (int)a[y * sa + x] - (int)b[y * sb + x]; };
    acc = satd_block_partial(diff, w, h, tid, 128);
  } else {
    const bool ssd = metric == XVCB200_METRIC_SSD || metric == -2;
    const bool fast = metric == XVCB200_METRIC_SAD_FAST;
    const int rows = fast ? h / 2 : h, rs = fast ? 2 : 1;   // every second row (sample_metric.cc:194-199)
    for (int i = tid; i < rows * w; i += 128) {
      const int y = (i / w) * rs, x = i % w;
      const int d = (int)a[y * sa + x] - (int)b[y * sb + x];
      // `diff * diff` is an int product in the reference (sample_metric.cc:309) widened to uint64
      acc += ssd ? (unsigned long long)(long long)(d * d) : (unsigned long long)abs(d);
    }
  }
  acc = warp_sum(acc);
  __shared__ unsigned long long part[4];
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    unsigned long long t = part[0] + part[1] + part[2] + part[3];
    const int s = bitdepth - 8;
    switch (metric) {
      case XVCB200_METRIC_SSD: t >>= 2 * s; break;
      case XVCB200_METRIC_SATD: t >>= s; break;
      case XVCB200_METRIC_SAD: t >>= s; break;
      case XVCB200_METRIC_SAD_FAST: t = (t * 2) >> s; break;
      default: break;   // raw sad / ssd
    }
    *out = t;
  }
}

cudaError_t launch_block_metric(cudaStream_t s, int metric, int bitdepth, int w, int h, int a_short, int b_short,
                                const void *a, int sa, const void *b, int sb, unsigned long long *d_out) {
  g_launch_count++;
  if (a_short && b_short)
    block_metric_kernel<int16_t, int16_t><<<1, 128, 0, s>>>(metric, bitdepth, w, h, (const int16_t *)a, sa, (const int16_t *)b, sb, d_out);
  else if (a_short)
    block_metric_kernel<int16_t, uint16_t><<<1, 128, 0, s>>>(metric, bitdepth, w, h, (const int16_t *)a, sa, (const uint16_t *)b, sb, d_out);
  else
    block_metric_kernel<uint16_t, uint16_t><<<1, 128, 0, s>>>(metric, bitdepth, w, h, (const uint16_t *)a, sa, (const uint16_t *)b, sb, d_out);
  return cudaGetLastError();
}

}  // namespace xvcb
