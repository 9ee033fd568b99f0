// Development probe: which tensor-map / box / coordinate combinations does cp.async.bulk.tensor.2d accept
// for the padded-plane layout of xvc_b200?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Maps { CUtensorMap m[4]; };

__global__ void probe(const __grid_constant__ Maps maps, int which, int box_w, int box_h, int x, int y, int nbox, uint16_t *out) {
  __shared__ __align__(128) uint16_t s[72 * 72];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbox * box_w * box_h * 2) : "memory");
    for (int k = 0; k < nbox; k++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(&s[k * box_h * box_w])), "l"(&maps.m[which]), "r"(x), "r"(y + k * box_h), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < nbox * box_h * box_w; i += blockDim.x) out[i] = s[i];
}

int main() {
  typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiled encode = (EncodeTiled)fn;
  const int pitch = 512, rows = 264;
  std::vector<uint16_t> h((size_t)pitch * rows);
  for (int y = 0; y < rows; y++) for (int x = 0; x < pitch; x++) h[(size_t)y * pitch + x] = (uint16_t)((y << 9) ^ x) & 0x3ff;
  uint16_t *d, *out;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&out, 72 * 72 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  const int boxes[4][2] = {{24, 8}, {72, 8}, {32, 8}, {64, 8}};
  const CUtensorMapL2promotion proms[2] = {CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE};
  for (int pr = 0; pr < 2; pr++) {
    Maps maps; memset(&maps, 0, sizeof(maps));
    for (int k = 0; k < 4; k++) {
      cuuint64_t dims[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
      cuuint32_t box[2] = {(cuuint32_t)boxes[k][0], (cuuint32_t)boxes[k][1]}; cuuint32_t es[2] = {1, 1};
      CUresult r = encode(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, proms[pr], CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("prom %d encode box %dx%d -> %d\n", pr, boxes[k][0], boxes[k][1], (int)r);
    }
    for (int k = 0; k < 4; k++)
      for (int x : {200, 201, 204, 207, 208, 3, 500})
        for (int y : {100, 101, 260}) {
          for (int nbox : {1, 2, 3}) {
          cudaMemset(out, 0xff, 72 * 72 * 2);
          probe<<<1, 128>>>(maps, k, boxes[k][0], boxes[k][1], x, y, nbox, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("prom %d box %d x %d y %d nbox %d: %s\n", pr, boxes[k][0], x, y, nbox, cudaGetErrorString(e)); return 1; }
          std::vector<uint16_t> o((size_t)nbox * boxes[k][0] * boxes[k][1]);
          cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int r = 0; r < nbox * boxes[k][1]; r++) for (int c = 0; c < boxes[k][0]; c++) {
            const int gx = x + c, gy = y + r;
            const uint16_t want = (gx < pitch && gy < rows) ? h[(size_t)gy * pitch + gx] : 0;
            bad += o[(size_t)r * boxes[k][0] + c] != want;
          }
          if (bad) printf("prom %d box %d x %d y %d nbox %d: %d mismatches\n", pr, boxes[k][0], x, y, nbox, bad);
          }
        }
    printf("prom %d done\n", pr);
  }
  printf("all ok\n");
  return 0;
}
