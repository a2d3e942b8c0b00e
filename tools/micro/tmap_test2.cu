// Micro test 2: simplest possible 2-D tensor-map load, linked against libcuda directly.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int bytes, uint8_t* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t b = saddr(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(saddr(sm)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(c1), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(saddr(sm)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 2;
  const int esz = argc > 2 ? atoi(argv[2]) : 1;     // element size 1 or 4
  const int bw = argc > 3 ? atoi(argv[3]) : 64, bh = argc > 4 ? atoi(argv[4]) : 64;  // box bytes x rows
  cuInit(0);
  const int B = 3, H = 224, row = 672;
  std::vector<uint8_t> h((size_t)B * H * row);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[3] = {(cuuint64_t)row / esz, (cuuint64_t)(rank == 2 ? B * H : H), (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)row, (cuuint64_t)H * row};
  cuuint32_t box[3] = {(cuuint32_t)bw / esz, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, dims, str, box, es,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("rank %d esz %d box %dx%d encode rc=%d | ", rank, esz, bw, bh, (int)r);
  const uint64_t* w = (const uint64_t*)&tm;
  const int bytes = bw * bh, c0 = 16 / esz, c1 = 100, c2 = 1;
  cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (rank == 2) k<2><<<1, 128, 65536>>>(tm, c0, c1, 0, bytes, o); else k<3><<<1, 128, 65536>>>(tm, c0, c1, c2, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) { printf("desc %016lx %016lx %016lx %016lx\n", w[0], w[1], w[2], w[3]); return 1; }
  std::vector<uint8_t> res(bytes); cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int x = 0; x < bw; ++x) {
    const int y = c1 + r2, xb = c0 * esz + x;
    const size_t im = rank == 2 ? 0 : c2;
    uint8_t want = (xb < row) ? h[(im * H + y) * row + xb] : 0;
    if (res[r2 * bw + x] != want) ++bad;
  }
  printf("mismatches %d of %d\n", bad, bytes);
  return 0;
}
