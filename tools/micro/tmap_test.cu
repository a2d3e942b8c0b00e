// Micro test: 3-D tensor-map box load of a uint8 NHWC batch (uint32 elements).  nvcc -arch=sm_100a
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
struct alignas(64) TMap { unsigned char b[128]; };
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <bool PARAM>
__global__ void k(const __grid_constant__ TMap tm, const TMap* gtm, int c0, int c1, int c2, int bytes, uint8_t* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ unsigned long long bar;
  const uint32_t b = saddr(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const TMap* t = PARAM ? &tm : gtm;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(saddr(sm)), "l"(reinterpret_cast<uint64_t>(t)), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(b) : "memory");
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int B = 3, H = 224, W = 224, C = 3, row = W * C;
  const int bw = argc > 1 ? atoi(argv[1]) : 304, bh = argc > 2 ? atoi(argv[2]) : 96;
  const int mode = argc > 3 ? atoi(argv[3]) : 0;
  std::vector<uint8_t> h((size_t)B * H * row);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 9));
  uint8_t *d, *o; cudaMalloc(&d, h.size()); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  TMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[3] = {(cuuint64_t)row / 4, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t str[2] = {(cuuint64_t)row, (cuuint64_t)H * row};
  cuuint32_t box[3] = {(cuuint32_t)bw / 4, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)fn)((CUtensorMap*)&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d box %dx%d mode %d\n", (int)r, bw, bh, mode);
  TMap* gtm; cudaMalloc(&gtm, sizeof(TMap)); cudaMemcpy(gtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  const int bytes = bw * bh, c0 = 10, c1 = 100, c2 = 1;
  cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (mode == 0) k<true><<<1, 128, 65536>>>(tm, gtm, c0, c1, c2, bytes, o);
  else k<false><<<1, 128, 65536>>>(tm, gtm, c0, c1, c2, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> res(bytes); cudaMemcpy(res.data(), o, bytes, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r2 = 0; r2 < bh; ++r2) for (int x = 0; x < bw; ++x) {
    const int y = c1 + r2, xb = c0 * 4 + x;
    uint8_t want = (y < H && xb < row) ? h[((size_t)c2 * H + y) * row + xb] : 0;
    if (res[r2 * bw + x] != want) ++bad;
  }
  printf("mismatches %d of %d\n", bad, bytes);
  return 0;
}
