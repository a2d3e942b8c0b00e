// plan_kernel / optab_kernel instantiation and their launchers (see chb_kernels.cuh).
#define CHB_WITH_PLAN 1
#include "chb_kernels.cuh"

namespace chb {

cudaError_t launch_optab(const DevOp* ops, uint8_t* optab, int n_ops, cudaStream_t stream) {
  optab_kernel<<<n_ops, 256, 0, stream>>>(ops, optab);
  return cudaGetLastError();
}

// The pass kernels run with the maximum shared-memory carve-out; give the plan kernel the same one
// so that the SMs do not have to be reconfigured between the kernels of a call.
cudaError_t configure_plan() {
  cudaError_t e = cudaFuncSetAttribute(plan_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(optab_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_plan(const KParams& p, int C, cudaStream_t stream) {
  plan_kernel<<<p.B, PLAN_NT, 0, stream>>>(p, C);
  return cudaGetLastError();
}

}  // namespace chb
