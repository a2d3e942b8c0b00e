// Launch planning and channel-count dispatch for the fused policy kernel.
#include "chb_internal.h"

namespace chb {

#define CHB_DECL(C)                                                                         \
  cudaError_t launch_policy_c##C(const KParams&, const LaunchInfo&, cudaStream_t);          \
  cudaError_t configure_c##C(size_t);                                                       \
  size_t smem_overhead_c##C();
CHB_DECL(1) CHB_DECL(2) CHB_DECL(3) CHB_DECL(4)
#undef CHB_DECL

size_t smem_overhead(int C) {
  switch (C) {
    case 1: return smem_overhead_c1();
    case 2: return smem_overhead_c2();
    case 3: return smem_overhead_c3();
    default: return smem_overhead_c4();
  }
}

LaunchInfo plan_launch(int B, int H, int W, int C, int num_sms, size_t smem_optin) {
  LaunchInfo li;
  const size_t img_pad = ((size_t)H * W * C + 127) / 128 * 128;
  const size_t over = smem_overhead(C);
  li.image_in_smem = img_pad + over <= smem_optin;
  li.smem = li.image_in_smem ? img_pad + over : over;
  li.block = 512;  // NT in chb_kernels.cuh (up to 128 registers per thread)
  // persistent CTAs, one per SM (the shared-memory carve-out allows no more), never more than images
  li.grid = B < num_sms ? B : num_sms;
  if (li.grid < 1) li.grid = 1;
  return li;
}

cudaError_t configure_kernels(size_t smem_optin) {
  cudaError_t e;
  if ((e = configure_c1(smem_optin)) != cudaSuccess) return e;
  if ((e = configure_c2(smem_optin)) != cudaSuccess) return e;
  if ((e = configure_c3(smem_optin)) != cudaSuccess) return e;
  return configure_c4(smem_optin);
}

cudaError_t launch_policy(const KParams& p, int C, const LaunchInfo& li, cudaStream_t stream) {
  switch (C) {
    case 1: return launch_policy_c1(p, li, stream);
    case 2: return launch_policy_c2(p, li, stream);
    case 3: return launch_policy_c3(p, li, stream);
    case 4: return launch_policy_c4(p, li, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace chb
