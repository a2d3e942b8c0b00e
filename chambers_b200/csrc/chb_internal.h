// Internal (non-ABI) declarations shared by chb_api.cu (host) and chb_kernels.cu (device).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chambers_aug.h"

namespace chb {

// Device-side form of one op layer: everything the reference computes at construction or per call
// on the host (float32 conversions, wrapped thresholds, the 8 projective coefficients for both
// outcomes of _randomly_negate_value) is resolved here once per (policy, H, W, batch_total).
struct DevOp {
  int32_t kind;        // chb_op_kind, or -1 for "no op in this slot"
  int32_t interp;      // chb_interpolation
  int32_t fill_mode;   // chb_fill_mode
  int32_t thr24;       // coin threshold on a 24-bit grid; 1<<24 = always (oracle/philox.py)
  int32_t ip0, ip1;    // Posterize: shift (0..8) | Solarize: thr (0..256) | SolarizeAdd: add, thr
                       // CutOut: half mask size, constant (0..255) | Contrast: blend constant
  float factor;        // float32(factor) of the blend ops / Sharpness
  int32_t blend_mode;  // BLEND_*
  float coef[2][8];    // [negate][8] projective coefficients (geometric ops)
  int32_t fill_u8;     // static_cast<uint8>(fill_value)
  int32_t _pad[3];
};
static_assert(sizeof(DevOp) == 112, "DevOp layout");

enum { BLEND_IMAGE2 = 0, BLEND_IMAGE1 = 1, BLEND_INTERP = 2, BLEND_EXTRAP = 3 };

struct KParams {
  const uint8_t* in;
  uint8_t* out;
  int B, H, W;
  const DevOp* ops;  // [T][K]
  int T, n_draws, K, elementwise;
  unsigned long long seed;
  unsigned int call_counter;
  unsigned long long image_index_base;
  const int32_t* replay;
  int32_t* record;
  uint8_t* scratch;               // gridDim.x * 2 * scratch_stride bytes
  unsigned long long scratch_stride;
  unsigned int* work_counter;     // dynamic image scheduler (zeroed before launch)
};

struct LaunchInfo {
  int grid, block;
  size_t smem;
  bool image_in_smem;
};

// Decides grid / block / smem for a (H, W, C) image on this device.
LaunchInfo plan_launch(int B, int H, int W, int C, int num_sms, size_t smem_optin);
// Launches the fused policy kernel.  Returns cudaGetLastError().
cudaError_t launch_policy(const KParams& p, int C, const LaunchInfo& li, cudaStream_t stream);
// Opt in to the large dynamic shared memory carve-out for every kernel instantiation.
cudaError_t configure_kernels(size_t smem_optin);
// Shared-memory bytes the kernel needs besides the image itself.
size_t smem_overhead(int C);

}  // namespace chb
