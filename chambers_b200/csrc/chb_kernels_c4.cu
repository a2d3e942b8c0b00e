// pass_kernel<4> instantiation (see chb_kernels.cuh).
#include "chb_kernels.cuh"
namespace chb {
CHB_DEFINE_CHANNEL_ENTRY(4)
}
