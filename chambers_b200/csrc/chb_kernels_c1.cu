// pass_kernel<1> instantiation (see chb_kernels.cuh).
#include "chb_kernels.cuh"
namespace chb {
CHB_DEFINE_CHANNEL_ENTRY(1)
}
