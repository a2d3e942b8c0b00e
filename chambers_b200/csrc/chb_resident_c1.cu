// resident_kernel<1> instantiation (see chb_resident.cuh).
#include "chb_resident.cuh"
namespace chb {
cudaError_t launch_resident_c1(const KParams& p, int grid, cudaStream_t stream) { return launch_resident_c<1>(p, grid, stream); }
void resident_splits_c1(KParams& p, int aux_bytes) { resident_splits_c<1>(p, aux_bytes); }
cudaError_t configure_resident_c1(int smem_bytes) { return configure_resident_c<1>(smem_bytes); }
}  // namespace chb
