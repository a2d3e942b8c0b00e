// resident_kernel<4> instantiation (see chb_resident.cuh).
#include "chb_resident.cuh"
namespace chb {
cudaError_t launch_resident_c4(const KParams& p, int grid, cudaStream_t stream) { return launch_resident_c<4>(p, grid, stream); }
void resident_splits_c4(KParams& p, int aux_bytes) { resident_splits_c<4>(p, aux_bytes); }
cudaError_t configure_resident_c4(int smem_bytes) { return configure_resident_c<4>(smem_bytes); }
}  // namespace chb
