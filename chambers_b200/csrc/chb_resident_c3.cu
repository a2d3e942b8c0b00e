// resident_kernel<3> instantiation (see chb_resident.cuh).
#include "chb_resident.cuh"
namespace chb {
cudaError_t launch_resident_c3(const KParams& p, int grid, cudaStream_t stream) { return launch_resident_c<3>(p, grid, stream); }
void resident_splits_c3(KParams& p, int aux_bytes) { resident_splits_c<3>(p, aux_bytes); }
cudaError_t configure_resident_c3(int smem_bytes) { return configure_resident_c<3>(smem_bytes); }
size_t resident_ctl_bytes() { return (sizeof(ResCtl) + 127) / 128 * 128; }
int resident_max_chunk_bytes() { return RES_CHUNK * RES_MAXCHUNK; }
}  // namespace chb
