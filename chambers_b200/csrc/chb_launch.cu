// Tile planning and channel-count dispatch for the pass kernel.
#include "chb_internal.h"

namespace chb {

#define CHB_DECL(C)                                                                          \
  cudaError_t launch_pass_c##C(const KParams&, const TMap&, const TMap&, int, cudaStream_t); \
  cudaError_t configure_c##C();                                                              \
  int ctas_per_sm_c##C();
CHB_DECL(1) CHB_DECL(2) CHB_DECL(3) CHB_DECL(4)
#undef CHB_DECL
#define CHB_DECL_RES(C)                                                        \
  cudaError_t launch_resident_c##C(const KParams&, int, cudaStream_t);         \
  cudaError_t configure_resident_c##C(int);                                    \
  void resident_splits_c##C(KParams&, int);
CHB_DECL_RES(1) CHB_DECL_RES(2) CHB_DECL_RES(3) CHB_DECL_RES(4)
#undef CHB_DECL_RES

// 2-D tiles of 64 x 64 pixels, ragged at the right / bottom edge (64-pixel columns keep every tile
// row a whole number of 16-byte units for C = 1..4, and 64 x 64, 64 x 32 and 32 x 32 tiles are all
// whole rounds of 256 threads x 4 pixels); the same tile count cuts an image into flat runs or row
// strips for the passes that have no spatial op.  224 x 224 -> 4 x 4 tiles; 512 x 512 -> 8 x 8.
TilePlan plan_tiles(int H, int W) {
  TilePlan t;
  t.tw = 64;
  t.tiles_x = W > 0 ? (W + 63) / 64 : 1;
  t.tiles_y = H > 0 ? (H + 63) / 64 : 1;
  t.th = 64;
  t.n_tiles = t.tiles_x * t.tiles_y;
  return t;
}

cudaError_t configure_plan();

cudaError_t configure_kernels() {
  cudaError_t e;
  if ((e = configure_plan()) != cudaSuccess) return e;
  if ((e = configure_c1()) != cudaSuccess) return e;
  if ((e = configure_c2()) != cudaSuccess) return e;
  if ((e = configure_c3()) != cudaSuccess) return e;
  return configure_c4();
}

int pass_ctas_per_sm(int C) {  // decided by the shared-memory footprint of pass_kernel<C>
  switch (C) {
    case 1: return ctas_per_sm_c1();
    case 2: return ctas_per_sm_c2();
    case 3: return ctas_per_sm_c3();
    default: return ctas_per_sm_c4();
  }
}

cudaError_t launch_pass(const KParams& p, const TMap& a, const TMap& b, int C, int grid, cudaStream_t stream) {
  switch (C) {
    case 1: return launch_pass_c1(p, a, b, grid, stream);
    case 2: return launch_pass_c2(p, a, b, grid, stream);
    case 3: return launch_pass_c3(p, a, b, grid, stream);
    case 4: return launch_pass_c4(p, a, b, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t configure_resident(int smem_bytes) {
  cudaError_t e;
  if ((e = configure_resident_c1(smem_bytes)) != cudaSuccess) return e;
  if ((e = configure_resident_c2(smem_bytes)) != cudaSuccess) return e;
  if ((e = configure_resident_c3(smem_bytes)) != cudaSuccess) return e;
  return configure_resident_c4(smem_bytes);
}

void resident_splits(KParams& p, int C, int aux_bytes) {
  switch (C) {
    case 1: resident_splits_c1(p, aux_bytes); break;
    case 2: resident_splits_c2(p, aux_bytes); break;
    case 3: resident_splits_c3(p, aux_bytes); break;
    default: resident_splits_c4(p, aux_bytes); break;
  }
}

cudaError_t launch_resident(const KParams& p, int C, int grid, cudaStream_t stream) {
  switch (C) {
    case 1: return launch_resident_c1(p, grid, stream);
    case 2: return launch_resident_c2(p, grid, stream);
    case 3: return launch_resident_c3(p, grid, stream);
    case 4: return launch_resident_c4(p, grid, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace chb
