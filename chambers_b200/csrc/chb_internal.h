// Internal (non-ABI) declarations shared by chb_api.cu (host) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
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

constexpr int MAXC = 4;
// Work items are ordered by the code they run (bin_of, chb_kernels.cuh): a pass CTA then
// executes long runs of the same tile executor instead of thrashing the instruction cache, and the
// long tiles are scheduled first.
constexpr int NBINS = 18;  // two groups (passes that are not / are the image's last) of 9 executor bins

// ---- per-image pass state (global memory; written by the plan kernel and by pass finalisers) ----
// The chain of one image is evaluated lazily.  Between passes the "virtual image" is
//     src bytes -> spatial list (nearest warps / CutOut masks, each carrying the final colour it
//     contributes) -> LUT l1 -> optional kernel K (Color | Sharpness | bilinear warp) -> LUT l2
// and every op only edits that state; a pass over the pixels happens only when an op needs the
// pixels themselves (a histogram, or a materialised neighbourhood).
enum { PASS_WRITE_OUT = 0, PASS_COUNT = 1, PASS_WRITE_SCRATCH = 2 };
enum { K_NONE = 0, K_COLOR = 1, K_SHARP = 2, K_BILINEAR = 3 };
enum { SP_GEOM = 0, SP_MASK = 1 };

struct Spatial {  // 72 bytes
  int32_t type, fill_mode;
  float t[8];
  int32_t y0, y1, x0, x1;
  int32_t color[MAXC];
};

struct alignas(16) TileState {  // what a tile needs; loaded whole by every tile CTA
  int32_t pass_kind, src_sel, dst_sel, n_sp;
  int32_t kmode, l1_id, l2_id, sp_fast;  // sp_fast: every entry is constant-fill (bbox staging is valid)
  float kfactor;
  // a LUT that is the same for every channel and of the form (x & m) ^ c (Invert, Posterize, a full
  // Solarize and their compositions) is applied with one logic op per word instead of lookups:
  // bit 16 = valid, bits 8..15 = m, bits 0..7 = c
  int32_t l1_aff, l2_aff;
  // l1 is exactly ONE blend op against a constant (Brightness: 0, Contrast: its mean constant) applied to
  // the identity: blend mode + 1 (0 = no), the constant, float32(factor).  The resident engine then
  // evaluates the blend arithmetically in its flat executor instead of looking 150 K bytes up.
  int32_t l1_blend;
  Spatial sp[CHB_MAX_CHAIN];
  Spatial kgeo;  // K_BILINEAR: the warp (t, fill_mode, color[0] = fill)
  int32_t l1_blend_const;
  float l1_blend_factor;
  uint8_t l1[MAXC][256];
  uint8_t l2[MAXC][256];
};
static_assert(sizeof(TileState) % 16 == 0, "TileState must be a whole number of 16-byte units");
static_assert(offsetof(TileState, l1) % 16 == 0, "the LUTs must start on a 16-byte boundary");

struct ProgRec {
  int32_t table_index, negate, cy, cx;
};

struct alignas(16) ImgState {
  TileState t;
  // finaliser-only part
  int32_t next_op, n_prog, hist_valid, _pad0;
  ProgRec prog[CHB_MAX_CHAIN];
  uint32_t tiles_done, _pad1[3];
  uint32_t color_cnt[CHB_MAX_CHAIN];  // COUNT passes: pixels that resolved to spatial entry k
  uint32_t hist[MAXC][256];           // COUNT passes: values entering the last LUT
};
static_assert(sizeof(ImgState) % 16 == 0, "ImgState must be a whole number of 16-byte units");

struct KParams {
  const uint8_t* in;
  uint8_t* out;
  int B, H, W;
  const DevOp* ops;      // [T][K]
  const uint8_t* optab;  // [T*K][256]: the value map of every point-wise table op
  int T, n_draws, K, elementwise;
  unsigned long long seed;
  unsigned int call_counter;
  unsigned long long image_index_base;
  const int32_t* replay;
  int32_t* record;
  uint8_t* scratch;                   // image i, buffer s in {1, 2}: scratch + (2 i + s - 1) * stride
  unsigned long long scratch_stride;
  ImgState* states;                   // [B]
  int* lists;                         // [NBINS][B]: the images, binned by the executor of their first pass
  unsigned int* counters;             // 32 words: [NBINS] bin lengths, then: work counter, ticket counter, continuation
                                      // entries allocated, images that may still publish a continuation, CTAs that have left
  int* cont;                          // continuation list: image + 1 per published pass, 0 = not yet (follows counters)
  int tiles_x, tiles_y, tw, th, n_tiles;
  int force_generic;                  // debugging: route every tile through the scalar executor
  int use_tmap, box_rows, box_bytes;  // gather tiles: one tensor-map box of box_rows x box_bytes per tile
  // per-launch constants of the tile planner
  int strip_rows;                     // rows of a strip tile: ceil(H / n_tiles)
  int flat_units, flat_upt;           // flat runs: units per image (48 bytes for C = 3, else 16) and per tile
  int nf_first;                       // bin layout: 1 = every non-final pass first (small batches), 0 = heavy executors first
  int n_flat_tiles;                   // flat runs per image (tiles of <= 256 units; == n_tiles unless the batch is 16-byte aligned)
  int flags;                          // 1: in is 16-byte aligned, 2: out is, 4: images are whole 16-byte units, 8: rows are
  unsigned long long* timeline;       // debug builds (-DCHB_TIMELINE): [1024 CTAs][2][16] words, else NULL
  int res_smem_bytes;                 // resident engine: dynamic shared memory of a CTA (control + policy + image + aux region)
  float* outf;                        //   fused ImageNetNormalization epilogue: float32 output (NULL: uint8 output to `out`)
  int norm_mode;                      //   chb_norm_mode of the epilogue (CHB_NORM_TF / CHB_NORM_TORCH)
  int res_lpt;                        //   1: small batches are claimed most-expensive-chain-first (CHB_LPT=0 disables)
  int res_rules;                      //   1: advance() materialises non-flat views in front of Sharpness / a histogram op
  int res_sharp_rows[4];              //   Sharpness: rows per sub-strip of the column walk, last pass cut into 1..4 row ranges
  int res_split_pct;                  //   an item is split when it would outlast this % of the average SM's load (CHB_SPLIT_PCT)
  int res_split;                      //   1: small batches may cut the last pass of an expensive image into row ranges (not when d_in == d_out)
};

// Opaque copy of a CUtensorMap (cuda.h), passed to the pass kernel as a __grid_constant__ parameter.
struct alignas(64) TMap {
  unsigned char bytes[128];
};

struct TilePlan {
  int tiles_x, tiles_y, tw, th, n_tiles;
};
TilePlan plan_tiles(int H, int W);

// Launchers (chb_kernels_c<N>.cu / chb_launch.cu).  Each returns cudaGetLastError().
cudaError_t launch_optab(const DevOp* ops, uint8_t* optab, int n_ops, cudaStream_t stream);
cudaError_t launch_plan(const KParams& p, int C, cudaStream_t stream);
cudaError_t launch_pass(const KParams& p, const TMap& tm_in, const TMap& tm_scr, int C, int grid, cudaStream_t stream);
cudaError_t configure_kernels();
int pass_ctas_per_sm(int C);
// Image-resident engine (chb_resident.cuh): one CTA per SM, the whole image in shared memory.
cudaError_t launch_resident(const KParams& p, int C, int grid, cudaStream_t stream);
cudaError_t configure_resident(int smem_bytes);
void resident_splits(KParams& p, int C, int aux_bytes);
// The layers either side of the policy path (chb_frontend.cu).
cudaError_t launch_normalize(const void* in, int in_is_f32, float* out, unsigned long long n, int C, int mode, int num_sms,
                             cudaStream_t stream);
cudaError_t launch_resize(const void* in, int in_is_f32, void* out, int B, int IH, int IW, int C, int OH, int OW, int nearest,
                          int num_sms, cudaStream_t stream);
size_t resident_ctl_bytes();  // shared memory in front of the image buffer
int resident_max_chunk_bytes();

}  // namespace chb
