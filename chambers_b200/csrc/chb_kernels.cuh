// Tile-parallel RandAugment / AutoAugment policy kernels for sm_100a.
//
// Work is cut far below one image so that a 256-image batch already fills 148 SMs evenly:
//   plan_kernel   one CTA per image: decodes the image's op schedule from Philox4x32-10 (or a
//                 replayed schedule), folds the chain into a per-image pass state (ImgState) and
//                 appends the image to the bin of the executor its first pass runs.
//   pass_kernel   ONE persistent launch per call: 2 CTAs per SM claim chunks of (image, tile) work from
//                 an atomic counter, bins in start order.  A tile is ~12 KB of the image.  Each CTA
//                 is a producer / consumer pipeline over a ring of four 16 KB units: one warp claims
//                 work and drives the TMA loads (cp.async.bulk[.tensor] + mbarrier), eight warps
//                 compute from shared memory.  A pass that is not the image's last (COUNT,
//                 WRITE_SCRATCH) ends with one CTA resuming the chain walk and publishing the image's
//                 next pass in a global list that every producer holds a ticket into, so later passes
//                 run inside the same launch, spread over all CTAs.
//
// The chain is evaluated lazily (chb_internal.h: ImgState).  Point-wise ops compose into 256-entry
// LUTs, nearest-neighbour warps and CutOut go on a spatial list, Color / Sharpness / a bilinear warp
// occupy the single "kernel" slot K.  Pixels are only touched by
//   WRITE_OUT      the one pass every image has: src -> spatial list -> l1 -> K -> l2 -> out;
//   COUNT          Equalize / AutoContrast need the histogram of the virtual image: tiles count into
//                  shared memory, the CTA that completes the pass turns the histogram into a LUT
//                  (warp-scan CDF) and resumes the chain walk;
//   WRITE_SCRATCH  an op needs a materialised neighbourhood of something that is itself a
//                  neighbourhood op (e.g. Rotate after Sharpness): the virtual image is written to
//                  a scratch image (L2-resident when its next pass follows soon) and becomes the new src.
// so a RandAugment image costs one HBM read and one HBM write unless its chain holds a histogram
// op (one extra read, served from L2 when the next pass follows soon) or a rare
// neighbourhood-of-neighbourhood pair.
//
// Tile executors (all bit-exact twins of each other; the scalar one is the fallback for odd shapes):
//   exec_flat     no spatial op: 48-byte (16-pixel) units, LDS.128 -> LUT/Color in registers -> STS.128 -> one TMA store
//   gather_warp   one or two spatial entries, WRITE pass: the source bounding box of a 64 x 64 tile is
//                 staged in shared memory (one tensor-map TMA box); each warp gathers whole rows and
//                 stores them itself (no CTA barrier)
//   exec_gather   the same staging for COUNT passes and longer spatial lists, CTA-wide
//   exec_sharp    Sharpness: row strip + halo staged in shared memory, sliding 3x3 window per word column
//   exec_gather_sharp   Sharpness of a gathered image
//   exec_generic  anything, one pixel per thread straight from global memory
#pragma once
#include <string.h>

#include "chb_device.cuh"

namespace chb {
namespace {

// Loads of the policy table (DevOp / value maps).  The tile engine reads them from global memory
// through the read-only path; the resident engine (chb_resident.cuh) stages the table in shared
// memory and defines CHB_LDP as a plain load before including this file.
#ifndef CHB_LDP
#define CHB_LDP(ptr) __ldg(ptr)
#endif

constexpr int NCONS = 256;                       // consumer threads of a pass CTA (8 warps)
constexpr int NT = NCONS + 32;                   // + one producer warp
constexpr int PLAN_NT = 128;                     // threads per plan CTA
constexpr int NU = 4;                            // pipeline units per CTA; a tile occupies one or two
constexpr int UNIT_BYTES = 16 * 1024;            // staged source bytes of one unit
// TileStates are prefetched one item ahead into a ring indexed by the item number and read in place
// by the consumers.  At most NU items are in flight (each holds a unit), so when item j's state is
// fetched (one iteration early) the oldest item that can still be in use is j - NU: NU + 2 buffers.
constexpr int NST = NU + 2;
constexpr int STATE_VECS = (int)(offsetof(ImgState, hist) / 16);   // everything but the histogram
constexpr int TILE_VECS = (int)(sizeof(TileState) / 16);
constexpr int FIN_BYTES = (int)sizeof(ImgState) + MAXC * 256 * 5;  // finaliser: state + hmap + etab
// one output staging tile: 64 x 64 pixels (a flat run never exceeds that either)
__host__ __device__ constexpr int ostage_bytes(int C) { return 4096 * C + 256; }
// R region: two output tiles | eight histogram copies of a COUNT pass | finaliser scratch
__host__ __device__ constexpr int r_bytes(int C) {
  return (2 * ostage_bytes(C) > FIN_BYTES ? 2 * ostage_bytes(C) : FIN_BYTES + 127) / 128 * 128;
}
static_assert(8 * (1 * 1024 + 16) <= r_bytes(1) && 8 * (2 * 1024 + 16) <= r_bytes(2) && 8 * (3 * 1024 + 16) <= r_bytes(3) &&
                  8 * (4 * 1024 + 16) <= r_bytes(4), "the histogram copies of exec_flat<COUNT> live in the R region");
static_assert(offsetof(ImgState, hist) % 16 == 0, "hist must start on a 16-byte boundary");
static_assert(offsetof(ImgState, next_op) == sizeof(TileState), "finaliser part follows the tile part");

struct Rect {
  int x0, x1, y0, y1;
};

// The value, but opaque to the optimiser at this point: everything an executor derives from it is
// computed inside the executor.  (Without it the set-up of the rarely used executors -- float
// conversions of the image size for the coordinate maps, row pitches -- is hoisted in front of the
// executor dispatch and runs for every tile.)
__device__ __forceinline__ int opaque(int v) {
  asm volatile("" : "+r"(v));
  return v;
}
// Where it pays was measured per executor (profiles/r01_v13_ab_experiments.txt, 5): the scalar, list
// and gather+sharpness executors yes; in the Sharpness strip executor it cost the flat path 5-15 %.
#ifndef CHB_OPQ_GEN
#define CHB_OPQ_GEN 1
#endif
#ifndef CHB_OPQ_LIST
#define CHB_OPQ_LIST 1
#endif
#ifndef CHB_OPQ_GS
#define CHB_OPQ_GS 1
#endif
#ifndef CHB_OPQ_SHARP
#define CHB_OPQ_SHARP 0
#endif
template <int ON>
__device__ __forceinline__ int opaque_if(int v) { return ON ? opaque(v) : v; }

// Debug timeline (-DCHB_TIMELINE, tools/timeline.py): every pass CTA leaves two 16-word records,
// [cta][producer | consumer][16], in KParams::timeline.  Stamps are %globaltimer (ns),
// accumulators are clock64 cycles.  Compiled out of the production library.
#ifdef CHB_TIMELINE
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL_DECL()                    \
  unsigned long long tl_acc[16];     \
  for (int i_ = 0; i_ < 16; ++i_) tl_acc[i_] = 0; \
  long long tl_t0 = 0;               \
  (void)tl_t0
#define TL_T0() tl_t0 = clock64()
#define TL_ACC(i) tl_acc[i] += (unsigned long long)(clock64() - tl_t0)
#define TL_ADD(i, v) tl_acc[i] += (unsigned long long)(v)
#define TL_STAMP(i) tl_acc[i] = tl_now()
#define TL_STAMP_ONCE(i) if (tl_acc[i] == 0) tl_acc[i] = tl_now()
#define TL_FLUSH(role, leader)                                                                             \
  if (p.timeline && (leader)) {                                                                            \
    unsigned long long* o_ = p.timeline + (((size_t)blockIdx.x) * 2 + (role)) * 16;       \
    for (int i_ = 0; i_ < 16; ++i_) o_[i_] = tl_acc[i_];                                                    \
  }
#else
#define TL_DECL()
#define TL_T0()
#define TL_ACC(i)
#define TL_ADD(i, v)
#define TL_STAMP(i)
#define TL_STAMP_ONCE(i)
#define TL_FLUSH(role, leader)
#endif

// Walks the spatial list from the last op to the first.  Returns -1 when the output pixel resolves
// to source pixel (x, y) (updated in place), else the index of the spatial entry whose colour it shows.
__device__ __forceinline__ int resolve(const Spatial* sp, int n_sp, int H, int W, int& x, int& y) {
  for (int k = n_sp - 1; k >= 0; --k) {
    const Spatial& e = sp[k];
    if (e.type == SP_MASK) {
      if (y >= e.y0 && y < e.y1 && x >= e.x0 && x < e.x1) return k;
    } else {
      float sx, sy;
      affine_source(e.t, x, y, sx, sy);
      if (e.fill_mode != CHB_FILL_CONSTANT) {
        sx = map_coordinate(sx, W, e.fill_mode);
        sy = map_coordinate(sy, H, e.fill_mode);
      }
      const float rx = round_half_away(sx), ry = round_half_away(sy);
      if (!(rx >= 0.0f && rx < (float)W && ry >= 0.0f && ry < (float)H)) return k;
      x = (int)rx;
      y = (int)ry;
    }
  }
  return -1;
}

// =========================================================================== chain walk (plan)
// Equalize / AutoContrast table of every channel from hmap[c][256] (counts of the virtual image's
// values), one warp per channel; result in etab[c][256].
__device__ void stat_tables(int kind, int C, const uint32_t* hmap, uint8_t* etab, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  if (warp < C) {
    const int c = warp;
    const uint32_t* hc = hmap + c * 256;
    int h[8];
    int sum = 0, last = -1, first = 256;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h[i] = (int)hc[lane * 8 + i];
      sum += h[i];
      if (h[i] != 0) {
        last = lane * 8 + i;
        first = min(first, lane * 8 + i);
      }
    }
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    int excl = incl - sum;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, d));
      first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, d));
    }
    if (kind == CHB_OP_EQUALIZE) {
      // tfa.image.equalize _scale_channel (oracle/ops.py equalize_lut), int32 arithmetic.
      const int last_count = (int)hc[last < 0 ? 0 : last];
      const int step = (total - last_count) / 255;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int v = lane * 8 + i;
        int e = v;
        if (step != 0) e = min(255, max(0, (excl + step / 2) / step));
        etab[c * 256 + v] = (uint8_t)e;
        excl += h[i];
      }
    } else {
      // AutoContrast, image_augmentations.py:69-86.
      const float lo = (float)first, hi = (float)last;
      float scale = 1.0f, offset = 0.0f;
      if (hi > lo) {
        scale = __fdiv_rn(255.0f, __fsub_rn(hi, lo));  // :72
        offset = __fmul_rn(-lo, scale);                // :73
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int v = lane * 8 + i;
        float xv = __fadd_rn(__fmul_rn((float)v, scale), offset);  // :84
        xv = fminf(fmaxf(xv, 0.0f), 255.0f);                       // :85
        etab[c * 256 + v] = (uint8_t)(int)xv;                      // :86
      }
    }
  }
}

__device__ __forceinline__ void reset_view(ImgState* s, int tid, int nt) {  // after a materialisation
  for (int t = tid; t < MAXC * 256; t += nt) {
    s->t.l1[t >> 8][t & 255] = (uint8_t)(t & 255);
    s->t.l2[t >> 8][t & 255] = (uint8_t)(t & 255);
  }
  if (tid == 0) {
    s->t.n_sp = 0; s->t.kmode = K_NONE; s->t.l1_id = 1; s->t.l2_id = 1; s->t.sp_fast = 1; s->t.l1_aff = 0; s->t.l2_aff = 0;
    s->t.l1_blend = 0;
    s->t.kfactor = 0.0f; s->hist_valid = 0;
  }
}

// Executor bin of an image's next pass (see NBINS).  Bins are laid out in the order the work should
// start, in six segments: {heavy, light 2-D tiles, flat runs} x {passes that are not the image's
// last, final passes}.  Heavy executors (their tiles cost microseconds each) must not be left for the
// tail; non-final passes come before final ones of the same weight because the image's next pass can
// only be published once they are done.  Within a segment the most expensive code comes first.
// Two layouts, chosen per call (KParams::nf_first):
//   large batches   0..3 heavy nf | 4..7 heavy f | 8..9 light nf | 10..12 flat nf | 13..14 light f | 15..17 flat f
//   small batches   0..3 heavy nf | 4..5 light nf | 6..8 flat nf | 9..12 heavy f  | 13..14 light f | 15..17 flat f
// (small batches: every non-final pass first -- the second stage of those images is the critical path
// of a 100 us kernel: 256 images 0.111 -> 0.104 ms; large batches lose 4 % that way,
// profiles/r01_v13_ab_experiments.txt 10)
constexpr int HEAVY_BINS = 4;
__host__ __device__ constexpr int seg_first(int k, bool nf) {
  return nf ? (k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 6 : k == 3 ? 9 : k == 4 ? 13 : 15)
            : (k == 0 ? 0 : k == 1 ? 4 : k == 2 ? 8 : k == 3 ? 10 : k == 4 ? 13 : 15);
}
__host__ __device__ constexpr int seg_last(int k, bool nf) {
  return nf ? (k == 0 ? 3 : k == 1 ? 5 : k == 2 ? 8 : k == 3 ? 12 : k == 4 ? 14 : 17)
            : (k == 0 ? 3 : k == 1 ? 7 : k == 2 ? 9 : k == 3 ? 12 : k == 4 ? 14 : 17);
}
// 0 heavy (one tile per claim), 1 light, 2 flat
__host__ __device__ constexpr int seg_kind(int k, bool nf) {
  return nf ? ((k == 0 || k == 3) ? 0 : (k == 1 || k == 4) ? 1 : 2) : (k < 2 ? 0 : (k == 2 || k == 4) ? 1 : 2);
}
__host__ __device__ constexpr bool seg_nonfinal_light(int k, bool nf) { return nf ? (k == 1 || k == 2) : (k == 2 || k == 3); }
constexpr int CONT_FLAT = 1 << 30;
__device__ __forceinline__ bool is_flat_bin(int bin, bool nf) {
  return bin >= 15 || (nf ? (bin >= 6 && bin <= 8) : (bin >= 10 && bin <= 12));
}
__device__ __forceinline__ int bin_of(const TileState& t, bool nf) {
  bool any_geom = false;
  for (int k = 0; k < t.n_sp; ++k) any_geom = any_geom || (t.sp[k].type == SP_GEOM);
  const bool count = t.pass_kind == PASS_COUNT;
  const bool fin = t.pass_kind == PASS_WRITE_OUT;
  int cost;  // executor cost class, most expensive first
  if (t.kmode == K_SHARP) cost = t.n_sp > 0 ? 0 : 1;
  else if (t.kmode == K_BILINEAR) cost = 2;
  else if (t.n_sp >= 2 && (any_geom || count)) cost = 3;
  else if (t.n_sp == 1 && (any_geom || count)) cost = t.kmode == K_COLOR ? 4 : 5;
  else if (count) cost = 6;
  else cost = t.kmode == K_COLOR ? 7 : 8;
  if (nf) {
    if (cost < HEAVY_BINS) return (fin ? 9 : 0) + cost;
    if (cost < 6) return (fin ? 13 : 4) + (cost - 4);
    return (fin ? 15 : 6) + (cost - 6);
  }
  if (cost < HEAVY_BINS) return (fin ? 4 : 0) + cost;
  if (cost < 6) return (fin ? 13 : 8) + (cost - 4);
  return (fin ? 15 : 10) + (cost - 6);
}
// Queues image `img` for its first pass.
__device__ __forceinline__ void enqueue_pass(const KParams& p, const TileState& t, int img) {
  const int bin = bin_of(t, p.nf_first != 0);
  const unsigned pos = atomicAdd(p.counters + bin, 1u);
  p.lists[(size_t)bin * p.B + pos] = img;
}

// CTA-cooperative chain walk.  *s lives in shared memory; starting at s->next_op every op is folded
// into the view until one needs the pixels.  On return s->t.pass_kind names the pass to run now; the
// walk resumes at the same op once that pass has finished.  hmap: MAXC*256 words, etab: MAXC*256
// bytes of shared scratch.  Needs nt >= 32 * C; `sync` is the barrier of the nt participating threads.
template <class Sync>
__device__ void advance(ImgState* s, ImgState* g, const KParams& p, int C, int H, int W, uint32_t* hmap,
                        uint8_t* etab, int tid, int nt, Sync sync) {
  for (;;) {
    sync();
    const int pi = s->next_op;
    const int n_prog = s->n_prog;
    const int kmode = s->t.kmode;
    const int n_sp = s->t.n_sp;
    const int hist_valid = s->hist_valid;
    const int src_sel = s->t.src_sel;
    ProgRec pe = {0, 0, 0, 0};
    if (pi < n_prog) pe = s->prog[pi];
    sync();
    if (pi >= n_prog) {
      if (tid == 0) s->t.pass_kind = PASS_WRITE_OUT;
      break;
    }
    const DevOp* op = p.ops + pe.table_index;
    const int kind = CHB_LDP(&op->kind);
    const bool frozen = (kmode == K_SHARP || kmode == K_BILINEAR);  // K has consumed the spatial list
    uint8_t(*last_lut)[256] = (kmode == K_NONE) ? s->t.l1 : s->t.l2;
    bool boundary = false;

    if (is_pointwise(kind)) {
      const uint8_t* tab = p.optab + (size_t)pe.table_index * 256;
      for (int t = tid; t < C * 256; t += nt) last_lut[t >> 8][t & 255] = CHB_LDP(tab + last_lut[t >> 8][t & 255]);
      if (!frozen)
        for (int t = tid; t < n_sp * C; t += nt) {
          const int k = t / C, c = t - k * C;
          s->t.sp[k].color[c] = CHB_LDP(tab + s->t.sp[k].color[c]);
        }
      if (tid == 0) {
        // the first op on an identity l1, and it is a blend against a constant: remember its closed form
        const bool blend_op = (kind == CHB_OP_BRIGHTNESS || kind == CHB_OP_CONTRAST);
        if (kmode == K_NONE) {
          s->t.l1_blend = (s->t.l1_id && blend_op) ? CHB_LDP(&op->blend_mode) + 1 : 0;
          s->t.l1_blend_const = (kind == CHB_OP_CONTRAST) ? CHB_LDP(&op->ip0) : 0;
          s->t.l1_blend_factor = CHB_LDP(&op->factor);
        }
        if (kmode == K_NONE) s->t.l1_id = 0; else s->t.l2_id = 0;
        s->next_op = pi + 1;
      }
    } else if (kind == CHB_OP_EQUALIZE || kind == CHB_OP_AUTOCONTRAST) {
      if (!hist_valid && p.res_rules && (kmode != K_NONE || n_sp > 0)) {
        // resident engine: the histogram of a view that holds a warp, a mask, Color or Sharpness is taken
        // while that view is MATERIALISED (one evaluation of the expensive part, tallied on the fly or by a
        // flat COUNT pass afterwards) instead of evaluating it once to count and once more to write
        boundary = true;
      } else if (!hist_valid) {
        // the histogram of the virtual image is needed: run a COUNT pass, then come back here.
        for (int t = tid; t < MAXC * 256; t += nt) (&g->hist[0][0])[t] = 0u;
        if (tid < CHB_MAX_CHAIN) s->color_cnt[tid] = 0u;
        if (tid == 0) { s->t.pass_kind = PASS_COUNT; s->tiles_done = 0u; }
        break;
      }
      if (!boundary) {
      // s->hist counts the values entering the last LUT; map them through it, add the spatial colours.
      for (int t = tid; t < MAXC * 256; t += nt) hmap[t] = 0u;
      sync();
      for (int t = tid; t < C * 256; t += nt) {
        const uint32_t cnt = s->hist[t >> 8][t & 255];
        if (cnt) atomicAdd(&hmap[(t & ~255) + last_lut[t >> 8][t & 255]], cnt);
      }
      if (!frozen)
        for (int t = tid; t < n_sp * C; t += nt) {
          const int k = t / C, c = t - k * C;
          const uint32_t cnt = s->color_cnt[k];
          if (cnt) atomicAdd(&hmap[c * 256 + s->t.sp[k].color[c]], cnt);
        }
      sync();
      stat_tables(kind, C, hmap, etab, tid);
      sync();
      for (int t = tid; t < C * 256; t += nt) last_lut[t >> 8][t & 255] = etab[(t & ~255) + last_lut[t >> 8][t & 255]];
      if (!frozen)
        for (int t = tid; t < n_sp * C; t += nt) {
          const int k = t / C, c = t - k * C;
          s->t.sp[k].color[c] = etab[c * 256 + s->t.sp[k].color[c]];
        }
      if (tid == 0) {
        if (kmode == K_NONE) { s->t.l1_id = 0; s->t.l1_blend = 0; } else s->t.l2_id = 0;
        // a presence-only histogram (resident engine, AutoContrast: hist_valid == 2) serves this op alone
        if (hist_valid == 2) s->hist_valid = 0;
        s->next_op = pi + 1;
      }
      }
    } else if (kind == CHB_OP_COLOR) {
      const int mode = CHB_LDP(&op->blend_mode);
      if (mode == BLEND_IMAGE2 || C != 3) {  // factor 1: blend returns the image itself (:30-31)
        if (tid == 0) s->next_op = pi + 1;
      } else if (kmode == K_NONE) {
        const float f = (mode == BLEND_IMAGE1) ? 0.0f : CHB_LDP(&op->factor);
        if (tid < n_sp) {  // spatial colours are pixels too
          int* col = s->t.sp[tid].color;
          uint32_t R, G, B;
          color_pixel_f((float)col[0], (float)col[1], (float)col[2], f, R, G, B);
          col[0] = (int)R; col[1] = (int)G; col[2] = (int)B;
        }
        if (tid == 0) { s->t.kmode = K_COLOR; s->t.kfactor = f; s->hist_valid = 0; s->next_op = pi + 1; }
      } else {
        boundary = true;
      }
    } else if (kind == CHB_OP_SHARPNESS) {
      const int mode = CHB_LDP(&op->blend_mode);
      if (mode == BLEND_IMAGE2) {
        if (tid == 0) s->next_op = pi + 1;
      } else if (kmode == K_NONE && p.res_rules && n_sp > 0) {
        // resident engine: Sharpness of a warped / masked view runs on the materialised view (a gather
        // into the scratch image or a paint in place, then the plain row walk) -- cheaper than gathering
        // a halo band per strip
        boundary = true;
      } else if (kmode == K_NONE) {
        if (tid == 0) {
          s->t.kmode = K_SHARP;
          s->t.kfactor = (mode == BLEND_IMAGE1) ? 0.0f : CHB_LDP(&op->factor);  // factor 0: deg exactly
          s->hist_valid = 0; s->next_op = pi + 1;
        }
      } else {
        boundary = true;
      }
    } else if (kind == CHB_OP_CUTOUT) {
      if (frozen) {
        boundary = true;
      } else if (tid == 0) {
        // tfa.image.random_cutout: rows [cy-h, cy+h) x cols [cx-h, cx+h), clipped (oracle/ops.py cutout)
        const int h = CHB_LDP(&op->ip0), colr = CHB_LDP(&op->ip1);
        Spatial& e = s->t.sp[n_sp];
        e.type = SP_MASK; e.fill_mode = CHB_FILL_CONSTANT;
        e.y0 = max(0, pe.cy - h); e.y1 = min(H, pe.cy + h);
        e.x0 = max(0, pe.cx - h); e.x1 = min(W, pe.cx + h);
        for (int c = 0; c < MAXC; ++c) e.color[c] = colr;
        s->t.n_sp = n_sp + 1; s->hist_valid = 0; s->next_op = pi + 1;
      }
    } else if (kind == CHB_OP_SHEAR_X || kind == CHB_OP_SHEAR_Y || kind == CHB_OP_TRANSLATE_X ||
               kind == CHB_OP_TRANSLATE_Y || kind == CHB_OP_ROTATE) {
      const float* tsrc = op->coef[pe.negate ? 1 : 0];
      const int fill_mode = CHB_LDP(&op->fill_mode), fill = CHB_LDP(&op->fill_u8);
      if (CHB_LDP(&op->interp) == CHB_INTERP_NEAREST) {
        if (frozen) {
          boundary = true;
        } else if (tid == 0) {
          Spatial& e = s->t.sp[n_sp];
          e.type = SP_GEOM; e.fill_mode = fill_mode;
          for (int q = 0; q < 8; ++q) e.t[q] = CHB_LDP(tsrc + q);
          for (int c = 0; c < MAXC; ++c) e.color[c] = fill;
          if (fill_mode != CHB_FILL_CONSTANT) s->t.sp_fast = 0;
          s->t.n_sp = n_sp + 1; s->hist_valid = 0; s->next_op = pi + 1;
        }
      } else {
        // bilinear warp: needs the plain neighbourhood of its input
        if (n_sp > 0 || kmode != K_NONE) {
          boundary = true;
        } else if (tid == 0) {
          Spatial& e = s->t.kgeo;
          e.type = SP_GEOM; e.fill_mode = fill_mode;
          for (int q = 0; q < 8; ++q) e.t[q] = CHB_LDP(tsrc + q);
          for (int c = 0; c < MAXC; ++c) e.color[c] = fill;
          s->t.kmode = K_BILINEAR; s->hist_valid = 0; s->next_op = pi + 1;
        }
      }
    } else {
      if (tid == 0) s->next_op = pi + 1;  // unknown kinds are rejected on the host
    }
    if (boundary) {
      if (tid == 0) {
        s->t.pass_kind = PASS_WRITE_SCRATCH;
        s->t.dst_sel = (src_sel == 1) ? 2 : 1;
        s->tiles_done = 0u;
      }
      break;
    }
  }
  sync();
  // Classify the LUTs of the pass that runs next (see TileState::l1_aff): hmap[0..1] collect the votes.
  if (s->t.l1_id && s->t.l2_id) {  // nothing to classify (uniform: read after the barrier above)
    if (tid == 0) { s->t.l1_aff = 0; s->t.l2_aff = 0; }
    sync();
    return;
  }
  if (tid < 2) hmap[tid] = 1u;
  sync();
  {
    const int c1 = s->t.l1[0][0], m1 = s->t.l1[0][255] ^ c1;
    const int c2 = s->t.l2[0][0], m2 = s->t.l2[0][255] ^ c2;
    bool ok1 = true, ok2 = true;
    for (int t = tid; t < C * 256; t += nt) {
      const int x = t & 255;
      ok1 = ok1 && (s->t.l1[t >> 8][x] == ((x & m1) ^ c1));
      ok2 = ok2 && (s->t.l2[t >> 8][x] == ((x & m2) ^ c2));
    }
    if (!ok1) hmap[0] = 0u;
    if (!ok2) hmap[1] = 0u;
    sync();
    if (tid == 0) {
      s->t.l1_aff = (hmap[0] && !s->t.l1_id) ? (0x10000 | (m1 << 8) | c1) : 0;
      s->t.l2_aff = (hmap[1] && !s->t.l2_id) ? (0x10000 | (m2 << 8) | c2) : 0;
    }
  }
  sync();
}

// Twin of oracle/philox.py decode_schedule for ONE image, executed by warp 0: every lane draws the
// Philox block of one slot, lane 0 assembles the chain.
__device__ void decode_image(const KParams& p, ImgState* s, uint32_t (*rnd)[4], uint32_t (*rndc)[4], int img,
                             int H, int W, int lane) {
  const unsigned long long own = p.image_index_base + (unsigned long long)img;
  const unsigned long long stream_img = p.elementwise ? own : ~0ull;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  const int n_slots = p.n_draws * (p.K + 1);
  if (!p.replay && lane < n_slots) {
    const uint4 w = philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter, (uint32_t)lane), key);
    uint4 wc = w;
    if (!p.elementwise) wc = philox4x32_10(make_uint4((uint32_t)own, (uint32_t)(own >> 32), p.call_counter, (uint32_t)lane), key);
    rnd[lane][0] = w.x; rnd[lane][1] = w.y; rnd[lane][2] = w.z; rnd[lane][3] = w.w;
    rndc[lane][2] = wc.z; rndc[lane][3] = wc.w;
  }
  __syncwarp();
  // one lane per (draw, sub-op) slot -- at most CHB_MAX_CHAIN of them -- in chain order; the applied ops
  // are compacted into s->prog with a ballot
  const int n_ops = p.n_draws * p.K;
  bool applied = false;
  int opi = 0, negate = 0, cy = 0, cx = 0;
  if (lane < n_ops) {
    const int i = lane / p.K, j = lane - i * p.K;
    const int slot0 = i * (p.K + 1);
    const size_t rbase = ((size_t)img * p.n_draws + i) * p.K * CHB_SCHED_FIELDS;
    int choice;
    if (p.replay) {
      choice = p.replay[rbase];
      if (choice < 0 || choice >= p.T) choice = 0;
    } else {
      choice = (int)__umulhi(rnd[slot0][0], (uint32_t)p.T);
    }
    opi = choice * p.K + j;
    const int kind = CHB_LDP(&p.ops[opi].kind);
    if (p.replay) {
      const int32_t* r = p.replay + rbase + (size_t)j * CHB_SCHED_FIELDS;
      applied = (r[1] != 0) && kind >= 0;
      negate = r[2] != 0;
      cy = r[3];
      cx = r[4];
    } else {
      const int slot = slot0 + 1 + j;
      applied = (kind >= 0) && ((int)(rnd[slot][0] >> 8) < CHB_LDP(&p.ops[opi].thr24));
      negate = rnd[slot][1] < 0x80000000u;
      cy = (int)__umulhi(rndc[slot][2], (uint32_t)H);
      cx = (int)__umulhi(rndc[slot][3], (uint32_t)W);
    }
    if (p.record) {
      int32_t* r = p.record + rbase + (size_t)j * CHB_SCHED_FIELDS;
      r[0] = choice; r[1] = applied ? 1 : 0; r[2] = negate; r[3] = cy; r[4] = cx;
    }
  }
  const unsigned mask = __ballot_sync(0xFFFFFFFFu, applied);
  const int pos = __popc(mask & ((1u << lane) - 1u));
  if (applied && pos < CHB_MAX_CHAIN) {
    s->prog[pos].table_index = opi; s->prog[pos].negate = negate; s->prog[pos].cy = cy; s->prog[pos].cx = cx;
  }
  if (lane == 0) s->n_prog = min(__popc(mask), CHB_MAX_CHAIN);
  __syncwarp();
}

#ifdef CHB_WITH_PLAN
// Value map of every point-wise table op, built once per uploaded policy.
__global__ void optab_kernel(const DevOp* ops, uint8_t* optab) {
  const DevOp op = ops[blockIdx.x];
  optab[(size_t)blockIdx.x * 256 + threadIdx.x] = (uint8_t)pointwise_value(op, (int)threadIdx.x);
}

__global__ void __launch_bounds__(PLAN_NT) plan_kernel(const KParams p, int C) {
  __shared__ ImgState s;
  __shared__ uint32_t hmap[MAXC * 256];
  __shared__ uint8_t etab[MAXC * 256];
  __shared__ uint32_t rnd[32][4], rndc[32][4];
  const int tid = threadIdx.x;
  const int img = blockIdx.x;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the pass kernel may set itself up while we plan
  // (the counters are zeroed by a memset node in front of this kernel)
  for (int i = tid; i < STATE_VECS; i += PLAN_NT) reinterpret_cast<uint4*>(&s)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid < 32) decode_image(p, &s, rnd, rndc, img, p.H, p.W, tid);
  reset_view(&s, tid, PLAN_NT);
  __syncthreads();
  advance(&s, p.states + img, p, C, p.H, p.W, hmap, etab, tid, PLAN_NT, [] { __syncthreads(); });
  for (int i = tid; i < STATE_VECS; i += PLAN_NT)
    reinterpret_cast<uint4*>(p.states + img)[i] = reinterpret_cast<const uint4*>(&s)[i];
  if (tid == 0) {
    if (s.t.pass_kind != PASS_WRITE_OUT) atomicAdd(p.counters + NBINS + 3, 1u);  // this image will publish a continuation
    enqueue_pass(p, s.t, img);
  }
}

#endif  // CHB_WITH_PLAN

// ============================================================================== pipeline units
// A pass CTA is a producer / consumer pipeline over a ring of NU units of 16 KB.  The producer warp
// claims (image, tile) items two ahead, prefetches each image's TileState with a TMA bulk copy,
// decides how the tile will be executed, takes one or two consecutive units for it and issues the
// TMA loads of the tile's source bytes -- one bulk copy for a flat run or a row strip, one 3-D
// tensor-map box (cp.async.bulk.tensor) for the bounding box of a gather tile -- all completing on
// the first unit's `full` mbarrier.  The 8 consumer warps wait on `full`, compute from shared
// memory into an output staging tile and hand that to the TMA again (bulk stores), then release the
// units through their `empty` mbarriers.  Up to four flat tiles or two gather tiles are in flight
// per CTA, two CTAs per SM.
enum { CLS_END = 0, CLS_FLAT = 1, CLS_GATHER = 2, CLS_SHARP = 3, CLS_GENERIC = 4, CLS_GATHER_SHARP = 5, CLS_SKIP = 6 };

struct alignas(16) SlotInfo {  // written by the producer, read by the consumers
  int32_t cls, img, tile, pass_kind;
  int32_t bx0, bx1, by0, by1;        // GATHER: source bounding box (inclusive); SHARP: by0 = first staged row
  int32_t bxb0, rowb, pitch, rows;   // GATHER: first staged byte of a row, staged bytes per row, pitch, rows
  int32_t x0, x1, y0, y1;            // region of the tile (box or strip); FLAT: x0 / x1 = first / last unit
  uint32_t fillc[2];                 // GATHER: the colour bytes of the last / last-but-one spatial entry
  int32_t paint;                     // FLAT: the spatial list is masks only; paint them over the result
  int32_t span;                      // units this tile occupies (1 or 2)
  int32_t sharp_rows;                // SHARP classes: rows per sub-strip of the column walk
  int32_t first, last;               // COUNT: first / last tile of a group that shares one histogram flush
  int32_t st_idx;                    // which prefetched TileState (PassSmem::stg) belongs to this tile
  uint32_t expected;                 // COUNT / WRITE_SCRATCH: chunks of this image that report to tiles_done
  int32_t rpi;                       // warp-autonomous gather: tile rows per warp step (32 / quads per row)
  int32_t n_steps;                   //   steps of the tile: ceil(th / rpi)
  uint32_t inv_qpr;                  //   ceil(1024 / quads per tile row): lane / d == (lane * inv) >> 10 (GATHER_SHARP: ceil(2^32 / halo width))
  uint32_t _pad0;
  int32_t _pad;
  const uint8_t* src;                // source image of this pass (the batch or a scratch image)
  uint8_t* dst;                      // destination image (unused by COUNT passes)
};

template <int C>
struct alignas(128) PassSmem {
  unsigned long long full[NU], empty[NU], stbar[NST];
  int32_t ctl[4];
  uint32_t color_cnt[CHB_MAX_CHAIN];
  uint32_t hist[MAXC][256];  // tile-local counts of a COUNT pass
  SlotInfo info[NU];
  TileState stg[NST];        // ring of prefetched TileStates, indexed by item number (see NST)
  alignas(128) uint8_t data[NU][UNIT_BYTES];
  alignas(128) uint8_t r[r_bytes(C)];
};

template <int C>
struct TC {
  const KParams* p;
  PassSmem<C>* sm;
  const TileState* t;
  const SlotInfo* info;
  uint32_t data;       // shared address of the slot's staged source bytes
  uint32_t ostage;     // shared address of this tile's output staging tile
  uint32_t* nstore;    // stores issued so far by this CTA's consumers (staging tiles alternate per store)
  uint32_t r;          // shared address of the R region
  const uint8_t* src;  // source image of this pass
  uint8_t* dst;        // destination image (unused by COUNT passes)
  int H, W, HW, img_bytes, tid, lane, tile;
  uint32_t l1a, l2a;   // shared addresses of the two LUTs
};

template <int C>
__device__ __forceinline__ void count_value(const TC<C>& c, int ch, uint32_t v) { atomicAdd(&c.sm->hist[ch][v & 255u], 1u); }

// Every consumer thread that issues bulk stores (warp 0) calls this before the barrier that
// precedes its next store: the previous store has then finished reading its staging tile, and
// since staging tiles alternate per store, the tile the next storing item writes is free once that
// barrier is passed.
__device__ __forceinline__ void stores_drained(int tid) {
  if (tid < 32) bulk_wait_read0();
}
// The barrier in front of a store: drain, sync, flip the staging tile for the next storing item.
template <int C>
__device__ __forceinline__ void store_barrier(const TC<C>& c) {
  fence_proxy_async();  // this thread's shared-memory writes -> visible to the TMA
  stores_drained(c.tid);
  cons_sync();
  ++(*c.nstore);
}

// =========================================================================== scalar executor
template <int C, bool COUNT>
__device__ void exec_generic(const TC<C>& c) {
  const TileState& t = *c.t;
  const int H = opaque_if<CHB_OPQ_GEN>(c.H), W = opaque_if<CHB_OPQ_GEN>(c.W);
  const int rx0 = c.info->x0, ry0 = c.info->y0;
  const int rw = c.info->x1 - rx0;
  const int n = rw * (c.info->y1 - ry0);
  const int kmode = t.kmode;
  const float f = t.kfactor;
  const uint8_t* src = c.src;
  for (int i = c.tid; i < n; i += NCONS) {
    const int ry = i / rw;
    const int y = ry0 + ry, x = rx0 + (i - ry * rw);
    int v[C];
    if (kmode == K_NONE || kmode == K_COLOR) {
      int sx = x, sy = y;
      const int k = resolve(t.sp, t.n_sp, H, W, sx, sy);
      if (k >= 0) {
        if (COUNT) { atomicAdd(&c.sm->color_cnt[k], 1u); continue; }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = t.sp[k].color[ch];
      } else {
        const uint8_t* px = src + ((size_t)sy * W + sx) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = ldg_pixel(px + ch);
        if (kmode == K_NONE) {
          if (!COUNT) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = t.l1[ch][v[ch]];
          }
        } else if (C == 3) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = t.l1[ch][v[ch]];
          uint32_t R, G, B;
          color_pixel_f((float)v[0], (float)v[1], (float)v[2], f, R, G, B);
          v[0] = (int)R; v[1] = (int)G; v[2] = (int)B;
          if (!COUNT) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = t.l2[ch][v[ch]];
          }
        }
        if (COUNT) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) count_value(c, ch, (uint32_t)v[ch]);
          continue;
        }
      }
    } else if (kmode == K_SHARP) {
      // tfa.image.sharpness (oracle/ops.py sharpness) on the virtual image spatial -> l1.
      const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
      const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
      auto vpre = [&](int xx, int yy, int ch) -> int {
        int sx = xx, sy = yy;
        const int k = resolve(t.sp, t.n_sp, H, W, sx, sy);
        if (k >= 0) return t.sp[k].color[ch];
        return t.l1[ch][ldg_pixel(src + ((size_t)sy * W + sx) * C + ch)];
      };
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        const int orig = vpre(x, y, ch);
        float deg = (float)orig;
        if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
          float acc = 0.0f;
          for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
              const float kk = (dy == 0 && dx == 0) ? k5 : k1;
              acc = __fadd_rn(acc, __fmul_rn((float)vpre(x + dx, y + dy, ch), kk));
            }
          deg = (float)(((int)acc) & 0xFF);
        }
        v[ch] = (int)sharp_blend(deg, (float)orig, f);
      }
      if (COUNT) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) count_value(c, ch, (uint32_t)v[ch]);
        continue;
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = t.l2[ch][v[ch]];
    } else {
      // bilinear warp (image_ops.h bilinear_interpolation; oracle/ops.py projective_transform) of l1(src).
      const Spatial& e = t.kgeo;
      const int fill = e.color[0];
      float sx, sy;
      affine_source(e.t, x, y, sx, sy);
      sx = map_coordinate(sx, W, e.fill_mode);
      sy = map_coordinate(sy, H, e.fill_mode);
      const float xf = floorf(sx), yf = floorf(sy);
      const float xc = __fadd_rn(xf, 1.0f), yc = __fadd_rn(yf, 1.0f);
      const float wxf = __fsub_rn(xc, sx), wxc = __fsub_rn(sx, xf);
      const float wyf = __fsub_rn(yc, sy), wyc = __fsub_rn(sy, yf);
      const bool inx0 = (xf >= 0.0f && xf < (float)W), inx1 = (xc >= 0.0f && xc < (float)W);
      const bool iny0 = (yf >= 0.0f && yf < (float)H), iny1 = (yc >= 0.0f && yc < (float)H);
      const int ix0 = inx0 ? (int)xf : 0, ix1 = inx1 ? (int)xc : 0;
      const int iy0 = iny0 ? (int)yf : 0, iy1 = iny1 ? (int)yc : 0;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        auto tap = [&](bool in, int iy, int ix) -> float {
          if (!in) return (float)fill;
          return (float)t.l1[ch][ldg_pixel(src + ((size_t)iy * W + ix) * C + ch)];
        };
        const float v00 = tap(iny0 && inx0, iy0, ix0), v01 = tap(iny0 && inx1, iy0, ix1);
        const float v10 = tap(iny1 && inx0, iy1, ix0), v11 = tap(iny1 && inx1, iy1, ix1);
        const float vyf = __fadd_rn(__fmul_rn(wxf, v00), __fmul_rn(wxc, v01));
        const float vyc = __fadd_rn(__fmul_rn(wxf, v10), __fmul_rn(wxc, v11));
        const float val = __fadd_rn(__fmul_rn(wyf, vyf), __fmul_rn(wyc, vyc));
        v[ch] = ((int)val) & 0xFF;
      }
      if (COUNT) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) count_value(c, ch, (uint32_t)v[ch]);
        continue;
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = t.l2[ch][v[ch]];
    }
    uint8_t* d = c.dst + ((size_t)y * W + x) * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) d[ch] = (uint8_t)v[ch];
  }
}

// ============================================================================== LUT helpers
// Maps the 4 bytes of w through per-channel 256-byte tables at shared addresses a0..a3 (the table
// of byte 0, 1, 2, 3).  All lanes of a warp use the same channel per lookup, so at most two words
// of a bank are in play (64 words per table): <= 2-way conflicts without replicating the table.
__device__ __forceinline__ uint32_t map_word(uint32_t w, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3) {
  uint32_t o = lds_u8(a0 + byte_of(w, 0));
  o = put_byte(o, lds_u8(a1 + byte_of(w, 1)), 1);
  o = put_byte(o, lds_u8(a2 + byte_of(w, 2)), 2);
  o = put_byte(o, lds_u8(a3 + byte_of(w, 3)), 3);
  return o;
}
// Unit of UW words whose byte 0 has channel 0: channel of byte b of word j is (4 j + b) % C.
template <int C, int UW>
__device__ __forceinline__ void map_unit(uint32_t* w, uint32_t lut) {
#pragma unroll
  for (int j = 0; j < UW; ++j)
    w[j] = map_word(w[j], lut + ((4 * j + 0) % C) * 256, lut + ((4 * j + 1) % C) * 256, lut + ((4 * j + 2) % C) * 256,
                    lut + ((4 * j + 3) % C) * 256);
}
// One uint4 whose byte 0 has channel ph (runtime).
template <int C>
__device__ __forceinline__ uint4 map_vec_phase(uint4 v, uint32_t lut, int ph) {
  uint32_t a[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = lut + (uint32_t)((ph + i) % C) * 256u;  // only a[0..C) distinct
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (C == 3)
      w[j] = map_word(w[j], a[(4 * j) % 3], a[(4 * j + 1) % 3], a[(4 * j + 2) % 3], a[(4 * j + 3) % 3]);
    else
      w[j] = map_word(w[j], a[0], a[1 % C], a[2 % C], a[3 % C]);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// =============================================================================== flat executor
// No spatial op pending, K in {none, Color}: the tile is a contiguous run of units (48 bytes = 16
// pixels for C == 3, else 16 bytes) whose channel phase is a compile-time constant.  The run sits in
// the slot (one TMA load); every thread takes one unit per round with LDS.128 (a quarter-warp's
// 48-byte-strided vectors fall into distinct 16-byte bank groups), transforms it in registers and
// writes it to the output staging tile, which leaves as one TMA store.
template <int C, bool COUNT>
__device__ void exec_flat(const TC<C>& c) {
  constexpr int UW = (C == 3) ? 12 : 4;
  constexpr int UB = UW * 4;
  const TileState& t = *c.t;
  const int u0 = c.info->x0, u1 = c.info->x1;
  const int nloc = u1 - u0;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF) * 0x01010101u, ac1 = (uint32_t)(t.l1_aff & 0xFF) * 0x01010101u;
  const float f = t.kfactor;
  // COUNT: eight copies of the u32 histogram in the R region, copy = lane & 7 (a value shared by a
  // whole warp is a 4-way same-address conflict at worst; copies are skewed by 4 banks so that the
  // same value in different copies falls in different banks).  Zeroed on the chunk's first tile,
  // summed into sm->hist on its last: 3 instructions per byte and nothing per tile, against 7 per
  // byte plus a zero / reduce pass per tile for the lane-private byte counters.
  constexpr uint32_t HC_COPY_BYTES = (uint32_t)C * 1024u + 16u;
  const uint32_t hl = c.r + (uint32_t)(c.lane & 7) * HC_COPY_BYTES;
  if (COUNT && c.info->first) {
    stores_drained(c.tid);  // the R region may still be feeding a store of the previous item
    cons_sync();
    for (uint32_t i = (uint32_t)c.tid * 16u; i < 8u * HC_COPY_BYTES; i += NCONS * 16u) sts_v4(c.r + i, make_uint4(0u, 0u, 0u, 0u));
    cons_sync();
  }
  for (int base = 0; base < nloc; base += NCONS) {  // one unit per thread per round
    const int u = base + c.tid;
    if (u < nloc) {
      uint32_t w[UW];
#pragma unroll
      for (int q = 0; q < UW / 4; ++q) {
        const uint4 v = lds_v4(c.data + u * UB + q * 16);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      if (kmode == K_NONE) {
        if (!COUNT && use1) {
          if (aff1) {
#pragma unroll
            for (int j = 0; j < UW; ++j) w[j] = (w[j] & am1) ^ ac1;
          } else {
            map_unit<C, UW>(w, c.l1a);
          }
        }
      } else if (C == 3) {
        if (use1) map_unit<C, UW>(w, c.l1a);
        uint32_t o[UW];
#pragma unroll
        for (int px = 0; px < 16; ++px) {
          const int b0 = px * 3, b1 = px * 3 + 1, b2 = px * 3 + 2;
          uint32_t R, G, B;
          color_pixel_f(byte_to_float(w[b0 >> 2], b0 & 3), byte_to_float(w[b1 >> 2], b1 & 3),
                        byte_to_float(w[b2 >> 2], b2 & 3), f, R, G, B);
          o[b0 >> 2] = ((b0 & 3) == 0) ? R : put_byte(o[b0 >> 2], R, b0 & 3);
          o[b1 >> 2] = ((b1 & 3) == 0) ? G : put_byte(o[b1 >> 2], G, b1 & 3);
          o[b2 >> 2] = ((b2 & 3) == 0) ? B : put_byte(o[b2 >> 2], B, b2 & 3);
        }
#pragma unroll
        for (int j = 0; j < UW; ++j) w[j] = o[j];
        if (!COUNT && use2) map_unit<C, UW>(w, c.l2a);
      }
      if (COUNT) {
#pragma unroll
        for (int j = 0; j < UW; ++j)
#pragma unroll
          for (int b = 0; b < 4; ++b) reds_add(hl + (uint32_t)(((4 * j + b) % C) * 1024) + (byte_of(w[j], b) << 2), 1u);
      } else {
#pragma unroll
        for (int q = 0; q < UW / 4; ++q)
          sts_v4(c.ostage + u * UB + q * 16, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
      }
    }
  }
  if (COUNT && c.info->last) {
    cons_sync();
    for (int i = c.tid; i < C * 256; i += NCONS) {
      uint32_t sum = 0;
#pragma unroll
      for (uint32_t k = 0; k < 8u; ++k) sum += lds_u32(c.r + k * HC_COPY_BYTES + (uint32_t)i * 4u);
      (&c.sm->hist[0][0])[i] += sum;
    }
    cons_sync();
  }
  if (!COUNT) {
    // CutOut rectangles (the spatial list holds nothing else in this class), in list order
    const int P0 = u0 * (UB / C), P1 = u1 * (UB / C);  // pixel range of this tile
    for (int k = 0; k < c.info->paint; ++k) {
      cons_sync();
      const Spatial& e = t.sp[k];
      const int rw = (e.x1 - e.x0) * C;
      if (rw <= 0 || nloc <= 0) continue;
      const int ylo = max(e.y0, P0 / c.W), yhi = min(e.y1, (P1 - 1) / c.W + 1);
      const int n = (yhi - ylo) * rw;
      for (int i = c.tid; i < n; i += NCONS) {
        const int ry = i / rw, rb = i - ry * rw;
        const int bidx = ((ylo + ry) * c.W + e.x0) * C + rb - P0 * C;  // byte index within the tile
        if (bidx >= 0 && bidx < nloc * UB)
          asm volatile("st.shared.u8 [%0], %1;" ::"r"(c.ostage + (uint32_t)bidx), "r"(e.color[rb % C]) : "memory");
      }
    }
    store_barrier(c);
    if (c.tid == 0 && nloc > 0) {
      bulk_store(c.dst + (size_t)u0 * UB, c.ostage, (uint32_t)nloc * UB);
      bulk_commit();
    }
  }
  // ragged tail of the image (whole pixels, fewer than one unit): last tile, one pixel per thread
  if (c.tile == c.p->n_flat_tiles - 1) {
    const int p0 = (c.img_bytes / UB) * UB / C;
    for (int pix = p0 + c.tid; pix < c.HW; pix += NCONS) {
      int v[C];
      if (!COUNT && c.info->paint) {
        int sx = pix % c.W, sy = pix / c.W;
        const int k = resolve(t.sp, t.n_sp, c.H, c.W, sx, sy);
        if (k >= 0) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) c.dst[(size_t)pix * C + ch] = (uint8_t)t.sp[k].color[ch];
          continue;
        }
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = ldg_pixel(c.src + (size_t)pix * C + ch);
      if (kmode == K_NONE) {
        if (!COUNT && use1) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = t.l1[ch][v[ch]];
        }
      } else if (C == 3) {
        if (use1) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = t.l1[ch][v[ch]];
        }
        uint32_t R, G, B;
        color_pixel_f((float)v[0], (float)v[1], (float)v[2], f, R, G, B);
        v[0] = (int)R; v[1] = (int)G; v[2] = (int)B;
        if (!COUNT && use2) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = t.l2[ch][v[ch]];
        }
      }
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        if (COUNT) count_value(c, ch, (uint32_t)v[ch]);
        else c.dst[(size_t)pix * C + ch] = (uint8_t)v[ch];
      }
    }
  }
}

// ============================================================================= gather executor
// round(v) (half away from zero) as a source index, together with its bounds test 0 <= round(v) < n,
// in five instructions.  fl_rz(v + (2^22 + 0.5)) lands in [2^22, 2^23) exactly when v + 0.5 is in
// [0, 2^22); the ulp there is 0.5 and round-toward-zero cuts v + 0.5 down to a multiple of 0.5, so
// the mantissa >> 1 is floor(v + 0.5) = round(v) for v >= 0.  -0.5 < v < 0 rounds to -0 (index 0,
// like std::round); v == -0.5 rounds away to -1 and is excluded by the explicit test; anything
// smaller or >= 2^22 (or NaN) produces an index >= n.  Needs n < 2^22 (checked on the host).
__device__ __forceinline__ bool src_index(float v, int n, int& idx) {
  const uint32_t u = __float_as_uint(__fadd_rz(v, 4194304.5f));
  idx = (int)((u - 0x4A800000u) >> 1);
  return (v > -0.5f) && ((uint32_t)idx < (uint32_t)n);
}

// Division-free walk over the 4-pixel quads of a tw x th tile: thread t starts at quad t and
// advances by NCONS quads per round.
struct QuadWalk {
  int rq, ry, drq, dry, qpr;
  __device__ __forceinline__ QuadWalk(int tid, int tw) {
    qpr = tw >> 2;
    ry = tid / qpr; rq = tid - ry * qpr;
    dry = NCONS / qpr; drq = NCONS - dry * qpr;
  }
  __device__ __forceinline__ void next() {
    rq += drq; ry += dry;
    if (rq >= qpr) { rq -= qpr; ++ry; }
  }
};

// Spatial list pending (constant-fill nearest warps and masks), K in {none, Color}.  The producer
// has staged the source bounding box of the tile (one tensor-map box); pixels are gathered from it
// four at a time into the output staging tile.
//
// Fast form: the list holds ONE or TWO entries (all a RandAugment(N=2) chain can produce).  Entry
// "a" = sp[n_sp-1] was applied last and is evaluated first on the output coordinate; if it is a warp,
// the x-dependent products of the four pixels of a thread's quad column stay in registers for the
// whole tile (a thread keeps its column when the quads per row divide the thread count), so its
// coordinates cost four float adds plus the 5-instruction index/bounds step each.  Entry "b" =
// sp[n_sp-2], if any, is evaluated on the integer coordinate that "a" produced.
template <int C, bool COUNT, bool TWO>
__device__ void gather_fast(const TC<C>& c) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  const int H = c.H, W = c.W;
  const int tw = in.x1 - in.x0, th = in.y1 - in.y0;
  const int n_sp = t.n_sp;
  const Spatial& ea = t.sp[n_sp - 1];
  const Spatial& eb = t.sp[TWO ? n_sp - 2 : 0];
  constexpr bool two = TWO;
  const bool a_geom = !TWO || ea.type == SP_GEOM;  // the one-entry form is only used for a warp
  const bool b_geom = TWO && eb.type == SP_GEOM;
  const float t0 = ea.t[0], t1 = ea.t[1], t2 = ea.t[2], t3 = ea.t[3], t4 = ea.t[4], t5 = ea.t[5];
  const float u0 = eb.t[0], u1 = eb.t[1], u2 = eb.t[2], u3 = eb.t[3], u4 = eb.t[4], u5 = eb.t[5];
  const int ay0 = ea.y0, ay1 = ea.y1, ax0 = ea.x0, ax1 = ea.x1;
  const int by0m = eb.y0, by1m = eb.y1, bx0m = eb.x0, bx1m = eb.x1;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool plain = (kmode == K_NONE) && (COUNT || !use1);  // staged bytes are the result
  const float f = t.kfactor;
  const uint32_t fill_a = in.fillc[0], fill_b = in.fillc[1];
  const uint32_t fa_addr = smem_addr(&in.fillc[0]), fb_addr = smem_addr(&in.fillc[1]);  // a miss is just another address
  const int pitch = in.pitch;
  // staged byte of source pixel (ix, iy), channel ch: base0 + iy * pitch + ix * C + ch
  const uint32_t base0 = c.data - (uint32_t)(in.by0 * pitch + in.bxb0);
  QuadWalk q(c.tid, tw);
  const bool fixed_col = (q.drq == 0);
  float ax[4], bx[4];
  auto xterms = [&](int rq) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float fx = small_uint_to_float((uint32_t)(in.x0 + (rq << 2) + i));
      ax[i] = __fmul_rn(t0, fx);
      bx[i] = __fmul_rn(t3, fx);
    }
  };
  xterms(q.rq);
  uint32_t n_fill_a = 0, n_fill_b = 0;
  for (; q.ry < th; q.next()) {
    if (!fixed_col) xterms(q.rq);
    const int y = in.y0 + q.ry;
    const int xq = in.x0 + (q.rq << 2);
    const float fy = small_uint_to_float((uint32_t)y);
    const float t1y = __fmul_rn(t1, fy), t4y = __fmul_rn(t4, fy);
    uint32_t o[C];
#pragma unroll
    for (int w = 0; w < C; ++w) o[w] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int ix = xq + i, iy = y;
      bool hit_a, hit_b = false;  // the pixel shows the colour of entry a / b
      if (a_geom) {
        int jx, jy;
        const bool inx = src_index(__fadd_rn(__fadd_rn(ax[i], t1y), t2), W, jx);
        const bool iny = src_index(__fadd_rn(__fadd_rn(bx[i], t4y), t5), H, jy);
        hit_a = !(inx && iny);
        ix = jx; iy = jy;
      } else {
        hit_a = (iy >= ay0) && (iy < ay1) && (ix >= ax0) && (ix < ax1);
      }
      if (two && !hit_a) {
        if (b_geom) {
          const float gx = small_uint_to_float((uint32_t)ix), gy = small_uint_to_float((uint32_t)iy);
          int jx, jy;
          const bool inx = src_index(__fadd_rn(__fadd_rn(__fmul_rn(u0, gx), __fmul_rn(u1, gy)), u2), W, jx);
          const bool iny = src_index(__fadd_rn(__fadd_rn(__fmul_rn(u3, gx), __fmul_rn(u4, gy)), u5), H, jy);
          hit_b = !(inx && iny);
          ix = jx; iy = jy;
        } else {
          hit_b = (iy >= by0m) && (iy < by1m) && (ix >= bx0m) && (ix < bx1m);
        }
      }
      const bool inside = !(hit_a || hit_b);
      const uint32_t a = inside ? base0 + (uint32_t)(iy * pitch + ix * C) : (hit_a ? fa_addr : fb_addr);
      uint32_t v[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
      if (!plain) {
        if (kmode == K_NONE) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
        } else if (C == 3) {
          if (use1) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
          }
          color_pixel_f(small_uint_to_float(v[0]), small_uint_to_float(v[1]), small_uint_to_float(v[2]), f, v[0], v[1], v[2]);
          if (!COUNT && use2) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l2a + ch * 256 + v[ch]);
          }
        }
        if (!COUNT) {
          const uint32_t fillv = hit_a ? fill_a : fill_b;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = inside ? v[ch] : byte_of(fillv, ch);
        }
      }
      if (COUNT) {
        if (inside) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) count_value(c, ch, v[ch]);
        } else if (hit_a) {
          ++n_fill_a;
        } else {
          ++n_fill_b;
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int bi = i * C + ch;
          o[bi >> 2] = ((bi & 3) == 0) ? (v[ch] & 255u) : put_byte(o[bi >> 2], v[ch], bi & 3);
        }
      }
    }
    if (!COUNT) {
      const uint32_t oa = c.ostage + (uint32_t)((q.ry * tw + (q.rq << 2)) * C);
#pragma unroll
      for (int w = 0; w < C; ++w) sts_u32(oa + 4 * w, o[w]);
    }
  }
  if (COUNT) {
    if (n_fill_a) atomicAdd(&c.sm->color_cnt[n_sp - 1], n_fill_a);
    if (TWO && n_fill_b) atomicAdd(&c.sm->color_cnt[n_sp - 2], n_fill_b);
  }
}

// ------------------------------------------------------------------ warp-autonomous executors
// WRITE passes of gather tiles need no CTA-wide barrier: every consumer warp takes its own rows of
// the tile, stores them itself and moves on to the next tile on its own.  (The same was tried for
// flat runs three ways -- each warp storing its 1.5 KB slice with its own TMA store, in place or
// through private staging, or straight from registers with 16-byte vector stores -- and lost 7-25 %
// against one 12 KB TMA store per tile behind a CTA barrier: profiles/r01_v10, r01_v13.)

// Gather tile with one or two spatial entries, WRITE pass (the per-pixel arithmetic is gather_fast's).
// A warp takes whole rows of the tile: 4 pixels (C words) per lane, 32 / (tw / 4) rows per step, steps
// 8 apart; the words go straight to global memory.  PLAIN: the staged bytes are the result (no LUT,
// no Color), so the 4 x C byte loads of a quad issue back to back.
template <int C, bool TWO, bool PLAIN>
__device__ __forceinline__ void gather_warp(const TC<C>& c, int warp) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  const int H = c.H, W = c.W;
  const int tw = in.x1 - in.x0, th = in.y1 - in.y0;
  const int n_sp = t.n_sp;
  const Spatial& ea = t.sp[n_sp - 1];
  const Spatial& eb = t.sp[TWO ? n_sp - 2 : 0];
  const bool a_geom = !TWO || ea.type == SP_GEOM;
  const bool b_geom = TWO && eb.type == SP_GEOM;
  const float t0 = ea.t[0], t1 = ea.t[1], t2 = ea.t[2], t3 = ea.t[3], t4 = ea.t[4], t5 = ea.t[5];
  const float u0 = eb.t[0], u1 = eb.t[1], u2 = eb.t[2], u3 = eb.t[3], u4 = eb.t[4], u5 = eb.t[5];
  const int ay0 = ea.y0, ay1 = ea.y1, ax0 = ea.x0, ax1 = ea.x1;
  const int by0m = eb.y0, by1m = eb.y1, bx0m = eb.x0, bx1m = eb.x1;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF), ac1 = (uint32_t)(t.l1_aff & 0xFF);
  const float f = t.kfactor;
  const uint32_t fill_a = in.fillc[0], fill_b = in.fillc[1];
  const uint32_t fa_addr = smem_addr(&in.fillc[0]), fb_addr = smem_addr(&in.fillc[1]);  // a miss is just another address
  const int pitch = in.pitch;
  const uint32_t base0 = c.data - (uint32_t)(in.by0 * pitch + in.bxb0);
  const int qpr = tw >> 2;              // quads per row: 4, 8, 12 or 16 for C == 3 (rows are whole 16-byte vectors)
  const int rpi = in.rpi;               // rows per step: 32 / qpr
  // lane -> (row of the step, quad of the row) and -> (row, vector) of the store, without divisions:
  // lane / d == (lane * ceil(1024 / d)) >> 10 for lane < 32, d <= 32 (the producer computed the factors)
  const int sub = (int)(((uint32_t)c.lane * in.inv_qpr) >> 10), rq = c.lane - sub * qpr;
  const bool active = sub < rpi;
  const int xq = in.x0 + (rq << 2);
  float ax[4], bx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float fx = small_uint_to_float((uint32_t)(xq + i));
    ax[i] = __fmul_rn(t0, fx);
    bx[i] = __fmul_rn(t3, fx);
  }
  const int n_steps = in.n_steps;
  for (int g = warp; g < n_steps; g += NCONS / 32) {
    const int ry = g * rpi + sub;
    if (active && ry < th) {
      const int y = in.y0 + ry;
      const float fy = small_uint_to_float((uint32_t)y);
      const float t1y = __fmul_rn(t1, fy), t4y = __fmul_rn(t4, fy);
      uint32_t adr[4];
      uint32_t hit = 0;  // bit i: pixel i shows entry a's colour, bit 4 + i: entry b's
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int ix = xq + i, iy = y;
        bool hit_a, hit_b = false;
        if (a_geom) {
          int jx, jy;
          const bool inx = src_index(__fadd_rn(__fadd_rn(ax[i], t1y), t2), W, jx);
          const bool iny = src_index(__fadd_rn(__fadd_rn(bx[i], t4y), t5), H, jy);
          hit_a = !(inx && iny);
          ix = jx; iy = jy;
        } else {
          hit_a = (iy >= ay0) && (iy < ay1) && (ix >= ax0) && (ix < ax1);
        }
        if (TWO && !hit_a) {
          if (b_geom) {
            const float gx = small_uint_to_float((uint32_t)ix), gy = small_uint_to_float((uint32_t)iy);
            int jx, jy;
            const bool inx = src_index(__fadd_rn(__fadd_rn(__fmul_rn(u0, gx), __fmul_rn(u1, gy)), u2), W, jx);
            const bool iny = src_index(__fadd_rn(__fadd_rn(__fmul_rn(u3, gx), __fmul_rn(u4, gy)), u5), H, jy);
            hit_b = !(inx && iny);
            ix = jx; iy = jy;
          } else {
            hit_b = (iy >= by0m) && (iy < by1m) && (ix >= bx0m) && (ix < bx1m);
          }
        }
        adr[i] = !(hit_a || hit_b) ? base0 + (uint32_t)(iy * pitch + ix * C) : (hit_a ? fa_addr : fb_addr);
        if (!PLAIN) hit |= (hit_a ? 1u : 0u) << i | (hit_b ? 16u : 0u) << i;
      }
      uint32_t v[4][C];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(adr[i] + ch);
      if (!PLAIN) {
        if (kmode == K_NONE) {
          if (aff1) {  // (x & m) ^ c, the same for every channel: no lookups
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[i][ch] = (v[i][ch] & am1) ^ ac1;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(c.l1a + ch * 256 + v[i][ch]);
          }
        } else if (C == 3) {
          if (use1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(c.l1a + ch * 256 + v[i][ch]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            color_pixel_f(small_uint_to_float(v[i][0]), small_uint_to_float(v[i][1 % C]), small_uint_to_float(v[i][2 % C]), f,
                          v[i][0], v[i][1 % C], v[i][2 % C]);
          if (use2) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(c.l2a + ch * 256 + v[i][ch]);
          }
        }
        // the spatial colours already are final colours
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ha = (hit >> i) & 1u, hb = (hit >> (4 + i)) & 1u;
          const uint32_t fillv = ha ? fill_a : fill_b;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[i][ch] = (ha || hb) ? byte_of(fillv, ch) : v[i][ch];
        }
      }
      uint32_t o[C];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int bi = i * C + ch;
          o[bi >> 2] = ((bi & 3) == 0) ? (v[i][ch] & 255u) : put_byte(o[bi >> 2], v[i][ch], bi & 3);
        }
      // the lane's 4 pixels (C words) go straight to global memory: against staging the step's rows in
      // shared memory and storing 16-byte vectors this was 4 % faster (no __syncwarp between steps)
      uint32_t* gp = reinterpret_cast<uint32_t*>(c.dst + ((size_t)y * W + xq) * C);
#pragma unroll
      for (int w = 0; w < C; ++w) __stcg(gp + w, o[w]);
    }
  }
}

template <int C>
__device__ __forceinline__ void exec_gather_warp(const TC<C>& c, int warp) {
  const TileState& t = *c.t;
  const bool plain = (t.kmode == K_NONE) && t.l1_id;
  if (t.n_sp == 2) {
    if (plain) gather_warp<C, true, true>(c, warp); else gather_warp<C, true, false>(c, warp);
  } else {
    if (plain) gather_warp<C, false, true>(c, warp); else gather_warp<C, false, false>(c, warp);
  }
}

// General form: any list of constant-fill warps and masks.
template <int C, bool COUNT>
__device__ void gather_list(const TC<C>& c) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  const int H = opaque_if<CHB_OPQ_LIST>(c.H), W = opaque_if<CHB_OPQ_LIST>(c.W);
  const int n_sp = t.n_sp;
  const int bx0 = in.bx0, by0 = in.by0;
  const uint32_t box_w = (uint32_t)(in.bx1 - bx0), box_h = (uint32_t)(in.by1 - by0);
  const bool box_empty = in.rows <= 0;
  const int tw = in.x1 - in.x0, th = in.y1 - in.y0;
  const int pitch = in.pitch;
  const uint32_t base0 = c.data - (uint32_t)(by0 * pitch + in.bxb0);
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  for (QuadWalk q(c.tid, tw); q.ry < th; q.next()) {
    const int y = in.y0 + q.ry;
    int x = in.x0 + (q.rq << 2);
    uint32_t o[C];
#pragma unroll
    for (int w = 0; w < C; ++w) o[w] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i, ++x) {
      int sx = x, sy = y;
      const int k = resolve(t.sp, n_sp, H, W, sx, sy);
      uint32_t v[C];
      if (k < 0) {
        if (!box_empty && (uint32_t)(sx - bx0) <= box_w && (uint32_t)(sy - by0) <= box_h) {
          const uint32_t a = base0 + (uint32_t)(sy * pitch + sx * C);
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
        } else {  // never taken if the box is right; keeps a box error from becoming a wrong pixel
          const uint8_t* px = c.src + ((size_t)sy * W + sx) * C;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = ldg_pixel(px + ch);
        }
        if (kmode == K_NONE) {
          if (!COUNT && use1) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
          }
        } else if (C == 3) {
          if (use1) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
          }
          color_pixel_f(small_uint_to_float(v[0]), small_uint_to_float(v[1]), small_uint_to_float(v[2]), f, v[0], v[1], v[2]);
          if (!COUNT && use2) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l2a + ch * 256 + v[ch]);
          }
        }
        if (COUNT) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) count_value(c, ch, v[ch]);
        }
      } else {
        if (COUNT) {
          atomicAdd(&c.sm->color_cnt[k], 1u);
        } else {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = (uint32_t)t.sp[k].color[ch];
        }
      }
      if (!COUNT) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int bi = i * C + ch;
          o[bi >> 2] = ((bi & 3) == 0) ? (v[ch] & 255u) : put_byte(o[bi >> 2], v[ch], bi & 3);
        }
      }
    }
    if (!COUNT) {
      const uint32_t oa = c.ostage + (uint32_t)((q.ry * tw + (q.rq << 2)) * C);
#pragma unroll
      for (int w = 0; w < C; ++w) sts_u32(oa + 4 * w, o[w]);
    }
  }
}

// Copies a tw x th output staging tile (top-left pixel (x0, y0)) to the image with 16-byte vectors:
// a row of the tile is a whole number of them, a warp writes 512 contiguous-by-row bytes per store.
// (One TMA bulk store per row was measured first: 64 small stores per tile drained so slowly that
// the wait before reusing the staging tile became the top stall, profiles/r01_v5.)
template <int C>
__device__ __forceinline__ void store_tile_rows(const TC<C>& c, uint32_t ostage, int x0, int y0, int tw, int th) {
  const int rowbytes = c.W * C;
  uint8_t* dbase = c.dst + ((size_t)y0 * c.W + x0) * C;
  const int rb = tw * C;
  const int n16 = rb >> 4;
  int rr = c.tid / n16, q = c.tid - rr * n16;
  const int drr = NCONS / n16, dq = NCONS - drr * n16;
  for (; rr < th;) {
    *reinterpret_cast<uint4*>(dbase + (size_t)rr * rowbytes + (q << 4)) = lds_v4(ostage + rr * rb + (q << 4));
    q += dq; rr += drr;
    if (q >= n16) { q -= n16; ++rr; }
  }
}

template <int C, bool COUNT>
__device__ void exec_gather(const TC<C>& c) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  // (WRITE passes of the one- and two-entry forms run warp-autonomously: exec_gather_warp)
  if (COUNT && t.n_sp == 1 && t.sp[0].type == SP_GEOM) gather_fast<C, true, false>(c);
  else if (COUNT && t.n_sp == 2) gather_fast<C, true, true>(c);
  else gather_list<C, COUNT>(c);
  if (!COUNT) {
    store_barrier(c);  // also flips the staging tile: the next item writes the other one
    store_tile_rows(c, c.ostage, in.x0, in.y0, in.x1 - in.x0, in.y1 - in.y0);
  }
}

// ========================================================================== sharpness executors
// tfa.image.sharpness (oracle/ops.py sharpness) on rows held in shared memory.  One thread owns one
// word column (4 output bytes) of a run of rows and walks down it with the float32 products
// (value x 1/13) of the 4 + 2C-byte windows of the previous, current and next row in registers, so
// every input byte is converted and multiplied once per column instead of nine times.  The row
// loop is unrolled by three so that the three windows rotate by renaming, and border bytes are a
// select, not a branch: the walk is straight-line float adds in the oracle's row-major order.
//   col      shared address of this column's word in the row ABOVE the first output row
//   pitch    bytes between rows
//   n        output rows
//   has_prev / has_next   the words left / right of the column exist (they hold the neighbours)
//   bmask    bit b set: byte b of the column lies in the first / last pixel of the IMAGE row
//   emit(r, word)         receives the sharpened word of output row r (0-based)
template <int C, class Emit>
__device__ __forceinline__ void sharp_walk(uint32_t col, int pitch, int n, bool has_prev, bool has_next, uint32_t bmask,
                                           float f, Emit emit) {
  constexpr int NB = 4 + 2 * C;  // window bytes per row
  const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
  const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
  float pa[NB], pb[NB], pc[NB];  // products of three consecutive rows
  float ca[4], cb[4], cc[4];     // float values of their centre bytes
  auto load_row = [&](uint32_t ra, float* pr, float* cv) {
    const uint32_t w1 = lds_u32(ra);
    const uint32_t w0 = has_prev ? lds_u32(ra - 4) : 0u;
    const uint32_t w2 = has_next ? lds_u32(ra + 4) : 0u;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int wb = 4 - C + j;  // byte index in the 12-byte (w0, w1, w2) window
      const uint32_t wsel = (wb < 4) ? w0 : (wb < 8) ? w1 : w2;
      const float fv = byte_to_float(wsel, wb & 3);
      pr[j] = __fmul_rn(fv, k1);
      if (j >= C && j < C + 4) cv[j - C] = fv;
    }
  };
  // rows above / at / below: window index b + C is the byte itself, b and b + 2C its left / right neighbours
  auto step = [&](const float* up, const float* mid, const float* midc, const float* dn, int r) {
    uint32_t o = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float orig = midc[b];
      float acc = up[b];
      acc = __fadd_rn(acc, up[b + C]);
      acc = __fadd_rn(acc, up[b + 2 * C]);
      acc = __fadd_rn(acc, mid[b]);
      acc = __fadd_rn(acc, __fmul_rn(orig, k5));
      acc = __fadd_rn(acc, mid[b + 2 * C]);
      acc = __fadd_rn(acc, dn[b]);
      acc = __fadd_rn(acc, dn[b + C]);
      acc = __fadd_rn(acc, dn[b + 2 * C]);
      float deg = __fadd_rn(__fadd_rz(acc, 8388608.0f), -8388608.0f);  // float(trunc(acc))
      deg = ((bmask >> b) & 1u) ? orig : deg;                           // border pixels keep the original
      const uint32_t res = sharp_blend(deg, orig, f);
      o = (b == 0) ? res : put_byte(o, res, b);
    }
    emit(r, o);
  };
  load_row(col, pa, ca);
  load_row(col + pitch, pb, cb);
  uint32_t ra = col + 2 * pitch;
  for (int r = 0; r < n; r += 3) {
    load_row(ra, pc, cc);
    step(pa, pb, cb, pc, r);
    if (r + 1 < n) {
      load_row(ra + pitch, pa, ca);
      step(pb, pc, cc, pa, r + 1);
    }
    if (r + 2 < n) {
      load_row(ra + 2 * pitch, pb, cb);
      step(pc, pa, ca, pb, r + 2);
    }
    ra += 3 * pitch;
  }
}

// Splits `inner` rows into sub-strips so that (columns x sub-strips) keeps every thread busy:
// minimise rounds * (rows per sub-strip + 2 halo rows).  Returns rows per sub-strip.  Runs once per
// tile on the producer warp (the result travels in SlotInfo::sharp_rows).
__device__ __forceinline__ int sharp_split(int columns, int inner) {
  int best_s = 1, best_cost = 0x7FFFFFFF;
  for (int S = 1; S <= 8 && S <= inner; ++S) {
    const int rounds = (columns * S + NCONS - 1) / NCONS;
    const int cost = rounds * ((inner + S - 1) / S + 2);
    if (cost < best_cost) { best_cost = cost; best_s = S; }
  }
  return (inner + best_s - 1) / best_s;
}

// Sharpness of a gathered image (K == Sharpness with a spatial list pending): the tile's virtual
// pre-image (spatial list -> l1) is gathered with a one-pixel halo into shared memory, half a tile
// at a time, and sharpened from there with the column walk.
// R region: [halo tile | output half 0 | output half 1].  A halo row is laid out so that the
// tile's first pixel starts on a word: 4 - C pad bytes, the left halo pixel, the tile, the right halo.
template <int C, bool COUNT>
__device__ void exec_gather_sharp(const TC<C>& c) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  const int H = opaque_if<CHB_OPQ_GS>(c.H), W = opaque_if<CHB_OPQ_GS>(c.W);
  const int n_sp = t.n_sp;
  const int bx0 = in.bx0, by0 = in.by0;
  const uint32_t box_w = (uint32_t)(in.bx1 - bx0), box_h = (uint32_t)(in.by1 - by0);
  const bool box_empty = in.rows <= 0;
  const int pitch = in.pitch;
  const uint32_t base0 = c.data - (uint32_t)(by0 * pitch + in.bxb0);
  const bool use1 = !t.l1_id, use2 = !t.l2_id;  // identity LUTs are not even copied into the unit
  const float f = t.kfactor;
  const int tw = in.x1 - in.x0, th = in.y1 - in.y0;
  const int hh = (th + 1) >> 1;              // rows per half
  const int vw = tw + 2;                     // halo tile width in pixels
  const int vpitch = (4 + (tw + 1) * C + 3) & ~3;
  const uint32_t vbuf = c.r;
  const uint32_t vbytes = (uint32_t)(((hh + 2) * vpitch + 15) & ~15);
  const uint32_t obytes = (uint32_t)(hh * tw * C);
  const int wpt = (tw * C) >> 2;             // output words per tile row
  const uint32_t inv_vw = in.inv_qpr;        // ceil(2^32 / vw), from the producer
  const bool one = (n_sp == 1) && !box_empty;
  const Spatial& e0 = t.sp[0];
  const bool e_geom = e0.type == SP_GEOM;
  const float g0 = e0.t[0], g1 = e0.t[1], g2 = e0.t[2], g3 = e0.t[3], g4 = e0.t[4], g5 = e0.t[5];
  const int my0 = e0.y0, my1 = e0.y1, mx0 = e0.x0, mx1 = e0.x1;
  const uint32_t fill_addr = smem_addr(&in.fillc[0]);  // the (final) colour bytes of the entry
  stores_drained(c.tid);  // the whole R region is used here
  cons_sync();
  for (int half = 0; half < 2; ++half) {
    const int ya = in.y0 + half * hh, yb = min(in.y1, ya + hh);
    if (yb <= ya) break;
    // phase 1: virtual pre-image on [x0-1, x1] x [ya-1, yb]
    const int nv = (yb - ya + 2) * vw;
    for (int i = c.tid; i < nv; i += NCONS) {
      const int vy = (int)__umulhi((uint32_t)i, inv_vw), vx = i - vy * vw;  // i / vw, exact for i < 2^16
      const int y = ya - 1 + vy, x = in.x0 - 1 + vx;
      uint32_t v[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = 0;
      if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
        if (one) {
          // one constant-fill entry (all a two-op chain can put in front of Sharpness): the exact
          // index step of the gather tiles; the staged box covers every source pixel
          int sx = x, sy = y;
          bool hit;
          if (e_geom) {
            const float fx = small_uint_to_float((uint32_t)x), fy = small_uint_to_float((uint32_t)y);
            const bool inx = src_index(__fadd_rn(__fadd_rn(__fmul_rn(g0, fx), __fmul_rn(g1, fy)), g2), W, sx);
            const bool iny = src_index(__fadd_rn(__fadd_rn(__fmul_rn(g3, fx), __fmul_rn(g4, fy)), g5), H, sy);
            hit = !(inx && iny);
          } else {
            hit = (y >= my0) && (y < my1) && (x >= mx0) && (x < mx1);
          }
          const uint32_t a = hit ? fill_addr : base0 + (uint32_t)(sy * pitch + sx * C);
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
          if (use1 && !hit) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
          }
        } else {
          int sx = x, sy = y;
          const int k = resolve(t.sp, n_sp, H, W, sx, sy);
          if (k < 0) {
            if (!box_empty && (uint32_t)(sx - bx0) <= box_w && (uint32_t)(sy - by0) <= box_h) {
              const uint32_t a = base0 + (uint32_t)(sy * pitch + sx * C);
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
            } else {
              const uint8_t* px = c.src + ((size_t)sy * W + sx) * C;
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[ch] = ldg_pixel(px + ch);
            }
            if (use1) {
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
            }
          } else {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = (uint32_t)t.sp[k].color[ch];
          }
        }
      }
      const uint32_t va = vbuf + (uint32_t)(vy * vpitch + 4 - C + vx * C);
#pragma unroll
      for (int ch = 0; ch < C; ++ch) asm volatile("st.shared.u8 [%0], %1;" ::"r"(va + ch), "r"(v[ch]) : "memory");
    }
    cons_sync();
    // phase 2: sharpen tw x (yb - ya) pixels out of the halo tile
    const uint32_t obuf = c.r + vbytes + half * obytes;
    auto emit_row = [&](int y, int xw, uint32_t o) {  // y: image row; xw: word column of the tile
      const int ph = (C == 3) ? (xw % 3) : 0;         // tile rows start at channel 0 (x0 * C is a multiple of 4 * C ... of 3)
      if (COUNT) {
#pragma unroll
        for (int b = 0; b < 4; ++b) count_value(c, (ph + b) % C, byte_of(o, b));
      } else {
        if (use2)
          o = map_word(o, c.l2a + (uint32_t)((ph + 0) % C) * 256u, c.l2a + (uint32_t)((ph + 1) % C) * 256u,
                       c.l2a + (uint32_t)((ph + 2) % C) * 256u, c.l2a + (uint32_t)((ph + 3) % C) * 256u);
        sts_u32(obuf + (uint32_t)((y - ya) * (tw * C) + (xw << 2)), o);
      }
    };
    // image rows 0 and H-1 keep the original
    for (int y = ya; y < yb; ++y)
      if (y == 0 || y == H - 1)
        for (int xw = c.tid; xw < wpt; xw += NCONS) emit_row(y, xw, lds_u32(vbuf + (uint32_t)((y - ya + 1) * vpitch + 4 + (xw << 2))));
    const int in0 = max(ya, 1), in1 = min(yb, H - 1);
    const int inner = in1 - in0;
    if (inner > 0) {
      const int R = in.sharp_rows;
      const int n_strips = (inner + R - 1) / R;
      const int n_items = wpt * n_strips;
      for (int item = c.tid; item < n_items; item += NCONS) {
        const int strip = item / wpt, xw = item - strip * wpt;
        const int y_begin = in0 + strip * R, y_end = min(in1, y_begin + R);
        uint32_t bmask = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xb = in.x0 * C + (xw << 2) + b;  // byte index in the image row
          if (xb < C || xb >= W * C - C) bmask |= 1u << b;
        }
        const uint32_t col = vbuf + (uint32_t)((y_begin - 1 - (ya - 1)) * vpitch + 4 + (xw << 2));
        sharp_walk<C>(col, vpitch, y_end - y_begin, true, true, bmask, f,
                      [&](int r, uint32_t o) { emit_row(y_begin + r, xw, o); });
      }
    }
    cons_sync();
    if (!COUNT) store_tile_rows(c, obuf, in.x0, ya, tw, yb - ya);
  }
  cons_sync();  // the next item's staging tile overlaps these buffers
}

// K == Sharpness with no spatial op pending: the tile is a strip of whole rows; the strip plus one
// halo row each side sits in the unit(s) (one TMA load).  l1 is applied to the staged rows in place,
// the column walk writes the output staging tile, which leaves as one TMA store (a strip is
// contiguous in the image).
template <int C, bool COUNT>
__device__ void exec_sharp(const TC<C>& c) {
  const TileState& t = *c.t;
  const SlotInfo& in = *c.info;
  const int H = opaque_if<CHB_OPQ_SHARP>(c.H);
  const int row = opaque_if<CHB_OPQ_SHARP>(c.W) * C;
  const int sr0 = in.by0, nrows = in.rows;
  const int y0 = in.y0, y1 = in.y1;
  const uint32_t stage = c.data;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  if (nrows <= 0 || y1 <= y0) return;
  if (use1) {  // byte 16 i of the image has channel (16 i) % C
    const int total = (nrows * row) >> 4;
    const int i0 = (sr0 * row) >> 4;
    for (int i = c.tid; i < total; i += NCONS) {
      const uint4 v = map_vec_phase<C>(lds_v4(stage + (i << 4)), c.l1a, (C == 3) ? ((i0 + i) % 3) : 0);
      sts_v4(stage + (i << 4), v);
    }
    cons_sync();
  }
  const int wpr = row >> 2;  // words per row
  auto emit = [&](int y, int xw, int ph, uint32_t o) {  // ph: channel of byte 0 of word xw
    if (COUNT) {
#pragma unroll
      for (int b = 0; b < 4; ++b) count_value(c, (ph + b) % C, byte_of(o, b));
    } else {
      if (use2)
        o = map_word(o, c.l2a + (uint32_t)((ph + 0) % C) * 256u, c.l2a + (uint32_t)((ph + 1) % C) * 256u,
                     c.l2a + (uint32_t)((ph + 2) % C) * 256u, c.l2a + (uint32_t)((ph + 3) % C) * 256u);
      sts_u32(c.ostage + (uint32_t)((y - y0) * row + (xw << 2)), o);
    }
  };
  // first and last image row: every pixel is border -> blend(orig, orig) == orig
  if (y0 == 0)
    for (int xw = c.tid; xw < wpr; xw += NCONS)
      emit(0, xw, (C == 3) ? (xw % 3) : 0, lds_u32(stage + (uint32_t)((0 - sr0) * row + (xw << 2))));
  if (H > 1 && y0 <= H - 1 && y1 > H - 1)
    for (int xw = c.tid; xw < wpr; xw += NCONS)
      emit(H - 1, xw, (C == 3) ? (xw % 3) : 0, lds_u32(stage + (uint32_t)((H - 1 - sr0) * row + (xw << 2))));
  const int in0 = max(y0, 1), in1 = min(y1, H - 1);
  const int inner = in1 - in0;
  if (inner > 0) {
    const int R = in.sharp_rows;
    const int n_strips = (inner + R - 1) / R;
    const int n_items = wpr * n_strips;
    for (int item = c.tid; item < n_items; item += NCONS) {
      const int strip = item / wpr, xw = item - strip * wpr;
      const int y_begin = in0 + strip * R, y_end = min(in1, y_begin + R);
      const int xb0 = xw << 2;
      uint32_t bmask = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (xb0 + b < C || xb0 + b >= row - C) bmask |= 1u << b;
      const int ph = (C == 3) ? (xw % 3) : 0;  // 4 == 1 mod 3
      const uint32_t col = stage + (uint32_t)((y_begin - 1 - sr0) * row + xb0);
      sharp_walk<C>(col, row, y_end - y_begin, xw > 0, xw + 1 < wpr, bmask, f,
                    [&](int r, uint32_t o) { emit(y_begin + r, xw, ph, o); });
    }
  }
  if (!COUNT) {
    store_barrier(c);
    if (c.tid == 0) {
      bulk_store(c.dst + (size_t)y0 * row, c.ostage, (uint32_t)((y1 - y0) * row));
      bulk_commit();
    }
  } else if (use1) {
    fence_proxy_async();  // the unit was written through the generic proxy; the TMA refills it next
  }
}

// =================================================================================== producer
// 3-D tensor-map load (box of W*C/4 x H x images uint32 elements) completing on an mbarrier.
__device__ __forceinline__ void tensor_load_3d(uint32_t dst_smem, const TMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// What the producer decides for one tile.
struct TilePlanD {
  SlotInfo in;
  uint32_t tx_bytes;
  int src_sel;
  bool box_overflow;  // a gather tile whose source box does not fit: worth retrying on a part of the tile
};

// Decides how a tile is executed.  Every lane of the producer warp computes the same plan.  This
// runs once per tile on a single warp, so everything that does not depend on the tile (strip
// height, flat units per tile, alignment of the batch) comes precomputed in KParams.
// `split` > 0 plans sub-rectangle `sub` of the tile's box cut into 2^split x-aligned row bands (split = 1)
// or 2 x 2 quarters (split = 2): used when the source box of the whole tile does not fit two units.
template <int C>
__device__ __forceinline__ void plan_tile(const KParams& p, const TileState& t, int img, int tile, TilePlanD& d,
                                          int split = 0, int sub = 0) {
  constexpr int UB = (C == 3) ? 48 : 16;
  const int H = p.H, W = p.W;
  const int pass_kind = t.pass_kind;
  const bool count = pass_kind == PASS_COUNT;
  const int kmode = t.kmode, n_sp = t.n_sp;
  const bool src16 = (t.src_sel != 0) || (p.flags & 1);                       // scratch images are 256-byte aligned
  const bool dst16 = count || (pass_kind != PASS_WRITE_OUT) || (p.flags & 2);
  const bool fast = !p.force_generic;
  SlotInfo& in = d.in;
  in.cls = CLS_GENERIC; in.img = img; in.tile = tile; in.pass_kind = pass_kind;
  in.bx0 = 0; in.bx1 = -1; in.by0 = 0; in.by1 = -1; in.bxb0 = 0; in.rowb = 0; in.pitch = 16; in.rows = 0;
  in.span = 1; in.sharp_rows = 1; in.fillc[0] = 0; in.fillc[1] = 0; in.paint = 0;
  in.first = 1; in.last = 1; in.st_idx = 0; in.expected = 1u; in.rpi = 1; in.n_steps = 0; in.inv_qpr = 1024u; in._pad0 = 0u; in._pad = 0;
  in.src = nullptr; in.dst = nullptr;
  d.tx_bytes = 0;
  d.src_sel = t.src_sel;
  d.box_overflow = false;
  const bool plainK = (kmode == K_NONE || kmode == K_COLOR);

  // ---- the common case first: no spatial op, a flat run of units
  if (plainK && n_sp == 0 && fast && (p.flags & 4) && src16 && dst16) {
    const int u0 = min(p.flat_units, tile * p.flat_upt), u1 = min(p.flat_units, u0 + p.flat_upt);
    in.cls = CLS_FLAT;
    in.x0 = u0; in.x1 = u1; in.y0 = 0; in.y1 = 0;
    d.tx_bytes = (uint32_t)(u1 - u0) * UB;  // <= 4096 * C <= one unit
    return;
  }

  const int rowbytes = W * C;
  const bool rows16 = (p.flags & 8) && src16 && dst16;
  // tile -> region.  The partition depends only on per-image state, so all tiles of a pass agree.
  Rect strip, box;
  strip.x0 = 0; strip.x1 = W; strip.y0 = min(H, tile * p.strip_rows); strip.y1 = min(H, strip.y0 + p.strip_rows);
  {
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    box.x0 = min(W, tx * p.tw); box.x1 = min(W, box.x0 + p.tw);
    box.y0 = min(H, ty * p.th); box.y1 = min(H, box.y0 + p.th);
    if (split == 1) {  // two row bands
      const int hy = box.y0 + (p.th >> 1);
      if (sub == 0) box.y1 = min(box.y1, hy); else box.y0 = min(box.y1, hy);
    } else if (split == 2) {  // four quarters
      const int hy = box.y0 + (p.th >> 1), hx = box.x0 + (p.tw >> 1);
      if (sub & 2) box.y0 = min(box.y1, hy); else box.y1 = min(box.y1, hy);
      if (sub & 1) box.x0 = min(box.x1, hx); else box.x1 = min(box.x1, hx);
    }
  }
  d.box_overflow = false;
  Rect reg = (n_sp > 0) ? box : strip;
  bool masks_only = n_sp > 0;
  for (int k = 0; k < n_sp; ++k) masks_only = masks_only && (t.sp[k].type == SP_MASK);
  for (int ch = 0; ch < C; ++ch) {
    in.fillc[0] |= (uint32_t)(t.sp[max(n_sp - 1, 0)].color[ch] & 255) << (8 * ch);
    in.fillc[1] |= (uint32_t)(t.sp[max(n_sp - 2, 0)].color[ch] & 255) << (8 * ch);
  }

  // Source bounding box of output rectangle q: push its corners back through every warp of the
  // list.  Affine maps take extremes at corners; one pixel of margin per stage covers the rounding
  // of that stage.  The box is fetched as ONE tensor-map box of p.box_rows x p.box_bytes starting at
  // the 16-byte aligned byte bxb0 of row by0; returns false if the bounding box does not fit in it.
  auto source_box = [&](const Rect q) -> bool {
    if (!p.use_tmap) return false;
    float minx = (float)q.x0, maxx = (float)(q.x1 - 1), miny = (float)q.y0, maxy = (float)(q.y1 - 1);
    bool empty = (q.x1 <= q.x0 || q.y1 <= q.y0);
    for (int k = n_sp - 1; k >= 0; --k) {
      const Spatial& e = t.sp[k];
      if (e.type != SP_GEOM || empty) continue;
      float lx = 3.0e38f, hx = -3.0e38f, ly = 3.0e38f, hy = -3.0e38f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float px = (j & 1) ? maxx : minx, py = (j & 2) ? maxy : miny;
        const float sx = e.t[0] * px + e.t[1] * py + e.t[2];
        const float sy = e.t[3] * px + e.t[4] * py + e.t[5];
        lx = fminf(lx, sx); hx = fmaxf(hx, sx); ly = fminf(ly, sy); hy = fmaxf(hy, sy);
      }
      minx = fmaxf(lx - 1.0f, 0.0f); maxx = fminf(hx + 1.0f, (float)(W - 1));
      miny = fmaxf(ly - 1.0f, 0.0f); maxy = fminf(hy + 1.0f, (float)(H - 1));
      if (!(minx <= maxx && miny <= maxy)) empty = true;  // also catches NaN
    }
    if (empty) return true;  // every pixel of the tile shows a fill colour: nothing to stage
    in.bx0 = max(0, (int)floorf(minx)); in.bx1 = min(W - 1, (int)ceilf(maxx));
    in.by0 = max(0, (int)floorf(miny)); in.by1 = min(H - 1, (int)ceilf(maxy));
    in.bxb0 = (in.bx0 * C) & ~15;  // the TMA wants a 16-byte aligned box start
    if (in.by1 - in.by0 + 1 > p.box_rows || (in.bx1 + 1) * C - in.bxb0 > p.box_bytes) {
      in.bx0 = 0; in.bx1 = -1; in.by0 = 0; in.by1 = -1; in.bxb0 = 0;
      d.box_overflow = true;
      return false;
    }
    in.rows = p.box_rows; in.rowb = p.box_bytes; in.pitch = p.box_bytes;
    return true;
  };

  if (plainK && (n_sp == 0 || (masks_only && !count))) {
    if (fast && (p.flags & 4) && src16 && dst16) {  // CutOut only: a flat run with the rectangles painted over it
      const int u0 = min(p.flat_units, tile * p.flat_upt), u1 = min(p.flat_units, u0 + p.flat_upt);
      in.cls = CLS_FLAT;
      // paint only if a rectangle reaches the rows of this run (most tiles of a CutOut image are plain)
      if (u1 > u0) {
        constexpr int PPU = UB / C;
        const int yf = (u0 * PPU) / W, yl = (u1 * PPU - 1) / W;
        bool hit = false;
        for (int k = 0; k < n_sp; ++k) hit = hit || (yl >= t.sp[k].y0 && yf < t.sp[k].y1 && t.sp[k].x1 > t.sp[k].x0);
        in.paint = hit ? n_sp : 0;
      }
      reg.x0 = u0; reg.x1 = u1; reg.y0 = 0; reg.y1 = 0;
      d.tx_bytes = (uint32_t)(u1 - u0) * UB;
    } else {
      reg = strip;  // all tiles of the pass take this branch together
    }
  } else if (plainK && fast && rows16 && t.sp_fast) {
    if (source_box(box)) {
      in.cls = CLS_GATHER;
      d.tx_bytes = (uint32_t)(in.rows * in.rowb);
    }
  } else if (kmode == K_SHARP && n_sp == 0 && fast && rows16) {
    const int sr0 = max(0, strip.y0 - 1), sr1 = min(H, strip.y1 + 1);
    const int nrows = (strip.y1 > strip.y0) ? sr1 - sr0 : 0;
    if ((long long)nrows * rowbytes <= 2 * UNIT_BYTES && (long long)(strip.y1 - strip.y0) * rowbytes <= ostage_bytes(C)) {
      in.cls = CLS_SHARP;
      in.by0 = sr0; in.rows = nrows;
      in.sharp_rows = sharp_split(rowbytes >> 2, max(1, min(strip.y1, H - 1) - max(strip.y0, 1)));
      d.tx_bytes = (uint32_t)(nrows * rowbytes);
    }
  } else if (kmode == K_SHARP && n_sp > 0 && fast && rows16 && t.sp_fast) {
    Rect halo = box;
    halo.x0 = max(0, box.x0 - 1); halo.x1 = min(W, box.x1 + 1);
    halo.y0 = max(0, box.y0 - 1); halo.y1 = min(H, box.y1 + 1);
    if (box.x1 > box.x0 && box.y1 > box.y0 && source_box(halo)) {
      in.cls = CLS_GATHER_SHARP;
      in.sharp_rows = sharp_split(((box.x1 - box.x0) * C) >> 2, max(1, (box.y1 - box.y0 + 1) >> 1));
      d.tx_bytes = (uint32_t)(in.rows * in.rowb);
    }
  }
  in.x0 = reg.x0; in.x1 = reg.x1; in.y0 = reg.y0; in.y1 = reg.y1;
  in.span = (d.tx_bytes > (uint32_t)UNIT_BYTES) ? 2 : 1;
}

// Re-planning of one part of a split tile.
template <int C>
__device__ __forceinline__ TilePlanD plan_tile_part(const KParams& p, const TileState& t, int img, int tile, int split, int sub) {
  TilePlanD d;
  plan_tile<C>(p, t, img, tile, d, split, sub);
  return d;
}

// ================================================================================ pass kernel
// (Making the rarely used executors -- scalar, list gather, COUNT forms of the Sharpness executors --
// real calls was measured twice to shrink the hot instruction footprint: the calls cost the flat
// path 10-20 % (identity 69 -> 54 %, LUT ops 51 -> 44 % of the copy peak, profiles/r01_v12_cold_calls_rejected.json)
// and did not move the small-batch time, so everything stays inlined.)
template <int C, bool COUNT>
__device__ __forceinline__ void run_tile(const TC<C>& c) {
  switch (c.info->cls) {
    case CLS_FLAT: exec_flat<C, COUNT>(c); break;
    case CLS_GATHER: exec_gather<C, COUNT>(c); break;
    case CLS_SHARP: exec_sharp<C, COUNT>(c); break;
    case CLS_GATHER_SHARP: exec_gather_sharp<C, COUNT>(c); break;
    default: exec_generic<C, COUNT>(c); break;
  }
}

template <int C>
__global__ void __launch_bounds__(NT, (sizeof(PassSmem<C>) + 1024) * 2 <= 227 * 1024 ? 2 : 1)
pass_kernel(const KParams p, const __grid_constant__ TMap tm_in, const __grid_constant__ TMap tm_scr) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  PassSmem<C>* sm = reinterpret_cast<PassSmem<C>*>(smem_raw);
  const int tid = threadIdx.x;
  TL_DECL();
  TL_STAMP(0);
  // Programmatic dependent launch: set up while the plan kernel still runs, do not read what it
  // wrote before it has completed.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int H = p.H, W = p.W;
  const int img_bytes = H * W * C;
  const uint32_t full0 = smem_addr(&sm->full[0]), empty0 = smem_addr(&sm->empty[0]);
  const uint32_t stbar0 = smem_addr(&sm->stbar[0]);
  if (tid == 0) {
    for (int u = 0; u < NU; ++u) { mbar_init(full0 + 8 * u, 1); mbar_init(empty0 + 8 * u, NCONS / 32); }
    for (int b = 0; b < NST; ++b) mbar_init(stbar0 + 8 * b, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  TL_STAMP(1);
  if (tid >= NCONS) {
    // ------------------------------------------------------------------ producer warp
    const int lane = tid - NCONS;
    // first passes of all images: the bins back to back
    // (one lane per bin reads its length -- every thread of every CTA reading the same few words made
    // that cache line a hot spot worth up to 15 us of start-up -- and a warp scan turns them into ends)
    unsigned bin_end[NBINS];
    unsigned n_entries = 0;
    {
      unsigned incl = (lane < NBINS) ? __ldcg(p.counters + lane) : 0u;
  #pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
      }
  #pragma unroll
      for (int b = 0; b < NBINS; ++b) bin_end[b] = __shfl_sync(0xFFFFFFFFu, incl, b);
      n_entries = bin_end[NBINS - 1];
    }
    // A claim is a *chunk* of one entry: one tile for the heavy bins, G consecutive tiles for the
    // light ones (one atomic, one image lookup and one TileState fetch per chunk).  G grows with the
    // work per CTA so that small batches keep enough chunks per CTA for the dynamic schedule to
    // balance the tail.  The bins form six segments (bin_of): heavy, light 2-D tiles and flat runs
    // (which have their own tile count, p.n_flat_tiles), each for non-final and final passes.
    const unsigned n_tiles_u = (unsigned)p.n_tiles, n_flat_u = (unsigned)p.n_flat_tiles;
    unsigned G = 1;
    const unsigned per_cta = (n_entries * n_tiles_u) / gridDim.x;  // tiles per CTA
#ifndef CHB_G2_MIN
#define CHB_G2_MIN 16u
#endif
#ifndef CHB_G4_MIN
#define CHB_G4_MIN 48u
#endif
    if (per_cta >= CHB_G4_MIN) G = 4; else if (per_cta >= CHB_G2_MIN) G = 2;
    // tiles per image and tiles per chunk of segment k.  At large batches a pass that is not the image's
    // last is claimed whole: the CTA then needs no election to know that it completed the pass and
    // finalises from its own shared-memory histogram (see the consumer loop).
#ifndef CHB_WHOLE_NF_MIN
#define CHB_WHOLE_NF_MIN 96u
#endif
#ifndef CHB_WHOLE_IMG_MIN
#define CHB_WHOLE_IMG_MIN 6u
#endif
    const bool whole_nf = per_cta >= CHB_WHOLE_NF_MIN && n_entries >= CHB_WHOLE_IMG_MIN * gridDim.x;
    const bool nf = p.nf_first != 0;
    auto seg_tiles = [&](int k) -> unsigned { return seg_kind(k, nf) == 2 ? n_flat_u : n_tiles_u; };
    auto seg_g = [&](int k) -> unsigned {
      if (seg_kind(k, nf) == 0) return 1u;
      if (whole_nf && seg_nonfinal_light(k, nf)) return seg_tiles(k);
      return G;
    };
    constexpr int NSEG = 6;
    unsigned seg_entry0[NSEG], seg_begin[NSEG + 1];
    seg_begin[0] = 0;
#pragma unroll
    for (int k = 0; k < NSEG; ++k) {
      const unsigned cp = (seg_tiles(k) + seg_g(k) - 1u) / seg_g(k);
      seg_entry0[k] = seg_first(k, nf) ? bin_end[seg_first(k, nf) - 1] : 0u;
      seg_begin[k + 1] = seg_begin[k] + (bin_end[seg_last(k, nf)] - seg_entry0[k]) * cp;
    }
    const unsigned n_chunks = seg_begin[NSEG];
    TL_STAMP(13);

    unsigned* work = p.counters + NBINS;
    struct Chunk { int img, t0, t1; unsigned expected; int local; };
    // image and tiles of a global chunk (img = -1 past the end)
    auto global_chunk = [&](unsigned chunk) -> Chunk {
      Chunk c = {-1, 0, 0, 1u, 0};
      if (chunk >= n_chunks) return c;
      unsigned sbeg = 0, sent = seg_entry0[0], g = seg_g(0), tiles = seg_tiles(0);
#pragma unroll
      for (int k = 1; k < NSEG; ++k)
        if (chunk >= seg_begin[k]) { sbeg = seg_begin[k]; sent = seg_entry0[k]; g = seg_g(k); tiles = seg_tiles(k); }
      const unsigned cp = (tiles + g - 1u) / g;
      const unsigned loc = chunk - sbeg;
      const unsigned e_in = loc / cp;
      const unsigned ci = loc - e_in * cp;
      const unsigned entry = sent + e_in;
      c.t0 = (int)(ci * g);
      c.t1 = (int)min(tiles, ci * g + g);
      c.expected = cp;
      int bin = 0;
      unsigned first = 0;
#pragma unroll
      for (int b = 0; b < NBINS - 1; ++b)
        if (entry >= bin_end[b]) { bin = b + 1; first = bin_end[b]; }
      c.img = __ldcg(p.lists + (size_t)bin * p.B + (entry - first));
      return c;
    };
    auto fetch_state = [&](int img, int buf) {
      if (lane == 0) {
        mbar_arrive_expect_tx(stbar0 + 8 * buf, (uint32_t)sizeof(TileState));
        bulk_load(smem_addr(&sm->stg[buf]), p.states + img, (uint32_t)sizeof(TileState), stbar0 + 8 * buf);
      }
    };
    const Chunk none = {-1, 0, 0, 1u, 0};
    // Continuations: when a CTA completes a COUNT / WRITE_SCRATCH pass of an image it publishes the
    // image in p.cont (entry = image + 1; 0 = not yet published).  Every producer holds one *ticket*
    // into that list (ticket -> entry ticket / cpi_c, tiles (ticket % cpi_c) * Gc ...) and polls it
    // once per chunk; a published ticket goes into the pipeline ahead of the global queue, so the
    // next pass of an image is spread over all CTAs as soon as it exists.  p.counters[NBINS + 3]
    // counts the images that may still publish: when it is zero an unpublished ticket is void.
    // (large batches: a ticket is a whole image, so a CTA runs 16+ tiles of one executor before its code changes again)
    unsigned Gc = 1u;
    while (Gc * 2u <= n_tiles_u && Gc * 16u <= per_cta) Gc *= 2u;  // at most an eighth of a CTA's share
    const unsigned cpi_c = (n_tiles_u + Gc - 1u) / Gc;
    unsigned* cwork = p.counters + NBINS + 1;
    volatile unsigned* pending = p.counters + NBINS + 3;
    volatile int* cont = p.cont;
    // Pipeline: c0 is issued now (its state fetch is in flight), c1's state is fetched while c0 is
    // issued, c2 is resolved (image lookup) meanwhile, and the claim of the chunk after that is in
    // flight.  Every chunk and ticket is claimed dynamically, also the first ones: a CTA that has
    // not started yet (the SMs may be shared with another stream's kernel) must not own any work
    // that running CTAs wait for.
    Chunk c0, c1, c2;
    unsigned chunk_next;  // claimed, not yet resolved
    unsigned ticket;
    {
      // three separate claims (in flight together): all CTAs start at once, so the first chunks --
      // the most expensive tiles of the batch -- are dealt round-robin instead of three to a CTA
      unsigned v0 = 0, v1 = 0, v2 = 0, tv = 0;
      if (lane == 0) { v0 = atomicAdd(work, 1u); v1 = atomicAdd(work, 1u); v2 = atomicAdd(work, 1u); tv = atomicAdd(cwork, 1u); }
      v0 = __shfl_sync(0xFFFFFFFFu, v0, 0);
      v1 = __shfl_sync(0xFFFFFFFFu, v1, 0);
      chunk_next = __shfl_sync(0xFFFFFFFFu, v2, 0);
      ticket = __shfl_sync(0xFFFFFFFFu, tv, 0);
      TL_STAMP(5);
      c0 = global_chunk(v0);
      c1 = global_chunk(v1);
    }
    bool g_done = false;      // the global queue is exhausted
    bool ticket_wait = false; // a new ticket has been requested (traw)
    unsigned traw = 0;
    int polled = 0;           // what the last poll of the ticket's entry returned (lane 0; 0 = not published)
    uint32_t ring = 0;        // next state buffer to allocate
    uint32_t sb0 = 0, sb1 = 0;
    uint32_t st_par = 0;      // parity to wait for on stbar[b]
    uint32_t pu = 0;          // next unit (monotonic)
    uint32_t empty_par = 0xF; // parity to wait for on empty[u]: a fresh barrier passes parity 1
    auto take_unit = [&](uint32_t u) {  // wait until the consumers have released unit u
      TL_T0();
      mbar_wait(empty0 + 8 * u, (empty_par >> u) & 1u);
      TL_ACC(8);
      empty_par ^= 1u << u;
    };
    auto alloc_state = [&]() -> uint32_t { const uint32_t b = ring; ring = (ring + 1 == NST) ? 0u : ring + 1; return b; };
    // (bit 30 of an entry: the pass is a flat run, which has n_flat_tiles tiles -- the tickets of the
    // chunks beyond them are void)
    auto cont_chunk = [&](int entry_value) -> Chunk {
      Chunk c;
      const unsigned ci = ticket % cpi_c;
      const bool flat = (entry_value & CONT_FLAT) != 0;
      const unsigned tiles = flat ? n_flat_u : n_tiles_u;
      c.img = (entry_value & ~CONT_FLAT) - 1; c.t0 = (int)(ci * Gc); c.t1 = (int)min(tiles, ci * Gc + Gc);
      c.expected = (tiles + Gc - 1u) / Gc; c.local = 1;
      if (c.t0 >= c.t1) c.img = -1;
      return c;
    };
    if (c0.img >= 0) { sb0 = alloc_state(); fetch_state(c0.img, (int)sb0); }
    for (;;) {
      if (c1.img >= 0) {
        sb1 = alloc_state();
        if (c1.local) { __threadfence(); fence_proxy_async_all(); }  // published by another CTA's finaliser (generic proxy)
        fetch_state(c1.img, (int)sb1);
      }
      // two ahead: a published continuation first, else the next global chunk
      bool claimed = false;
      unsigned raw = 0;
      {
        const int pv = __shfl_sync(0xFFFFFFFFu, polled, 0);
        if (pv > 0) {
          c2 = cont_chunk(pv);
          if (lane == 0) traw = atomicAdd(cwork, 1u);
          ticket_wait = true;
          polled = 0;
        } else if (!g_done) {
          c2 = global_chunk(chunk_next);
          if (c2.img < 0) {
            g_done = true;
          } else {
            claimed = true;
            if (lane == 0) raw = atomicAdd(work, 1u);  // consumed at the end of this iteration
          }
        } else {
          c2 = none;
        }
        // poll the ticket for the next iteration (not while a new ticket is on its way)
        if (!ticket_wait && lane == 0) polled = cont[ticket / cpi_c];
      }
      if (c0.img >= 0) {
        TL_T0();
        mbar_wait(stbar0 + 8 * sb0, (st_par >> sb0) & 1u);
        TL_ACC(9);
        TL_STAMP_ONCE(6);
        st_par ^= 1u << sb0;
        const TileState& st = sm->stg[sb0];
        const int img0 = c0.img;
        for (int tile = c0.t0; tile < c0.t1; ++tile) {
          TL_T0();
          TilePlanD d;
          plan_tile<C>(p, st, img0, tile, d);
          // takes the unit(s), hands the plan to the consumers and issues the loads of one tile
          auto emit_tile = [&]() {
            if (d.in.span == 2 && (pu & (NU - 1)) == NU - 1) {  // a double tile may not wrap: pad the ring
              const uint32_t us = pu & (NU - 1);
              take_unit(us);
              if (lane == 0) { sm->info[us].cls = CLS_SKIP; sm->info[us].span = 1; mbar_arrive(full0 + 8 * us); }
              ++pu;
            }
            const uint32_t u = pu & (NU - 1);
            take_unit(u);
            if (d.in.span == 2) take_unit(u + 1);
            pu += d.in.span;
            d.in.st_idx = (int)sb0;  // the consumers read the prefetched state in place
            d.in.expected = c0.expected;
            if (lane == 0) {
              const size_t img_off = (size_t)img0 * img_bytes;
              const uint8_t* src = (d.src_sel == 0) ? p.in + img_off
                                                   : p.scratch + (size_t)(2 * (size_t)img0 + (d.src_sel - 1)) * p.scratch_stride;
              d.in.src = src;
              d.in.dst = (d.in.pass_kind == PASS_WRITE_OUT)
                             ? p.out + img_off
                             : p.scratch + (size_t)(2 * (size_t)img0 + (st.dst_sel - 1)) * p.scratch_stride;
              if (d.in.cls == CLS_GATHER_SHARP) {  // halo tile width -> exact division by multiplication
                const unsigned vw = (unsigned)(d.in.x1 - d.in.x0 + 2);
                d.in.inv_qpr = (uint32_t)((0x100000000ull + vw - 1ull) / vw);
              }
              if (d.in.cls == CLS_GATHER) {  // split of the tile over lanes and warp steps (exec_gather_warp)
                const int tw = d.in.x1 - d.in.x0, th = d.in.y1 - d.in.y0;
                const int qpr = max(1, tw >> 2);
                const int rpi = max(1, 32 / qpr);
                d.in.rpi = rpi;
                d.in.n_steps = (th + rpi - 1) / rpi;
                d.in.inv_qpr = (uint32_t)((1024 + qpr - 1) / qpr);
              }
              sm->info[u] = d.in;
              const uint32_t fb = full0 + 8 * u;
              const uint32_t dst = smem_addr(sm->data[u]);
              if (d.tx_bytes == 0) {
                mbar_arrive(fb);
              } else {
                mbar_arrive_expect_tx(fb, d.tx_bytes);
                if (d.in.cls == CLS_FLAT) {
                  constexpr int UB = (C == 3) ? 48 : 16;
                  bulk_load(dst, src + (size_t)d.in.x0 * UB, d.tx_bytes, fb);
                } else if (d.in.cls == CLS_SHARP) {
                  bulk_load(dst, src + (size_t)d.in.by0 * (W * C), d.tx_bytes, fb);
                } else if (d.src_sel == 0) {
                  tensor_load_3d(dst, &tm_in, d.in.bxb0 >> 2, d.in.by0, img0, fb);
                } else {
                  tensor_load_3d(dst, &tm_scr, d.in.bxb0 >> 2, d.in.by0, 2 * img0 + (d.src_sel - 1), fb);
                }
              }
            }
            __syncwarp();
            TL_STAMP_ONCE(7);
          };
          // A gather tile whose source box does not fit two units is issued as 2 / 4 parts.  The
          // first part of the chunk's first tile and the last part of its last tile carry the
          // first / last flags: the chunk shares one histogram flush and counts once towards the
          // completion handshake of a COUNT / WRITE_SCRATCH pass.
          int n_sub = 1, split = 0;
          if (d.box_overflow) {
            for (split = 1; split <= 2; ++split) {
              bool ok = true;
              for (int sub = 0; sub < (1 << split) && ok; ++sub) {
                d = plan_tile_part<C>(p, st, img0, tile, split, sub);
                ok = !d.box_overflow;
              }
              if (ok) break;
            }
            if (split > 2) split = 0;  // no luck: the scalar executor takes the whole tile
            n_sub = 1 << split;
            if (split == 0) d = plan_tile_part<C>(p, st, img0, tile, 0, 0);
          }
          TL_ACC(11);
          for (int sub = 0; sub < n_sub; ++sub) {
            if (split > 0) d = plan_tile_part<C>(p, st, img0, tile, split, sub);
            d.in.first = (tile == c0.t0 && sub == 0); d.in.last = (tile == c0.t1 - 1 && sub == n_sub - 1);
            emit_tile();
          }
        }
      } else if (c1.img < 0 && c2.img < 0 && !ticket_wait) {
        // Nothing left to issue: the global queue is exhausted.  Wait for the ticket: it is either
        // published, or void once no image can publish any more.
        TL_T0();
        int pv = 0;
        if (lane == 0) {
          for (;;) {
            pv = cont[ticket / cpi_c];
            if (pv > 0) break;
            if (*pending == 0u) {
              __threadfence();
              pv = cont[ticket / cpi_c];
              break;
            }
            __nanosleep(100);
          }
          polled = pv;
        }
        pv = __shfl_sync(0xFFFFFFFFu, pv, 0);
        TL_ACC(12);
        if (pv <= 0) {
          const uint32_t u = pu & (NU - 1);
          take_unit(u);
          if (lane == 0) { sm->info[u].cls = CLS_END; mbar_arrive(full0 + 8 * u); }
          break;
        }
      }
      TL_T0();
      if (claimed) chunk_next = __shfl_sync(0xFFFFFFFFu, raw, 0);
      if (ticket_wait) { ticket = __shfl_sync(0xFFFFFFFFu, traw, 0); ticket_wait = false; }
      TL_ACC(10);
      c0 = c1; sb0 = sb1; c1 = c2;
    }
    TL_STAMP(4);
    TL_FLUSH(0, lane == 0);
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  TC<C> c;
  uint32_t nstore = 0;
  c.nstore = &nstore;
  c.p = &p; c.sm = sm; c.r = smem_addr(sm->r);
  c.H = H; c.W = W; c.HW = H * W; c.img_bytes = img_bytes; c.tid = tid; c.lane = tid & 31;
  const int warp = tid >> 5, lane = tid & 31;
  uint32_t cu = 0;        // next unit (monotonic)
  uint32_t full_par = 0;  // parity to wait for on full[u]
  for (;;) {
    const uint32_t u = cu & (NU - 1);
    TL_T0();
    mbar_wait(full0 + 8 * u, (full_par >> u) & 1u);
    TL_ACC(6);
    full_par ^= 1u << u;
    const SlotInfo& info = sm->info[u];
    const int cls = info.cls;
    if (cls == CLS_END) break;
    TL_STAMP_ONCE(2);
    if (cls == CLS_SKIP) {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * u);
      ++cu;
      continue;
    }
    const int span = info.span;
    const unsigned expected = info.expected;  // chunks of the image that report once each, on their last tile
    const int img = info.img, pass_kind = info.pass_kind;
    const TileState& st = sm->stg[info.st_idx];
    c.t = &st; c.info = &info; c.tile = info.tile;
    c.data = smem_addr(sm->data[u]);
    c.ostage = c.r + (nstore & 1u) * ostage_bytes(C);
    c.l1a = smem_addr(&st.l1[0][0]); c.l2a = smem_addr(&st.l2[0][0]);
    c.src = info.src;
    c.dst = info.dst;
    ImgState* g = p.states + img;
    const int is_last = info.last;
    TL_T0();
    if (pass_kind == PASS_COUNT) {
      if (info.first) {
        for (int i = tid; i < MAXC * 256; i += NCONS) (&sm->hist[0][0])[i] = 0u;
        if (tid < CHB_MAX_CHAIN) sm->color_cnt[tid] = 0u;
        cons_sync();
      }
      run_tile<C, true>(c);
      if (is_last && expected != 1u) {
        cons_sync();
        for (int i = tid; i < C * 256; i += NCONS) {
          const uint32_t v = (&sm->hist[0][0])[i];
          if (v) atomicAdd(&g->hist[0][0] + i, v);
        }
        if (tid < CHB_MAX_CHAIN && sm->color_cnt[tid]) atomicAdd(&g->color_cnt[tid], sm->color_cnt[tid]);
      }
    } else if (cls == CLS_GATHER && (st.n_sp == 2 || (st.n_sp == 1 && st.sp[0].type == SP_GEOM))) {
      exec_gather_warp<C>(c, warp);
    } else {
      run_tile<C, false>(c);
    }
    TL_ACC(7 + cls);  // cls 1..5 -> slots 8..12
    TL_ADD(5, 1);
    TL_STAMP(3);
    // This warp is done with the unit(s): state, info and staged bytes.  What the loop needs from
    // the plan from here on is read again from shared memory instead of being kept alive across the
    // executor: at the 96-register cap those values were spilled, and their reloads from local
    // memory were 13 % of the gather tiles' stall samples and cost the flat path 10-20 %
    // (profiles/r01_v13_ab_experiments.txt, 9).  The second __syncwarp() matters: lane 0's arrival
    // releases the unit, and without it the other lanes' reads were not ordered before that release --
    // a lane could read the *next* tile's pass_kind, take the finaliser branch alone and hang the CTA.
    __syncwarp();
    const uint32_t ia = smem_addr(&sm->info[cu & (NU - 1)]);
    const int span_r = (int)lds_u32(ia + (uint32_t)offsetof(SlotInfo, span));
    const int last_r = (int)lds_u32(ia + (uint32_t)offsetof(SlotInfo, last));
    const int pass_r = (int)lds_u32(ia + (uint32_t)offsetof(SlotInfo, pass_kind));
    const int img_r = (int)lds_u32(ia + (uint32_t)offsetof(SlotInfo, img));
    const unsigned expected_r = lds_u32(ia + (uint32_t)offsetof(SlotInfo, expected));
    __syncwarp();  // every lane has its copy before lane 0 lets go of the unit
    if (lane == 0) {
      const uint32_t ur = cu & (NU - 1);
      mbar_arrive(empty0 + 8 * ur);
      if (span_r == 2) mbar_arrive(empty0 + 8 * (ur + 1));
    }
    cu += span_r;
    if (pass_r == PASS_WRITE_OUT || !last_r) continue;

    // ---- COUNT / WRITE_SCRATCH: the CTA that completes the image's last chunk resumes the chain walk
    TL_T0();
    // the finaliser scratch below aliases the output staging tiles; a scratch image must also
    // have landed in global memory before the pass is reported complete
    if (pass_r == PASS_WRITE_SCRATCH) {
      if (lane == 0) { bulk_wait_all0(); fence_proxy_async_all(); }
    } else if (tid < 32) {
      bulk_wait_read0();
    }
    cons_sync();
    ImgState* gr = p.states + img_r;
    const bool sole = expected_r == 1u;  // this CTA ran the whole pass: no election, the counts are in shared memory
    if (!sole) {
      if (tid == 0) {
        __threadfence();  // cumulative: publishes the histogram atomics of the whole CTA (ordered by the barrier)
        const int elected = (atomicAdd(&gr->tiles_done, 1u) == expected_r - 1u) ? 1 : 0;
        if (elected) __threadfence();  // acquire side: only the finaliser reads what the others published
        sm->ctl[1] = elected;
      }
      cons_sync();
      if (!sm->ctl[1]) { TL_ACC(13); continue; }
    }
    // the slot of the image's next pass in the continuation list: asked for now, needed at the very end
    // (one global round trip less on the critical path of a two-stage image)
    unsigned cont_pos = 0;
    if (tid == 0) cont_pos = atomicAdd(p.counters + NBINS + 2, 1u);
    ImgState* fs = reinterpret_cast<ImgState*>(sm->r);
    uint32_t* hmap = reinterpret_cast<uint32_t*>(sm->r + sizeof(ImgState));
    uint8_t* etab = sm->r + sizeof(ImgState) + MAXC * 256 * 4;
    if (sole && pass_r == PASS_COUNT) {
      for (int i = tid; i < STATE_VECS; i += NCONS)
        reinterpret_cast<uint4*>(fs)[i] = __ldcg(reinterpret_cast<const uint4*>(gr) + i);
      for (int i = tid; i < MAXC * 256; i += NCONS) (&fs->hist[0][0])[i] = (i < C * 256) ? (&sm->hist[0][0])[i] : 0u;
      cons_sync();
      if (tid < CHB_MAX_CHAIN) fs->color_cnt[tid] = sm->color_cnt[tid];
    } else {
      for (int i = tid; i < (int)(sizeof(ImgState) / 16); i += NCONS)
        reinterpret_cast<uint4*>(fs)[i] = __ldcg(reinterpret_cast<const uint4*>(gr) + i);
    }
    cons_sync();
    if (pass_r == PASS_COUNT) {
      if (tid == 0) fs->hist_valid = 1;
    } else {
      if (tid == 0) fs->t.src_sel = fs->t.dst_sel;
      reset_view(fs, tid, NCONS);
    }
    cons_sync();
    advance(fs, gr, p, C, H, W, hmap, etab, tid, NCONS, [] { cons_sync(); });
    for (int i = tid; i < STATE_VECS; i += NCONS)
      reinterpret_cast<uint4*>(gr)[i] = reinterpret_cast<const uint4*>(fs)[i];
    const int fs_next_pass = fs->t.pass_kind;
    const int fs_next_flat = (p.n_flat_tiles != p.n_tiles && is_flat_bin(bin_of(fs->t, p.nf_first != 0), p.nf_first != 0)) ? CONT_FLAT : 0;
    // Publish the image's next pass: whichever CTAs hold the tickets of its chunks fetch the state
    // just written (and, after a WRITE_SCRATCH pass, the scratch image) with TMA loads.
    __threadfence();
    fence_proxy_async_all();
    cons_sync();  // the R region is free again, the state is in global memory
    if (tid == 0) {
      const bool more = fs_next_pass != PASS_WRITE_OUT;  // the image will publish again
      // (every thread fenced its part of the state before the barrier above)
      *reinterpret_cast<volatile int*>(p.cont + cont_pos) = (img_r + 1) | fs_next_flat;
      if (!more) {
        __threadfence();
        atomicSub(p.counters + NBINS + 3, 1u);
      }
    }
    TL_ACC(13);
    TL_ADD(14, 1);
  }
  if (tid < 32) {
    bulk_wait_all0();  // every store of this CTA has landed before the grid retires
    // The last CTA out leaves the counters and the continuation list zeroed for the next call on this
    // workspace (the host then needs no memset node in front of every call).  Every other CTA has
    // stopped reading them: its producer sent END after its last look, and this is past END.
    unsigned last = 0;
    if (lane == 0) {
      __threadfence();
      last = (atomicAdd(p.counters + NBINS + 4, 1u) == gridDim.x - 1u) ? 1u : 0u;
    }
    last = __shfl_sync(0xFFFFFFFFu, last, 0);
    if (last) {
      __threadfence();
      const unsigned n_cont = __ldcg(p.counters + NBINS + 2);
      for (unsigned i = (unsigned)lane; i < n_cont; i += 32u) p.cont[i] = 0;
      p.counters[lane] = 0u;  // 32 counter words
    }
  }
  TL_STAMP(4);
  TL_FLUSH(1, tid == 0);
}

template <int C>
cudaError_t launch_pass_c(const KParams& p, const TMap& tm_in, const TMap& tm_scr, int grid, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = sizeof(PassSmem<C>);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // may start before the previous kernel has drained
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, pass_kernel<C>, p, tm_in, tm_scr);
}

template <int C>
cudaError_t configure_c() {
  cudaError_t e = cudaFuncSetAttribute(pass_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PassSmem<C>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pass_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int C>
int ctas_per_sm_c() { return (sizeof(PassSmem<C>) + 1024) * 2 <= 227 * 1024 ? 2 : 1; }

}  // namespace

// Per-channel-count entry points, one translation unit each (chb_kernels_c<N>.cu) so the four
// instantiations compile in parallel.
#define CHB_DEFINE_CHANNEL_ENTRY(C)                                                                           \
  cudaError_t launch_pass_c##C(const KParams& p, const TMap& a, const TMap& b, int grid, cudaStream_t stream) { \
    return launch_pass_c<C>(p, a, b, grid, stream);                                                           \
  }                                                                                                           \
  cudaError_t configure_c##C() { return configure_c<C>(); }                                                   \
  int ctas_per_sm_c##C() { return ctas_per_sm_c<C>(); }

}  // namespace chb
