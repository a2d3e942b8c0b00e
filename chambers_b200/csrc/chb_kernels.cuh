// Tile-parallel RandAugment / AutoAugment policy kernels for sm_100a.
//
// Work is cut far below one image so that a 256-image batch already fills 148 SMs evenly:
//   plan_kernel   one CTA per image: decodes the image's op schedule from Philox4x32-10 (or a
//                 replayed schedule) and folds the chain into a per-image pass state (ImgState).
//   pass_kernel   persistent CTAs pull (image, tile) items of one *level* from an atomic counter.
//                 A tile is ~10 KB of the image; level 0 is every image's first pass, level n the
//                 n-th extra pass of the few images that need one.
//
// The chain is evaluated lazily (chb_internal.h: ImgState).  Point-wise ops compose into 256-entry
// LUTs, nearest-neighbour warps and CutOut go on a spatial list, Color / Sharpness / a bilinear warp
// occupy the single "kernel" slot K.  Pixels are only touched by
//   WRITE_OUT      the one pass every image has: src -> spatial list -> l1 -> K -> l2 -> out;
//   COUNT          Equalize / AutoContrast need the histogram of the virtual image: tiles count into
//                  shared memory, flush with global atomics, the last tile of the image turns the
//                  histogram into a LUT (warp-scan CDF) and resumes the chain walk;
//   WRITE_SCRATCH  an op needs a materialised neighbourhood of something that is itself a
//                  neighbourhood op (e.g. Rotate after Sharpness): the virtual image is written to
//                  an L2-resident scratch image and becomes the new src.
// so a RandAugment image costs one HBM read and one HBM write unless its chain holds a histogram
// op (one extra read, served from L2 for batches that fit) or a rare neighbourhood-of-neighbourhood
// pair.
//
// Tile executors (all bit-exact twins of each other; the scalar one is the fallback for odd shapes):
//   exec_flat     no spatial op: 48-byte (16-pixel) units, LDG.128 -> LUT/Color in registers -> STG.128
//   exec_gather   spatial list: the source bounding box of a 64 x 56 tile is staged in shared memory
//                 with 16-byte loads, pixels are gathered from it, results leave through a staged
//                 tile with 16-byte stores
//   exec_sharp    Sharpness: row strip + halo staged in shared memory, sliding 3x3 window per word column
//   exec_generic  anything, one pixel per thread straight from global memory
#pragma once
#include "chb_device.cuh"

namespace chb {
namespace {

constexpr int NT = 256;        // threads per pass CTA
constexpr int PLAN_NT = 128;   // threads per plan CTA
constexpr int PASS_MIN_CTAS = 4;
constexpr int BIG_BYTES = 48 * 1024;             // staging / histogram region of a pass CTA
constexpr int LHIST_CH_BYTES = 64 * 32 * 4;      // lane-private u8x4 histogram of one channel
constexpr int CTL_BYTES = 16;
constexpr size_t PASS_SMEM = sizeof(ImgState) + CTL_BYTES + BIG_BYTES;
constexpr int STATE_VECS = (int)(offsetof(ImgState, hist) / 16);   // everything but the histogram
constexpr int TILE_VECS = (int)(sizeof(TileState) / 16);
static_assert(offsetof(ImgState, hist) % 16 == 0, "hist must start on a 16-byte boundary");
static_assert(offsetof(ImgState, next_op) == sizeof(TileState), "finaliser part follows the tile part");

struct Rect {
  int x0, x1, y0, y1;
};

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
    s->t.n_sp = 0; s->t.kmode = K_NONE; s->t.l1_id = 1; s->t.l2_id = 1; s->t.sp_fast = 1;
    s->t.kfactor = 0.0f; s->hist_valid = 0;
  }
}

// CTA-cooperative chain walk.  *s lives in shared memory; starting at s->next_op every op is folded
// into the view until one needs the pixels.  On return s->t.pass_kind names the pass to run now; the
// walk resumes at the same op once that pass has finished.  hmap: MAXC*256 words, etab: MAXC*256
// bytes of shared scratch.  Needs nt >= 32 * C.
__device__ void advance(ImgState* s, ImgState* g, const KParams& p, int C, int H, int W, uint32_t* hmap,
                        uint8_t* etab, int tid, int nt) {
  for (;;) {
    __syncthreads();
    const int pi = s->next_op;
    const int n_prog = s->n_prog;
    const int kmode = s->t.kmode;
    const int n_sp = s->t.n_sp;
    const int hist_valid = s->hist_valid;
    const int src_sel = s->t.src_sel;
    ProgRec pe = {0, 0, 0, 0};
    if (pi < n_prog) pe = s->prog[pi];
    __syncthreads();
    if (pi >= n_prog) {
      if (tid == 0) s->t.pass_kind = PASS_WRITE_OUT;
      break;
    }
    const DevOp* op = p.ops + pe.table_index;
    const int kind = __ldg(&op->kind);
    const bool frozen = (kmode == K_SHARP || kmode == K_BILINEAR);  // K has consumed the spatial list
    uint8_t(*last_lut)[256] = (kmode == K_NONE) ? s->t.l1 : s->t.l2;
    bool boundary = false;

    if (is_pointwise(kind)) {
      const uint8_t* tab = p.optab + (size_t)pe.table_index * 256;
      for (int t = tid; t < C * 256; t += nt) last_lut[t >> 8][t & 255] = __ldg(tab + last_lut[t >> 8][t & 255]);
      if (!frozen)
        for (int t = tid; t < n_sp * C; t += nt) {
          const int k = t / C, c = t - k * C;
          s->t.sp[k].color[c] = __ldg(tab + s->t.sp[k].color[c]);
        }
      if (tid == 0) {
        if (kmode == K_NONE) s->t.l1_id = 0; else s->t.l2_id = 0;
        s->next_op = pi + 1;
      }
    } else if (kind == CHB_OP_EQUALIZE || kind == CHB_OP_AUTOCONTRAST) {
      if (!hist_valid) {
        // the histogram of the virtual image is needed: run a COUNT pass, then come back here.
        for (int t = tid; t < MAXC * 256; t += nt) (&g->hist[0][0])[t] = 0u;
        if (tid < CHB_MAX_CHAIN) s->color_cnt[tid] = 0u;
        if (tid == 0) { s->t.pass_kind = PASS_COUNT; s->tiles_done = 0u; }
        break;
      }
      // s->hist counts the values entering the last LUT; map them through it, add the spatial colours.
      for (int t = tid; t < MAXC * 256; t += nt) hmap[t] = 0u;
      __syncthreads();
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
      __syncthreads();
      stat_tables(kind, C, hmap, etab, tid);
      __syncthreads();
      for (int t = tid; t < C * 256; t += nt) last_lut[t >> 8][t & 255] = etab[(t & ~255) + last_lut[t >> 8][t & 255]];
      if (!frozen)
        for (int t = tid; t < n_sp * C; t += nt) {
          const int k = t / C, c = t - k * C;
          s->t.sp[k].color[c] = etab[c * 256 + s->t.sp[k].color[c]];
        }
      if (tid == 0) {
        if (kmode == K_NONE) s->t.l1_id = 0; else s->t.l2_id = 0;
        s->next_op = pi + 1;
      }
    } else if (kind == CHB_OP_COLOR) {
      const int mode = __ldg(&op->blend_mode);
      if (mode == BLEND_IMAGE2 || C != 3) {  // factor 1: blend returns the image itself (:30-31)
        if (tid == 0) s->next_op = pi + 1;
      } else if (kmode == K_NONE) {
        const float f = (mode == BLEND_IMAGE1) ? 0.0f : __ldg(&op->factor);
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
      const int mode = __ldg(&op->blend_mode);
      if (mode == BLEND_IMAGE2) {
        if (tid == 0) s->next_op = pi + 1;
      } else if (kmode == K_NONE) {
        if (tid == 0) {
          s->t.kmode = K_SHARP;
          s->t.kfactor = (mode == BLEND_IMAGE1) ? 0.0f : __ldg(&op->factor);  // factor 0: deg exactly
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
        const int h = __ldg(&op->ip0), colr = __ldg(&op->ip1);
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
      const int fill_mode = __ldg(&op->fill_mode), fill = __ldg(&op->fill_u8);
      if (__ldg(&op->interp) == CHB_INTERP_NEAREST) {
        if (frozen) {
          boundary = true;
        } else if (tid == 0) {
          Spatial& e = s->t.sp[n_sp];
          e.type = SP_GEOM; e.fill_mode = fill_mode;
          for (int q = 0; q < 8; ++q) e.t[q] = __ldg(tsrc + q);
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
          for (int q = 0; q < 8; ++q) e.t[q] = __ldg(tsrc + q);
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
  __syncthreads();
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
  if (lane == 0) {
    int n = 0;
    for (int i = 0; i < p.n_draws; ++i) {
      const int slot0 = i * (p.K + 1);
      const size_t rbase = ((size_t)img * p.n_draws + i) * p.K * CHB_SCHED_FIELDS;
      int choice;
      if (p.replay) {
        choice = p.replay[rbase];
        if (choice < 0 || choice >= p.T) choice = 0;
      } else {
        choice = (int)__umulhi(rnd[slot0][0], (uint32_t)p.T);
      }
      for (int j = 0; j < p.K; ++j) {
        const int opi = choice * p.K + j;
        const int kind = __ldg(&p.ops[opi].kind);
        int applied, negate, cy, cx;
        if (p.replay) {
          const int32_t* r = p.replay + rbase + (size_t)j * CHB_SCHED_FIELDS;
          applied = (r[1] != 0) && kind >= 0;
          negate = r[2] != 0;
          cy = r[3];
          cx = r[4];
        } else {
          const int slot = slot0 + 1 + j;
          applied = (kind >= 0) && ((int)(rnd[slot][0] >> 8) < __ldg(&p.ops[opi].thr24));
          negate = rnd[slot][1] < 0x80000000u;
          cy = (int)__umulhi(rndc[slot][2], (uint32_t)H);
          cx = (int)__umulhi(rndc[slot][3], (uint32_t)W);
        }
        if (p.record) {
          int32_t* r = p.record + rbase + (size_t)j * CHB_SCHED_FIELDS;
          r[0] = choice; r[1] = applied; r[2] = negate; r[3] = cy; r[4] = cx;
        }
        if (applied && n < CHB_MAX_CHAIN) {
          s->prog[n].table_index = opi; s->prog[n].negate = negate; s->prog[n].cy = cy; s->prog[n].cx = cx;
          ++n;
        }
      }
    }
    s->n_prog = n;
  }
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
  if (img == 0)
    for (int i = tid; i < 2 * p.max_levels; i += PLAN_NT) p.counters[i] = 0u;
  for (int i = tid; i < STATE_VECS; i += PLAN_NT) reinterpret_cast<uint4*>(&s)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid < 32) decode_image(p, &s, rnd, rndc, img, p.H, p.W, tid);
  reset_view(&s, tid, PLAN_NT);
  __syncthreads();
  advance(&s, p.states + img, p, C, p.H, p.W, hmap, etab, tid, PLAN_NT);
  for (int i = tid; i < STATE_VECS; i += PLAN_NT)
    reinterpret_cast<uint4*>(p.states + img)[i] = reinterpret_cast<const uint4*>(&s)[i];
}

#endif  // CHB_WITH_PLAN

// ================================================================================ tile context
template <int C>
struct TC {
  const KParams* p;
  ImgState* s;
  uint32_t big;        // shared address of the staging region
  const uint8_t* src;  // source image of this pass
  uint8_t* dst;        // destination image (unused by COUNT passes)
  int H, W, HW, img_bytes, tid, lane, tile;
  uint32_t l1a, l2a;   // shared addresses of the two LUTs
};

__device__ __forceinline__ void count_value(ImgState* s, int c, uint32_t v) { atomicAdd(&s->hist[c][v & 255u], 1u); }

// =========================================================================== scalar executor
template <int C, bool COUNT>
__device__ void exec_generic(const TC<C>& c, const Rect r) {
  const TileState& t = c.s->t;
  const int H = c.H, W = c.W;
  const int rw = r.x1 - r.x0;
  const int n = rw * (r.y1 - r.y0);
  const int kmode = t.kmode;
  const float f = t.kfactor;
  const uint8_t* src = c.src;
  for (int i = c.tid; i < n; i += NT) {
    const int ry = i / rw;
    const int y = r.y0 + ry, x = r.x0 + (i - ry * rw);
    int v[C];
    if (kmode == K_NONE || kmode == K_COLOR) {
      int sx = x, sy = y;
      const int k = resolve(t.sp, t.n_sp, H, W, sx, sy);
      if (k >= 0) {
        if (COUNT) { atomicAdd(&c.s->color_cnt[k], 1u); continue; }
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = t.sp[k].color[ch];
      } else {
        const uint8_t* px = src + ((size_t)sy * W + sx) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = __ldg(px + ch);
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
          for (int ch = 0; ch < C; ++ch) count_value(c.s, ch, (uint32_t)v[ch]);
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
        return t.l1[ch][__ldg(src + ((size_t)sy * W + sx) * C + ch)];
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
        for (int ch = 0; ch < C; ++ch) count_value(c.s, ch, (uint32_t)v[ch]);
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
          return (float)t.l1[ch][__ldg(src + ((size_t)iy * W + ix) * C + ch)];
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
        for (int ch = 0; ch < C; ++ch) count_value(c.s, ch, (uint32_t)v[ch]);
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

// Lane-private histogram: word (v >> 2) * 32 + lane of channel c holds four u8 counters (bins
// v & ~3 .. v | 3) fed only by lane `lane` of any warp, so a lane always hits its own bank and no
// two lanes of an instruction ever share an address.  A counter may take at most 255 increments
// between two reductions.
__device__ __forceinline__ void lhist_add(uint32_t ch_lane_base, uint32_t v) {
  reds_add(ch_lane_base + ((v & 0xFCu) << 5), 1u << ((v & 3u) << 3));
}
template <int C>
__device__ __forceinline__ void lhist_zero(uint32_t big, int tid) {
  for (int i = tid; i < C * LHIST_CH_BYTES / 16; i += NT) sts_v4(big + i * 16, make_uint4(0u, 0u, 0u, 0u));
}
// Adds the lane-private counters into s->hist (shared u32 bins).
template <int C>
__device__ __forceinline__ void lhist_reduce(ImgState* s, uint32_t big, int tid) {
  const int lane = tid & 31;
  for (int t = tid; t < C * 64; t += NT) {
    const int ch = t >> 6, row = t & 63;
    const uint32_t base = big + ch * LHIST_CH_BYTES + row * 128;
    uint32_t lo = 0, hi = 0;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) {
      const uint32_t w = lds_u32(base + (((l + lane) & 31) << 2));
      lo += w & 0x00FF00FFu;
      hi += (w >> 8) & 0x00FF00FFu;
    }
    uint32_t* h = &s->hist[ch][row * 4];
    h[0] += lo & 0xFFFFu; h[1] += hi & 0xFFFFu; h[2] += lo >> 16; h[3] += hi >> 16;
  }
}

// =============================================================================== flat executor
// No spatial op pending, K in {none, Color}: the tile is a contiguous run of units (48 bytes = 16
// pixels for C == 3, else 16 bytes) whose channel phase is a compile-time constant.
template <int C, bool COUNT>
__device__ void exec_flat(const TC<C>& c) {
  constexpr int UW = (C == 3) ? 12 : 4;
  constexpr int UB = UW * 4;
  const TileState& t = c.s->t;
  const int n_tiles = c.p->n_tiles;
  const int n_units = c.img_bytes / UB;
  const int upt = (n_units + n_tiles - 1) / n_tiles;
  const int u0 = min(n_units, c.tile * upt), u1 = min(n_units, u0 + upt);
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  const uint32_t hl = c.big + (c.lane << 2);

  for (int base = u0; base < u1; base += NT) {  // one unit per thread per round
    const int u = base + c.tid;
    if (COUNT) {
      lhist_zero<C>(c.big, c.tid);
      __syncthreads();
    }
    if (u < u1) {
      uint32_t w[UW];
      const uint8_t* sp = c.src + (size_t)u * UB;
#pragma unroll
      for (int q = 0; q < UW / 4; ++q) {
        const uint4 v = ldg_stream(sp + q * 16);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      if (kmode == K_NONE) {
        if (!COUNT && use1) map_unit<C, UW>(w, c.l1a);
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
          for (int b = 0; b < 4; ++b) lhist_add(hl + ((4 * j + b) % C) * LHIST_CH_BYTES, byte_of(w[j], b));
      } else {
        uint8_t* dp = c.dst + (size_t)u * UB;
#pragma unroll
        for (int q = 0; q < UW / 4; ++q)
          *reinterpret_cast<uint4*>(dp + q * 16) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      }
    }
    if (COUNT) {
      __syncthreads();
      lhist_reduce<C>(c.s, c.big, c.tid);
      __syncthreads();
    }
  }
  // ragged tail of the image (whole pixels, fewer than one unit): last tile, one pixel per thread
  if (c.tile == n_tiles - 1) {
    const int p0 = n_units * UB / C;
    for (int pix = p0 + c.tid; pix < c.HW; pix += NT) {
      int v[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = c.src[(size_t)pix * C + ch];
      if (kmode == K_NONE) {
        if (!COUNT && use1) {
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
#pragma unroll
      for (int ch = 0; ch < C; ++ch) {
        if (COUNT) count_value(c.s, ch, (uint32_t)v[ch]);
        else c.dst[(size_t)pix * C + ch] = (uint8_t)v[ch];
      }
    }
  }
}

// ============================================================================= gather executor
// Spatial list pending (constant-fill nearest warps and masks), K in {none, Color}.  Returns false
// (nothing done) when the tile's source bounding box does not fit the staging region.
// Requires W * C % 16 == 0 (rows are whole 16-byte units) and 16-byte aligned images.
template <int C, bool COUNT>
__device__ bool exec_gather(const TC<C>& c, const Rect r) {
  const TileState& t = c.s->t;
  const int H = c.H, W = c.W;
  const int n_sp = t.n_sp;
  // ---- source bounding box: push the tile's corners back through every warp of the list.  Affine
  // maps take extremes at corners; one pixel of margin per stage covers the rounding of that stage.
  float minx = (float)r.x0, maxx = (float)(r.x1 - 1), miny = (float)r.y0, maxy = (float)(r.y1 - 1);
  bool empty = false;
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
  int bx0 = 0, bx1 = -1, by0 = 0, by1 = -1;  // inclusive
  if (!empty) {
    bx0 = max(0, (int)floorf(minx)); bx1 = min(W - 1, (int)ceilf(maxx));
    by0 = max(0, (int)floorf(miny)); by1 = min(H - 1, (int)ceilf(maxy));
  }
  const int rows = by1 - by0 + 1;
  const int rowbytes = W * C;
  const int bxb0 = (bx0 * C) & ~15;
  const int bxb1 = min(rowbytes, ((bx1 + 1) * C + 15) & ~15);
  const int rowb = rows > 0 ? bxb1 - bxb0 : 0;
  const int pitch = rowb + 16;  // consecutive rows start 4 banks apart
  const int tw = r.x1 - r.x0, th = r.y1 - r.y0;
  const int out_bytes = COUNT ? 0 : tw * th * C;
  if ((long long)max(rows, 0) * pitch + out_bytes > BIG_BYTES) return false;
  const uint32_t ostage = c.big;
  const uint32_t stage = c.big + out_bytes;

  // ---- stage the bounding box
  if (rows > 0) {
    const int n16 = rowb >> 4;
    const int total = rows * n16;
    const uint8_t* sbase = c.src + (size_t)by0 * rowbytes + bxb0;
    for (int i = c.tid; i < total; i += NT) {
      const int rr = i / n16, q = i - rr * n16;
      sts_v4(stage + rr * pitch + (q << 4), ldg_stream(sbase + (size_t)rr * rowbytes + (q << 4)));
    }
  }
  __syncthreads();

  // ---- gather: one item = 4 consecutive pixels of a tile row
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  const bool single = (n_sp == 1 && t.sp[0].type == SP_GEOM);
  float t0 = 1.f, t1 = 0.f, t2 = 0.f, t3 = 0.f, t4 = 1.f, t5 = 0.f;
  uint32_t fill0 = 0;
  if (single) {
    const Spatial& e = t.sp[0];
    t0 = e.t[0]; t1 = e.t[1]; t2 = e.t[2]; t3 = e.t[3]; t4 = e.t[4]; t5 = e.t[5];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) fill0 |= (uint32_t)(e.color[ch] & 255) << (8 * ch);
  }
  const int qpr = tw >> 2;
  const int n_items = qpr * th;
  const uint32_t box_w = (uint32_t)(bx1 - bx0), box_h = (uint32_t)(by1 - by0);
  for (int it = c.tid; it < n_items; it += NT) {
    const int ry = it / qpr, rq = it - ry * qpr;
    const int y = r.y0 + ry;
    int x = r.x0 + (rq << 2);
    uint32_t o[C];
#pragma unroll
    for (int q = 0; q < C; ++q) o[q] = 0;
    const float fy = small_uint_to_float((uint32_t)y);
    const float t1y = __fmul_rn(t1, fy), t4y = __fmul_rn(t4, fy);
#pragma unroll
    for (int i = 0; i < 4; ++i, ++x) {
      int sx = x, sy = y, k = -1;
      if (single) {
        const float fx = small_uint_to_float((uint32_t)x);
        sx = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(t0, fx), t1y), t2));
        sy = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(t3, fx), t4y), t5));
        if (!((unsigned)sx < (unsigned)W && (unsigned)sy < (unsigned)H)) k = 0;
      } else {
        k = resolve(t.sp, n_sp, H, W, sx, sy);
      }
      uint32_t v[C];
      if (k < 0) {
        if ((uint32_t)(sx - bx0) <= box_w && (uint32_t)(sy - by0) <= box_h) {
          const uint32_t a = stage + (uint32_t)((sy - by0) * pitch + sx * C - bxb0);
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
        } else {  // never taken if the box is right; keeps a box error from becoming a wrong pixel
          const uint8_t* px = c.src + ((size_t)sy * W + sx) * C;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = __ldg(px + ch);
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
          for (int ch = 0; ch < C; ++ch) count_value(c.s, ch, v[ch]);
        }
      } else {
        if (COUNT) {
          atomicAdd(&c.s->color_cnt[k], 1u);
        } else if (single) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = byte_of(fill0, ch);
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
      const uint32_t oa = ostage + (uint32_t)((ry * tw + (rq << 2)) * C);
#pragma unroll
      for (int q = 0; q < C; ++q) sts_u32(oa + 4 * q, o[q]);
    }
  }
  if (!COUNT) {
    __syncthreads();
    const int n16 = (tw * C) >> 4;
    const int total = th * n16;
    uint8_t* dbase = c.dst + ((size_t)r.y0 * W + r.x0) * C;
    for (int i = c.tid; i < total; i += NT) {
      const int rr = i / n16, q = i - rr * n16;
      *reinterpret_cast<uint4*>(dbase + (size_t)rr * rowbytes + (q << 4)) = lds_v4(ostage + rr * (tw * C) + (q << 4));
    }
  }
  return true;
}

// ========================================================================== sharpness executor
// K == Sharpness with no spatial op pending: the tile is a strip of whole rows.  The strip plus one
// halo row each side is staged in shared memory with l1 already applied; a thread then owns one word
// column (4 bytes wide) of a sub-strip and walks down it, keeping the float32 products of the last
// two rows of its 4 + 2C-byte window in registers, so every input byte is converted and multiplied
// once per column instead of nine times.  Requires W * C % 16 == 0.  Returns false if the strip does
// not fit the staging region.
template <int C, bool COUNT>
__device__ bool exec_sharp(const TC<C>& c, const Rect r) {
  constexpr int NB = 4 + 2 * C;  // window bytes per row
  const TileState& t = c.s->t;
  const int H = c.H;
  const int row = c.W * C;
  const int sr0 = max(0, r.y0 - 1), sr1 = min(H, r.y1 + 1);
  const int nrows = sr1 - sr0;
  if (nrows <= 0) return true;
  if ((long long)nrows * row > BIG_BYTES) return false;
  const uint32_t stage = c.big;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  {  // stage rows [sr0, sr1) through l1; byte 16 i of the image has channel (16 i) % C
    const int total = (nrows * row) >> 4;
    const uint8_t* sbase = c.src + (size_t)sr0 * row;
    const int i0 = (sr0 * row) >> 4;
    for (int i = c.tid; i < total; i += NT) {
      uint4 v = ldg_stream(sbase + ((size_t)i << 4));
      if (use1) v = map_vec_phase<C>(v, c.l1a, (C == 3) ? ((i0 + i) % 3) : 0);
      sts_v4(stage + (i << 4), v);
    }
  }
  __syncthreads();
  const int wpr = row >> 2;  // words per row
  auto emit = [&](int y, int xw, uint32_t o) {
    const int ph = (C == 3) ? (xw % 3) : 0;  // channel of byte 0 of word xw (4 == 1 mod 3)
    if (COUNT) {
#pragma unroll
      for (int b = 0; b < 4; ++b) count_value(c.s, (ph + b) % C, byte_of(o, b));
    } else {
      if (use2)
        o = map_word(o, c.l2a + (uint32_t)((ph + 0) % C) * 256u, c.l2a + (uint32_t)((ph + 1) % C) * 256u,
                     c.l2a + (uint32_t)((ph + 2) % C) * 256u, c.l2a + (uint32_t)((ph + 3) % C) * 256u);
      *reinterpret_cast<uint32_t*>(c.dst + (size_t)y * row + ((size_t)xw << 2)) = o;
    }
  };
  // first and last image row: every pixel is border -> blend(orig, orig) == orig
  if (r.y0 == 0 && r.y1 > 0)
    for (int xw = c.tid; xw < wpr; xw += NT) emit(0, xw, lds_u32(stage + (uint32_t)((0 - sr0) * row + (xw << 2))));
  if (H > 1 && r.y0 <= H - 1 && r.y1 > H - 1)
    for (int xw = c.tid; xw < wpr; xw += NT) emit(H - 1, xw, lds_u32(stage + (uint32_t)((H - 1 - sr0) * row + (xw << 2))));
  const int in0 = max(r.y0, 1), in1 = min(r.y1, H - 1);
  const int inner = in1 - in0;
  if (inner <= 0) return true;
  const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
  const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
  // split the inner rows into S sub-strips so that (word columns x sub-strips) keeps every thread
  // busy: minimise rounds * (rows per sub-strip + 2 halo rows).
  int best_s = 1;
  long best_cost = 1L << 60;
  for (int S = 1; S <= 16 && S <= inner; ++S) {
    const long rounds = ((long)wpr * S + NT - 1) / NT;
    const long cost = rounds * ((inner + S - 1) / S + 2);
    if (cost < best_cost) { best_cost = cost; best_s = S; }
  }
  const int R = (inner + best_s - 1) / best_s;
  const int n_strips = (inner + R - 1) / R;
  const int n_items = wpr * n_strips;
  for (int item = c.tid; item < n_items; item += NT) {
    const int strip = item / wpr;
    const int xw = item - strip * wpr;
    const int y_begin = in0 + strip * R;
    const int y_end = min(in1, y_begin + R);
    const int xb0 = xw << 2;
    bool border[4];  // is the byte in the first / last pixel of the row?
#pragma unroll
    for (int b = 0; b < 4; ++b) border[b] = (xb0 + b < C) || (xb0 + b >= row - C);
    const bool has_prev = xw > 0, has_next = xw + 1 < wpr;
    float pa[NB], pb[NB];  // products (x k1) of rows y-1 and y
    float ctr[4];          // float values of the centre bytes of row y
    auto load_row = [&](int yy, float* pr, float* cvals) {
      const uint32_t ra = stage + (uint32_t)((yy - sr0) * row + xb0);
      const uint32_t w1 = lds_u32(ra);
      const uint32_t w0 = has_prev ? lds_u32(ra - 4) : 0u;
      const uint32_t w2 = has_next ? lds_u32(ra + 4) : 0u;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int wb = 4 - C + j;  // byte index in the 12-byte (w0, w1, w2) window
        const uint32_t wsel = (wb < 4) ? w0 : (wb < 8) ? w1 : w2;
        const float fv = byte_to_float(wsel, wb & 3);
        pr[j] = __fmul_rn(fv, k1);
        if (cvals != nullptr && j >= C && j < C + 4) cvals[j - C] = fv;
      }
    };
    load_row(y_begin - 1, pa, nullptr);
    load_row(y_begin, pb, ctr);
    for (int y = y_begin; y < y_end; ++y) {
      float pc[NB], nctr[4];
      load_row(y + 1, pc, nctr);
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float orig = ctr[b];
        float deg = orig;
        if (!border[b]) {
          // window index b + C is the byte itself, b / b + 2C its left / right neighbours
          float acc = pa[b];
          acc = __fadd_rn(acc, pa[b + C]);
          acc = __fadd_rn(acc, pa[b + 2 * C]);
          acc = __fadd_rn(acc, pb[b]);
          acc = __fadd_rn(acc, __fmul_rn(orig, k5));
          acc = __fadd_rn(acc, pb[b + 2 * C]);
          acc = __fadd_rn(acc, pc[b]);
          acc = __fadd_rn(acc, pc[b + C]);
          acc = __fadd_rn(acc, pc[b + 2 * C]);
          deg = __fadd_rn(__fadd_rz(acc, 8388608.0f), -8388608.0f);  // float(trunc(acc))
        }
        const uint32_t res = sharp_blend(deg, orig, f);
        o = (b == 0) ? res : put_byte(o, res, b);
      }
      emit(y, xw, o);
#pragma unroll
      for (int j = 0; j < NB; ++j) { pa[j] = pb[j]; pb[j] = pc[j]; }
#pragma unroll
      for (int b = 0; b < 4; ++b) ctr[b] = nctr[b];
    }
  }
  return true;
}

// ================================================================================ pass kernel
template <int C, bool COUNT>
__device__ void run_tile(const TC<C>& c) {
  const KParams& p = *c.p;
  const TileState& t = c.s->t;
  const int H = c.H, W = c.W;
  const int kmode = t.kmode;
  const bool rows16 = ((W * C) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.src) & 15) == 0 &&
                      (COUNT || (reinterpret_cast<uintptr_t>(c.dst) & 15) == 0);
  const bool fast = !p.force_generic;
  const int tile = c.tile;
  // tile -> region.  The partition depends only on per-image state, so all tiles of a pass agree.
  Rect strip;
  {
    const int R = (H + p.n_tiles - 1) / p.n_tiles;
    strip.x0 = 0; strip.x1 = W; strip.y0 = min(H, tile * R); strip.y1 = min(H, strip.y0 + R);
  }
  Rect box;
  {
    const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
    box.x0 = min(W, tx * p.tw); box.x1 = min(W, box.x0 + p.tw);
    box.y0 = min(H, ty * p.th); box.y1 = min(H, box.y0 + p.th);
  }
  if (kmode == K_NONE || kmode == K_COLOR) {
    if (t.n_sp == 0) {
      const bool ok = fast && (c.img_bytes & 15) == 0 && (reinterpret_cast<uintptr_t>(c.src) & 15) == 0 &&
                      (COUNT || (reinterpret_cast<uintptr_t>(c.dst) & 15) == 0);
      if (ok) exec_flat<C, COUNT>(c);
      else exec_generic<C, COUNT>(c, strip);
    } else {
      bool done = false;
      if (fast && rows16 && t.sp_fast) done = exec_gather<C, COUNT>(c, box);
      if (!done) exec_generic<C, COUNT>(c, box);
    }
  } else if (kmode == K_SHARP) {
    if (t.n_sp == 0) {
      bool done = false;
      if (fast && rows16) done = exec_sharp<C, COUNT>(c, strip);
      if (!done) exec_generic<C, COUNT>(c, strip);
    } else {
      exec_generic<C, COUNT>(c, box);
    }
  } else {
    exec_generic<C, COUNT>(c, strip);
  }
}

template <int C>
__global__ void __launch_bounds__(NT, PASS_MIN_CTAS) pass_kernel(const KParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ImgState* s = reinterpret_cast<ImgState*>(smem_raw);
  volatile int* ctl = reinterpret_cast<volatile int*>(smem_raw + sizeof(ImgState));
  uint8_t* bigp = smem_raw + sizeof(ImgState) + CTL_BYTES;
  const int tid = threadIdx.x;
  const int L = p.level;
  const int H = p.H, W = p.W;
  const int img_bytes = H * W * C;
  const unsigned n_entries = (L == 0) ? (unsigned)p.B : __ldcg(p.counters + L);
  const unsigned long long n_items = (unsigned long long)n_entries * (unsigned)p.n_tiles;

  TC<C> c;
  c.p = &p; c.s = s; c.big = smem_addr(bigp);
  c.H = H; c.W = W; c.HW = H * W; c.img_bytes = img_bytes; c.tid = tid; c.lane = tid & 31;
  c.l1a = smem_addr(&s->t.l1[0][0]); c.l2a = smem_addr(&s->t.l2[0][0]);

  for (;;) {
    __syncthreads();  // the previous item is completely done with shared memory
    if (tid == 0) ctl[0] = (int)atomicAdd(p.counters + p.max_levels + L, 1u);
    __syncthreads();
    const unsigned item = (unsigned)ctl[0];
    if (item >= n_items) break;
    const int entry = (int)(item / (unsigned)p.n_tiles);
    const int tile = (int)(item - (unsigned)entry * (unsigned)p.n_tiles);
    const int img = (L == 0) ? entry : __ldcg(p.lists + (size_t)L * p.B + entry);
    ImgState* g = p.states + img;
    for (int i = tid; i < TILE_VECS; i += NT)
      reinterpret_cast<uint4*>(s)[i] = __ldcg(reinterpret_cast<const uint4*>(g) + i);
    __syncthreads();
    const int pass_kind = s->t.pass_kind;
    const size_t img_off = (size_t)img * img_bytes;
    const int src_sel = s->t.src_sel;
    c.src = (src_sel == 0) ? p.in + img_off : p.scratch + (size_t)(2 * (size_t)img + (src_sel - 1)) * p.scratch_stride;
    c.dst = (pass_kind == PASS_WRITE_OUT)
                ? p.out + img_off
                : p.scratch + (size_t)(2 * (size_t)img + (s->t.dst_sel - 1)) * p.scratch_stride;
    c.tile = tile;
    if (pass_kind == PASS_COUNT) {
      for (int i = tid; i < MAXC * 256; i += NT) (&s->hist[0][0])[i] = 0u;
      if (tid < CHB_MAX_CHAIN) s->color_cnt[tid] = 0u;
      __syncthreads();
      run_tile<C, true>(c);
      __syncthreads();
      for (int i = tid; i < C * 256; i += NT) {
        const uint32_t v = (&s->hist[0][0])[i];
        if (v) atomicAdd(&g->hist[0][0] + i, v);
      }
      if (tid < CHB_MAX_CHAIN && s->color_cnt[tid]) atomicAdd(&g->color_cnt[tid], s->color_cnt[tid]);
    } else {
      run_tile<C, false>(c);
    }
    if (pass_kind == PASS_WRITE_OUT) continue;

    // ---- COUNT / WRITE_SCRATCH: the last tile of the image resumes the chain walk
    __threadfence();
    __syncthreads();
    if (tid == 0) ctl[1] = (atomicAdd(&g->tiles_done, 1u) == (unsigned)p.n_tiles - 1u) ? 1 : 0;
    __syncthreads();
    if (!ctl[1]) continue;
    __threadfence();
    for (int i = tid; i < (int)(sizeof(ImgState) / 16) - TILE_VECS; i += NT)
      reinterpret_cast<uint4*>(s)[TILE_VECS + i] = __ldcg(reinterpret_cast<const uint4*>(g) + TILE_VECS + i);
    __syncthreads();
    if (pass_kind == PASS_COUNT) {
      if (tid == 0) s->hist_valid = 1;
    } else {
      if (tid == 0) s->t.src_sel = s->t.dst_sel;
      reset_view(s, tid, NT);
    }
    __syncthreads();
    advance(s, g, p, C, H, W, reinterpret_cast<uint32_t*>(bigp), bigp + MAXC * 256 * 4, tid, NT);
    for (int i = tid; i < STATE_VECS; i += NT)
      reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(s)[i];
    if (tid == 0 && L + 1 < p.max_levels) {
      const unsigned pos = atomicAdd(p.counters + L + 1, 1u);
      p.lists[(size_t)(L + 1) * p.B + pos] = img;
    }
  }
}

template <int C>
cudaError_t launch_pass_c(const KParams& p, int grid, cudaStream_t stream) {
  pass_kernel<C><<<grid, NT, PASS_SMEM, stream>>>(p);
  return cudaGetLastError();
}

template <int C>
cudaError_t configure_c() {
  return cudaFuncSetAttribute(pass_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PASS_SMEM);
}

}  // namespace

// Per-channel-count entry points, one translation unit each (chb_kernels_c<N>.cu) so the four
// instantiations compile in parallel.
#define CHB_DEFINE_CHANNEL_ENTRY(C)                                                                   \
  cudaError_t launch_pass_c##C(const KParams& p, int grid, cudaStream_t stream) {                     \
    return launch_pass_c<C>(p, grid, stream);                                                         \
  }                                                                                                   \
  cudaError_t configure_c##C() { return configure_c<C>(); }

}  // namespace chb
