// Fused RandAugment / AutoAugment policy kernel for sm_100a.
//
// One persistent CTA per SM walks the batch image by image.  An image that fits is bulk-copied
// (TMA, cp.async.bulk + mbarrier) into shared memory ONCE, the whole per-image op chain is evaluated
// against that resident copy, and the result is written to HBM ONCE -- one HBM read + one HBM write
// per image for any chain (SURVEY.md section 7 / BASELINE.json north_star).
//
// The chain is evaluated lazily.  The state of the "virtual image" is
//     raw bytes  +  one 256-entry LUT per channel  +  a list of spatial ops (nearest-neighbour
//     affine gathers and CutOut rectangles, each carrying the colour it contributes)
// and every op only edits that state:
//   * point-wise ops (Invert, Posterize, Solarize, SolarizeAdd, Brightness, Contrast) compose into
//     the LUT (and map the spatial colours): 768 table entries instead of H*W*C pixels;
//   * Equalize / AutoContrast histogram the virtual image (lane-replicated, bank-conflict-free
//     shared-memory atomics; warp-scan CDF) and compose their per-channel LUT the same way;
//   * Color is per pixel across channels: one in-place pass with the pending LUT fused in;
//   * nearest-neighbour Rotate / Shear / Translate and CutOut are appended to the spatial list;
//   * Sharpness and bilinear warps need a materialised neighbourhood: the virtual image is written
//     to a per-CTA scratch (L2-resident) and read back.
// The final pass pulls every output pixel through the spatial list, the raw image and the LUT and
// stores 16-byte vectors.
//
// Semantics follow /root/reference/chambers/augmentations/image_augmentations.py (cited per op) and
// the oracle in /oracle (which this file never calls).  All float32 arithmetic that feeds a
// truncation uses explicit round-to-nearest intrinsics so that no FMA contraction can change a
// result (TensorFlow's CPU kernels round after every op).
#include <cuda_runtime.h>
#include <stdint.h>

#include "chb_internal.h"

namespace chb {

namespace {

constexpr int MAXC = 4;
constexpr int HIST_WORDS_PER_CH = 128 * 32;  // 128 bin pairs x 32 lane replicas, u16x2 packed
constexpr int TAB_WORDS = 256 * 32;          // replicated LUT table: word = {lut0,lut1,lut2,lut3}[v]

enum { SP_GEOM = 0, SP_MASK = 1 };

struct Spatial {
  int type;
  int fill_mode;
  float t[8];
  int y0, y1, x0, x1;
  int color[MAXC];
};

struct ProgEntry {
  int op, negate, cy, cx;
};

struct Small {
  unsigned long long mbar;
  int n_prog, n_sp, lut_identity, next_img;
  ProgEntry prog[CHB_MAX_CHAIN];
  Spatial sp[CHB_MAX_CHAIN];
  uint8_t lut[MAXC][256];
  uint8_t etab[MAXC][256];
  unsigned int hmap[MAXC][256];
  DevOp ops[CHB_MAX_TABLE_OPS];
};

__host__ __device__ constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ constexpr size_t big_region_bytes(int C) {
  return (size_t)(C * HIST_WORDS_PER_CH > TAB_WORDS ? C * HIST_WORDS_PER_CH : TAB_WORDS) * 4;
}

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_addr(bar)),
      "r"(phase)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                          unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_addr(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint4 ld_cg(const uint4* p) { return __ldcg(p); }
__device__ __forceinline__ void st_stream(uint4* p, uint4 v) { __stcs(p, v); }

// -------------------------------------------------------------------------------------- Philox
// Philox4x32-10, Salmon et al. SC'11; twin of oracle/philox.py (checked against Random123's KATs).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// ------------------------------------------------------------------------- per-value semantics
// chambers' blend, image_augmentations.py:10-49, on one value (image1 = degenerate, image2 = x).
__device__ __forceinline__ int blend_value(int i1, int i2, float f, int mode) {
  if (mode == BLEND_IMAGE2) return i2;  // factor == 1.0, :30-31
  if (mode == BLEND_IMAGE1) return i1;  // factor == 0.0, :28-29
  const float a = (float)i1, b = (float)i2;
  float temp = __fadd_rn(a, __fmul_rn(f, __fsub_rn(b, a)));  // :36-40, one rounding per op
  if (mode == BLEND_EXTRAP) temp = fminf(fmaxf(temp, 0.0f), 255.0f);  // :49
  return ((int)temp) & 0xFF;  // truncating cast :45 / :49
}

// Point-wise ops as functions of one uint8 value.
__device__ __forceinline__ int pointwise_value(const DevOp& op, int v) {
  switch (op.kind) {
    case CHB_OP_INVERT:  // :113
      return 255 - v;
    case CHB_OP_POSTERIZE:  // :172-173, shift pre-clamped like TF's shift functors
      return op.ip0 >= 8 ? 0 : ((v >> op.ip0) << op.ip0);
    case CHB_OP_SOLARIZE:  // :193, threshold pre-wrapped to uint8
      return v < op.ip0 ? v : 255 - v;
    case CHB_OP_SOLARIZE_ADD: {  // :213-215
      const int a = min(255, max(0, v + op.ip0));
      return v < op.ip1 ? a : v;
    }
    case CHB_OP_BRIGHTNESS:  // :284-285 blend(zeros, x, factor)
      return blend_value(0, v, op.factor, op.blend_mode);
    case CHB_OP_CONTRAST:  // :253-265; the degenerate image is a constant
      return blend_value(op.ip0, v, op.factor, op.blend_mode);
    default:
      return v;
  }
}

// Color (:233-235): blend(grayscale(x) broadcast, x, factor); tf.image.rgb_to_grayscale restated
// (oracle/ops.py rgb_to_grayscale).
__device__ __forceinline__ void color_pixel(int& r, int& g, int& b, float f, int mode) {
  const float k = __int_as_float(0x3b808081);  // float32(1/255)
  const float fr = __fmul_rn((float)r, k), fg = __fmul_rn((float)g, k), fb = __fmul_rn((float)b, k);
  float s = __fmul_rn(fr, __int_as_float(0x3e99096c));           // 0.2989
  s = __fadd_rn(s, __fmul_rn(fg, __int_as_float(0x3f1645a2)));   // 0.5870
  s = __fadd_rn(s, __fmul_rn(fb, __int_as_float(0x3de978d5)));   // 0.1140
  const int gray = ((int)__fmul_rn(s, 255.5f)) & 0xFF;
  r = blend_value(gray, r, f, mode);
  g = blend_value(gray, g, f, mode);
  b = blend_value(gray, b, f, mode);
}

// std::round (half away from zero) on float32, exactly (oracle/ops.py round_half_away).
__device__ __forceinline__ float round_half_away(float v) {
  const float r = truncf(v);
  const float d = __fsub_rn(v, r);
  return r + (d >= 0.5f ? 1.0f : 0.0f) - (d <= -0.5f ? 1.0f : 0.0f);
}

// image_ops.h MapCoordinate for the non-constant fill modes (oracle/ops.py _map_coordinate).
__device__ __forceinline__ float map_coordinate(float c, int n, int mode) {
  if (mode == CHB_FILL_CONSTANT) return c;
  const float hi = (float)(n - 1);
  if (mode == CHB_FILL_NEAREST) return fminf(fmaxf(c, 0.0f), hi);
  float o = c;
  if (mode == CHB_FILL_REFLECT) {
    if (c < 0.0f) {
      if (n <= 1) {
        o = 0.0f;
      } else {
        const float sz2 = (float)(2 * n);
        float v = c;
        if (v < sz2) v = __fadd_rn(__fmul_rn(sz2, truncf(__fdiv_rn(-v, sz2))), v);
        o = (v < (float)(-n)) ? __fadd_rn(v, sz2) : __fsub_rn(-v, 1.0f);
      }
    } else if (c > hi) {
      if (n <= 1) {
        o = 0.0f;
      } else {
        const float sz2 = (float)(2 * n);
        float w = __fsub_rn(c, __fmul_rn(sz2, truncf(__fdiv_rn(c, sz2))));
        o = (w >= (float)n) ? __fsub_rn(__fsub_rn(sz2, w), 1.0f) : w;
      }
    }
  } else {  // wrap
    if (c < 0.0f) {
      if (n <= 1) {
        o = 0.0f;
      } else {
        const float sz = hi;
        o = __fadd_rn(c, __fmul_rn((float)n, __fadd_rn(truncf(__fdiv_rn(-c, sz)), 1.0f)));
      }
    } else if (c > hi) {
      if (n <= 1) {
        o = 0.0f;
      } else {
        const float sz = hi;
        o = __fsub_rn(c, __fmul_rn((float)n, truncf(__fdiv_rn(c, sz))));
      }
    }
  }
  return fminf(fmaxf(o, 0.0f), hi);
}

// Source coordinate of output pixel (x, y): ProjectiveGenerator (image_ops.h) with t6 = t7 = 0.
__device__ __forceinline__ void affine_source(const float* t, int x, int y, float& sx, float& sy) {
  const float fx = (float)x, fy = (float)y;
  sx = __fadd_rn(__fadd_rn(__fmul_rn(t[0], fx), __fmul_rn(t[1], fy)), t[2]);
  sy = __fadd_rn(__fadd_rn(__fmul_rn(t[3], fx), __fmul_rn(t[4], fy)), t[5]);
}

// Walks the spatial list from the last op to the first.  Returns -1 when the output pixel resolves
// to raw pixel (x, y) (updated in place), else the index of the spatial entry whose colour it shows.
__device__ __forceinline__ int pull_resolve(const Small* s, int n_sp, int H, int W, int& x, int& y) {
  for (int k = n_sp - 1; k >= 0; --k) {
    const Spatial& e = s->sp[k];
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

template <int C>
struct Unit {
  static constexpr int BYTES = (C == 3) ? 48 : 16;
  static constexpr int WORDS = BYTES / 4;
  static constexpr int VECS = BYTES / 16;
};

__device__ __forceinline__ int get_byte(uint32_t w, int i) { return (w >> (8 * i)) & 0xFF; }

// The kernel's view of one image while its chain is interpreted.
template <int C, bool SMEM>
struct Ctx {
  Small* s;
  uint32_t* big;       // histogram replicas / replicated LUT table (aliased)
  uint8_t* simg;       // shared-memory image (SMEM only)
  const uint8_t* raw;  // current raw image: simg, the input image, or a scratch buffer
  uint8_t* scrA;
  uint8_t* scrB;
  int H, W, HW, img_bytes;
  int tid, nthreads, lane;
  bool tab_valid;

  // In the shared-memory variant every read of the image is provably a shared-space load.
  __device__ __forceinline__ const uint8_t* img_src() const { return SMEM ? simg : raw; }
  __device__ __forceinline__ int raw_at(int idx) const { return img_src()[idx]; }

  __device__ __forceinline__ int lut_byte(int c, int v) const {
    // bank-conflict-free lookup: every lane reads its own replica (bank == lane).
    return (big[v * 32 + lane] >> (8 * c)) & 0xFF;
  }

  // (Re)build the replicated table from the compact per-channel LUTs.
  __device__ void build_table() {
    for (int i = tid; i < TAB_WORDS; i += nthreads) {
      const int v = i >> 5;
      uint32_t w = 0;
#pragma unroll
      for (int c = 0; c < C; ++c) w |= (uint32_t)s->lut[c][v] << (8 * c);
      big[i] = w;
    }
    __syncthreads();
    tab_valid = true;
  }

  __device__ uint8_t* free_scratch() const { return (raw == scrA) ? scrB : scrA; }

  // ---- pass: dst[i] = LUT[channel(i)][raw[i]] for the whole image (no spatial ops pending).
  __device__ void map_pass(uint8_t* dst, bool identity) {
    using U = Unit<C>;
    if (!identity && !tab_valid) build_table();
    const int n_units = img_bytes / U::BYTES;
    const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int u = tid; u < n_units; u += nthreads) {
      uint32_t w[U::WORDS];
      const uint4* src = reinterpret_cast<const uint4*>(img_src() + (size_t)u * U::BYTES);
#pragma unroll
      for (int q = 0; q < U::VECS; ++q) {
        const uint4 v = src[q];
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      if (!identity) {
#pragma unroll
        for (int j = 0; j < U::WORDS; ++j) {
          uint32_t o = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int c = (4 * j + b) % C;
            o |= (uint32_t)lut_byte(c, get_byte(w[j], b)) << (8 * b);
          }
          w[j] = o;
        }
      }
      uint8_t* d = dst + (size_t)u * U::BYTES;
      if (dst_vec) {
#pragma unroll
        for (int q = 0; q < U::VECS; ++q)
          reinterpret_cast<uint4*>(d)[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < U::WORDS; ++j)
#pragma unroll
          for (int b = 0; b < 4; ++b) d[4 * j + b] = (uint8_t)get_byte(w[j], b);
      }
    }
    for (int i = n_units * U::BYTES + tid; i < img_bytes; i += nthreads) {
      const int v = raw_at(i);
      dst[i] = (uint8_t)(identity ? v : lut_byte(i % C, v));
    }
  }

  // ---- pass: Color with the pending LUT fused in; dst may alias raw (per-pixel op).
  __device__ void color_pass(uint8_t* dst, float f, int mode, bool identity) {
    static_assert(C == 3 || C != 3, "");
    if (!identity && !tab_valid) build_table();
    if (C != 3) return;
    using U = Unit<3>;
    const int n_units = img_bytes / U::BYTES;
    for (int u = tid; u < n_units; u += nthreads) {
      uint32_t w[U::WORDS];
      const uint4* src = reinterpret_cast<const uint4*>(img_src() + (size_t)u * U::BYTES);
#pragma unroll
      for (int q = 0; q < U::VECS; ++q) {
        const uint4 v = src[q];
        w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      uint32_t o[U::WORDS];
#pragma unroll
      for (int j = 0; j < U::WORDS; ++j) o[j] = 0;
#pragma unroll
      for (int px = 0; px < 16; ++px) {
        int ch[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int bi = px * 3 + c;
          int v = get_byte(w[bi >> 2], bi & 3);
          if (!identity) v = lut_byte(c, v);
          ch[c] = v;
        }
        color_pixel(ch[0], ch[1], ch[2], f, mode);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int bi = px * 3 + c;
          o[bi >> 2] |= (uint32_t)ch[c] << (8 * (bi & 3));
        }
      }
      uint4* d = reinterpret_cast<uint4*>(dst + (size_t)u * U::BYTES);
#pragma unroll
      for (int q = 0; q < U::VECS; ++q) d[q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    }
    const int first_px = n_units * 16;
    for (int px = first_px + tid; px < HW; px += nthreads) {
      int ch[3];
      for (int c = 0; c < 3; ++c) {
        const int v = raw_at(px * 3 + c);
        ch[c] = identity ? v : lut_byte(c, v);
      }
      color_pixel(ch[0], ch[1], ch[2], f, mode);
      for (int c = 0; c < 3; ++c) dst[px * 3 + c] = (uint8_t)ch[c];
    }
  }

  // ---- pass: write the virtual image (spatial list + raw + LUT) to dst, 16 pixels per thread.
  __device__ void pull_pass(uint8_t* dst, bool identity) {
    if (!identity && !tab_valid) build_table();
    const int n_sp = s->n_sp;
    const int n_pu = (HW + 15) >> 4;
    const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int pu = tid; pu < n_pu; pu += nthreads) {
      const int pix0 = pu << 4;
      int y = pix0 / W;
      int x = pix0 - y * W;
      uint32_t o[4 * C];
#pragma unroll
      for (int j = 0; j < 4 * C; ++j) o[j] = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (pix0 + i < HW) {
          int sx = x, sy = y;
          const int k = pull_resolve(s, n_sp, H, W, sx, sy);
          const int base = (sy * W + sx) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            int v;
            if (k < 0) {
              v = raw_at(base + c);
              if (!identity) v = lut_byte(c, v);
            } else {
              v = s->sp[k].color[c];
            }
            const int bi = i * C + c;
            o[bi >> 2] |= (uint32_t)v << (8 * (bi & 3));
          }
        }
        if (++x == W) { x = 0; ++y; }
      }
      uint8_t* d = dst + (size_t)pix0 * C;
      if (dst_vec && pix0 + 16 <= HW) {
#pragma unroll
        for (int q = 0; q < C; ++q)
          st_stream(reinterpret_cast<uint4*>(d) + q, make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
      } else {
        const int nb = min(16, HW - pix0) * C;
#pragma unroll
        for (int b = 0; b < 16 * C; ++b)
          if (b < nb) d[b] = (uint8_t)get_byte(o[b >> 2], b & 3);
      }
    }
  }

  // ---- histogram of the virtual image into s->hmap[c][256].
  __device__ __forceinline__ void hist_add(int c, int v) {
    atomicAdd(&big[c * HIST_WORDS_PER_CH + ((v >> 1) << 5) + lane], 1u << ((v & 1) << 4));
  }

  __device__ void histogram(bool identity) {
    using U = Unit<C>;
    const int n_sp = s->n_sp;
    // A pending LUT and a pull pass both need the table, which aliases the histogram: when spatial
    // ops are pending, count RAW values for resolved pixels and final colours separately, then map.
    for (int i = tid; i < C * HIST_WORDS_PER_CH; i += nthreads) big[i] = 0;
    for (int i = tid; i < MAXC * 256; i += nthreads) (&s->hmap[0][0])[i] = 0;
    tab_valid = false;
    __syncthreads();
    if (n_sp == 0) {
      const int n_units = img_bytes / U::BYTES;
      for (int u = tid; u < n_units; u += nthreads) {
        const uint4* src = reinterpret_cast<const uint4*>(img_src() + (size_t)u * U::BYTES);
#pragma unroll
        for (int q = 0; q < U::VECS; ++q) {
          const uint4 v = src[q];
          const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) hist_add((16 * q + 4 * j + b) % C, get_byte(w[j], b));
        }
      }
      for (int i = n_units * U::BYTES + tid; i < img_bytes; i += nthreads) hist_add(i % C, raw_at(i));
    } else {
      for (int pix = tid; pix < HW; pix += nthreads) {
        int y = pix / W;
        int x = pix - y * W;
        const int k = pull_resolve(s, n_sp, H, W, x, y);
        if (k < 0) {
          const int base = (y * W + x) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) hist_add(c, raw_at(base + c));
        } else {
          // colours are already final values: count them straight into the mapped histogram.
#pragma unroll
          for (int c = 0; c < C; ++c) atomicAdd(&s->hmap[c][s->sp[k].color[c]], 1u);
        }
      }
    }
    __syncthreads();
    // reduce the 32 lane replicas, mapping raw values through the pending LUT.
    for (int t = tid; t < C * 256; t += nthreads) {
      const int c = t >> 8, v = t & 255;
      const uint32_t* row = big + c * HIST_WORDS_PER_CH + ((v >> 1) << 5);
      const int sh = (v & 1) << 4;
      unsigned int total = 0;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) total += (row[(l + lane) & 31] >> sh) & 0xFFFFu;
      if (total) atomicAdd(&s->hmap[c][identity ? v : s->lut[c][v]], total);
    }
    __syncthreads();
  }

  // ---- Equalize / AutoContrast: per-channel table from the histogram, composed into the LUT.
  __device__ void stat_op(int kind, bool identity) {
    histogram(identity);
    const int warp = tid >> 5;
    if (warp < C) {
      const int c = warp;
      int h[8];
      int sum = 0, last = -1, first = 256;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[i] = (int)s->hmap[c][lane * 8 + i];
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
        const int last_count = (int)s->hmap[c][last < 0 ? 0 : last];
        const int step = (total - last_count) / 255;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int v = lane * 8 + i;
          int e = v;
          if (step != 0) e = min(255, max(0, (excl + step / 2) / step));
          s->etab[c][v] = (uint8_t)e;
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
          s->etab[c][v] = (uint8_t)(int)xv;                          // :86
        }
      }
    }
    __syncthreads();
    for (int t = tid; t < C * 256; t += nthreads) {
      const int c = t >> 8, v = t & 255;
      s->lut[c][v] = s->etab[c][s->lut[c][v]];
    }
    for (int t = tid; t < s->n_sp * C; t += nthreads) {
      const int k = t / C, c = t - k * C;
      s->sp[k].color[c] = s->etab[c][s->sp[k].color[c]];
    }
    __syncthreads();
  }

  // ---- Sharpness (tfa.image.sharpness; oracle/ops.py sharpness): raw holds the materialised input.
  __device__ void sharpness_pass(uint8_t* dst, float f, int mode) {
    const int row = W * C;
    const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
    const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
    const int n16 = (img_bytes + 15) >> 4;
    const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int u = tid; u < n16; u += nthreads) {
      const int i0 = u << 4;
      int y = i0 / row;
      int xb = i0 - y * row;
      uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
      for (int b = 0; b < 16; ++b) {
        const int i = i0 + b;
        if (i < img_bytes) {
          const int orig = raw_at(i);
          int deg = orig;
          const int x = xb / C;
          if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
            float acc = 0.0f;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                const float kk = (dy == 0 && dx == 0) ? k5 : k1;
                acc = __fadd_rn(acc, __fmul_rn((float)raw_at(i + dy * row + dx * C), kk));
              }
            deg = ((int)acc) & 0xFF;
          }
          int r;
          if (mode == BLEND_IMAGE1) {
            r = deg;
          } else {
            const float a = (float)deg;
            float t = __fadd_rn(a, __fmul_rn(f, __fsub_rn((float)orig, a)));
            t = rintf(fminf(fmaxf(t, 0.0f), 255.0f));  // TFA blend: clip, round half-to-even
            r = (int)t;
          }
          o[b >> 2] |= (uint32_t)r << (8 * (b & 3));
        }
        if (++xb == row) { xb = 0; ++y; }
      }
      if (dst_vec && i0 + 16 <= img_bytes) {
        st_stream(reinterpret_cast<uint4*>(dst + i0), make_uint4(o[0], o[1], o[2], o[3]));
      } else {
        const int nb = min(16, img_bytes - i0);
#pragma unroll
        for (int b = 0; b < 16; ++b)
          if (b < nb) dst[i0 + b] = (uint8_t)get_byte(o[b >> 2], b & 3);
      }
    }
  }

  // ---- bilinear warp (image_ops.h bilinear_interpolation; oracle/ops.py projective_transform).
  __device__ void bilinear_pass(uint8_t* dst, const float* t, int fill_mode, int fill, bool identity) {
    if (!identity && !tab_valid) build_table();
    for (int pix = tid; pix < HW; pix += nthreads) {
      const int y = pix / W, x = pix - y * W;
      float sx, sy;
      affine_source(t, x, y, sx, sy);
      sx = map_coordinate(sx, W, fill_mode);
      sy = map_coordinate(sy, H, fill_mode);
      const float xf = floorf(sx), yf = floorf(sy);
      const float xc = __fadd_rn(xf, 1.0f), yc = __fadd_rn(yf, 1.0f);
      const float wxf = __fsub_rn(xc, sx), wxc = __fsub_rn(sx, xf);
      const float wyf = __fsub_rn(yc, sy), wyc = __fsub_rn(sy, yf);
      const bool inx0 = (xf >= 0.0f && xf < (float)W), inx1 = (xc >= 0.0f && xc < (float)W);
      const bool iny0 = (yf >= 0.0f && yf < (float)H), iny1 = (yc >= 0.0f && yc < (float)H);
      const int ix0 = inx0 ? (int)xf : 0, ix1 = inx1 ? (int)xc : 0;
      const int iy0 = iny0 ? (int)yf : 0, iy1 = iny1 ? (int)yc : 0;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        auto tap = [&](bool in, int iy, int ix) -> float {
          if (!in) return (float)fill;
          const int v = raw_at((iy * W + ix) * C + c);
          return (float)(identity ? v : lut_byte(c, v));
        };
        const float v00 = tap(iny0 && inx0, iy0, ix0), v01 = tap(iny0 && inx1, iy0, ix1);
        const float v10 = tap(iny1 && inx0, iy1, ix0), v11 = tap(iny1 && inx1, iy1, ix1);
        const float vyf = __fadd_rn(__fmul_rn(wxf, v00), __fmul_rn(wxc, v01));
        const float vyc = __fadd_rn(__fmul_rn(wxf, v10), __fmul_rn(wxc, v11));
        const float val = __fadd_rn(__fmul_rn(wyf, vyf), __fmul_rn(wyc, vyc));
        dst[pix * C + c] = (uint8_t)(((int)val) & 0xFF);
      }
    }
  }

  // ---- bring a global buffer back as the raw image.
  __device__ void adopt(const uint8_t* buf) {
    __threadfence();
    __syncthreads();
    if (SMEM) {
      const int n16 = img_bytes >> 4;
      for (int i = tid; i < n16; i += nthreads)
        reinterpret_cast<uint4*>(simg)[i] = ld_cg(reinterpret_cast<const uint4*>(buf) + i);
      for (int i = (n16 << 4) + tid; i < img_bytes; i += nthreads) simg[i] = __ldcg(buf + i);
      raw = simg;
    } else {
      raw = buf;
    }
    __syncthreads();
  }

  __device__ void reset_lut() {
    for (int t = tid; t < C * 256; t += nthreads) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
    if (tid == 0) { s->lut_identity = 1; s->n_sp = 0; }
    tab_valid = false;
    __syncthreads();
  }

  // ---- write the virtual image to a scratch buffer and adopt it (LUT and spatial list reset).
  __device__ void materialize() {
    uint8_t* dst = free_scratch();
    const bool identity = s->lut_identity != 0;
    if (s->n_sp > 0) pull_pass(dst, identity); else map_pass(dst, identity);
    adopt(dst);
    reset_lut();
  }
};

// ---------------------------------------------------------------------------- schedule decode
// Twin of oracle/philox.py decode_schedule for ONE image; executed by one thread.
__device__ void decode_image(const KParams& p, const DevOp* ops, Small* s, int img, int H, int W) {
  const unsigned long long own = p.image_index_base + (unsigned long long)img;
  const unsigned long long stream_img = p.elementwise ? own : ~0ull;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  int n = 0;
  for (int i = 0; i < p.n_draws; ++i) {
    const uint32_t slot0 = (uint32_t)(i * (p.K + 1));
    int choice;
    const size_t rbase = ((size_t)img * p.n_draws + i) * p.K * CHB_SCHED_FIELDS;
    if (p.replay) {
      choice = p.replay[rbase];
      if (choice < 0 || choice >= p.T) choice = 0;
    } else {
      const uint4 w = philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter, slot0), key);
      choice = (int)__umulhi(w.x, (uint32_t)p.T);
    }
    for (int j = 0; j < p.K; ++j) {
      const int opi = choice * p.K + j;
      const DevOp& op = ops[opi];
      int applied, negate, cy, cx;
      if (p.replay) {
        const int32_t* r = p.replay + rbase + (size_t)j * CHB_SCHED_FIELDS;
        applied = (r[1] != 0) && op.kind >= 0;
        negate = r[2] != 0;
        cy = r[3];
        cx = r[4];
      } else {
        const uint32_t slot = slot0 + 1 + (uint32_t)j;
        const uint4 w = philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter, slot), key);
        uint4 wc = w;
        if (!p.elementwise) wc = philox4x32_10(make_uint4((uint32_t)own, (uint32_t)(own >> 32), p.call_counter, slot), key);
        applied = (op.kind >= 0) && ((int)(w.x >> 8) < op.thr24);
        negate = w.y < 0x80000000u;
        cy = (int)__umulhi(wc.z, (uint32_t)H);
        cx = (int)__umulhi(wc.w, (uint32_t)W);
      }
      if (p.record) {
        int32_t* r = p.record + rbase + (size_t)j * CHB_SCHED_FIELDS;
        r[0] = choice; r[1] = applied; r[2] = negate; r[3] = cy; r[4] = cx;
      }
      if (applied && n < CHB_MAX_CHAIN) {
        s->prog[n].op = opi; s->prog[n].negate = negate; s->prog[n].cy = cy; s->prog[n].cx = cx;
        ++n;
      }
    }
  }
  s->n_prog = n;
}

// ------------------------------------------------------------------------------------- kernel
template <int C, bool SMEM>
__global__ void __launch_bounds__(1024, 1) policy_kernel(const KParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int H = p.H, W = p.W;
  const int img_bytes = H * W * C;
  const size_t img_pad = SMEM ? align_up((size_t)img_bytes, 128) : 0;
  uint8_t* simg = smem_raw;
  uint32_t* big = reinterpret_cast<uint32_t*>(smem_raw + img_pad);
  Small* s = reinterpret_cast<Small*>(smem_raw + img_pad + big_region_bytes(C));

  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int n_table = p.T * p.K;
  const bool ops_in_smem = n_table <= CHB_MAX_TABLE_OPS;
  if (ops_in_smem) {
    const int nwords = n_table * (int)(sizeof(DevOp) / 4);
    for (int i = tid; i < nwords; i += nthreads)
      reinterpret_cast<uint32_t*>(s->ops)[i] = reinterpret_cast<const uint32_t*>(p.ops)[i];
  }
  const DevOp* ops = ops_in_smem ? s->ops : p.ops;
  if (tid == 0) {
    mbar_init(&s->mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  Ctx<C, SMEM> cx;
  cx.s = s; cx.big = big; cx.simg = simg;
  cx.scrA = p.scratch + (size_t)blockIdx.x * 2 * p.scratch_stride;
  cx.scrB = cx.scrA + p.scratch_stride;
  cx.H = H; cx.W = W; cx.HW = H * W; cx.img_bytes = img_bytes;
  cx.tid = tid; cx.nthreads = nthreads; cx.lane = tid & 31;
  uint32_t phase = 0;

  int img = blockIdx.x;
  while (img < p.B) {
    const uint8_t* in_img = p.in + (size_t)img * img_bytes;
    uint8_t* out_img = p.out + (size_t)img * img_bytes;
    const bool in_vec = ((reinterpret_cast<uintptr_t>(in_img) & 15) == 0);
    bool tma_pending = false;

    // 1. start the image on its way into shared memory, then decode the schedule under it.
    if (SMEM) {
      if (in_vec && (img_bytes & 15) == 0) {
        if (tid == 0) {
          fence_proxy_async();
          mbar_expect_tx(&s->mbar, (uint32_t)img_bytes);
          const uint32_t chunk = 32768;
          for (uint32_t off = 0; off < (uint32_t)img_bytes; off += chunk)
            bulk_load(simg + off, in_img + off, min(chunk, (uint32_t)img_bytes - off), &s->mbar);
        }
        tma_pending = true;
      } else {
        for (int i = tid; i < img_bytes; i += nthreads) simg[i] = in_img[i];
      }
      cx.raw = simg;
    } else {
      if (in_vec) {
        cx.raw = in_img;
      } else {
        for (int i = tid; i < img_bytes; i += nthreads) cx.scrA[i] = in_img[i];
        __threadfence();
        cx.raw = cx.scrA;
      }
    }
    if (tid == 0) {
      decode_image(p, ops, s, img, H, W);
      s->next_img = (int)gridDim.x + (int)atomicAdd(p.work_counter, 1u);
      s->n_sp = 0;
      s->lut_identity = 1;
    }
    for (int t = tid; t < C * 256; t += nthreads) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
    cx.tab_valid = false;
    __syncthreads();

    const int n_prog = s->n_prog;
    bool emitted = false;
    auto ensure_loaded = [&]() {
      if (tma_pending) {
        mbar_wait(&s->mbar, phase);
        phase ^= 1;
        tma_pending = false;
      }
    };

    // 2. interpret the chain (uniform across the CTA).
    for (int pi = 0; pi < n_prog; ++pi) {
      const ProgEntry pe = s->prog[pi];
      const DevOp& op = ops[pe.op];
      const bool last = (pi == n_prog - 1);
      const int kind = op.kind;
      switch (kind) {
        case CHB_OP_INVERT: case CHB_OP_POSTERIZE: case CHB_OP_SOLARIZE: case CHB_OP_SOLARIZE_ADD:
        case CHB_OP_BRIGHTNESS: case CHB_OP_CONTRAST: {
          for (int t = tid; t < C * 256; t += nthreads)
            s->lut[t >> 8][t & 255] = (uint8_t)pointwise_value(op, s->lut[t >> 8][t & 255]);
          for (int t = tid; t < s->n_sp * C; t += nthreads) {
            const int k = t / C, c = t - k * C;
            s->sp[k].color[c] = pointwise_value(op, s->sp[k].color[c]);
          }
          cx.tab_valid = false;
          __syncthreads();
          if (tid == 0) s->lut_identity = 0;
          __syncthreads();
        } break;
        case CHB_OP_COLOR: {
          if (op.blend_mode == BLEND_IMAGE2 || C != 3) break;
          ensure_loaded();
          const bool identity = s->lut_identity != 0;
          uint8_t* dst = SMEM ? simg : cx.free_scratch();
          cx.color_pass(dst, op.factor, op.blend_mode, identity);
          if (!SMEM) { __threadfence(); cx.raw = dst; }
          __syncthreads();
          // the LUT is now baked in; spatial colours are pixels too.
          for (int t = tid; t < C * 256; t += nthreads) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
          if (tid < s->n_sp) {
            int* col = s->sp[tid].color;
            color_pixel(col[0], col[1], col[2], op.factor, op.blend_mode);
          }
          if (tid == 0) s->lut_identity = 1;
          cx.tab_valid = false;
          __syncthreads();
        } break;
        case CHB_OP_AUTOCONTRAST: case CHB_OP_EQUALIZE: {
          ensure_loaded();
          cx.stat_op(kind, s->lut_identity != 0);
          if (tid == 0) s->lut_identity = 0;
          cx.tab_valid = false;
          __syncthreads();
        } break;
        case CHB_OP_CUTOUT: {
          const int h = op.ip0;
          const int y0 = max(0, pe.cy - h), y1 = min(H, pe.cy + h);
          const int x0 = max(0, pe.cx - h), x1 = min(W, pe.cx + h);
          if (SMEM && s->n_sp == 0) {
            // no gather pending: bake the LUT (if any) and paint the rectangle into the image.
            ensure_loaded();
            if (!s->lut_identity) {
              cx.map_pass(simg, false);
              __syncthreads();
              cx.reset_lut();
            }
            const int rw = (x1 - x0) * C, rh = y1 - y0;
            if (rw > 0 && rh > 0)
              for (int i = tid; i < rw * rh; i += nthreads) {
                const int ry = i / rw, rb = i - ry * rw;
                simg[((y0 + ry) * W + x0) * C + rb] = (uint8_t)op.ip1;
              }
            __syncthreads();
          } else {
            if (tid == 0) {
              Spatial& e = s->sp[s->n_sp];
              e.type = SP_MASK; e.fill_mode = 0;
              e.y0 = y0; e.y1 = y1; e.x0 = x0; e.x1 = x1;
              for (int c = 0; c < MAXC; ++c) e.color[c] = op.ip1;
              s->n_sp = s->n_sp + 1;
            }
            __syncthreads();
          }
        } break;
        case CHB_OP_SHEAR_X: case CHB_OP_SHEAR_Y: case CHB_OP_TRANSLATE_X: case CHB_OP_TRANSLATE_Y:
        case CHB_OP_ROTATE: {
          const float* t = op.coef[pe.negate ? 1 : 0];
          if (op.interp == CHB_INTERP_NEAREST) {
            if (tid == 0) {
              Spatial& e = s->sp[s->n_sp];
              e.type = SP_GEOM; e.fill_mode = op.fill_mode;
              for (int q = 0; q < 8; ++q) e.t[q] = t[q];
              for (int c = 0; c < MAXC; ++c) e.color[c] = op.fill_u8;
              s->n_sp = s->n_sp + 1;
            }
            __syncthreads();
          } else {
            ensure_loaded();
            if (s->n_sp > 0) cx.materialize();
            uint8_t* dst = last ? out_img : cx.free_scratch();
            cx.bilinear_pass(dst, t, op.fill_mode, op.fill_u8, s->lut_identity != 0);
            if (last) { emitted = true; } else { cx.adopt(dst); cx.reset_lut(); }
          }
        } break;
        case CHB_OP_SHARPNESS: {
          if (op.blend_mode == BLEND_IMAGE2) break;
          ensure_loaded();
          if (s->n_sp > 0) {
            cx.materialize();
          } else if (!s->lut_identity) {
            if (SMEM) {
              cx.map_pass(simg, false);
              __syncthreads();
              cx.reset_lut();
            } else {
              cx.materialize();
            }
          }
          uint8_t* dst = last ? out_img : cx.free_scratch();
          cx.sharpness_pass(dst, op.factor, op.blend_mode);
          if (last) { emitted = true; } else { cx.adopt(dst); cx.reset_lut(); }
        } break;
        default:
          break;
      }
    }

    // 3. one write of the finished image.
    if (!emitted) {
      ensure_loaded();
      const bool identity = s->lut_identity != 0;
      if (s->n_sp > 0) cx.pull_pass(out_img, identity); else cx.map_pass(out_img, identity);
    }
    __syncthreads();
    img = s->next_img;
    __syncthreads();
  }

  // last CTA out re-arms the work counter for the next launch.
  if (tid == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(p.work_counter + 1, 1u);
    if (done == gridDim.x - 1) {
      p.work_counter[0] = 0;
      p.work_counter[1] = 0;
      __threadfence();
    }
  }
}

template <int C>
cudaError_t launch_c(const KParams& p, const LaunchInfo& li, cudaStream_t stream) {
  if (li.image_in_smem)
    policy_kernel<C, true><<<li.grid, li.block, li.smem, stream>>>(p);
  else
    policy_kernel<C, false><<<li.grid, li.block, li.smem, stream>>>(p);
  return cudaGetLastError();
}

template <int C>
cudaError_t configure_c(size_t smem_optin) {
  cudaError_t e = cudaFuncSetAttribute(policy_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(policy_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin);
}

}  // namespace

// Per-channel-count entry points, one translation unit each (chb_kernels_c<N>.cu) so the four
// instantiations compile in parallel.
#define CHB_DEFINE_CHANNEL_ENTRY(C)                                                                  \
  cudaError_t launch_policy_c##C(const KParams& p, const LaunchInfo& li, cudaStream_t stream) {      \
    return launch_c<C>(p, li, stream);                                                               \
  }                                                                                                  \
  cudaError_t configure_c##C(size_t smem_optin) { return configure_c<C>(smem_optin); }               \
  size_t smem_overhead_c##C() { return big_region_bytes(C) + align_up(sizeof(Small), 128) + 128; }

}  // namespace chb
