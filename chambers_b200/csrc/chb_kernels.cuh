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
// The final pass either maps 16-byte vectors through the LUT straight to global memory, or gathers
// every output pixel through the spatial list into a shared-memory staging tile that a TMA bulk
// store (cp.async.bulk.global.shared) drains while the next tile is computed.
//
// All hot loops have small bodies (4 pixels / 16 bytes per trip): the first version unrolled 16
// pixels per trip and ncu showed instruction-cache misses ("no_instruction") as its top stall
// (profiles/r01_v2_*).
//
// Semantics follow /root/reference/chambers/augmentations/image_augmentations.py (cited per op) and
// the oracle in /oracle (which this file never calls).  All float32 arithmetic that feeds a
// truncation uses explicit round-to-nearest intrinsics so that no FMA contraction can change a
// result (TensorFlow's CPU kernels round after every op).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "chb_internal.h"

namespace chb {

namespace {

constexpr int NT = 512;                      // threads per CTA (up to 128 registers each)
constexpr int MAXC = 4;
constexpr int HIST_WORDS_PER_CH = 128 * 32;  // 128 bin pairs x 32 lane replicas, u16x2 packed
constexpr int TAB_WORDS = 256 * 32;          // replicated LUT table: word = {lut0,lut1,lut2,lut3}[v]
constexpr int STAGE_PIX = NT * 4;            // pixels per staging tile: one 4-pixel group per thread

enum { SP_GEOM = 0, SP_MASK = 1 };

struct Spatial {
  int type;
  int fill_mode;
  float t[8];
  int y0, y1, x0, x1;
  int color[MAXC];
};

struct ProgEntry {
  DevOp op;  // private copy: the interpreter never goes back to global memory for parameters
  int negate, cy, cx, table_index;
};

struct Small {
  unsigned long long mbar;
  int n_prog, n_sp, lut_identity, next_img;
  uint32_t rnd[32][4];   // schedule decode scratch: Philox words per slot (stream / own image)
  uint32_t rndc[32][4];
  ProgEntry prog[CHB_MAX_CHAIN];
  Spatial sp[CHB_MAX_CHAIN];
  uint8_t lut[MAXC][256];
  uint8_t etab[MAXC][256];
  unsigned int hmap[MAXC][256];
};

__host__ __device__ constexpr size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
__host__ __device__ constexpr size_t big_region_bytes(int C) {
  return (size_t)(C * HIST_WORDS_PER_CH > TAB_WORDS ? C * HIST_WORDS_PER_CH : TAB_WORDS) * 4;
}
__host__ __device__ constexpr size_t stage_bytes(int C) { return (size_t)STAGE_PIX * C; }  // one buffer
__host__ __device__ constexpr size_t smem_overhead_bytes(int C) {
  return big_region_bytes(C) + 2 * stage_bytes(C) + align_up(sizeof(Small), 128) + 128;
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
// TMA 1-D bulk copy shared -> global, tracked by bulk async-groups.
__device__ __forceinline__ void bulk_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_addr(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {  // <= N groups still reading their shared source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {  // <= N groups not yet complete (writes visible)
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Explicit shared-space accesses on 32-bit shared addresses.  Going through generic pointers made
// the compiler rebuild a shared::cluster address (S2R SR_CgaCtaId + LEA) in front of every access.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) {  // fire-and-forget shared atomic add
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
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
// Table-building form (768 entries per op, not hot).
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
// (oracle/ops.py rgb_to_grayscale).  Integer form for the few spatial colours.
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

__device__ __forceinline__ int get_byte(uint32_t w, int i) { return (w >> (8 * i)) & 0xFF; }
// Byte i (compile-time) of w, zero-extended: one PRMT.
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int i) { return __byte_perm(w, 0u, 0x4440u | (uint32_t)i); }
// acc with byte i (compile-time) replaced by the low byte of v: one PRMT.
__device__ __forceinline__ uint32_t put_byte(uint32_t acc, uint32_t v, int i) {
  return __byte_perm(acc, v, i == 0 ? 0x3214u : i == 1 ? 0x3240u : i == 2 ? 0x3410u : 0x4210u);
}
// float(byte i of w) without a conversion-pipe instruction: 0x4B0000vv is 2^23 + v, minus 2^23.
__device__ __forceinline__ float byte_to_float(uint32_t w, int i) {
  return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u | (uint32_t)i)), -8388608.0f);
}
__device__ __forceinline__ float small_uint_to_float(uint32_t v) {  // v < 2^23, exact
  return __fadd_rn(__uint_as_float(0x4B000000u | v), -8388608.0f);
}
// trunc(t) for 0 <= t < 2^23 as the low bits of fl_rz(t + 2^23) (ulp there is 1).
__device__ __forceinline__ uint32_t trunc_bits(float t) { return __float_as_uint(__fadd_rz(t, 8388608.0f)); }
// rint (half-to-even) for 0 <= t < 2^22 as the low bits of fl_rn(t + 1.5 * 2^23).
__device__ __forceinline__ uint32_t rint_bits(float t) { return __float_as_uint(__fadd_rn(t, 12582912.0f)); }
__device__ __forceinline__ float clamp255(float t) { return fminf(fmaxf(t, 0.0f), 255.0f); }

// std::round(v) (half away from zero) as an int, exact for |v| < 2^23; saturates beyond.
__device__ __forceinline__ int round_half_away_i(float v) {
  const int i = __float2int_rz(__fadd_rn(v, copysignf(0.5f, v)));
  return fabsf(v) < 0.5f ? 0 : i;
}
// std::round on float32 (oracle/ops.py round_half_away), float result.
__device__ __forceinline__ float round_half_away(float v) {
  const float r = truncf(v);
  const float d = __fsub_rn(v, r);
  return r + (d >= 0.5f ? 1.0f : 0.0f) - (d <= -0.5f ? 1.0f : 0.0f);
}

// Hot-loop form of chambers' blend on floats holding integers: trunc(clip(a + f * (b - a))).  The
// clip is a no-op for 0 <= f <= 1 (image_augmentations.py:43-45), so one formula serves both modes,
// and f == 0 yields a exactly.
__device__ __forceinline__ uint32_t blend_trunc(float a, float b, float f) {
  return trunc_bits(clamp255(__fadd_rn(a, __fmul_rn(f, __fsub_rn(b, a)))));
}
// Color on float channels; returns the three result bytes in the low bits of R, G, B.
__device__ __forceinline__ void color_pixel_f(float r, float g, float b, float f, uint32_t& R, uint32_t& G, uint32_t& B) {
  const float k = __int_as_float(0x3b808081);  // float32(1/255)
  float s = __fmul_rn(__fmul_rn(r, k), __int_as_float(0x3e99096c));
  s = __fadd_rn(s, __fmul_rn(__fmul_rn(g, k), __int_as_float(0x3f1645a2)));
  s = __fadd_rn(s, __fmul_rn(__fmul_rn(b, k), __int_as_float(0x3de978d5)));
  const float gray = __fadd_rn(__fadd_rz(__fmul_rn(s, 255.5f), 8388608.0f), -8388608.0f);  // float(trunc(.))
  R = blend_trunc(gray, r, f);
  G = blend_trunc(gray, g, f);
  B = blend_trunc(gray, b, f);
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

// The kernel's view of one image while its chain is interpreted.
template <int C, bool SMEM>
struct Ctx {
  Small* s;
  uint32_t* big;       // histogram replicas / replicated LUT table (aliased)
  uint8_t* stage;      // 2 x stage_bytes(C) staging tiles for TMA stores
  uint8_t* simg;       // shared-memory image (SMEM only)
  const uint8_t* raw;  // current raw image: simg, the input image, or a scratch buffer
  uint8_t* scrA;
  uint8_t* scrB;
  int H, W, HW, img_bytes;
  int tid, lane;
  uint32_t stage_seq;  // staging tiles issued so far by this CTA
  bool tab_valid;

  // In the shared-memory variant every read of the image is provably a shared-space load.
  __device__ __forceinline__ const uint8_t* img_src() const { return SMEM ? simg : raw; }
  __device__ __forceinline__ int raw_at(int idx) const { return img_src()[idx]; }

  __device__ __forceinline__ const uint8_t* lane_tab() const {
    return reinterpret_cast<const uint8_t*>(big) + (lane << 2);
  }
  // bank-conflict-free lookup: every lane reads its own replica (bank == lane); LEA + LDS.U8.
  __device__ __forceinline__ uint32_t lut_at(const uint8_t* tab, int c, uint32_t v) const { return tab[(v << 7) + c]; }
  __device__ __forceinline__ uint32_t lut_byte(int c, uint32_t v) const { return lut_at(lane_tab(), c, v); }

  // (Re)build the replicated table from the compact per-channel LUTs.
  __device__ void build_table() {
    for (int i = tid; i < TAB_WORDS; i += NT) {
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
  // 16 bytes per trip; the channel of byte b of unit u is (u + b) mod 3 for C == 3, so the three
  // per-position channel offsets rotate with u instead of unrolling 48 bytes.
  __device__ void map_pass(uint8_t* dst, bool identity) {
    if (!identity && !tab_valid) build_table();
    const int n_units = img_bytes >> 4;
    const bool dst_vec = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const uint8_t* src = img_src();
    const uint8_t* tab = lane_tab();
    if (identity) {
      if (dst_vec) {
        for (int u = tid; u < n_units; u += NT)
          st_stream(reinterpret_cast<uint4*>(dst) + u, reinterpret_cast<const uint4*>(src)[u]);
      } else {
        for (int i = tid; i < (n_units << 4); i += NT) dst[i] = src[i];
      }
    } else {
      for (int u = tid; u < n_units; u += NT) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[u];
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int ph = (C == 3) ? (u % 3) : 0;  // channel of byte 0 of this unit
        const uint8_t* t0 = tab + ((C == 3) ? ph : 0);
        const uint8_t* t1 = tab + ((C == 3) ? (ph == 2 ? 0 : ph + 1) : (1 % C));
        const uint8_t* t2 = tab + ((C == 3) ? (ph == 0 ? 2 : ph - 1) : (2 % C));
        const uint8_t* t3 = tab + ((C == 3) ? ph : (3 % C));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t o;
          if (C == 3) {
            // byte index 4j+b -> channel offset pointer t[(4j+b) % 3]
            const uint8_t* q0 = ((4 * j) % 3 == 0) ? t0 : ((4 * j) % 3 == 1) ? t1 : t2;
            const uint8_t* q1 = ((4 * j + 1) % 3 == 0) ? t0 : ((4 * j + 1) % 3 == 1) ? t1 : t2;
            const uint8_t* q2 = ((4 * j + 2) % 3 == 0) ? t0 : ((4 * j + 2) % 3 == 1) ? t1 : t2;
            const uint8_t* q3 = ((4 * j + 3) % 3 == 0) ? t0 : ((4 * j + 3) % 3 == 1) ? t1 : t2;
            o = q0[byte_of(w[j], 0) << 7];
            o = put_byte(o, q1[byte_of(w[j], 1) << 7], 1);
            o = put_byte(o, q2[byte_of(w[j], 2) << 7], 2);
            o = put_byte(o, q3[byte_of(w[j], 3) << 7], 3);
          } else {
            o = t0[byte_of(w[j], 0) << 7];
            o = put_byte(o, t1[byte_of(w[j], 1) << 7], 1);
            o = put_byte(o, t2[byte_of(w[j], 2) << 7], 2);
            o = put_byte(o, t3[byte_of(w[j], 3) << 7], 3);
          }
          w[j] = o;
        }
        uint8_t* d = dst + ((size_t)u << 4);
        if (dst_vec) {
          *reinterpret_cast<uint4*>(d) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) d[4 * j + b] = (uint8_t)get_byte(w[j], b);
        }
      }
    }
    for (int i = (n_units << 4) + tid; i < img_bytes; i += NT) {
      const int v = src[i];
      dst[i] = (uint8_t)(identity ? v : lut_byte(i % C, v));
    }
  }

  // ---- pass: Color with the pending LUT fused in; dst may alias raw (per-pixel op).
  // 4 pixels (12 bytes, three words) per trip.
  __device__ void color_pass(uint8_t* dst, float f, int mode, bool identity) {
    if (!identity && !tab_valid) build_table();
    if (C != 3) return;
    const uint8_t* src = img_src();
    const uint8_t* tab = lane_tab();
    const int n_groups = HW >> 2;
    for (int g = tid; g < n_groups; g += NT) {
      const uint32_t* sp = reinterpret_cast<const uint32_t*>(src) + 3 * g;
      const uint32_t w[3] = {sp[0], sp[1], sp[2]};
      uint32_t o[3] = {0, 0, 0};
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        float ch[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int bi = px * 3 + c;
          if (identity) ch[c] = byte_to_float(w[bi >> 2], bi & 3);
          else ch[c] = small_uint_to_float(lut_at(tab, c, byte_of(w[bi >> 2], bi & 3)));
        }
        uint32_t res[3];
        color_pixel_f(ch[0], ch[1], ch[2], f, res[0], res[1], res[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int bi = px * 3 + c;
          o[bi >> 2] = put_byte(o[bi >> 2], res[c], bi & 3);
        }
      }
      uint32_t* dp = reinterpret_cast<uint32_t*>(dst) + 3 * g;
      dp[0] = o[0]; dp[1] = o[1]; dp[2] = o[2];
    }
    for (int px = (n_groups << 2) + tid; px < HW; px += NT) {
      int ch[3];
      for (int c = 0; c < 3; ++c) {
        const int v = src[px * 3 + c];
        ch[c] = identity ? v : (int)lut_byte(c, v);
      }
      color_pixel(ch[0], ch[1], ch[2], f, mode);
      for (int c = 0; c < 3; ++c) dst[px * 3 + c] = (uint8_t)ch[c];
    }
  }

  // ---- gather passes.  The virtual image is produced tile by tile (STAGE_PIX pixels, one 4-pixel
  // group per thread) into a shared-memory staging buffer; one elected thread hands each finished
  // tile to the TMA (bulk store shared -> global) and the CTA moves on to the other buffer.
  // `stage_seq` counts tiles across images, so a store may still be draining its buffer while the
  // next image is already being loaded and computed; a buffer is reused only after the store issued
  // two tiles earlier has finished reading it.  `complete`: also wait until the data is in global
  // memory (needed when this CTA reads it back).
  template <typename GroupFn>
  __device__ void staged_emit(uint8_t* dst, bool complete, GroupFn&& group_fn) {
    const int n_tiles = (HW + STAGE_PIX - 1) / STAGE_PIX;
    const bool use_tma = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    for (int tile = 0; tile < n_tiles; ++tile, ++stage_seq) {
      uint8_t* buf = stage + (size_t)(stage_seq & 1u) * stage_bytes(C);
      if (tid == 0) bulk_wait_read<1>();  // the store issued two tiles ago has released this buffer
      __syncthreads();
      const int pix0 = tile * STAGE_PIX;
      const int npix = min(STAGE_PIX, HW - pix0);
      const int p = pix0 + (tid << 2);
      if ((tid << 2) < npix) {
        uint32_t o[C];
        group_fn(p, min(4, HW - p), o);
        uint32_t* bp = reinterpret_cast<uint32_t*>(buf) + tid * C;
#pragma unroll
        for (int q = 0; q < C; ++q) bp[q] = o[q];
      }
      const uint32_t bytes = (uint32_t)npix * C;
      const uint32_t bulk = use_tma ? (bytes & ~15u) : 0u;
      if (bulk) fence_proxy_async();  // generic-proxy writes -> visible to the async proxy
      __syncthreads();
      uint8_t* d = dst + (size_t)pix0 * C;
      if (tid == 0) {
        if (bulk) bulk_store(d, buf, bulk);
        bulk_commit();  // one (possibly empty) group per tile keeps the wait arithmetic uniform
      }
      if (bulk < bytes) {  // unaligned dst / ragged tail: plain copies
        for (uint32_t i = bulk + tid; i < bytes; i += NT) d[i] = buf[i];
        __syncthreads();
      }
    }
    if (complete) {
      if (tid == 0) bulk_wait_all<0>();
      __syncthreads();
    }
  }

  // Fast gather: the spatial list is one or two constant-fill nearest-neighbour affine warps (the
  // only form the policies produce).  Coefficients live in registers; ops whose matrix leaves a
  // coordinate alone (Shear / Translate) skip that coordinate.
  // MODE 0: general; 1: source row == output row (ShearX, TranslateX); 2: source column == output column.
  template <int MODE>
  __device__ void pull_affine(uint8_t* dst, bool complete, bool identity, bool two) {
    const Spatial& ea = s->sp[s->n_sp - 1];  // applied last -> evaluated first
    const float t0 = ea.t[0], t1 = ea.t[1], t2 = ea.t[2], t3 = ea.t[3], t4 = ea.t[4], t5 = ea.t[5];
    float u0 = 1.f, u1 = 0.f, u2 = 0.f, u3 = 0.f, u4 = 1.f, u5 = 0.f;
    uint32_t fill_a = 0, fill_b = 0;  // colours packed one byte per channel
#pragma unroll
    for (int c = 0; c < C; ++c) fill_a |= (uint32_t)ea.color[c] << (8 * c);
    if (two) {
      const Spatial& eb = s->sp[s->n_sp - 2];
      u0 = eb.t[0]; u1 = eb.t[1]; u2 = eb.t[2]; u3 = eb.t[3]; u4 = eb.t[4]; u5 = eb.t[5];
#pragma unroll
      for (int c = 0; c < C; ++c) fill_b |= (uint32_t)eb.color[c] << (8 * c);
    }
    const uint8_t* src = img_src();
    const uint8_t* tab = lane_tab();
    const int Wl = W, Hl = H;
    staged_emit(dst, complete, [&](int p, int n, uint32_t* o) {
      int y = p / Wl;
      int x = p - y * Wl;
#pragma unroll
      for (int q = 0; q < C; ++q) o[q] = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < n) {
          const float fx = small_uint_to_float((uint32_t)x), fy = small_uint_to_float((uint32_t)y);
          int ix = x, iy = y;
          if (MODE != 2) ix = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(t0, fx), __fmul_rn(t1, fy)), t2));
          if (MODE != 1) iy = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(t3, fx), __fmul_rn(t4, fy)), t5));
          bool inside = (unsigned)ix < (unsigned)Wl && (unsigned)iy < (unsigned)Hl;
          uint32_t fill = fill_a;
          if (two && inside) {
            const float gx = small_uint_to_float((uint32_t)ix), gy = small_uint_to_float((uint32_t)iy);
            ix = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(u0, gx), __fmul_rn(u1, gy)), u2));
            iy = round_half_away_i(__fadd_rn(__fadd_rn(__fmul_rn(u3, gx), __fmul_rn(u4, gy)), u5));
            inside = (unsigned)ix < (unsigned)Wl && (unsigned)iy < (unsigned)Hl;
            fill = fill_b;
          }
          const uint8_t* px = src + (iy * Wl + ix) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            uint32_t v = byte_of(fill, c);
            if (inside) {
              v = px[c];
              if (!identity) v = lut_at(tab, c, v);
            }
            const int bi = i * C + c;
            o[bi >> 2] = put_byte(o[bi >> 2], v, bi & 3);
          }
        }
        if (++x == Wl) { x = 0; ++y; }
      }
    });
  }

  // General gather: any spatial list (CutOut rectangles, non-constant fill modes, 3+ warps).
  __device__ void pull_generic(uint8_t* dst, bool complete, bool identity) {
    const int n_sp = s->n_sp;
    const uint8_t* src = img_src();
    const uint8_t* tab = lane_tab();
    const int Wl = W, Hl = H;
    const Small* sl = s;
    staged_emit(dst, complete, [&](int p, int n, uint32_t* o) {
      int y = p / Wl;
      int x = p - y * Wl;
#pragma unroll
      for (int q = 0; q < C; ++q) o[q] = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (i < n) {
          int sx = x, sy = y;
          const int k = pull_resolve(sl, n_sp, Hl, Wl, sx, sy);
          const uint8_t* px = src + (sy * Wl + sx) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            uint32_t v;
            if (k < 0) {
              v = px[c];
              if (!identity) v = lut_at(tab, c, v);
            } else {
              v = (uint32_t)sl->sp[k].color[c];
            }
            const int bi = i * C + c;
            o[bi >> 2] = put_byte(o[bi >> 2], v, bi & 3);
          }
        }
        if (++x == Wl) { x = 0; ++y; }
      }
    });
  }

  // ---- pass: write the virtual image (spatial list + raw + LUT) to dst.
  __device__ void pull_pass(uint8_t* dst, bool identity, bool complete) {
    if (!identity && !tab_valid) build_table();
    const int n_sp = s->n_sp;
    bool fast = n_sp >= 1 && n_sp <= 2;
    for (int k = 0; k < n_sp && fast; ++k)
      fast = s->sp[k].type == SP_GEOM && s->sp[k].fill_mode == CHB_FILL_CONSTANT;
    if (fast) {
      const float* t = s->sp[n_sp - 1].t;
      const bool row_id = (t[3] == 0.0f && t[4] == 1.0f && t[5] == 0.0f);
      const bool col_id = (t[0] == 1.0f && t[1] == 0.0f && t[2] == 0.0f);
      if (row_id) pull_affine<1>(dst, complete, identity, n_sp == 2);
      else if (col_id) pull_affine<2>(dst, complete, identity, n_sp == 2);
      else pull_affine<0>(dst, complete, identity, n_sp == 2);
      return;
    }
    pull_generic(dst, complete, identity);
  }

  // ---- histogram of the virtual image into s->hmap[c][256].
  // Replica layout: word (v >> 1) * 32 + lane holds the u16 counts of bins v & ~1 and v | 1 for
  // the pixels seen by lane `lane` of any warp: every lane always hits its own bank.
  __device__ __forceinline__ void hist_add(uint8_t* hb, uint32_t v) {  // hb: channel base + lane * 4
    atomicAdd(reinterpret_cast<unsigned int*>(hb + ((v & 0xFEu) << 6)), 1u << ((v & 1u) << 4));
  }

  __device__ void histogram(bool identity) {
    const int n_sp = s->n_sp;
    for (int i = tid; i < C * HIST_WORDS_PER_CH; i += NT) big[i] = 0;
    for (int i = tid; i < MAXC * 256; i += NT) (&s->hmap[0][0])[i] = 0;
    tab_valid = false;
    __syncthreads();
    uint8_t* hb = reinterpret_cast<uint8_t*>(big) + (lane << 2);
    const uint8_t* src = img_src();
    if (n_sp == 0) {
      const int n_units = img_bytes >> 4;
      for (int u = tid; u < n_units; u += NT) {
        const uint4 v = reinterpret_cast<const uint4*>(src)[u];
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        const int ph = (C == 3) ? (u % 3) : 0;
        uint8_t* h0 = hb + (size_t)((C == 3) ? ph : 0) * (HIST_WORDS_PER_CH * 4);
        uint8_t* h1 = hb + (size_t)((C == 3) ? (ph == 2 ? 0 : ph + 1) : (1 % C)) * (HIST_WORDS_PER_CH * 4);
        uint8_t* h2 = hb + (size_t)((C == 3) ? (ph == 0 ? 2 : ph - 1) : (2 % C)) * (HIST_WORDS_PER_CH * 4);
        uint8_t* h3 = hb + (size_t)((C == 3) ? ph : (3 % C)) * (HIST_WORDS_PER_CH * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (C == 3) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const int r = (4 * j + b) % 3;
              hist_add(r == 0 ? h0 : r == 1 ? h1 : h2, byte_of(w[j], b));
            }
          } else {
            hist_add(h0, byte_of(w[j], 0));
            hist_add(h1, byte_of(w[j], 1));
            hist_add(h2, byte_of(w[j], 2));
            hist_add(h3, byte_of(w[j], 3));
          }
        }
      }
      for (int i = (n_units << 4) + tid; i < img_bytes; i += NT)
        hist_add(hb + (size_t)(i % C) * (HIST_WORDS_PER_CH * 4), src[i]);
    } else {
      // spatial ops pending: count RAW values of resolved pixels (mapped through the LUT below) and
      // the already-final colours of fill / mask pixels straight into the mapped histogram.
      for (int pix = tid; pix < HW; pix += NT) {
        int y = pix / W;
        int x = pix - y * W;
        const int k = pull_resolve(s, n_sp, H, W, x, y);
        if (k < 0) {
          const uint8_t* px = src + (y * W + x) * C;
#pragma unroll
          for (int c = 0; c < C; ++c) hist_add(hb + (size_t)c * (HIST_WORDS_PER_CH * 4), px[c]);
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) atomicAdd(&s->hmap[c][s->sp[k].color[c]], 1u);
        }
      }
    }
    __syncthreads();
    // reduce the 32 lane replicas, mapping raw values through the pending LUT.
    for (int t = tid; t < C * 256; t += NT) {
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
    for (int t = tid; t < C * 256; t += NT) {
      const int c = t >> 8, v = t & 255;
      s->lut[c][v] = s->etab[c][s->lut[c][v]];
    }
    for (int t = tid; t < s->n_sp * C; t += NT) {
      const int k = t / C, c = t - k * C;
      s->sp[k].color[c] = s->etab[c][s->sp[k].color[c]];
    }
    if (tid == 0) s->lut_identity = 0;
    tab_valid = false;
    __syncthreads();
  }

  // ---- Sharpness (tfa.image.sharpness; oracle/ops.py sharpness): raw holds the materialised input.
  // out = rint(clip(deg + f * (orig - deg))), deg = trunc(sum of the 9 float32 products in row-major
  // order, starting from 0), border pixels keep deg = orig.
  __device__ __forceinline__ uint32_t sharp_blend(float deg, float orig, float f) {
    return rint_bits(clamp255(__fadd_rn(deg, __fmul_rn(f, __fsub_rn(orig, deg)))));
  }

  // Sliding-column form for rows that are a whole number of words: a thread owns one word column
  // (4 bytes wide) of a strip of rows and walks down it, keeping the float32 products of the last
  // two rows of its 4 + 2C-byte window in registers, so every input byte is converted and multiplied
  // once per column instead of nine times.  A warp's 32 lanes own 32 adjacent word columns: loads
  // are conflict-free and the 4-byte stores coalesce into full 128-byte lines.
  __device__ void sharpness_columns(uint8_t* dst, float f) {
    constexpr int NB = 4 + 2 * C;  // window bytes per row
    const int row = W * C;
    const int wpr = row >> 2;      // words per row
    const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
    const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
    const uint8_t* src = img_src();
    // first and last row: every pixel is border -> blend(orig, orig) == orig
    for (int i = tid; i < wpr; i += NT) {
      reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
      if (H > 1)
        reinterpret_cast<uint32_t*>(dst)[(size_t)(H - 1) * wpr + i] =
            reinterpret_cast<const uint32_t*>(src)[(size_t)(H - 1) * wpr + i];
    }
    const int inner = H - 2;
    if (inner <= 0) return;
    // split the inner rows into S strips so that (word columns x strips) keeps every thread busy:
    // minimise rounds * (rows per strip + 2 halo rows) over a small range of S.
    int best_s = 1;
    long best_cost = 1L << 60;
    for (int S = 1; S <= 64 && S <= inner; ++S) {
      const long rounds = ((long)wpr * S + NT - 1) / NT;
      const long cost = rounds * ((inner + S - 1) / S + 2);
      if (cost < best_cost) { best_cost = cost; best_s = S; }
    }
    const int R = (inner + best_s - 1) / best_s;
    const int n_strips = (inner + R - 1) / R;
    const int n_items = wpr * n_strips;
    for (int item = tid; item < n_items; item += NT) {
      const int strip = item / wpr;
      const int xw = item - strip * wpr;
      const int y_begin = 1 + strip * R;
      const int y_end = min(H - 1, y_begin + R);
      const int xb0 = xw << 2;
      // per byte of the word: is it in the first / last pixel of the row?
      bool border[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) border[b] = (xb0 + b < C) || (xb0 + b >= row - C);
      const bool has_prev = xw > 0, has_next = xw + 1 < wpr;
      float pa[NB], pb[NB];  // products (x k1) of rows y-1 and y
      float ctr[4];          // float values of the centre bytes of row y
      auto load_row = [&](int yy, float* p, float* cvals) {
        const uint32_t* rp = reinterpret_cast<const uint32_t*>(src) + (size_t)yy * wpr + xw;
        const uint32_t w1 = rp[0];
        const uint32_t w0 = has_prev ? rp[-1] : 0u;
        const uint32_t w2 = has_next ? rp[1] : 0u;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const int wb = 4 - C + j;  // byte index in the 12-byte (w0, w1, w2) window
          const uint32_t wsel = (wb < 4) ? w0 : (wb < 8) ? w1 : w2;
          const float fv = byte_to_float(wsel, wb & 3);
          p[j] = __fmul_rn(fv, k1);
          if (cvals != nullptr && j >= C && j < C + 4) cvals[j - C] = fv;
        }
      };
      load_row(y_begin - 1, pa, nullptr);
      load_row(y_begin, pb, ctr);
#pragma unroll 3
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
          const uint32_t r = sharp_blend(deg, orig, f);
          o = (b == 0) ? (r & 0xFFu) : put_byte(o, r, b);
        }
        reinterpret_cast<uint32_t*>(dst)[(size_t)y * wpr + xw] = o;
#pragma unroll
        for (int j = 0; j < NB; ++j) { pa[j] = pb[j]; pb[j] = pc[j]; }
#pragma unroll
        for (int b = 0; b < 4; ++b) ctr[b] = nctr[b];
      }
    }
  }

  __device__ void sharpness_pass(uint8_t* dst, float f, int mode) {
    const int row = W * C;
    if (mode == BLEND_IMAGE1) f = 0.0f;  // factor 0: deg + 0 * (orig - deg) == deg exactly
    if ((row & 3) == 0 && H >= 1 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0 &&
        (reinterpret_cast<uintptr_t>(img_src()) & 3) == 0) {
      sharpness_columns(dst, f);
      return;
    }
    const float k1 = __int_as_float(0x3d9d89d9);
    const float k5 = __int_as_float(0x3ec4ec4f);
    for (int i = tid; i < img_bytes; i += NT) {
      const int y = i / row;
      const int xb = i - y * row;
      const int x = xb / C;
      const int orig = raw_at(i);
      float deg = (float)orig;
      if (y > 0 && y < H - 1 && x > 0 && x < W - 1) {
        float acc = 0.0f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const float kk = (dy == 0 && dx == 0) ? k5 : k1;
            acc = __fadd_rn(acc, __fmul_rn((float)raw_at(i + dy * row + dx * C), kk));
          }
        deg = (float)(((int)acc) & 0xFF);
      }
      dst[i] = (uint8_t)(sharp_blend(deg, (float)orig, f) & 0xFFu);
    }
  }

  // ---- bilinear warp (image_ops.h bilinear_interpolation; oracle/ops.py projective_transform).
  __device__ void bilinear_pass(uint8_t* dst, const float* t, int fill_mode, int fill, bool identity) {
    if (!identity && !tab_valid) build_table();
    for (int pix = tid; pix < HW; pix += NT) {
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
          return (float)(identity ? v : (int)lut_byte(c, v));
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
      for (int i = tid; i < n16; i += NT)
        reinterpret_cast<uint4*>(simg)[i] = ld_cg(reinterpret_cast<const uint4*>(buf) + i);
      for (int i = (n16 << 4) + tid; i < img_bytes; i += NT) simg[i] = __ldcg(buf + i);
      raw = simg;
    } else {
      raw = buf;
    }
    __syncthreads();
  }

  __device__ void reset_lut() {
    for (int t = tid; t < C * 256; t += NT) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
    if (tid == 0) { s->lut_identity = 1; s->n_sp = 0; }
    tab_valid = false;
    __syncthreads();
  }

  // ---- write the virtual image to a scratch buffer and adopt it (LUT and spatial list reset).
  __device__ void materialize() {
    uint8_t* dst = free_scratch();
    const bool identity = s->lut_identity != 0;
    if (s->n_sp > 0) pull_pass(dst, identity, true); else map_pass(dst, identity);
    adopt(dst);
    reset_lut();
  }
};

// ---------------------------------------------------------------------------- schedule decode
// Twin of oracle/philox.py decode_schedule for ONE image, executed by warp 0: every lane draws the
// Philox block of one slot, lane 0 assembles the chain, the lanes then copy the chosen DevOps.
__device__ void decode_image(const KParams& p, Small* s, int img, int H, int W, int lane) {
  const unsigned long long own = p.image_index_base + (unsigned long long)img;
  const unsigned long long stream_img = p.elementwise ? own : ~0ull;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  const int n_slots = p.n_draws * (p.K + 1);
  if (!p.replay && lane < n_slots) {
    const uint4 w = philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter, (uint32_t)lane), key);
    uint4 wc = w;
    if (!p.elementwise) wc = philox4x32_10(make_uint4((uint32_t)own, (uint32_t)(own >> 32), p.call_counter, (uint32_t)lane), key);
    s->rnd[lane][0] = w.x; s->rnd[lane][1] = w.y; s->rnd[lane][2] = w.z; s->rnd[lane][3] = w.w;
    s->rndc[lane][2] = wc.z; s->rndc[lane][3] = wc.w;
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
        choice = (int)__umulhi(s->rnd[slot0][0], (uint32_t)p.T);
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
          applied = (kind >= 0) && ((int)(s->rnd[slot][0] >> 8) < __ldg(&p.ops[opi].thr24));
          negate = s->rnd[slot][1] < 0x80000000u;
          cy = (int)__umulhi(s->rndc[slot][2], (uint32_t)H);
          cx = (int)__umulhi(s->rndc[slot][3], (uint32_t)W);
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
  const int n = s->n_prog;
  constexpr int OPW = (int)(sizeof(DevOp) / 4);
  for (int e = 0; e < n; ++e)
    if (lane < OPW)
      reinterpret_cast<uint32_t*>(&s->prog[e].op)[lane] =
          __ldg(reinterpret_cast<const uint32_t*>(&p.ops[s->prog[e].table_index]) + lane);
  __syncwarp();
}

// ------------------------------------------------------------------------------------- kernel
template <int C, bool SMEM>
__global__ void __launch_bounds__(NT, 1) policy_kernel(const KParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int H = p.H, W = p.W;
  const int img_bytes = H * W * C;
  const size_t img_pad = SMEM ? align_up((size_t)img_bytes, 128) : 0;
  uint8_t* simg = smem_raw;
  uint32_t* big = reinterpret_cast<uint32_t*>(smem_raw + img_pad);
  uint8_t* stage = smem_raw + img_pad + big_region_bytes(C);
  Small* s = reinterpret_cast<Small*>(smem_raw + img_pad + big_region_bytes(C) + 2 * stage_bytes(C));

  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&s->mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
  }
  __syncthreads();

  Ctx<C, SMEM> cx;
  cx.s = s; cx.big = big; cx.stage = stage; cx.simg = simg;
  cx.scrA = p.scratch + (size_t)blockIdx.x * 2 * p.scratch_stride;
  cx.scrB = cx.scrA + p.scratch_stride;
  cx.H = H; cx.W = W; cx.HW = H * W; cx.img_bytes = img_bytes;
  cx.tid = tid; cx.lane = tid & 31; cx.stage_seq = 0;
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
        for (int i = tid; i < img_bytes; i += NT) simg[i] = in_img[i];
      }
      cx.raw = simg;
    } else {
      if (in_vec) {
        cx.raw = in_img;
      } else {
        for (int i = tid; i < img_bytes; i += NT) cx.scrA[i] = in_img[i];
        __threadfence();
        cx.raw = cx.scrA;
      }
    }
    if (tid < 32) {
      decode_image(p, s, img, H, W, tid);
      if (tid == 0) {
        s->next_img = (int)gridDim.x + (int)atomicAdd(p.work_counter, 1u);
        s->n_sp = 0;
        s->lut_identity = 1;
      }
    } else {
      for (int t = tid - 32; t < C * 256; t += NT - 32) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
    }
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
      const ProgEntry& pe = s->prog[pi];
      const DevOp& op = pe.op;
      const bool last = (pi == n_prog - 1);
      const int kind = op.kind;
      switch (kind) {
        case CHB_OP_INVERT: case CHB_OP_POSTERIZE: case CHB_OP_SOLARIZE: case CHB_OP_SOLARIZE_ADD:
        case CHB_OP_BRIGHTNESS: case CHB_OP_CONTRAST: {
          for (int t = tid; t < C * 256; t += NT)
            s->lut[t >> 8][t & 255] = (uint8_t)pointwise_value(op, s->lut[t >> 8][t & 255]);
          for (int t = tid; t < s->n_sp * C; t += NT) {
            const int k = t / C, c = t - k * C;
            s->sp[k].color[c] = pointwise_value(op, s->sp[k].color[c]);
          }
          if (tid == 0) s->lut_identity = 0;
          cx.tab_valid = false;
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
          for (int t = tid; t < C * 256; t += NT) s->lut[t >> 8][t & 255] = (uint8_t)(t & 255);
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
              for (int i = tid; i < rw * rh; i += NT) {
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
      if (s->n_sp > 0) cx.pull_pass(out_img, identity, false); else cx.map_pass(out_img, identity);
    }
    __syncthreads();
    img = s->next_img;
    __syncthreads();
  }

  // last CTA out re-arms the work counter for the next launch.
  if (tid == 0) {
    bulk_wait_read<0>();  // staging buffers must outlive the stores that read them
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
    policy_kernel<C, true><<<li.grid, NT, li.smem, stream>>>(p);
  else
    policy_kernel<C, false><<<li.grid, NT, li.smem, stream>>>(p);
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
  size_t smem_overhead_c##C() { return smem_overhead_bytes(C); }

}  // namespace chb
