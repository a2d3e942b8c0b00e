// Device-side building blocks shared by the plan and pass kernels (chb_kernels.cuh): PTX wrappers,
// the Philox4x32-10 generator and the per-value / per-pixel semantics of every op.
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

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
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
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void reds_add(uint32_t a, uint32_t v) {  // fire-and-forget shared atomic add
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// ---- mbarrier + TMA (1-D bulk copies; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "CHB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CHB_DONE_%=;\n\t"
      "bra CHB_WAIT_%=;\n\t"
      "CHB_DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// global -> shared, completion (bytes) signalled on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_store(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {  // this thread's stores have finished reading shared memory
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all0() {  // this thread's stores are complete
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {  // generic-proxy smem writes -> visible to the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Same for every state space: global bytes written through the generic proxy (or by bulk stores)
// before a TMA load of them, and the other way round.
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Barrier of the NCONS consumer threads of a pass CTA (the producer warp never joins it).
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Image bytes read through the generic proxy go to L2 (ld.global.cg): a scratch image may have been
// written by another SM earlier in the same kernel, and L1 is not coherent.
__device__ __forceinline__ int ldg_pixel(const uint8_t* p) { return (int)__ldcg(p); }
// Streaming global accesses: image bytes are touched once per pass, keep them out of L1.
__device__ __forceinline__ uint4 ldg_stream(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

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
// Table-building form (256 entries per op, not hot).
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
__device__ __forceinline__ bool is_pointwise(int kind) {
  return kind == CHB_OP_INVERT || kind == CHB_OP_POSTERIZE || kind == CHB_OP_SOLARIZE ||
         kind == CHB_OP_SOLARIZE_ADD || kind == CHB_OP_BRIGHTNESS || kind == CHB_OP_CONTRAST;
}

// Color (:233-235): blend(grayscale(x) broadcast, x, factor); tf.image.rgb_to_grayscale restated
// (oracle/ops.py rgb_to_grayscale).  Integer form for the few spatial colours and the scalar paths.
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
  R = blend_trunc(gray, r, f) & 0xFFu;
  G = blend_trunc(gray, g, f) & 0xFFu;
  B = blend_trunc(gray, b, f) & 0xFFu;
}

// Sharpness blend (tfa.image.sharpness -> tfa blend; oracle/ops.py sharpness):
// rint(clip(deg + f * (orig - deg))).
__device__ __forceinline__ uint32_t sharp_blend(float deg, float orig, float f) {
  return rint_bits(clamp255(__fadd_rn(deg, __fmul_rn(f, __fsub_rn(orig, deg))))) & 0xFFu;
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

// ImageNetNormalization.call, image_augmentations.py:629-682 (chb_frontend.cu and the fused write epilogue of
// the resident engine).
// float32 steps of the reference, one rounding per op (TensorFlow's CPU kernels do not contract):
//   tf     :660-665   x / 127.5 - 1.0
//   torch  :652-657   ((x / 255.0) - mean[c]) / std[c]
//   caffe  :647-650   x[..., ::-1] - mean[c]        (c indexes the REVERSED channels)
__device__ __forceinline__ float norm_value(float x, int mode, int c) {
  if (mode == CHB_NORM_TF) return __fsub_rn(__fdiv_rn(x, 127.5f), 1.0f);
  if (mode == CHB_NORM_TORCH) {
    const float mean = c == 0 ? 0.485f : c == 1 ? 0.456f : 0.406f;
    const float sd = c == 0 ? 0.229f : c == 1 ? 0.224f : 0.225f;
    return __fdiv_rn(__fsub_rn(__fdiv_rn(x, 255.0f), mean), sd);
  }
  const float mean = c == 0 ? 103.939f : c == 1 ? 116.779f : 123.68f;
  return __fsub_rn(x, mean);
}

}  // namespace
}  // namespace chb
