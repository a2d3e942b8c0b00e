// The two chambers layers either side of the policy path:
//   ImageNetNormalization  image_augmentations.py:620-682   uint8 / float32 NHWC -> float32 NHWC
//   ResizingMinMax         image_augmentations.py:685-748   bilinear / nearest resize keeping the aspect ratio
// Both are memory-bound element-wise kernels: one input byte becomes four output bytes.
#include "chb_device.cuh"

namespace chb {
namespace {

constexpr int NORM_NT = 256;

// uint8 input: every output is one of 256 values per channel -> a 3 x 256 float table per CTA, built
// with the exact IEEE divisions above; the hot loop is one coalesced word load, four table reads and
// one coalesced 16-byte store per thread (thread j owns output float4 j = input word j).
//   C == 3: the channel of byte 0 of word j is j % 3 (4 == 1 mod 3).  caffe reads the pixel's
//   channels reversed: output element e of channel c comes from input byte e + 2 - 2c, i.e. from the
//   8-byte window around the word.
__global__ void __launch_bounds__(NORM_NT) normalize_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                              unsigned long long n_bytes, int C, int mode) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 3 * 256; i += NORM_NT) lut[i >> 8][i & 255] = norm_value((float)(i & 255), mode, i >> 8);
  __syncthreads();
  const unsigned long long n_words = n_bytes >> 2;
  const uint32_t* in32 = reinterpret_cast<const uint32_t*>(in);
  float4* out4 = reinterpret_cast<float4*>(out);
  const unsigned long long stride = (unsigned long long)gridDim.x * NORM_NT;
  for (unsigned long long j = (unsigned long long)blockIdx.x * NORM_NT + threadIdx.x; j < n_words; j += stride) {
    const uint32_t w = __ldg(in32 + j);
    float r[4];
    if (mode == CHB_NORM_CAFFE) {
      const uint32_t wp = j > 0 ? __ldg(in32 + j - 1) : 0u;
      uint32_t wn = 0u;
      if (j + 1 < n_words) {
        wn = __ldg(in32 + j + 1);
      } else {  // the bytes after the last whole word (fewer than four)
        for (unsigned long long t = (j + 1) << 2, k = 0; t < n_bytes; ++t, ++k) wn |= (uint32_t)in[t] << (8 * k);
      }
      const int ph = (int)(j % 3ull);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c = (ph + b) % 3;
        const int rel = b + 2 - 2 * c;  // -2 .. 5, relative to the word's first byte
        const uint32_t src = rel < 0 ? wp : rel < 4 ? w : wn;
        r[b] = lut[c][(src >> (8 * (rel & 3))) & 255u];
      }
    } else {
      const int ph = (C == 3) ? (int)(j % 3ull) : 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) r[b] = lut[(mode == CHB_NORM_TF) ? 0 : (ph + b) % 3][(w >> (8 * b)) & 255u];
    }
    __stcs(out4 + j, make_float4(r[0], r[1], r[2], r[3]));  // written once, read by the next layer: streaming store
  }
  // tail: fewer than four bytes, whole pixels (n_bytes is a multiple of C)
  if (blockIdx.x == 0 && threadIdx.x < (int)(n_bytes & 3ull)) {
    const unsigned long long e = (n_words << 2) + threadIdx.x;
    const int c = (int)(e % (unsigned long long)C);
    const unsigned long long src = (mode == CHB_NORM_CAFFE) ? e + (unsigned long long)(C - 1 - 2 * c) : e;
    out[e] = lut[(mode == CHB_NORM_TF) ? 0 : c][in[src]];
  }
}

// Unaligned buffers: one element per thread.
__global__ void __launch_bounds__(NORM_NT) normalize_u8_scalar_kernel(const uint8_t* __restrict__ in, float* __restrict__ out,
                                                                     unsigned long long n_bytes, int C, int mode) {
  const unsigned long long stride = (unsigned long long)gridDim.x * NORM_NT;
  for (unsigned long long e = (unsigned long long)blockIdx.x * NORM_NT + threadIdx.x; e < n_bytes; e += stride) {
    const int c = (int)(e % (unsigned long long)C);
    const unsigned long long src = (mode == CHB_NORM_CAFFE) ? e + (unsigned long long)(C - 1 - 2 * c) : e;
    out[e] = norm_value((float)in[src], mode, c);
  }
}

// float32 input (the reference casts whatever it is given, :642-682).
__global__ void __launch_bounds__(NORM_NT) normalize_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                               unsigned long long n, int C, int mode) {
  const unsigned long long stride = (unsigned long long)gridDim.x * NORM_NT;
  for (unsigned long long e = (unsigned long long)blockIdx.x * NORM_NT + threadIdx.x; e < n; e += stride) {
    const int c = (int)(e % (unsigned long long)C);
    const unsigned long long src = (mode == CHB_NORM_CAFFE) ? e + (unsigned long long)(C - 1 - 2 * c) : e;
    out[e] = norm_value(__ldg(in + src), mode, c);
  }
}

// ------------------------------------------------------------------------------------ resize
// tf.image.resize(images, [oh, ow], method) as Keras' Resizing layer calls it (image_ops_impl
// resize_images_v2 -> ResizeBilinear / ResizeNearestNeighbor with half_pixel_centers=True,
// antialias=False); float32 output for bilinear, the input dtype for nearest.
//   scale = in / out (float32);  src = (dst + 0.5) * scale - 0.5
//   bilinear: lower = max(floor(src), 0), upper = min(ceil(src), in - 1), lerp = src - floor(src);
//             top = tl + (tr - tl) * xl; bottom = bl + (br - bl) * xl; out = top + (bottom - top) * yl
//   nearest : index = min(floor((dst + 0.5) * scale), in - 1)
struct ResizeParams {
  int B, IH, IW, C, OH, OW;
  float hs, ws;  // IH / OH, IW / OW in float32
};

__device__ __forceinline__ void bilinear_taps(int o, float scale, int in_size, int& lo, int& hi, float& lerp) {
  const float src = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), 0.5f);
  const float fl = floorf(src);
  lo = max((int)fl, 0);
  hi = min((int)ceilf(src), in_size - 1);
  lerp = __fsub_rn(src, fl);
}

template <typename T>
__global__ void __launch_bounds__(NORM_NT) resize_bilinear_kernel(const T* __restrict__ in, float* __restrict__ out, ResizeParams p) {
  const unsigned long long n = (unsigned long long)p.B * p.OH * p.OW * p.C;
  const unsigned long long stride = (unsigned long long)gridDim.x * NORM_NT;
  for (unsigned long long e = (unsigned long long)blockIdx.x * NORM_NT + threadIdx.x; e < n; e += stride) {
    const int c = (int)(e % (unsigned)p.C);
    unsigned long long r = e / (unsigned)p.C;
    const int ox = (int)(r % (unsigned)p.OW); r /= (unsigned)p.OW;
    const int oy = (int)(r % (unsigned)p.OH);
    const int b = (int)(r / (unsigned)p.OH);
    int y0, y1, x0, x1;
    float yl, xl;
    bilinear_taps(oy, p.hs, p.IH, y0, y1, yl);
    bilinear_taps(ox, p.ws, p.IW, x0, x1, xl);
    const T* img = in + (size_t)b * p.IH * p.IW * p.C;
    const float tl = (float)img[((size_t)y0 * p.IW + x0) * p.C + c], tr = (float)img[((size_t)y0 * p.IW + x1) * p.C + c];
    const float bl = (float)img[((size_t)y1 * p.IW + x0) * p.C + c], br = (float)img[((size_t)y1 * p.IW + x1) * p.C + c];
    const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
    const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
    out[e] = __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
  }
}

template <typename T>
__global__ void __launch_bounds__(NORM_NT) resize_nearest_kernel(const T* __restrict__ in, T* __restrict__ out, ResizeParams p) {
  const unsigned long long n = (unsigned long long)p.B * p.OH * p.OW * p.C;
  const unsigned long long stride = (unsigned long long)gridDim.x * NORM_NT;
  for (unsigned long long e = (unsigned long long)blockIdx.x * NORM_NT + threadIdx.x; e < n; e += stride) {
    const int c = (int)(e % (unsigned)p.C);
    unsigned long long r = e / (unsigned)p.C;
    const int ox = (int)(r % (unsigned)p.OW); r /= (unsigned)p.OW;
    const int oy = (int)(r % (unsigned)p.OH);
    const int b = (int)(r / (unsigned)p.OH);
    const int iy = min((int)floorf(__fmul_rn(__fadd_rn((float)oy, 0.5f), p.hs)), p.IH - 1);
    const int ix = min((int)floorf(__fmul_rn(__fadd_rn((float)ox, 0.5f), p.ws)), p.IW - 1);
    out[e] = in[(((size_t)b * p.IH + iy) * p.IW + ix) * p.C + c];
  }
}

int grid_for(unsigned long long items, int num_sms) {
  unsigned long long g = (items + NORM_NT - 1) / NORM_NT;
  const unsigned long long cap = (unsigned long long)num_sms * 8ull;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace

cudaError_t launch_normalize(const void* in, int in_is_f32, float* out, unsigned long long n, int C, int mode, int num_sms,
                             cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (in_is_f32) {
    normalize_f32_kernel<<<grid_for(n, num_sms), NORM_NT, 0, stream>>>(static_cast<const float*>(in), out, n, C, mode);
  } else if ((((uintptr_t)in) & 3) == 0 && (((uintptr_t)out) & 15) == 0) {
    normalize_u8_kernel<<<grid_for((n + 3) / 4, num_sms), NORM_NT, 0, stream>>>(static_cast<const uint8_t*>(in), out, n, C, mode);
  } else {
    normalize_u8_scalar_kernel<<<grid_for(n, num_sms), NORM_NT, 0, stream>>>(static_cast<const uint8_t*>(in), out, n, C, mode);
  }
  return cudaGetLastError();
}

cudaError_t launch_resize(const void* in, int in_is_f32, void* out, int B, int IH, int IW, int C, int OH, int OW, int nearest,
                          int num_sms, cudaStream_t stream) {
  ResizeParams p;
  p.B = B; p.IH = IH; p.IW = IW; p.C = C; p.OH = OH; p.OW = OW;
  p.hs = (float)IH / (float)OH;
  p.ws = (float)IW / (float)OW;
  const unsigned long long n = (unsigned long long)B * OH * OW * C;
  if (n == 0) return cudaSuccess;
  const int grid = grid_for(n, num_sms);
  if (nearest) {
    if (in_is_f32) resize_nearest_kernel<float><<<grid, NORM_NT, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), p);
    else resize_nearest_kernel<uint8_t><<<grid, NORM_NT, 0, stream>>>(static_cast<const uint8_t*>(in), static_cast<uint8_t*>(out), p);
  } else {
    if (in_is_f32) resize_bilinear_kernel<float><<<grid, NORM_NT, 0, stream>>>(static_cast<const float*>(in), static_cast<float*>(out), p);
    else resize_bilinear_kernel<uint8_t><<<grid, NORM_NT, 0, stream>>>(static_cast<const uint8_t*>(in), static_cast<float*>(out), p);
  }
  return cudaGetLastError();
}

}  // namespace chb
