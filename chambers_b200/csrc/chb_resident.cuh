// Image-resident policy kernel for sm_100a: ONE launch per call, one CTA per SM, one image per CTA
// at a time, the whole image in shared memory for as long as its op chain needs it.
//
//   resident_kernel<C>   Work items (images; in the smallest batches also row ranges of an expensive image's last
//                        pass) are claimed from an atomic counter, in a cost-sorted order for batches of up to 2048
//                        images (plan_order).  Per image: one thread issues the bulk loads (cp.async.bulk +
//                        mbarrier, 48 KB chunks, front to back) of the whole image into shared memory; while they
//                        fly, warp 0 decodes the image's schedule (Philox4x32-10 or replay) and the CTA folds the
//                        chain into the lazy per-image state (advance(), chb_kernels.cuh -- the same chain walk
//                        the tile engine uses, so both engines share every line that decides WHAT is computed).
//                        Then the image's passes run back to back on the resident source:
//                          COUNT          histogram of the (flat) view: 8 skewed copies in shared memory,
//                                         red.shared -- or, for AutoContrast behind a monotone table, a min / max
//                                         pass in registers; Equalize / AutoContrast table, chain walk resumed --
//                                         no second trip to HBM, no election, no ticket;
//                          WRITE_SCRATCH  a view is materialised when an op finds the kernel slot occupied, and
//                                         (KParams::res_rules) in front of Sharpness / a histogram op whenever it
//                                         holds a warp, a mask or an occupied slot: point-wise views and CutOut in
//                                         place, gathers and Sharpness through a per-CTA scratch image (L2-
//                                         resident) that is loaded back, tallied while written if a histogram op
//                                         comes next;
//                          WRITE_OUT      flat chains are transformed in place and leave as 48 KB bulk stores
//                                         (TMA, shared -> global); gathers, row-shift copies and Sharpness store
//                                         from registers; with the fused ImageNetNormalization epilogue all of
//                                         them write float32 instead.
//   An image costs exactly one HBM read and one HBM write whatever its chain.
//
// Eligibility (decided on the host, chb_api.cu resident_eligible): the image, the policy table and >= 10 KB of
// working space fit one SM's shared memory beside the 22 KB control block (images up to ~180 KB: 224 x 224 x 3 =
// 147 KB does; 512 x 512 x 3 does not and runs on the tile engine), rows are whole 16-byte units, every geometric
// op is nearest / constant-fill (what the policies use, augmentation_schemes.py:7-9).  Everything else stays on
// the tile engine (chb_kernels.cuh).
#pragma once
#define RES_CHUNK_BYTES 49152
// the policy table is staged in shared memory: plain loads (see chb_kernels.cuh)
#define CHB_LDP(ptr) (*(ptr))
#include "chb_kernels.cuh"

namespace chb {
namespace {

#ifndef CHB_RNT
#define CHB_RNT 1024
#endif
constexpr int RNT = CHB_RNT;           // threads of a resident CTA (1024: 32 warps, 64 registers each)
static_assert(RNT % 256 == 0, "whole warps, and flat steps of whole 16-byte units");
constexpr int RES_CHUNK = RES_CHUNK_BYTES;       // bytes per load chunk: one flat step of 1024 48-byte units (issuing a bulk copy costs
                                                 // the issuing thread ~150 ns: 13 chunks of 12 KB were 2 us of every image's start-up)
constexpr int RES_MAXCHUNK = 4;        // -> images of up to 192 KB
constexpr int RES_MIN_AUX = 10 * 1024;    // plan_order scratch (8.4 KB) / four histogram copies (<= 8 KB); the host checks it (resident_eligible)
constexpr int LPT_MAX = 2048;          // batches up to this size are claimed longest-chain-first

struct alignas(128) ResCtl {
  unsigned long long full[RES_MAXCHUNK];
  int32_t next_img, n_claimed, _p1, _p2;
  // work splits that depend only on the shape (computed once per launch by thread 0)
  int32_t sharp_rows[4];               // res_sharp: rows per sub-strip when the image's last pass is cut into 1..4 parts
  int32_t n_items, cost_sum, _p4[2];   // claim positions of this call (images, or image parts); sum of the cost estimates
  uint32_t mm[MAXC][2];                // presence-only COUNT pass (AutoContrast): min / max per channel
  int32_t mono[2], _p5[2];             // l1 is non-decreasing / non-increasing in every channel
  uint32_t fillc[2];                   // colour bytes of the last / last-but-one spatial entry ("a miss is just another address")
  uint32_t _p3[2];
  uint32_t rnd[32][4], rndc[32][4];
  uint32_t hmap[MAXC * 256];
  uint8_t etab[MAXC * 256];
  KParams kp;                          // the launch parameters with ops / optab pointing at the shared-memory copy
  uint16_t order[LPT_MAX];             // small batches: claim position -> image, most expensive chains first
  float nlut[3][256];                  // fused ImageNetNormalization epilogue: value -> float32, per channel
  ImgState st[2];                      // the current image's state | the next image's, prepared by warp 31 during the last pass
};

template <int C>
struct RC {
  const KParams* p;
  ResCtl* ctl;
  ImgState* st;        // the current image's state
  const TileState* t;  // == &st->t
  int nt;              // threads working on this pass: RNT, or RNT - 32 while warp 31 prepares the next image
  uint32_t img;        // shared address of the resident source image
  uint32_t aux;        // shared address of the auxiliary region (histogram copies | band buffer)
  int aux_bytes;
  uint32_t full0;      // shared address of the chunk barriers
  uint32_t par;        // parity the chunk barriers complete with for the load in flight
  uint8_t* dst;        // global destination of a WRITE pass
  int H, W, row, img_bytes, tid, lane;
  uint32_t l1a, l2a;   // shared addresses of the two LUTs
  uint32_t hcopy;      // shared address of this lane's histogram copy
  int ncopy;           // histogram copies: 8, 4, 2 or 1
  int tally;           // WRITE pass: also count the bytes written (the materialised view feeds a histogram op next)
  int y_lo, y_hi;      // WRITE passes: output rows this CTA produces (the whole image unless a small batch split the image's last pass)
  int sharp_rows;      // Sharpness: rows per sub-strip of the column walk for this row range
  float* dstf;         // last pass with the fused normalisation epilogue: float32 destination image (else NULL)
  uint32_t nlut;       //   shared address of the value -> float table
  int minmax;          // COUNT pass: only the smallest and largest value per channel are needed (AutoContrast on a monotone l1)
};

// Histogram of a COUNT pass (or the tally of a materialised view): `ncopy` copies of the u32 histogram in the
// aux region, copy = lane & (ncopy - 1), skewed by four banks each -- three instructions per byte (byte
// extract, address, red.shared).  Shared-memory atomics retire ~6-8 distinct addresses per clock whatever
// their banks: a conflict-free layout (32 lane-owned copies of packed 16-bit counters) was measured and
// bought nothing on noise, lost 35 % on constant images where ATOMS.POPC.INC merges the lanes of a copy, and
// cost three more instructions per byte (profiles/r02_ab_notes.md 3).
template <int C>
__host__ __device__ constexpr uint32_t hist_copy_bytes() { return (uint32_t)C * 1024u + 16u; }
template <int C>
__host__ __device__ constexpr uint32_t hist_bytes(int ncopy) { return hist_copy_bytes<C>() * (uint32_t)ncopy; }

template <int C>
__device__ __forceinline__ void hist_zero(const RC<C>& c) {
  const uint32_t n = hist_bytes<C>(c.ncopy);
  for (uint32_t i = (uint32_t)c.tid * 16u; i < n; i += RNT * 16u) sts_v4(c.aux + i, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();
}
template <int C>
__device__ __forceinline__ void hist_add(const RC<C>& c, int ch, uint32_t v) {  // v: a zero-extended byte
  reds_add(c.hcopy + (uint32_t)ch * 1024u + (v << 2), 1u);
}
// Sums the copies into st.hist (which advance() zeroed when it asked for the COUNT pass).
template <int C>
__device__ __forceinline__ void hist_reduce(const RC<C>& c) {
  __syncthreads();
  for (int i = c.tid; i < C * 256; i += RNT) {
    uint32_t sum = 0;
    for (int k = 0; k < c.ncopy; ++k) sum += lds_u32(c.aux + (uint32_t)k * hist_copy_bytes<C>() + (uint32_t)i * 4u);
    (&c.st->hist[0][0])[i] += sum;
  }
  __syncthreads();
}

template <int C>
__device__ __forceinline__ void wait_chunks(const RC<C>& c, int c0, int c1) {
  for (int k = c0; k <= c1; ++k) mbar_wait(c.full0 + 8u * (uint32_t)k, c.par);
}
template <int C>
__device__ __forceinline__ void wait_image(const RC<C>& c) {
  wait_chunks(c, 0, (c.img_bytes + RES_CHUNK - 1) / RES_CHUNK - 1);
}

// chambers' blend against a constant (image_augmentations.py:10-49; blend_value in chb_device.cuh) on the
// UW words of a unit, in float32 steps identical to the table builder's: d = v - a (exact), t = a + f * d
// (two roundings), clip for factors outside [0, 1], truncate.  v enters as the mantissa of 2^23 (one
// PRMT), so d is one add; for a == 0 (Brightness) the product f * v is one FFMA on 2^23 + v -- the exact
// product rounded once, the same float32 -- and the addition of 0 disappears.
template <int UW>
__device__ __forceinline__ void blend_unit(uint32_t* w, int extrap, int a, float f) {
  const float af = (float)a;
  const float na = -(8388608.0f + af);   // exact
  const float nf = -8388608.0f * f;      // exact
#pragma unroll
  for (int j = 0; j < UW; ++j) {
    uint32_t o = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const float xm = __uint_as_float(__byte_perm(w[j], 0x4B000000u, 0x7440u | (uint32_t)b));  // 2^23 + v
      float t;
      if (a == 0) t = __fmaf_rn(xm, f, nf);
      else t = __fadd_rn(af, __fmul_rn(f, __fadd_rn(xm, na)));
      if (extrap) t = fminf(fmaxf(t, 0.0f), 255.0f);
      const uint32_t r = trunc_bits(t);
      o = (b == 0) ? byte_of(r, 0) : put_byte(o, r, b);
    }
    w[j] = o;
  }
}

// Fused ImageNetNormalization epilogue (modes tf / torch: no channel reversal): the 4 bytes of output word
// `widx` of the image (channel of byte 0 = ph) leave as one float4 -- the table holds exactly the float32
// values the standalone layer would compute from the uint8 result.
template <int C>
__device__ __forceinline__ void store_norm_word(const RC<C>& c, size_t widx, int ph, uint32_t w) {
  float4 f;
  f.x = __uint_as_float(lds_u32(c.nlut + (uint32_t)(((ph + 0) % C) % 3) * 1024u + (byte_of(w, 0) << 2)));
  f.y = __uint_as_float(lds_u32(c.nlut + (uint32_t)(((ph + 1) % C) % 3) * 1024u + (byte_of(w, 1) << 2)));
  f.z = __uint_as_float(lds_u32(c.nlut + (uint32_t)(((ph + 2) % C) % 3) * 1024u + (byte_of(w, 2) << 2)));
  f.w = __uint_as_float(lds_u32(c.nlut + (uint32_t)(((ph + 3) % C) % 3) * 1024u + (byte_of(w, 3) << 2)));
  __stcs(reinterpret_cast<float4*>(c.dstf) + widx, f);
}

// Barrier of the threads that work on a pass: the whole CTA, or warps 0 .. 30 while warp 31 decodes and walks
// the NEXT image's chain (named barrier 2; warp 31 never joins it).
template <int C>
__device__ __forceinline__ void work_sync(const RC<C>& c) {
  if (c.nt == RNT) __syncthreads();
  else asm volatile("bar.sync 2, %0;" ::"n"(RNT - 32) : "memory");
}

// ================================================================================ flat executor
// No warp pending, K in {none, Color}: units of 48 bytes (16 pixels; 16 bytes for C != 3) are
// transformed in place as their load chunks arrive, one unit per thread per step of RNT units; a
// step leaves as one bulk store.  CutOut rectangles are painted over the step before it is stored.
//   store   0: in place only (materialisation)  1: bulk-store every step to c.dst
template <int C, bool COUNT>
__device__ __forceinline__ void res_flat(const RC<C>& c, int store) {
  constexpr int UW = (C == 3) ? 12 : 4;
  constexpr int UB = UW * 4;
  const TileState& t = *c.t;
  const int NT = c.nt;  // (COUNT passes always run on the whole CTA)
  const int upr = c.row / UB;                                  // units per row (rows are whole units)
  const int u_first = COUNT ? 0 : c.y_lo * upr;                // COUNT passes always see the whole image
  const int n_units = COUNT ? c.img_bytes / UB : c.y_hi * upr;  // (one past the last unit)
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF) * 0x01010101u, ac1 = (uint32_t)(t.l1_aff & 0xFF) * 0x01010101u;
  const float f = t.kfactor;
  const int blend1 = t.l1_blend, blend_a = t.l1_blend_const;  // l1 as one blend against a constant (closed form)
  const float blend_f = t.l1_blend_factor;
  const bool minmax = COUNT && c.minmax;
  uint32_t mn[C], mx[C];
#pragma unroll
  for (int ch = 0; ch < C; ++ch) { mn[ch] = 255u; mx[ch] = 0u; }
  const int n_paint = COUNT ? 0 : t.n_sp;  // this class holds masks only
  const bool touch = COUNT || use1 || kmode != K_NONE;  // else the staged bytes already are the result
  if (COUNT && !minmax) hist_zero(c);
  for (int base = u_first; base < n_units; base += NT) {
    const int wu = base + (c.tid & ~31);
    if (wu < n_units) {
      const int wl = min(wu + 31, n_units - 1);
      wait_chunks(c, (wu * UB) / RES_CHUNK, (wl * UB + UB - 1) / RES_CHUNK);
    }
    const int u = base + c.tid;
    if (u < n_units && touch) {
      const uint32_t ua = c.img + (uint32_t)u * UB;
      uint32_t w[UW];
#pragma unroll
      for (int q = 0; q < UW / 4; ++q) {
        const uint4 v = lds_v4(ua + q * 16);
        w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
      }
      if (kmode == K_NONE) {
        if (!COUNT && use1) {
          if (aff1) {
#pragma unroll
            for (int j = 0; j < UW; ++j) w[j] = (w[j] & am1) ^ ac1;
          } else if (blend1 == BLEND_EXTRAP + 1) {
            if (blend_a == 0) blend_unit<UW>(w, 1, 0, blend_f); else blend_unit<UW>(w, 1, blend_a, blend_f);
          } else if (blend1 == BLEND_INTERP + 1) {
            if (blend_a == 0) blend_unit<UW>(w, 0, 0, blend_f); else blend_unit<UW>(w, 0, blend_a, blend_f);
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
        if (minmax) {  // three ALU instructions per byte and no shared-memory traffic at all
#pragma unroll
          for (int j = 0; j < UW; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t v = byte_of(w[j], b);
              mn[(4 * j + b) % C] = min(mn[(4 * j + b) % C], v);
              mx[(4 * j + b) % C] = max(mx[(4 * j + b) % C], v);
            }
        } else {
#pragma unroll
          for (int j = 0; j < UW; ++j)
#pragma unroll
            for (int b = 0; b < 4; ++b) hist_add(c, (4 * j + b) % C, byte_of(w[j], b));
        }
      } else {
#pragma unroll
        for (int q = 0; q < UW / 4; ++q) sts_v4(ua + q * 16, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
      }
    }
    if (!COUNT) {
      const int u1 = min(n_units, base + NT);
      // CutOut rectangles, in list order, over the pixels [P0, P1) of this step
      const int P0 = base * (UB / C), P1 = u1 * (UB / C);
      for (int k = 0; k < n_paint; ++k) {
        work_sync(c);
        const Spatial& e = t.sp[k];
        const int rw = (e.x1 - e.x0) * C;
        if (rw <= 0) continue;
        const int ylo = max(e.y0, P0 / c.W), yhi = min(e.y1, (P1 - 1) / c.W + 1);
        const int n = (yhi - ylo) * rw;
        for (int i = c.tid; i < n; i += NT) {
          const int ry = i / rw, rb = i - ry * rw;
          const int pix_b = ((ylo + ry) * c.W + e.x0) * C + rb;  // byte index in the image
          if (pix_b >= P0 * C && pix_b < P1 * C)
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(c.img + (uint32_t)pix_b), "r"(e.color[rb % C]) : "memory");
        }
      }
      if (store && c.dstf) {
        // normalisation epilogue: the step's words are read back (consecutive lanes, consecutive words) and
        // leave as coalesced float4 stores
        work_sync(c);
        const int w0 = base * UW, w1 = u1 * UW;
        for (int wi = w0 + c.tid; wi < w1; wi += NT)
          store_norm_word(c, (size_t)wi, (C == 3) ? (wi % 3) : 0, lds_u32(c.img + ((uint32_t)wi << 2)));
      } else if (store) {
        fence_proxy_async();  // this thread's shared-memory writes -> visible to the TMA
        work_sync(c);
        if (c.tid == 0) {
          bulk_store(c.dst + (size_t)base * UB, c.img + (uint32_t)base * UB, (uint32_t)(u1 - base) * UB);
          bulk_commit();
        }
      }
    }
  }
  if (COUNT && minmax) {
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
      const uint32_t a = __reduce_min_sync(0xFFFFFFFFu, mn[ch]), b = __reduce_max_sync(0xFFFFFFFFu, mx[ch]);
      if (c.lane == 0) { atomicMin(&c.ctl->mm[ch][0], a); atomicMax(&c.ctl->mm[ch][1], b); }
    }
    __syncthreads();
    if (c.tid < C) {  // a histogram that holds exactly what AutoContrast reads: which values bound the range
      const uint32_t lo = c.ctl->mm[c.tid][0], hi = c.ctl->mm[c.tid][1];
      if (lo <= hi) { c.st->hist[c.tid][lo] += 1u; c.st->hist[c.tid][hi] += 1u; }
    }
    __syncthreads();
  } else if (COUNT) {
    hist_reduce(c);
  } else {
    work_sync(c);
  }
}

// ============================================================================== gather executors
// Division-free walk over the units of an image: thread t starts at unit t and advances by the pass's thread count.
struct UnitWalk {
  int ux, y, dux, dy, upr;
  __device__ __forceinline__ UnitWalk(int tid, int upr_, int nt) {
    upr = upr_;
    y = tid / upr; ux = tid - y * upr;
    dy = nt / upr; dux = nt - dy * upr;
  }
  __device__ __forceinline__ void next() {
    ux += dux; y += dy;
    if (ux >= upr) { ux -= upr; ++y; }
  }
};

// src_index (chb_kernels.cuh) without its integer epilogue: returns round(v) + SRC_K in `raw` -- the
// callers fold the constant into their base address -- and tests the range in the float domain:
// round-half-away(v) lies in [0, n) exactly when -0.5 < v < n - 0.5 (nmh = float(n) - 0.5, exact for
// n < 2^22).  fl_rz(v + (2^22 + 0.5)) has ulp 0.5 there, so (bits >> 1) - (0x4A800000 >> 1) = floor(v + 0.5).
constexpr uint32_t SRC_K = 0x4A800000u >> 1;
__device__ __forceinline__ bool src_raw(float v, float nmh, uint32_t& raw) {
  raw = __float_as_uint(__fadd_rz(v, 4194304.5f)) >> 1;
  return (v > -0.5f) && (v < nmh);
}

// src_index for BOTH coordinates of a pixel with one combined validity: the four range tests are joined
// with non-short-circuit ANDs so that they become one predicate chain and ONE select of the address
// (the separate tests compiled to four SELs per pixel, profiles/r02_ab_notes.md 8).
__device__ __forceinline__ bool src_index2(float vx, float vy, int W, int H, int& ix, int& iy) {
  const uint32_t ux = __float_as_uint(__fadd_rz(vx, 4194304.5f)), uy = __float_as_uint(__fadd_rz(vy, 4194304.5f));
  ix = (int)((ux - 0x4A800000u) >> 1);
  iy = (int)((uy - 0x4A800000u) >> 1);
  return (vx > -0.5f) & (vy > -0.5f) & ((uint32_t)ix < (uint32_t)W) & ((uint32_t)iy < (uint32_t)H);
}

// One or two spatial entries (all a RandAugment(N=2) chain can produce), K in {none, Color}: the
// per-pixel arithmetic of the tile engine's gather_warp (exact float32 coordinates, src_index), the
// source being the resident image.
template <int C, bool COUNT, bool TWO>
__device__ __forceinline__ void res_gather_fast(const RC<C>& c) {
  const TileState& t = *c.t;
  const int H = c.H, W = c.W;
  const int n_sp = t.n_sp;
  const Spatial& ea = t.sp[n_sp - 1];
  const Spatial& eb = t.sp[TWO ? n_sp - 2 : 0];
  const bool a_geom = !TWO || ea.type == SP_GEOM;
  const bool b_geom = TWO && eb.type == SP_GEOM;
  const float t0 = ea.t[0], t1 = ea.t[1], t2 = ea.t[2], t3 = ea.t[3], t4 = ea.t[4], t5 = ea.t[5];
  // (entry b's coefficients and rectangle in registers too: read through the reference they were six shared-memory
  // loads per pixel of the second stage)
  const float u0 = eb.t[0], u1 = eb.t[1], u2 = eb.t[2], u3 = eb.t[3], u4 = eb.t[4], u5 = eb.t[5];
  const int by0 = eb.y0, by1 = eb.y1, bx0 = eb.x0, bx1 = eb.x1;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF), ac1 = (uint32_t)(t.l1_aff & 0xFF);
  const bool plain = (kmode == K_NONE) && (COUNT || !use1);
  const float f = t.kfactor;
  const uint32_t fill_a = c.ctl->fillc[0], fill_b = c.ctl->fillc[1];
  const uint32_t fa_addr = smem_addr(&c.ctl->fillc[0]), fb_addr = smem_addr(&c.ctl->fillc[1]);
  const int pitch = c.row;
  const float wmh = __fadd_rn((float)W, -0.5f), hmh = __fadd_rn((float)H, -0.5f);
  // source byte of raw indices (rx, ry): img + (ry - K) * pitch + (rx - K) * C
  const uint32_t base_raw = c.img - SRC_K * (uint32_t)pitch - SRC_K * (uint32_t)C;
  uint32_t n_fill_a = 0, n_fill_b = 0;
  // A thread takes G quads (4 adjacent pixels, C words each) at a time.  G = 1: the lanes of a warp read
  // source pixels that are neighbours along the warp's direction, i.e. nearly consecutive shared-memory
  // words (with G = 4 for C = 3 -- one 48-byte unit per lane -- the lanes are 16 source pixels apart and
  // 61 % of the gather's wavefronts were bank conflicts, profiles/r02_ab_notes.md 4).
#ifndef CHB_GATHER_G
#define CHB_GATHER_G 1
#endif
#ifndef CHB_GATHER_RAW
#define CHB_GATHER_RAW 0
#endif
  constexpr int G = CHB_GATHER_G;
  const int y_end = COUNT ? H : c.y_hi;
  UnitWalk q(c.tid, W / (4 * G), c.nt);
  if (!COUNT) q.y += c.y_lo;
  for (; q.y < y_end; q.next()) {
    const int y = q.y, xu = q.ux * (4 * G);
    const float fy = small_uint_to_float((uint32_t)y);
    const float t1y = __fmul_rn(t1, fy), t4y = __fmul_rn(t4, fy);
#pragma unroll
    for (int g = 0; g < G; ++g) {
    const int x0 = xu + 4 * g;
    const float fx0 = small_uint_to_float((uint32_t)x0);
    uint32_t adr[4];
    uint32_t hit = 0;  // bit i: pixel i shows entry a's colour, bit 4 + i: entry b's
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int ix = x0 + i, iy = y;
      bool hit_a, hit_b = false;
#if CHB_GATHER_RAW
      uint32_t rx = (uint32_t)ix + SRC_K, ry = (uint32_t)iy + SRC_K;
      if (a_geom) {
        const float fx = small_uint_to_float((uint32_t)ix);
        const bool inx = src_raw(__fadd_rn(__fadd_rn(__fmul_rn(t0, fx), t1y), t2), wmh, rx);
        const bool iny = src_raw(__fadd_rn(__fadd_rn(__fmul_rn(t3, fx), t4y), t5), hmh, ry);
        hit_a = !(inx && iny);
      } else {
        hit_a = (iy >= ea.y0) && (iy < ea.y1) && (ix >= ea.x0) && (ix < ea.x1);
      }
      if (TWO && !hit_a) {
        ix = (int)(rx - SRC_K); iy = (int)(ry - SRC_K);
        if (b_geom) {
          const float gx = small_uint_to_float((uint32_t)ix), gy = small_uint_to_float((uint32_t)iy);
          const bool inx = src_raw(__fadd_rn(__fadd_rn(__fmul_rn(eb.t[0], gx), __fmul_rn(eb.t[1], gy)), eb.t[2]), wmh, rx);
          const bool iny = src_raw(__fadd_rn(__fadd_rn(__fmul_rn(eb.t[3], gx), __fmul_rn(eb.t[4], gy)), eb.t[5]), hmh, ry);
          hit_b = !(inx && iny);
        } else {
          hit_b = (iy >= eb.y0) && (iy < eb.y1) && (ix >= eb.x0) && (ix < eb.x1);
        }
      }
      adr[i] = !(hit_a || hit_b) ? base_raw + ry * (uint32_t)pitch + rx * (uint32_t)C : (hit_a ? fa_addr : fb_addr);
#else
      if (a_geom) {
        const float fx = (i == 0) ? fx0 : __fadd_rn(fx0, (float)i);  // exact: small integers
        int jx, jy;
        hit_a = !src_index2(__fadd_rn(__fadd_rn(__fmul_rn(t0, fx), t1y), t2), __fadd_rn(__fadd_rn(__fmul_rn(t3, fx), t4y), t5), W, H, jx, jy);
        ix = jx; iy = jy;
      } else {
        hit_a = (iy >= ea.y0) && (iy < ea.y1) && (ix >= ea.x0) && (ix < ea.x1);
      }
      if (TWO && !hit_a) {
        if (b_geom) {
          const float gx = small_uint_to_float((uint32_t)ix), gy = small_uint_to_float((uint32_t)iy);
          int jx, jy;
          hit_b = !src_index2(__fadd_rn(__fadd_rn(__fmul_rn(u0, gx), __fmul_rn(u1, gy)), u2),
                              __fadd_rn(__fadd_rn(__fmul_rn(u3, gx), __fmul_rn(u4, gy)), u5), W, H, jx, jy);
          ix = jx; iy = jy;
        } else {
          hit_b = (iy >= by0) & (iy < by1) & (ix >= bx0) & (ix < bx1);
        }
      }
      adr[i] = !(hit_a | hit_b) ? c.img + (uint32_t)(iy * pitch + ix * C) : (hit_a ? fa_addr : fb_addr);
#endif
      hit |= (hit_a ? 1u : 0u) << i | (hit_b ? 16u : 0u) << i;
    }
    uint32_t v[4][C];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(adr[i] + ch);
    if (!plain) {
      if (kmode == K_NONE) {
        if (aff1) {
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
        if (!COUNT && use2) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[i][ch] = lds_u8(c.l2a + ch * 256 + v[i][ch]);
        }
      }
      if (!COUNT) {  // the spatial colours already are final colours
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ha = (hit >> i) & 1u, hb = (hit >> (4 + i)) & 1u;
          const uint32_t fillv = ha ? fill_a : fill_b;
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[i][ch] = (ha || hb) ? byte_of(fillv, ch) : v[i][ch];
        }
      }
    }
    if (COUNT) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool ha = (hit >> i) & 1u, hb = (hit >> (4 + i)) & 1u;
        if (!(ha || hb)) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) hist_add(c, ch, v[i][ch]);
        } else if (ha) {
          ++n_fill_a;
        } else {
          ++n_fill_b;
        }
      }
    } else {
      if (c.tally) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int ch = 0; ch < C; ++ch) hist_add(c, ch, v[i][ch]);
      }
      uint32_t o[C];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
          const int bi = i * C + ch;
          o[bi >> 2] = ((bi & 3) == 0) ? v[i][ch] : put_byte(o[bi >> 2], v[i][ch], bi & 3);  // (every v is a zero-extended byte)
        }
      if (c.dstf) {
        const size_t w0 = (((size_t)y * W + x0) * C) >> 2;  // a quad is C whole words starting at channel 0
#pragma unroll
        for (int w = 0; w < C; ++w) store_norm_word(c, w0 + w, (4 * w) % C, o[w]);
      } else {
      uint32_t* gp = reinterpret_cast<uint32_t*>(c.dst + ((size_t)y * W + x0) * C);
      if (C == 4) {
        __stcg(reinterpret_cast<uint4*>(gp), make_uint4(o[0], o[1 % C], o[2 % C], o[3 % C]));
      } else if (C == 2) {
        __stcg(reinterpret_cast<uint2*>(gp), make_uint2(o[0], o[1 % C]));
      } else {
#pragma unroll
        for (int w = 0; w < C; ++w) __stcg(gp + w, o[w]);
      }
      }
    }
    }
  }
  if (COUNT) {
    if (n_fill_a) atomicAdd(&c.st->color_cnt[n_sp - 1], n_fill_a);
    if (TWO && n_fill_b) atomicAdd(&c.st->color_cnt[n_sp - 2], n_fill_b);
  }
}

// One warp whose source row is the output row shifted by a whole number of rows and whose source column
// is x + s for a per-row constant s -- ShearX (s = level * y), TranslateX (s = pixels), TranslateY
// (s = 0) -- K == none.  The exact coordinate of the reference is sx = fl(x + s) (one float32 rounding:
// the entry has t0 == 1, t3 == 0, t4 == 1 and t1 == 0 or t2 == 0), then round-half-away.  For integer
// x the value fl(x + s) - x depends only on the binade (and sign) of x + s: the rounding grid 2^(e-23)
// divides 1, so fl(x + s) = x + rnd_e(s): shifts only change at binade boundaries.  Every pixel still
// gets its exact index (one add + the index step: 6 instructions instead of the general warp's 16), and
// a quad whose four indices are consecutive -- all but the quads on a binade boundary or the image
// border -- is copied as C + 1 aligned words and a funnel shift instead of 4 * C byte loads.
template <int C, bool COUNT>
__device__ __forceinline__ void res_gather_rowshift(const RC<C>& c) {
  const TileState& t = *c.t;
  const int H = c.H, W = c.W;
  const Spatial& ea = t.sp[0];
  const float t1 = ea.t[1], t2 = ea.t[2];
  const int dy = (int)ea.t[5];
  const bool use1 = !COUNT && !t.l1_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF) * 0x01010101u, ac1 = (uint32_t)(t.l1_aff & 0xFF) * 0x01010101u;
  const uint32_t fill_a = c.ctl->fillc[0];  // the entry's final colour bytes
  const int pitch = c.row;
  uint32_t n_fill = 0;
  // the fill quad: colour bytes repeated (C words)
  uint32_t fq[C];
#pragma unroll
  for (int w = 0; w < C; ++w) {
    uint32_t x = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) x |= byte_of(fill_a, (4 * w + b) % C) << (8 * b);
    fq[w] = x;
  }
  const int y_end = COUNT ? H : c.y_hi;
  UnitWalk q(c.tid, W >> 2, c.nt);
  if (!COUNT) q.y += c.y_lo;
  for (; q.y < y_end; q.next()) {
    const int y = q.y, x0 = q.ux << 2;
    const int iy = y + dy;
    uint32_t o[C];
    if ((unsigned)iy >= (unsigned)H) {  // the whole row shows the fill colour
      if (COUNT) { n_fill += 4; continue; }
#pragma unroll
      for (int w = 0; w < C; ++w) o[w] = fq[w];
    } else {
      const float fy = small_uint_to_float((uint32_t)y);
      const float s = (t1 == 0.0f) ? t2 : __fmul_rn(t1, fy);  // (the other term adds +-0)
      // the exact source column of every pixel: one add and the index step each
      int j[4];
      bool in[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) in[i] = src_index(__fadd_rn(small_uint_to_float((uint32_t)(x0 + i)), s), W, j[i]);
      const bool all_in = in[0] && in[1] && in[2] && in[3];
      const bool none_in = !(in[0] || in[1] || in[2] || in[3]);
      if (all_in && j[1] - j[0] == 1 && j[2] - j[0] == 2 && j[3] - j[0] == 3) {
        // four consecutive source pixels: C + 1 aligned words and a funnel shift
        const uint32_t a = c.img + (uint32_t)(iy * pitch + j[0] * C);
        const uint32_t aw = a & ~3u, sh = (a & 3u) * 8u;
        uint32_t w[C + 1];
#pragma unroll
        for (int k = 0; k <= C; ++k) w[k] = lds_u32(aw + 4 * k);  // (the word after the quad may belong to the next row / the pad: its bytes are shifted out)
#pragma unroll
        for (int k = 0; k < C; ++k) o[k] = __funnelshift_r(w[k], w[k + 1], sh);
        if (COUNT) {
#pragma unroll
          for (int k = 0; k < C; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) hist_add(c, (4 * k + b) % C, byte_of(o[k], b));
          continue;
        }
        if (use1) {
          if (aff1) {
#pragma unroll
            for (int k = 0; k < C; ++k) o[k] = (o[k] & am1) ^ ac1;
          } else {
#pragma unroll
            for (int k = 0; k < C; ++k)
              o[k] = map_word(o[k], c.l1a + ((4 * k + 0) % C) * 256, c.l1a + ((4 * k + 1) % C) * 256, c.l1a + ((4 * k + 2) % C) * 256,
                              c.l1a + ((4 * k + 3) % C) * 256);
          }
        }
      } else if (none_in) {
        if (COUNT) { n_fill += 4; continue; }
#pragma unroll
        for (int w = 0; w < C; ++w) o[w] = fq[w];
      } else {
        // the quad straddles the image border or a float32 binade boundary of x + s: byte gathers
#pragma unroll
        for (int k = 0; k < C; ++k) o[k] = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t v[C];
          if (in[i]) {
            const uint32_t a = c.img + (uint32_t)(iy * pitch + j[i] * C);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
            if (COUNT) {
#pragma unroll
              for (int ch = 0; ch < C; ++ch) hist_add(c, ch, v[ch]);
            } else if (use1) {
#pragma unroll
              for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l1a + ch * 256 + v[ch]);
            }
          } else {
            if (COUNT) ++n_fill;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = byte_of(fill_a, ch);
          }
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            const int bi = i * C + ch;
            o[bi >> 2] = ((bi & 3) == 0) ? v[ch] : put_byte(o[bi >> 2], v[ch], bi & 3);
          }
        }
        if (COUNT) continue;
      }
    }
    if (!COUNT && c.tally) {
#pragma unroll
      for (int k = 0; k < C; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) hist_add(c, (4 * k + b) % C, byte_of(o[k], b));
    }
    if (!COUNT && c.dstf) {
      const size_t w0 = (((size_t)y * W + x0) * C) >> 2;
#pragma unroll
      for (int w = 0; w < C; ++w) store_norm_word(c, w0 + w, (4 * w) % C, o[w]);
      continue;
    }
    uint32_t* gp = reinterpret_cast<uint32_t*>(c.dst + ((size_t)y * W + x0) * C);
    if (C == 4) {
      __stcg(reinterpret_cast<uint4*>(gp), make_uint4(o[0], o[1 % C], o[2 % C], o[3 % C]));
    } else if (C == 2) {
      __stcg(reinterpret_cast<uint2*>(gp), make_uint2(o[0], o[1 % C]));
    } else {
#pragma unroll
      for (int w = 0; w < C; ++w) __stcg(gp + w, o[w]);
    }
  }
  if (COUNT && n_fill) atomicAdd(&c.st->color_cnt[0], n_fill);
}

// Does the single spatial entry qualify for res_gather_rowshift?
__device__ __forceinline__ bool is_rowshift(const Spatial& e) {
  const float t5 = e.t[5];
  return e.type == SP_GEOM && e.t[0] == 1.0f && e.t[3] == 0.0f && e.t[4] == 1.0f && (e.t[1] == 0.0f || e.t[2] == 0.0f) &&
         fabsf(t5) < 4194304.0f && t5 == truncf(t5);
}

// General form: any list of constant-fill warps and masks, K in {none, Color}; one pixel per thread
// at a time, byte stores (chains of three and more spatial ops are rare and this path is not tuned).
template <int C, bool COUNT>
__device__ __forceinline__ void res_gather_list(const RC<C>& c) {
  const TileState& t = *c.t;
  const int H = opaque(c.H), W = opaque(c.W);
  const int n_sp = t.n_sp;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  const int n_pix = (COUNT ? H : c.y_hi) * W;
  for (int i = (COUNT ? 0 : c.y_lo * W) + c.tid; i < n_pix; i += c.nt) {
    const int y = i / W;
    int sx = i - y * W, sy = y;
    const int k = resolve(t.sp, n_sp, H, W, sx, sy);
    uint32_t v[C];
    if (k < 0) {
      const uint32_t a = c.img + (uint32_t)(sy * c.row + sx * C);
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
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
        color_pixel_f(small_uint_to_float(v[0]), small_uint_to_float(v[1 % C]), small_uint_to_float(v[2 % C]), f, v[0], v[1 % C], v[2 % C]);
        if (!COUNT && use2) {
#pragma unroll
          for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(c.l2a + ch * 256 + v[ch]);
        }
      }
      if (COUNT) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) hist_add(c, ch, v[ch]);
      }
    } else {
      if (COUNT) {
        atomicAdd(&c.st->color_cnt[k], 1u);
      } else {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = (uint32_t)t.sp[k].color[ch];
      }
    }
    if (!COUNT) {
      if (c.tally) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) hist_add(c, ch, v[ch] & 255u);
      }
      if (c.dstf) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
          c.dstf[(size_t)i * C + ch] = __uint_as_float(lds_u32(c.nlut + (uint32_t)(ch % 3) * 1024u + ((v[ch] & 255u) << 2)));
        continue;
      }
      uint8_t* d = c.dst + (size_t)i * C;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) d[ch] = (uint8_t)v[ch];
    }
  }
}

template <int C, bool COUNT>
__device__ __forceinline__ void res_gather(const RC<C>& c) {
  const TileState& t = *c.t;
#ifndef CHB_ROWSHIFT
#define CHB_ROWSHIFT 1
#endif
  const bool rowshift = CHB_ROWSHIFT && t.n_sp == 1 && t.kmode == K_NONE && is_rowshift(t.sp[0]);
  wait_image(c);  // (waiting per source row inside the row-shift copy was measured: the test per quad cost more than the overlap gained)
  if (COUNT) hist_zero(c);
  if (rowshift) res_gather_rowshift<C, COUNT>(c);
  else if (t.n_sp == 1 && t.sp[0].type == SP_GEOM) res_gather_fast<C, COUNT, false>(c);
  else if (t.n_sp == 2) res_gather_fast<C, COUNT, true>(c);
  else res_gather_list<C, COUNT>(c);
  if (COUNT) hist_reduce(c);
}

// ========================================================================== sharpness executors
// l1 applied to the resident source in place (a Sharpness tap reads every byte nine times); the view
// then continues with l1 = identity.
template <int C>
__device__ __forceinline__ void res_bake_l1(const RC<C>& c) {
  TileState& t = c.st->t;
  if (t.l1_id) return;
  wait_image(c);
  const int total = c.img_bytes >> 4;
  for (int i = c.tid; i < total; i += c.nt) {
    const uint4 v = map_vec_phase<C>(lds_v4(c.img + ((uint32_t)i << 4)), c.l1a, (C == 3) ? (i % 3) : 0);
    sts_v4(c.img + ((uint32_t)i << 4), v);
  }
  work_sync(c);
  for (int i = c.tid; i < MAXC * 256; i += c.nt) t.l1[i >> 8][i & 255] = (uint8_t)(i & 255);
  if (c.tid == 0) { t.l1_id = 1; t.l1_aff = 0; }
  fence_proxy_async();  // the image buffer is refilled by the TMA later
  work_sync(c);
}

// Rows per sub-strip so that (word columns x sub-strips) fills the CTA: minimise rounds * (rows + 2).
inline int res_sharp_split(int columns, int inner) {
  if (inner <= 0) return 1;
  int best_s = 1, best_cost = 0x7FFFFFFF;
  for (int S = 1; S <= 64 && S <= inner; ++S) {
    const int rounds = (columns * S + RNT - 1) / RNT;
    const int cost = rounds * ((inner + S - 1) / S + 2);
    if (cost < best_cost) { best_cost = cost; best_s = S; }
  }
  return (inner + best_s - 1) / best_s;
}

// Work splits of the Sharpness executors: they depend only on the shape and the size of the aux
// region, so the host computes them once per call (KParams::res_*).
template <int C>
void resident_splits_c(KParams& p, int aux_bytes) {
  (void)aux_bytes;
  const int H = p.H, W = p.W;
  const int wpr = (W * C) >> 2;
  for (int n = 1; n <= 4; ++n) {  // the last pass of an image cut into n row ranges (small batches)
    const int rows = (H + n - 1) / n;
    const int inner = n == 1 ? H - 2 : rows;
    p.res_sharp_rows[n - 1] = res_sharp_split(wpr, inner);
  }
}

// tfa.image.sharpness (oracle/ops.py sharpness) down one word column (4 output bytes), streaming: the
// float32 chain of an output byte, (((((((p00 + p01) + p02) + p10) + c11 * 5/13) + p12) + p20) + p21) + p22
// with p = value * 1/13, is carried in two partial sums per byte -- T (the row above is in) and A (the
// byte's own row is in) -- so a row is loaded, converted and multiplied once and only 8 accumulators
// stay live across rows (the tile engine's sharp_walk keeps the 30 products of three rows: it needs
// 90+ registers, this one fits the 64 of a 1024-thread CTA).  Products are one FFMA each: the byte
// is placed in the mantissa of 2^23 (one PRMT), and fma(2^23 + v, k, -2^23 * k) rounds the exact
// v * k once -- the same float32 as float(v) * k.
//   col      shared address of this column's word in the row ABOVE the first output row
//   n        output rows; bmask bit b: byte b lies in the first / last pixel of the image row
template <int C, class Emit>
__device__ __forceinline__ void sharp_stream(uint32_t col, int pitch, int n, bool has_prev, bool has_next, uint32_t bmask,
                                             float f, Emit emit) {
  constexpr int NB = 4 + 2 * C;  // window bytes per row
  const float k1 = __int_as_float(0x3d9d89d9);  // float32(1)/float32(13)
  const float k5 = __int_as_float(0x3ec4ec4f);  // float32(5)/float32(13)
  const float nk1 = -8388608.0f * k1, nk5 = -8388608.0f * k5;  // exact (powers of two)
  float T[4], A[4];
  uint32_t w_mid = 0;  // centre word of the row whose outputs are being accumulated in A
  float p[NB], xm[4];
  uint32_t w1 = 0;
  auto load_row = [&](uint32_t ra) {
    w1 = lds_u32(ra);
    const uint32_t w0 = has_prev ? lds_u32(ra - 4) : 0u;
    const uint32_t w2 = has_next ? lds_u32(ra + 4) : 0u;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int wb = 4 - C + j;  // byte index in the 12-byte (w0, w1, w2) window
      const uint32_t wsel = (wb < 4) ? w0 : (wb < 8) ? w1 : w2;
      const float m = __uint_as_float(__byte_perm(wsel, 0x4B000000u, 0x7440u | (uint32_t)(wb & 3)));  // 2^23 + v
      p[j] = __fmaf_rn(m, k1, nk1);
      if (j >= C && j < C + 4) xm[j - C] = m;
    }
  };
  // row above the first output row: T only
  load_row(col);
#pragma unroll
  for (int b = 0; b < 4; ++b) T[b] = __fadd_rn(__fadd_rn(p[b], p[b + C]), p[b + 2 * C]);
  uint32_t ra = col + pitch;
  for (int r = 0; r <= n; ++r, ra += pitch) {
    load_row(ra);
    if (r > 0) {  // this row completes output row r - 1
      uint32_t o = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float acc = __fadd_rn(A[b], p[b]);
        acc = __fadd_rn(acc, p[b + C]);
        acc = __fadd_rn(acc, p[b + 2 * C]);
        const float orig = byte_to_float(w_mid, b);
        float deg = __fadd_rn(__fadd_rz(acc, 8388608.0f), -8388608.0f);  // float(trunc(acc))
        if (bmask) deg = ((bmask >> b) & 1u) ? orig : deg;                // border pixels keep the original (two word columns of a row)
        // tfa blend: rint(clip(deg + f * (orig - deg))); the byte is the low byte of the magic-number sum
        const uint32_t res = rint_bits(clamp255(__fadd_rn(deg, __fmul_rn(f, __fsub_rn(orig, deg)))));
        o = (b == 0) ? byte_of(res, 0) : put_byte(o, res, b);
      }
      emit(r - 1, o);
    }
    if (r < n) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        // the byte's own row: + left, + centre * 5/13, + right
        float acc = __fadd_rn(T[b], p[b]);
        acc = __fadd_rn(acc, __fmaf_rn(xm[b], k5, nk5));
        A[b] = __fadd_rn(acc, p[b + 2 * C]);
        // and it is the row above output row r + 1
        T[b] = __fadd_rn(__fadd_rn(p[b], p[b + C]), p[b + 2 * C]);
      }
      w_mid = w1;
    }
  }
}

// K == Sharpness, no spatial op pending: the column walk of the tile engine (sharp_walk) straight on
// the resident image; a thread owns one word column of a run of rows and stores its words itself
// (adjacent lanes, adjacent words).
template <int C, bool COUNT>
__device__ __forceinline__ void res_sharp(const RC<C>& c) {
  res_bake_l1(c);
  wait_image(c);
  if (COUNT) hist_zero(c);
  const TileState& t = *c.t;
  const int H = c.H, row = c.row;
  const bool use2 = !t.l2_id;
  const float f = t.kfactor;
  const int wpr = row >> 2;
  auto emit = [&](int y, int xw, int ph, uint32_t o) {
    if (COUNT) {
#pragma unroll
      for (int b = 0; b < 4; ++b) hist_add(c, (ph + b) % C, byte_of(o, b));
    } else {
      if (use2)
        o = map_word(o, c.l2a + (uint32_t)((ph + 0) % C) * 256u, c.l2a + (uint32_t)((ph + 1) % C) * 256u,
                     c.l2a + (uint32_t)((ph + 2) % C) * 256u, c.l2a + (uint32_t)((ph + 3) % C) * 256u);
      if (c.tally) {
#pragma unroll
        for (int b = 0; b < 4; ++b) hist_add(c, (ph + b) % C, byte_of(o, b));
      }
      if (c.dstf) store_norm_word(c, (size_t)y * (size_t)(row >> 2) + (size_t)xw, ph, o);
      else __stcg(reinterpret_cast<uint32_t*>(c.dst + (size_t)y * row) + xw, o);
    }
  };
  // first and last image row: every pixel is border -> blend(orig, orig) == orig
  const int y_lo = COUNT ? 0 : c.y_lo, y_hi = COUNT ? H : c.y_hi;
  if (y_lo == 0)
    for (int xw = c.tid; xw < wpr; xw += c.nt) emit(0, xw, (C == 3) ? (xw % 3) : 0, lds_u32(c.img + ((uint32_t)xw << 2)));
  if (H > 1 && y_hi == H)
    for (int xw = c.tid; xw < wpr; xw += c.nt)
      emit(H - 1, xw, (C == 3) ? (xw % 3) : 0, lds_u32(c.img + (uint32_t)((H - 1) * row + (xw << 2))));
  const int in0 = max(1, y_lo), in1 = min(H - 1, y_hi);
  const int inner = in1 - in0;
  if (inner > 0) {
    const int R = c.sharp_rows;
    const int n_strips = (inner + R - 1) / R;
    const int n_items = wpr * n_strips;
    for (int item = c.tid; item < n_items; item += c.nt) {
      const int strip = item / wpr, xw = item - strip * wpr;
      const int y_begin = in0 + strip * R, y_end = min(in1, y_begin + R);
      const int xb0 = xw << 2;
      uint32_t bmask = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (xb0 + b < C || xb0 + b >= row - C) bmask |= 1u << b;
      const int ph = (C == 3) ? (xw % 3) : 0;
      const uint32_t col = c.img + (uint32_t)((y_begin - 1) * row + xb0);
      sharp_stream<C>(col, row, y_end - y_begin, xw > 0, xw + 1 < wpr, bmask, f,
                      [&](int r, uint32_t o) { emit(y_begin + r, xw, ph, o); });
    }
  }
  if (COUNT) hist_reduce(c);
}

// ================================================================ claim order of small batches
// A CTA works on one image at a time, so a 256-image call is 1.7 images per SM and its duration is the
// most loaded SM's: which images an SM gets matters more than anything else.  Every CTA therefore
// decodes the op chain of EVERY image of a small batch (Philox again, or the replayed schedule), turns
// it into a cost estimate and sorts the work by decreasing cost -- the same order in every CTA, no
// communication -- and the claim counter indexes that order: the expensive chains start first, one
// per SM, and the cheap ones fill the gaps (LPT).  In the smallest batches (<= SPLIT_MAX images) the
// LAST pass of an image that alone would outlast the average SM is cut into 2..4 row ranges, each a
// work item of its own: every part loads the image and repeats the passes in front of the last one
// (they need the whole image), then writes only its rows.
// Only the SCHEDULE depends on the estimate; the pixels do not.
constexpr int SPLIT_MAX = 512;
constexpr int ORDER_IMG_BITS = 11;  // order entry: image | part << 11 | (parts - 1) << 13

// Estimated microseconds of one image at 224 x 224 x 3 (profiles/r02_timeline_256.txt), following the
// rules of the chain walk: `sunk` = loading, bookkeeping and every pass in front of the last one (paid
// by every part), `pend` = the evaluation of the final view (shared between the parts).
__device__ __forceinline__ void chain_cost(const KParams& p, const DevOp* ops, int img, int& sunk, int& pend) {
  const unsigned long long own = p.image_index_base + (unsigned long long)img;
  const unsigned long long stream_img = p.elementwise ? own : ~0ull;
  const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32));
  int n_geo = 0, n_mask = 0, k = 0, lut = 0;  // k: 0 none, 1 Color, 2 Sharpness
  sunk = 8; pend = 0;
  auto materialise = [&](int extra) { sunk += pend + 3 + extra; pend = 0; n_geo = n_mask = 0; k = 0; lut = 0; };
  for (int i = 0; i < p.n_draws; ++i) {
    const int slot0 = i * (p.K + 1);
    const size_t rbase = ((size_t)img * p.n_draws + i) * p.K * CHB_SCHED_FIELDS;
    int choice;
    if (p.replay) {
      choice = p.replay[rbase];
      if (choice < 0 || choice >= p.T) choice = 0;
    } else {
      choice = (int)__umulhi(philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter, (uint32_t)slot0), key).x,
                             (uint32_t)p.T);
    }
    for (int j = 0; j < p.K; ++j) {
      const DevOp& op = ops[choice * p.K + j];
      const int kind = op.kind;
      bool applied;
      if (p.replay) {
        applied = p.replay[rbase + (size_t)j * CHB_SCHED_FIELDS + 1] != 0 && kind >= 0;
      } else {
        const uint32_t r = philox4x32_10(make_uint4((uint32_t)stream_img, (uint32_t)(stream_img >> 32), p.call_counter,
                                                    (uint32_t)(slot0 + 1 + j)), key).x;
        applied = kind >= 0 && (int)(r >> 8) < op.thr24;
      }
      if (!applied) continue;
      switch (kind) {
        case CHB_OP_AUTOCONTRAST: case CHB_OP_EQUALIZE:
          if (n_geo > 0 || k == 2) materialise(4);            // tallied while written, reload
          else if (n_mask > 0 || k == 1) materialise(12);      // in place, flat COUNT
          else sunk += 12;                                     // flat COUNT
          pend += 5;                                           // flat apply
          break;
        case CHB_OP_COLOR:
          if (op.blend_mode == BLEND_IMAGE2) break;
          if (k != 0) materialise(0);
          k = 1; pend += 13;
          break;
        case CHB_OP_SHARPNESS:
          if (op.blend_mode == BLEND_IMAGE2) break;
          if (k != 0 || n_geo + n_mask > 0) materialise(0);
          k = 2; pend += lut ? 36 : 30;
          break;
        case CHB_OP_CUTOUT:
          if (k == 2) materialise(0);
          ++n_mask; pend += (n_geo > 0) ? 9 : 3;
          break;
        case CHB_OP_SHEAR_X: case CHB_OP_SHEAR_Y: case CHB_OP_TRANSLATE_X: case CHB_OP_TRANSLATE_Y: case CHB_OP_ROTATE:
          if (k == 2) materialise(0);
          ++n_geo; pend += (n_geo + n_mask > 1) ? 14 : 8;
          break;
        case CHB_OP_INVERT: case CHB_OP_POSTERIZE: case CHB_OP_SOLARIZE:
          break;
        default:  // Brightness, Contrast, SolarizeAdd: table lookups
          if (!lut) { lut = 1; pend += 5; }
          break;
      }
    }
  }
}

// Fills ctl->order / ctl->n_items for a batch of B <= LPT_MAX images.  `scratch`: 64 x 64 uint16 + 64 int
// of shared memory.
__device__ void plan_order(const KParams& p, const DevOp* ops, ResCtl* ctl, uint16_t* scratch, int tid) {
  constexpr int NBK = 64;                       // cost buckets (bucket 0 = most expensive)
  constexpr int NR = LPT_MAX / RNT;             // images per thread
  const int B = p.B;
  const int lane = tid & 31;
  bool split = p.res_split && B <= SPLIT_MAX;
  for (int i = tid; i < NBK * NBK; i += RNT) scratch[i] = 0;
  int sunk[NR], pend[NR];
  int mine = 0;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int i = r * RNT + tid;
    sunk[r] = pend[r] = 0;
    if (i < B) { chain_cost(p, ops, i, sunk[r], pend[r]); mine += sunk[r] + pend[r]; }
  }
  mine = __reduce_add_sync(0xFFFFFFFFu, mine);
  if (lane == 0 && mine) atomicAdd(&ctl->cost_sum, mine);
  __syncthreads();
  // an item should not outlast KParams::res_split_pct % of the average SM's share of the batch
  const int target = max(12, (int)((long long)ctl->cost_sum * p.res_split_pct / (100 * (long long)gridDim.x)));
  // pass 1: parts and bucket of every image; count per (32-item-id chunk, bucket)
  int bk[NR], parts[NR];
  int any_split = 0;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int i = r * RNT + tid;
    parts[r] = 1;
    if (i < B && split)
      while (parts[r] < 4 && sunk[r] + pend[r] / parts[r] > target && pend[r] / (parts[r] + 1) >= 15) ++parts[r];  // (only last passes of >= 30 us are worth a second load of the image: Sharpness)
    any_split |= parts[r] > 1;
  }
  split = __syncthreads_or(any_split) != 0;     // nothing to split (e.g. AutoAugment holds no Sharpness): the cheaper one-slot sort
  const int slots = split ? 4 : 1;              // item id = image * slots + part: the sort is stable in it
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int i = r * RNT + tid;
    bk[r] = 0;
    if (i < B) {
      const int cst = sunk[r] + pend[r] / parts[r];
      bk[r] = NBK - 1 - min(cst >> 1, NBK - 1);
      for (int q = 0; q < parts[r]; ++q) {
        const int e = ((i * slots + q) >> 5) * NBK + bk[r];
        atomicAdd(reinterpret_cast<unsigned int*>(scratch) + (e >> 1), (e & 1) ? 0x10000u : 1u);
      }
    }
  }
  __syncthreads();
  // exclusive prefix over (bucket major, chunk minor): a warp scans the chunks of one bucket at a time (lane l
  // owns chunks l, l + 32, l + 64: at most 96 chunks), then warp 0 scans the 64 bucket totals
  const int n_chunks = (B * slots + 31) >> 5;
  int* bucket_total = reinterpret_cast<int*>(scratch + NBK * NBK);  // (no static shared memory: the kernel takes the SM's whole carve-out dynamically)
  for (int b = tid >> 5; b < NBK; b += RNT / 32) {
    int v[3], run = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ch = lane + 32 * k;
      v[k] = ch < n_chunks ? (int)scratch[ch * NBK + b] : 0;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int incl = v[k];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += o;
      }
      const int ch = lane + 32 * k;
      if (ch < n_chunks) scratch[ch * NBK + b] = (uint16_t)(run + incl - v[k]);
      run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) bucket_total[b] = run;
  }
  __syncthreads();
  if (tid < 32) {
    const int a = bucket_total[2 * tid], b2 = bucket_total[2 * tid + 1];
    int incl = a + b2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (tid >= d) incl += o;
    }
    bucket_total[2 * tid] = incl - a - b2;
    bucket_total[2 * tid + 1] = incl - b2;
    if (tid == 31) ctl->n_items = incl;
  }
  __syncthreads();
  // pass 2: position = bucket start + items of the bucket in earlier chunks + rank among the chunk's item ids
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const int i = r * RNT + tid;
    const bool valid = i < B;
    if (split) {
      // a chunk of 32 item ids is 8 consecutive images = the 8 lanes lane & ~7 .. of this warp.  Items of the
      // same chunk and bucket with a smaller id: every part of the earlier images of the group that share the
      // bucket, plus this image's own earlier parts.
      if (__ballot_sync(0xFFFFFFFFu, valid) == 0u) continue;  // (warp-uniform: a batch of 256 fills 8 of the 32 warps)
      int base = 0;
      const int grp = lane & ~7;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int ob = __shfl_sync(0xFFFFFFFFu, bk[r], grp + o);
        const int op = __shfl_sync(0xFFFFFFFFu, valid ? parts[r] : 0, grp + o);
        if (grp + o < lane && ob == bk[r]) base += op;
      }
      if (valid)
        for (int q = 0; q < parts[r]; ++q) {
          const int id = i * slots + q;
          const int pos = bucket_total[bk[r]] + scratch[(id >> 5) * NBK + bk[r]] + base + q;
          ctl->order[pos] = (uint16_t)(i | (q << ORDER_IMG_BITS) | ((parts[r] - 1) << (ORDER_IMG_BITS + 2)));
        }
    } else {
      const unsigned act = __ballot_sync(0xFFFFFFFFu, valid);
      if (valid) {
        const unsigned same = __match_any_sync(act, bk[r]);
        const int rank = __popc(same & ((1u << lane) - 1u));
        const int pos = bucket_total[bk[r]] + scratch[(i >> 5) * NBK + bk[r]] + rank;
        ctl->order[pos] = (uint16_t)i;
      }
    }
  }
  __syncthreads();
}

// ================================================================================= the kernel
// With KParams::res_rules the chain walk materialises a view before Sharpness or a histogram op sees a
// spatial list or an occupied K slot, so COUNT passes only ever meet the flat point-wise view and
// Sharpness never has a warp pending; the general list gather stays as the safety net of the former.
template <int C>
__device__ __forceinline__ void res_count_pass(RC<C>& c) {
  const TileState& t = *c.t;
  if (t.kmode == K_NONE && t.n_sp == 0) res_flat<C, true>(c, 0);
  else if (t.kmode != K_SHARP) { wait_image(c); hist_zero(c); res_gather_list<C, true>(c); hist_reduce(c); }
  else __trap();
}

template <int C>
__device__ __forceinline__ void res_write_pass(RC<C>& c, int store) {
  const TileState& t = *c.t;
  bool any_geom = false;
  for (int k = 0; k < t.n_sp; ++k) any_geom = any_geom || (t.sp[k].type == SP_GEOM);
  if (t.kmode == K_SHARP) {
    if (t.n_sp > 0) __trap();
    res_sharp<C, false>(c);
  } else if (!any_geom) {
    res_flat<C, false>(c, store);
  } else {
    res_gather<C, false>(c);
  }
}

// The chain walk is a sequence of short dependent steps separated by barriers: it runs on the first
// CHB_WALK_NT threads with a named barrier (a 1024-thread barrier costs ~200 cycles a time, and 32 warps
// replicating the walk's scalar control flow were 15 % of all executed instructions of a mixed batch,
// profiles/r02_ncu_op_RandAugment_b2048.txt); the table loops (768 entries) still get enough threads.
#ifndef CHB_WALK_NT
#define CHB_WALK_NT 128
#endif
template <int C>
__device__ __forceinline__ void res_advance(ResCtl* ctl, ImgState* st, const KParams& pl, int H, int W, int tid) {
  if (tid < CHB_WALK_NT)
    advance(st, st, pl, C, H, W, ctl->hmap, ctl->etab, tid, CHB_WALK_NT,
            [] { asm volatile("bar.sync 1, %0;" ::"n"(CHB_WALK_NT) : "memory"); });
  __syncthreads();
}

// Schedule decode and chain walk of the NEXT image on one warp (warp 31), while the other 31 warps run the current
// image's last pass: slower than the cooperative walk (~5 us against ~3) but entirely off the critical path.  The
// first walk of an image never builds a histogram table (nothing has been counted yet), the one step that wants a
// warp per channel; hmap / etab are free during a last pass (only COUNT passes and the walk behind them use them).
template <int C>
__device__ __forceinline__ void res_bookkeeping_warp(ResCtl* ctl, ImgState* st, const KParams& pl, int image, int H, int W, int lane) {
  for (int i = lane; i < STATE_VECS; i += 32) reinterpret_cast<uint4*>(st)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  decode_image(pl, st, ctl->rnd, ctl->rndc, image, H, W, lane);
  reset_view(st, lane, 32);
  __syncwarp();
  advance(st, st, pl, C, H, W, ctl->hmap, ctl->etab, lane, 32, [] { __syncwarp(); });
}

template <int C>
__global__ void __launch_bounds__(RNT, 1) resident_kernel(const KParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  ResCtl* ctl = reinterpret_cast<ResCtl*>(smem_raw);
  const int tid = threadIdx.x;
  const int H = p.H, W = p.W;
  const int img_bytes = H * W * C;
  const int n_chunks = (img_bytes + RES_CHUNK - 1) / RES_CHUNK;
  // shared memory: control block | policy table (DevOp[n_ops], value maps [n_ops][256]) | image | aux region
  const int n_pol = p.T * p.K;
  const uint32_t pol_off = (uint32_t)((sizeof(ResCtl) + 127) / 128 * 128);
  const uint32_t img_off = pol_off + (uint32_t)((n_pol * (int)(sizeof(DevOp) + 256) + 127) / 128 * 128);
  const uint32_t aux_off = img_off + (uint32_t)((img_bytes + 127) / 128 * 128);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  RC<C> c;
  ImgState* cur = &ctl->st[0];  // the current image's state; the other buffer is where warp 31 prepares the next image
  c.p = &p; c.ctl = ctl; c.st = cur; c.t = &cur->t; c.nt = RNT;
  c.img = smem_addr(smem_raw) + img_off;
  c.aux = smem_addr(smem_raw) + aux_off;
  c.aux_bytes = p.res_smem_bytes - (int)aux_off;
  c.full0 = smem_addr(&ctl->full[0]);
  c.H = H; c.W = W; c.row = W * C; c.img_bytes = img_bytes; c.tid = tid; c.lane = tid & 31;
  c.l1a = smem_addr(&cur->t.l1[0][0]); c.l2a = smem_addr(&cur->t.l2[0][0]);
  int ncopy = 8;
  while (ncopy > 1 && hist_bytes<C>(ncopy) > (uint32_t)c.aux_bytes) ncopy >>= 1;
  c.par = 1;  // toggled to 0 by the first load
  c.dst = nullptr;
  if (tid == 0) {
    for (int k = 0; k < RES_MAXCHUNK; ++k) mbar_init(c.full0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    // shape-only work splits (computed on the host, resident_splits)
    for (int k = 0; k < 4; ++k) ctl->sharp_rows[k] = p.res_sharp_rows[k];
    ctl->n_items = p.B;
    ctl->cost_sum = 0;
  }
  // Programmatic dependent launch.  What this call alone owns is set up BEFORE waiting for the previous kernel of
  // the stream -- the policy table (uploaded and synchronised when the policy was first seen, chb_api.cu get_policy),
  // the normalisation table, and, unless the schedule is replayed from a buffer that kernel may have written, the
  // cost-sorted claim order -- so that a back-to-back call does this part in the shadow of its predecessor's tail.
  // Nothing the previous kernel wrote (the images, a replayed schedule, the work counter it left zeroed) is touched
  // before it has completed.
  // the policy table comes to shared memory once: the chain walk of every image reads it many times
  DevOp* s_ops = reinterpret_cast<DevOp*>(smem_raw + pol_off);
  uint8_t* s_optab = smem_raw + pol_off + (size_t)n_pol * sizeof(DevOp);
  for (int i = tid; i < n_pol * (int)(sizeof(DevOp) / 16); i += RNT)
    reinterpret_cast<uint4*>(s_ops)[i] = __ldg(reinterpret_cast<const uint4*>(p.ops) + i);
  for (int i = tid; i < n_pol * 16; i += RNT)
    reinterpret_cast<uint4*>(s_optab)[i] = __ldg(reinterpret_cast<const uint4*>(p.optab) + i);
  if (tid == 0) {  // what decode_image / advance see
    ctl->kp = p;
    ctl->kp.ops = s_ops;
    ctl->kp.optab = s_optab;
  }
  const KParams& pl = ctl->kp;
  if (p.outf)  // fused normalisation epilogue: the 3 x 256 float32 values of the standalone layer
    for (int i = tid; i < 3 * 256; i += RNT) ctl->nlut[i >> 8][i & 255] = norm_value((float)(i & 255), p.norm_mode, i >> 8);
  c.nlut = smem_addr(&ctl->nlut[0][0]);
  c.dstf = nullptr;
  __syncthreads();
  // small batches: positions map to images through the cost-sorted order (all images share one chain in
  // batch mode: nothing to sort)
  const bool lpt = p.B <= LPT_MAX && p.elementwise && p.res_lpt && (p.B > (int)gridDim.x || p.res_split);
  int part = 0, n_parts = 1;
  auto image_of = [&](int pos) {     // claim position -> image (and the row range of its last pass)
    part = 0; n_parts = 1;
    if (pos >= ctl->n_items) return p.B;  // (n_items == p.B unless images were split)
    if (!lpt) return pos;
    const int e = (int)ctl->order[pos];
    part = (e >> ORDER_IMG_BITS) & 3;
    n_parts = ((e >> (ORDER_IMG_BITS + 2)) & 3) + 1;
    return e & ((1 << ORDER_IMG_BITS) - 1);
  };
  // schedule decode + chain walk of an image up to its first pass
  auto bookkeeping = [&](int image) {
    for (int i = tid; i < STATE_VECS; i += RNT) reinterpret_cast<uint4*>(cur)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid < 32) decode_image(pl, cur, ctl->rnd, ctl->rndc, image, H, W, tid);
    reset_view(cur, tid, RNT);
    __syncthreads();
    res_advance<C>(ctl, cur, pl, H, W, tid);
  };
  // A CTA's FIRST item is position blockIdx.x (no claim needed: later claims start behind the grid), so when the
  // schedule comes from the RNG and is not recorded, its bookkeeping runs before the wait as well.
  const bool early = !p.replay && !p.record;
  int img = p.B;
  if (early) {
    if (lpt) plan_order(pl, s_ops, ctl, reinterpret_cast<uint16_t*>(smem_raw + aux_off), tid);
    img = image_of((int)blockIdx.x);
    if (img < p.B) bookkeeping(img);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (!early) {
    if (lpt) plan_order(pl, s_ops, ctl, reinterpret_cast<uint16_t*>(smem_raw + aux_off), tid);
    img = image_of((int)blockIdx.x);
  }
  bool booked = early;  // the current image's bookkeeping is already done
  uint8_t* scratch = p.scratch + (size_t)blockIdx.x * p.scratch_stride;
  // Thread 0 issues the chunks IN ORDER: the executors consume them front to back, and chunks issued
  // by different lanes of a warp were served in an order that made every step of a flat image wait for
  // the tail of the load (profiles/r02_ab_notes.md 1: identity 102 % -> 79 % of the copy peak).
  auto issue_load = [&](const uint8_t* src) {
    if (tid == 32) {  // (not thread 0: that one drives the claims and the bulk stores)
      for (int k = 0; k < n_chunks; ++k) {
        const uint32_t bytes = (uint32_t)min(RES_CHUNK, img_bytes - k * RES_CHUNK);
        mbar_arrive_expect_tx(c.full0 + 8 * k, bytes);
        bulk_load(c.img + (uint32_t)k * RES_CHUNK, src + (size_t)k * RES_CHUNK, bytes, c.full0 + 8 * k);
      }
    }
  };
#ifdef CHB_TIMELINE
  // debug builds: one record of 4 words per processed image -- (cta << 32 | image), start, first pass start, end
  // (ns, %globaltimer) -- appended to KParams::timeline; word 0 counts the records, words 1-3: kernel entry of
  // the first CTA to get there, unused, unused.
  unsigned long long tl_t0 = 0, tl_t1 = 0;
  if (p.timeline && tid == 0) atomicMin(p.timeline + 1, tl_now());
#endif
  while (img < p.B) {
#ifdef CHB_TIMELINE
    if (tid == 0) tl_t0 = tl_now();
#endif
    c.par ^= 1u;
    issue_load(p.in + (size_t)img * img_bytes);
    // schedule decode + chain walk up to the first pass (the loads are in flight)
    if (!booked) bookkeeping(img);
    booked = false;
    uint8_t* out_img = p.outf ? nullptr : p.out + (size_t)img * img_bytes;
#ifdef CHB_TIMELINE
    if (tid == 0) tl_t1 = tl_now();
#endif
    for (;;) {
      const TileState& t = cur->t;
      const int pass_kind = t.pass_kind;
      if (tid == 0) {
        uint32_t fa = 0, fb = 0;
        for (int ch = 0; ch < C; ++ch) {
          fa |= (uint32_t)(t.sp[max(t.n_sp - 1, 0)].color[ch] & 255) << (8 * ch);
          fb |= (uint32_t)(t.sp[max(t.n_sp - 2, 0)].color[ch] & 255) << (8 * ch);
        }
        ctl->fillc[0] = fa; ctl->fillc[1] = fb;
      }
      c.ncopy = ncopy;
      c.hcopy = c.aux + (uint32_t)(c.lane & (ncopy - 1)) * hist_copy_bytes<C>();
      __syncthreads();
      c.tally = 0;
      c.minmax = 0;
      if (pass_kind == PASS_COUNT) {
        // AutoContrast reads only the smallest and the largest value of each channel (image_augmentations.py:
        // 69-70).  If the table in front of it (l1) is monotone in every channel, the extremes of l1's INPUT
        // bytes decide them: a min / max pass in registers replaces 150 K shared-memory atomics.  The
        // resulting histogram is presence-only and marked so (hist_valid == 2: it serves this op alone).
        bool want_mm = false;
        if (t.kmode == K_NONE && t.n_sp == 0 && cur->next_op < cur->n_prog)
          want_mm = s_ops[cur->prog[cur->next_op].table_index].kind == CHB_OP_AUTOCONTRAST;
        if (want_mm) {
          if (tid < 2) ctl->mono[tid] = 1;
          if (tid < MAXC) { ctl->mm[tid][0] = 255u; ctl->mm[tid][1] = 0u; }
          __syncthreads();
          for (int i = tid; i < C * 256; i += RNT)
            if ((i & 255) != 255) {
              const int a = t.l1[i >> 8][i & 255], b = t.l1[i >> 8][(i & 255) + 1];
              if (a > b) ctl->mono[0] = 0;
              if (a < b) ctl->mono[1] = 0;
            }
          __syncthreads();
          c.minmax = (ctl->mono[0] || ctl->mono[1]) ? 1 : 0;
        }
        res_count_pass<C>(c);
        if (tid == 0) cur->hist_valid = c.minmax ? 2 : 1;
        __syncthreads();
        res_advance<C>(ctl, cur, pl, H, W, tid);
        continue;
      }
      // WRITE_OUT, or WRITE_SCRATCH: point-wise views (and CutOut rectangles, painted over them)
      // materialise in the resident image itself, the rare neighbourhood-of-neighbourhood chains go
      // through this CTA's scratch image and come back through the TMA
      const bool last = pass_kind == PASS_WRITE_OUT;
      // The next image is claimed when this one's last pass starts: the round trip hides behind the pass, and
      // -- unlike a claim at the image's start -- a CTA with a long chain does not sit on a second image that
      // an idle CTA could have taken (profiles/r02_timeline_256.txt: the tail of a 256-image call was the
      // pre-claimed image behind the longest chain).
      if (last && tid == 0) ctl->n_claimed = (int)(atomicAdd(p.counters, 1u) + gridDim.x);  // (positions 0 .. grid-1 are the CTAs' first items)
      bool any_geom = false;
      for (int k = 0; k < t.n_sp; ++k) any_geom = any_geom || (t.sp[k].type == SP_GEOM);
      const bool in_place = !last && (t.kmode == K_NONE || t.kmode == K_COLOR) && !any_geom;
      c.dst = last ? out_img : scratch;
      c.dstf = (last && p.outf) ? p.outf + (size_t)img * img_bytes : nullptr;
      c.y_lo = last ? (H * part) / n_parts : 0;
      c.y_hi = last ? (H * (part + 1)) / n_parts : H;
      c.sharp_rows = ctl->sharp_rows[last ? n_parts - 1 : 0];
      // a view materialised for a histogram op is counted while it is written (gathers and Sharpness:
      // their bytes pass through registers anyway); in-place materialisations are counted by a flat
      // COUNT pass afterwards
      bool tally = false;
      if (!last && !in_place && cur->next_op < cur->n_prog) {
        const int kind = s_ops[cur->prog[cur->next_op].table_index].kind;
        tally = (kind == CHB_OP_EQUALIZE || kind == CHB_OP_AUTOCONTRAST);
      }
      if (tally) {
        c.tally = 1;
        wait_image(c);
        hist_zero(c);
        for (int i = tid; i < MAXC * 256; i += RNT) (&cur->hist[0][0])[i] = 0u;
      }
      res_write_pass<C>(c, last ? 1 : 0);
      if (last) break;
      if (tally) hist_reduce(c);
      if (!in_place) {
        __threadfence();
        fence_proxy_async_all();
        __syncthreads();
        c.par ^= 1u;
        issue_load(scratch);
      }
      reset_view(cur, tid, RNT);
      __syncthreads();
      if (tally && tid == 0) cur->hist_valid = 1;  // reset_view cleared it: the tallied bytes ARE the new source
      res_advance<C>(ctl, cur, pl, H, W, tid);
    }
    // every thread is done with the resident source; bulk stores out of it have read their bytes;
    // in-place writes (generic proxy) are ordered before the TMA refills the buffer
    fence_proxy_async();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();
#ifdef CHB_TIMELINE
    if (p.timeline && tid == 0) {
      const unsigned long long slot = atomicAdd(p.timeline, 1ull);
      if (slot < 8000ull) {
        unsigned long long* r = p.timeline + 4 + slot * 4;
        r[0] = ((unsigned long long)blockIdx.x << 32) | (unsigned)img; r[1] = tl_t0; r[2] = tl_t1; r[3] = tl_now();
      }
    }
#endif
    img = image_of(ctl->n_claimed);
    __syncthreads();
  }
  if (tid == 0) {
    bulk_wait_all0();
    // The last CTA out leaves the work counter zeroed for the next call on this workspace.
    __threadfence();
    if (atomicAdd(p.counters + 1, 1u) == gridDim.x - 1u) {
      __threadfence();
      p.counters[0] = 0u;
      p.counters[1] = 0u;
    }
  }
}

template <int C>
cudaError_t launch_resident_c(const KParams& p, int grid, cudaStream_t stream) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(RNT);
  cfg.dynamicSmemBytes = (size_t)p.res_smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, resident_kernel<C>, p);
}

template <int C>
cudaError_t configure_resident_c(int smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(resident_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(resident_kernel<C>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

}  // namespace

}  // namespace chb
