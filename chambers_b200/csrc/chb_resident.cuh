// Image-resident policy kernel for sm_100a: ONE launch per call, one CTA per SM, one image per CTA
// at a time, the whole image in shared memory for as long as its op chain needs it.
//
//   resident_kernel<C>   claims images from an atomic counter.  Per image: thread 0 issues the bulk
//                        loads (cp.async.bulk + mbarrier, 12 KB chunks) of the whole image into shared
//                        memory; while they fly, warp 0 decodes the image's schedule (Philox4x32-10 or
//                        replay) and the CTA folds the chain into the lazy per-image state (advance(),
//                        chb_kernels.cuh -- the same chain walk the tile engine uses, so both engines
//                        share every line that decides WHAT is computed).  Then the image's passes run
//                        back to back on the resident source:
//                          COUNT          histogram of the virtual image (8 skewed copies in shared
//                                         memory, red.shared), Equalize / AutoContrast table, chain walk
//                                         resumed -- no second trip to HBM, no election, no ticket;
//                          WRITE_SCRATCH  Color materialises in place; the rare neighbourhood-of-
//                                         neighbourhood chains go through a per-CTA scratch image
//                                         (L2-resident) and are loaded back;
//                          WRITE_OUT      flat chains are transformed in place and leave as 48 KB bulk
//                                         stores (TMA, shared -> global); gathers and Sharpness store
//                                         from registers.
//   An image costs exactly one HBM read and one HBM write whatever its chain.
//
// Eligibility (decided on the host, chb_api.cu): the image plus ~30 KB of working set fit one SM's
// shared memory (224 x 224 x 3 = 147 KB does; 512 x 512 x 3 does not and runs on the tile engine),
// rows are whole 16-byte units, every geometric op is nearest / constant-fill (what the policies
// use, augmentation_schemes.py:7-9).  Everything else stays on the tile engine (chb_kernels.cuh).
#pragma once
// the policy table is staged in shared memory: plain loads (see chb_kernels.cuh)
#define CHB_LDP(ptr) (*(ptr))
#include "chb_kernels.cuh"

namespace chb {
namespace {

constexpr int RNT = 1024;              // threads of a resident CTA (32 warps, 64 registers each)
constexpr int RES_CHUNK = 12288;       // bytes per load chunk: 256 units of 48 bytes / 768 of 16
constexpr int RES_MAXCHUNK = 16;       // -> images of up to 192 KB
constexpr int RES_MIN_AUX = 16 * 1024;

struct alignas(128) ResCtl {
  unsigned long long full[RES_MAXCHUNK];
  int32_t next_img, n_claimed, n_prefetched, _p2;
  // work splits that depend only on the shape (computed once per launch by thread 0)
  int32_t sharp_rows;                  // res_sharp: rows per sub-strip
  int32_t gs_band[2], gs_rows[2];      // res_gather_sharp, WRITE / COUNT: output rows per band, rows per sub-strip
  int32_t _p4[3];
  uint32_t fillc[2];                   // colour bytes of the last / last-but-one spatial entry ("a miss is just another address")
  uint32_t _p3[2];
  uint32_t rnd[32][4], rndc[32][4];
  uint32_t hmap[MAXC * 256];
  uint8_t etab[MAXC * 256];
  KParams kp;                          // the launch parameters with ops / optab pointing at the shared-memory copy
  ImgState st;
};

template <int C>
struct RC {
  const KParams* p;
  ResCtl* ctl;
  const TileState* t;
  uint32_t img;        // shared address of the resident source image
  uint32_t aux;        // shared address of the auxiliary region (histogram copies | band buffer)
  int aux_bytes;
  uint32_t full0;      // shared address of the chunk barriers
  uint32_t par;        // parity the chunk barriers complete with for the load in flight
  uint8_t* dst;        // global destination of a WRITE pass
  int H, W, row, img_bytes, tid, lane;
  uint32_t l1a, l2a;   // shared addresses of the two LUTs
  uint32_t hcopy;      // shared address of this lane's histogram copy (word 0)
  int ncopy;           // histogram copies: 32, 16, 8 or 4
  uint32_t hshift;     // log2(ncopy * 4): byte shift of a counter-pair row
};

// Histogram of a COUNT pass: `ncopy` copies of packed 16-bit counters in the aux region, copy =
// lane & (ncopy - 1).  Counter of (channel ch, value v) in copy k: half (v & 1) of the 32-bit word
//     ((ch * 128 + (v >> 1)) * ncopy + k)
// so with ncopy == 32 every lane of a warp owns one shared-memory bank: a red.shared of 32 random
// values is ONE conflict-free wavefront (the 8 skewed u32 copies of the first version averaged 3.5:
// profiles/r02_v1_ncu_op_Equalize.txt, 52 % of the wavefronts were bank conflicts and mio_throttle
// was the top stall).  A copy counts at most (pixels / ncopy) * (32 / 32) ... <= 4 x 6656 values for
// ncopy >= 4 (images of <= 192 KB), so 16 bits never overflow.  ncopy = 32 needs C x 16 KB.
template <int C>
__host__ __device__ constexpr uint32_t hist_bytes(int ncopy) { return (uint32_t)C * 512u * (uint32_t)ncopy; }

template <int C>
__device__ __forceinline__ void hist_zero(const RC<C>& c) {
  const uint32_t n = hist_bytes<C>(c.ncopy);
  for (uint32_t i = (uint32_t)c.tid * 16u; i < n; i += RNT * 16u) sts_v4(c.aux + i, make_uint4(0u, 0u, 0u, 0u));
  __syncthreads();
}
template <int C>
__device__ __forceinline__ void hist_add(const RC<C>& c, int ch, uint32_t v) {  // v: a zero-extended byte
  reds_add(c.hcopy + (((uint32_t)ch * 128u + (v >> 1)) << c.hshift), 1u + (v & 1u) * 0xFFFFu);
}
// Sums the copies into st.hist (which advance() zeroed when it asked for the COUNT pass).
template <int C>
__device__ __forceinline__ void hist_reduce(const RC<C>& c) {
  __syncthreads();
  for (int i = c.tid; i < C * 256; i += RNT) {
    const uint32_t rowa = c.aux + ((uint32_t)(i >> 1) << c.hshift);
    const uint32_t sh = (uint32_t)(i & 1) * 16u;
    uint32_t sum = 0;
    for (int k = 0; k < c.ncopy; ++k) {
      const uint32_t kk = (uint32_t)(k + c.lane) & (uint32_t)(c.ncopy - 1);  // staggered: a row of copies is one bank per lane
      sum += (lds_u32(rowa + kk * 4u) >> sh) & 0xFFFFu;
    }
    (&c.ctl->st.hist[0][0])[i] += sum;
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

// ================================================================================ flat executor
// No warp pending, K in {none, Color}: units of 48 bytes (16 pixels; 16 bytes for C != 3) are
// transformed in place as their load chunks arrive, one unit per thread per step of RNT units; a
// step leaves as one bulk store.  CutOut rectangles are painted over the step before it is stored.
//   store   0: in place only (materialisation)  1: bulk-store every step to c.dst
__device__ __forceinline__ void bulk_wait_read1() {  // all but this thread's most recent store have finished reading shared memory
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

template <int C, bool COUNT>
__device__ __forceinline__ void res_flat(const RC<C>& c, int store) {
  constexpr int UW = (C == 3) ? 12 : 4;
  constexpr int UB = UW * 4;
  const TileState& t = *c.t;
  const int n_units = c.img_bytes / UB;
  const int kmode = t.kmode;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const bool aff1 = (t.l1_aff & 0x10000) != 0;
  const uint32_t am1 = (uint32_t)((t.l1_aff >> 8) & 0xFF) * 0x01010101u, ac1 = (uint32_t)(t.l1_aff & 0xFF) * 0x01010101u;
  const float f = t.kfactor;
  const int n_paint = COUNT ? 0 : t.n_sp;  // this class holds masks only
  const bool touch = COUNT || use1 || kmode != K_NONE;  // else the staged bytes already are the result
  // WRITE_OUT: the chunks of a step are free once the step's store has read them -- the next image's
  // chunks go into them right away (thread 0 issues both, in order), so a flat image's store and the
  // next image's load overlap the remaining steps instead of starting at the image boundary.
  const uint8_t* next_src = nullptr;
  int n_released = 0;  // thread 0: chunks of the next image already issued
  if (!COUNT && store && c.tid == 0 && c.ctl->n_claimed < c.p->B)
    next_src = c.p->in + (size_t)c.ctl->n_claimed * (size_t)c.img_bytes;
  const int n_chunks = (c.img_bytes + RES_CHUNK - 1) / RES_CHUNK;
  auto release_upto = [&](int bytes_done) {  // thread 0: every chunk that lies below bytes_done is free
    while (n_released < n_chunks && min((n_released + 1) * RES_CHUNK, c.img_bytes) <= bytes_done) {
      const uint32_t bytes = (uint32_t)min(RES_CHUNK, c.img_bytes - n_released * RES_CHUNK);
      mbar_arrive_expect_tx(c.full0 + 8u * (uint32_t)n_released, bytes);
      bulk_load(c.img + (uint32_t)n_released * RES_CHUNK, next_src + (size_t)n_released * RES_CHUNK, bytes,
                c.full0 + 8u * (uint32_t)n_released);
      ++n_released;
    }
  };
  if (COUNT) hist_zero(c);
  for (int base = 0; base < n_units; base += RNT) {
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
          for (int b = 0; b < 4; ++b) hist_add(c, (4 * j + b) % C, byte_of(w[j], b));
      } else {
#pragma unroll
        for (int q = 0; q < UW / 4; ++q) sts_v4(ua + q * 16, make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]));
      }
    }
    if (!COUNT) {
      const int u1 = min(n_units, base + RNT);
      // CutOut rectangles, in list order, over the pixels [P0, P1) of this step
      const int P0 = base * (UB / C), P1 = u1 * (UB / C);
      for (int k = 0; k < n_paint; ++k) {
        __syncthreads();
        const Spatial& e = t.sp[k];
        const int rw = (e.x1 - e.x0) * C;
        if (rw <= 0) continue;
        const int ylo = max(e.y0, P0 / c.W), yhi = min(e.y1, (P1 - 1) / c.W + 1);
        const int n = (yhi - ylo) * rw;
        for (int i = c.tid; i < n; i += RNT) {
          const int ry = i / rw, rb = i - ry * rw;
          const int pix_b = ((ylo + ry) * c.W + e.x0) * C + rb;  // byte index in the image
          if (pix_b >= P0 * C && pix_b < P1 * C)
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(c.img + (uint32_t)pix_b), "r"(e.color[rb % C]) : "memory");
        }
      }
      if (store) {
        fence_proxy_async();  // this thread's shared-memory writes -> visible to the TMA
        __syncthreads();
        if (c.tid == 0) {
          bulk_store(c.dst + (size_t)base * UB, c.img + (uint32_t)base * UB, (uint32_t)(u1 - base) * UB);
          bulk_commit();
          if (next_src) {
            bulk_wait_read1();        // the previous step's store has read its bytes
            release_upto(base * UB);
          }
        }
      }
    }
  }
  if (!COUNT && store && c.tid == 0 && next_src) {
    bulk_wait_read0();
    release_upto(c.img_bytes);
    c.ctl->n_prefetched = n_released;
  }
  if (COUNT) hist_reduce(c);
  else __syncthreads();
}

// ============================================================================== gather executors
// Division-free walk over the units of an image: thread t starts at unit t and advances by RNT.
struct UnitWalk {
  int ux, y, dux, dy, upr;
  __device__ __forceinline__ UnitWalk(int tid, int upr_) {
    upr = upr_;
    y = tid / upr; ux = tid - y * upr;
    dy = RNT / upr; dux = RNT - dy * upr;
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

// One or two spatial entries (all a RandAugment(N=2) chain can produce), K in {none, Color}: the
// per-pixel arithmetic of the tile engine's gather_warp (exact float32 coordinates, src_index), the
// source being the resident image.  A thread takes one unit (16 pixels for C = 3) at a time, four
// pixels in flight, and stores the unit's 48 bytes from registers.
template <int C, bool COUNT, bool TWO>
__device__ __forceinline__ void res_gather_fast(const RC<C>& c) {
  constexpr int UW = (C == 3) ? 12 : 4;
  constexpr int PPU = UW * 4 / C;  // pixels per unit: 16, 8, 16, 4
  const TileState& t = *c.t;
  const int H = c.H, W = c.W;
  const int n_sp = t.n_sp;
  const Spatial& ea = t.sp[n_sp - 1];
  const Spatial& eb = t.sp[TWO ? n_sp - 2 : 0];
  const bool a_geom = !TWO || ea.type == SP_GEOM;
  const bool b_geom = TWO && eb.type == SP_GEOM;
  const float t0 = ea.t[0], t1 = ea.t[1], t2 = ea.t[2], t3 = ea.t[3], t4 = ea.t[4], t5 = ea.t[5];
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
  for (UnitWalk q(c.tid, W / PPU); q.y < H; q.next()) {
    const int y = q.y, x0 = q.ux * PPU;
    const float fy = small_uint_to_float((uint32_t)y);
    const float t1y = __fmul_rn(t1, fy), t4y = __fmul_rn(t4, fy);
    uint32_t o[UW];
#pragma unroll
    for (int g = 0; g < PPU / 4; ++g) {
      uint32_t adr[4];
      uint32_t hit = 0;  // bit i: pixel i shows entry a's colour, bit 4 + i: entry b's
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int ix = x0 + 4 * g + i, iy = y;
        bool hit_a, hit_b = false;
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
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int ch = 0; ch < C; ++ch) {
            const int bi = i * C + ch;
            const int wi = g * C + (bi >> 2);
            o[wi] = ((bi & 3) == 0) ? v[i][ch] : put_byte(o[wi], v[i][ch], bi & 3);  // (every v is a zero-extended byte)
          }
      }
    }
    if (!COUNT) {
      uint4* gp = reinterpret_cast<uint4*>(c.dst + ((size_t)y * W + x0) * C);
#pragma unroll
      for (int qv = 0; qv < UW / 4; ++qv) __stcg(gp + qv, make_uint4(o[4 * qv], o[4 * qv + 1], o[4 * qv + 2], o[4 * qv + 3]));
    }
  }
  if (COUNT) {
    if (n_fill_a) atomicAdd(&c.ctl->st.color_cnt[n_sp - 1], n_fill_a);
    if (TWO && n_fill_b) atomicAdd(&c.ctl->st.color_cnt[n_sp - 2], n_fill_b);
  }
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
  const int n_pix = H * W;
  for (int i = c.tid; i < n_pix; i += RNT) {
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
        atomicAdd(&c.ctl->st.color_cnt[k], 1u);
      } else {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v[ch] = (uint32_t)t.sp[k].color[ch];
      }
    }
    if (!COUNT) {
      uint8_t* d = c.dst + (size_t)i * C;
#pragma unroll
      for (int ch = 0; ch < C; ++ch) d[ch] = (uint8_t)v[ch];
    }
  }
}

template <int C, bool COUNT>
__device__ __forceinline__ void res_gather(const RC<C>& c) {
  const TileState& t = *c.t;
  wait_image(c);
  if (COUNT) hist_zero(c);
  if (t.n_sp == 1 && t.sp[0].type == SP_GEOM) res_gather_fast<C, COUNT, false>(c);
  else if (t.n_sp == 2) res_gather_fast<C, COUNT, true>(c);
  else res_gather_list<C, COUNT>(c);
  if (COUNT) hist_reduce(c);
}

// ========================================================================== sharpness executors
// l1 applied to the resident source in place (a Sharpness tap reads every byte nine times); the view
// then continues with l1 = identity.
template <int C>
__device__ __forceinline__ void res_bake_l1(const RC<C>& c) {
  TileState& t = c.ctl->st.t;
  if (t.l1_id) return;
  wait_image(c);
  const int total = c.img_bytes >> 4;
  for (int i = c.tid; i < total; i += RNT) {
    const uint4 v = map_vec_phase<C>(lds_v4(c.img + ((uint32_t)i << 4)), c.l1a, (C == 3) ? (i % 3) : 0);
    sts_v4(c.img + ((uint32_t)i << 4), v);
  }
  __syncthreads();
  for (int i = c.tid; i < MAXC * 256; i += RNT) t.l1[i >> 8][i & 255] = (uint8_t)(i & 255);
  if (c.tid == 0) { t.l1_id = 1; t.l1_aff = 0; }
  fence_proxy_async();  // the image buffer is refilled by the TMA later
  __syncthreads();
}

// Rows per sub-strip so that (word columns x sub-strips) fills the CTA: minimise rounds * (rows + 2).
// Runs on one thread once per launch (the results live in ResCtl).
__device__ __forceinline__ int res_sharp_split(int columns, int inner) {
  if (inner <= 0) return 1;
  int best_s = 1, best_cost = 0x7FFFFFFF;
  for (int S = 1; S <= 64 && S <= inner; ++S) {
    const int rounds = (columns * S + RNT - 1) / RNT;
    const int cost = rounds * ((inner + S - 1) / S + 2);
    if (cost < best_cost) { best_cost = cost; best_s = S; }
  }
  return (inner + best_s - 1) / best_s;
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
  float T[4], A[4], xm_mid[4];
  float p[NB], xm[4];
  auto load_row = [&](uint32_t ra) {
    const uint32_t w1 = lds_u32(ra);
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
        const float orig = __fadd_rn(xm_mid[b], -8388608.0f);
        float deg = __fadd_rn(__fadd_rz(acc, 8388608.0f), -8388608.0f);  // float(trunc(acc))
        deg = ((bmask >> b) & 1u) ? orig : deg;                           // border pixels keep the original
        const uint32_t res = sharp_blend(deg, orig, f);
        o = (b == 0) ? res : put_byte(o, res, b);
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
        xm_mid[b] = xm[b];
        // and it is the row above output row r + 1
        T[b] = __fadd_rn(__fadd_rn(p[b], p[b + C]), p[b + 2 * C]);
      }
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
      __stcg(reinterpret_cast<uint32_t*>(c.dst + (size_t)y * row) + xw, o);
    }
  };
  // first and last image row: every pixel is border -> blend(orig, orig) == orig
  for (int xw = c.tid; xw < wpr; xw += RNT) emit(0, xw, (C == 3) ? (xw % 3) : 0, lds_u32(c.img + ((uint32_t)xw << 2)));
  if (H > 1)
    for (int xw = c.tid; xw < wpr; xw += RNT)
      emit(H - 1, xw, (C == 3) ? (xw % 3) : 0, lds_u32(c.img + (uint32_t)((H - 1) * row + (xw << 2))));
  const int inner = H - 2;
  if (inner > 0) {
    const int R = c.ctl->sharp_rows;
    const int n_strips = (inner + R - 1) / R;
    const int n_items = wpr * n_strips;
    for (int item = c.tid; item < n_items; item += RNT) {
      const int strip = item / wpr, xw = item - strip * wpr;
      const int y_begin = 1 + strip * R, y_end = min(H - 1, y_begin + R);
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

// K == Sharpness with a spatial list pending (e.g. Rotate -> Sharpness): the virtual pre-image
// (spatial list -> l1) of a band of rows plus one halo row each side is gathered from the resident
// source into the aux region, then sharpened from there with the column walk.  A halo row is laid
// out so that the row's first pixel starts on a word: 4 - C pad bytes, the left halo pixel, the
// row, the right halo pixel.
template <int C, bool COUNT>
__device__ __forceinline__ void res_gather_sharp(RC<C> c) {
  wait_image(c);
  const TileState& t = *c.t;
  const int H = c.H, W = c.W, row = c.row;
  const int n_sp = t.n_sp;
  const bool use1 = !t.l1_id, use2 = !t.l2_id;
  const float f = t.kfactor;
  const int vw = W + 2;
  const int vpitch = (4 + (W + 1) * C + 3) & ~3;
  uint32_t vbuf = c.aux;
  if (COUNT) {  // four histogram copies in front of the band buffer
    c.ncopy = 4; c.hshift = 4u; c.hcopy = c.aux + (uint32_t)(c.lane & 3) * 4u;
    hist_zero(c);
    vbuf += hist_bytes<C>(4);
  }
  const int band = c.ctl->gs_band[COUNT ? 1 : 0];  // output rows per band (aux capacity)
  const int band_rows = c.ctl->gs_rows[COUNT ? 1 : 0];
  const int wpr = row >> 2;
  const Spatial& e0 = t.sp[0];
  const bool one = (n_sp == 1);
  const bool e_geom = e0.type == SP_GEOM;
  const uint32_t fill_addr = smem_addr(&c.ctl->fillc[0]);
  const uint32_t inv_vw = (uint32_t)((0x100000000ull + (unsigned)vw - 1ull) / (unsigned)vw);
  auto emit = [&](int y, int xw, uint32_t o) {
    const int ph = (C == 3) ? (xw % 3) : 0;
    if (COUNT) {
#pragma unroll
      for (int b = 0; b < 4; ++b) hist_add(c, (ph + b) % C, byte_of(o, b));
    } else {
      if (use2)
        o = map_word(o, c.l2a + (uint32_t)((ph + 0) % C) * 256u, c.l2a + (uint32_t)((ph + 1) % C) * 256u,
                     c.l2a + (uint32_t)((ph + 2) % C) * 256u, c.l2a + (uint32_t)((ph + 3) % C) * 256u);
      __stcg(reinterpret_cast<uint32_t*>(c.dst + (size_t)y * row) + xw, o);
    }
  };
  for (int ya = 0; ya < H; ya += band) {
    const int yb = min(H, ya + band);
    // phase 1: virtual pre-image on [-1, W] x [ya-1, yb]
    const int nv = (yb - ya + 2) * vw;
    for (int i = c.tid; i < nv; i += RNT) {
      const int vy = (int)__umulhi((uint32_t)i, inv_vw), vx = i - vy * vw;  // i / vw, exact for i < 2^16 ... 2^22
      const int y = ya - 1 + vy, x = vx - 1;
      uint32_t v[C];
#pragma unroll
      for (int ch = 0; ch < C; ++ch) v[ch] = 0;
      if ((unsigned)x < (unsigned)W && (unsigned)y < (unsigned)H) {
        if (one) {
          int sx = x, sy = y;
          bool hit;
          if (e_geom) {
            const float fx = small_uint_to_float((uint32_t)x), fy = small_uint_to_float((uint32_t)y);
            const bool inx = src_index(__fadd_rn(__fadd_rn(__fmul_rn(e0.t[0], fx), __fmul_rn(e0.t[1], fy)), e0.t[2]), W, sx);
            const bool iny = src_index(__fadd_rn(__fadd_rn(__fmul_rn(e0.t[3], fx), __fmul_rn(e0.t[4], fy)), e0.t[5]), H, sy);
            hit = !(inx && iny);
          } else {
            hit = (y >= e0.y0) && (y < e0.y1) && (x >= e0.x0) && (x < e0.x1);
          }
          const uint32_t a = hit ? fill_addr : c.img + (uint32_t)(sy * row + sx * C);
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
            const uint32_t a = c.img + (uint32_t)(sy * row + sx * C);
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v[ch] = lds_u8(a + ch);
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
    __syncthreads();
    // phase 2: sharpen W x (yb - ya) pixels out of the halo band
    for (int y = ya; y < yb; ++y)
      if (y == 0 || y == H - 1)
        for (int xw = c.tid; xw < wpr; xw += RNT) emit(y, xw, lds_u32(vbuf + (uint32_t)((y - ya + 1) * vpitch + 4 + (xw << 2))));
    const int in0 = max(ya, 1), in1 = min(yb, H - 1);
    const int inner = in1 - in0;
    if (inner > 0) {
      const int R = band_rows;
      const int n_strips = (inner + R - 1) / R;
      const int n_items = wpr * n_strips;
      for (int item = c.tid; item < n_items; item += RNT) {
        const int strip = item / wpr, xw = item - strip * wpr;
        const int y_begin = in0 + strip * R, y_end = min(in1, y_begin + R);
        uint32_t bmask = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xb = (xw << 2) + b;
          if (xb < C || xb >= row - C) bmask |= 1u << b;
        }
        const uint32_t col = vbuf + (uint32_t)((y_begin - 1 - (ya - 1)) * vpitch + 4 + (xw << 2));
        sharp_stream<C>(col, vpitch, y_end - y_begin, true, true, bmask, f,
                        [&](int r, uint32_t o) { emit(y_begin + r, xw, o); });
      }
    }
    __syncthreads();  // the next band overwrites the buffer
  }
  if (COUNT) hist_reduce(c);
}

// ================================================================================= the kernel
template <int C, bool COUNT>
__device__ __forceinline__ void res_run_pass(RC<C>& c, int store) {
  const TileState& t = *c.t;
  bool any_geom = false;
  for (int k = 0; k < t.n_sp; ++k) any_geom = any_geom || (t.sp[k].type == SP_GEOM);
  if (t.kmode == K_SHARP) {
    if (t.n_sp > 0) res_gather_sharp<C, COUNT>(c);
    else res_sharp<C, COUNT>(c);
  } else if (t.n_sp == 0 || (!any_geom && !COUNT)) {
    res_flat<C, COUNT>(c, store);
  } else {
    res_gather<C, COUNT>(c);
  }
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
  c.p = &p; c.ctl = ctl; c.t = &ctl->st.t;
  c.img = smem_addr(smem_raw) + img_off;
  c.aux = smem_addr(smem_raw) + aux_off;
  c.aux_bytes = p.res_smem_bytes - (int)aux_off;
  c.full0 = smem_addr(&ctl->full[0]);
  c.H = H; c.W = W; c.row = W * C; c.img_bytes = img_bytes; c.tid = tid; c.lane = tid & 31;
  c.l1a = smem_addr(&ctl->st.t.l1[0][0]); c.l2a = smem_addr(&ctl->st.t.l2[0][0]);
  int ncopy = 32;
  while (ncopy > 4 && hist_bytes<C>(ncopy) > (uint32_t)c.aux_bytes) ncopy >>= 1;
  c.par = 1;  // toggled to 0 by the first load
  c.dst = nullptr;
  if (tid == 0) {
    for (int k = 0; k < RES_MAXCHUNK; ++k) mbar_init(c.full0 + 8 * k, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    // shape-only work splits
    const int wpr = (W * C) >> 2;
    ctl->sharp_rows = res_sharp_split(wpr, H - 2);
    const int vpitch = (4 + (W + 1) * C + 3) & ~3;
    for (int k = 0; k < 2; ++k) {
      const int vbytes = c.aux_bytes - (k ? (int)hist_bytes<C>(4) : 0);
      int band = max(1, vbytes / vpitch - 2);
      // whole bands of equal height: the last band is not a sliver
      const int n_bands = (H + band - 1) / band;
      band = (H + n_bands - 1) / n_bands;
      ctl->gs_band[k] = band;
      ctl->gs_rows[k] = res_sharp_split(wpr, min(band, max(H - 2, 1)));
    }
    ctl->n_prefetched = 0;
  }
  // Programmatic dependent launch: nothing the previous kernel of the stream wrote (the images, a
  // replayed schedule, the policy table, the work counter it left zeroed) is touched before it has completed.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid == 0) ctl->next_img = (int)atomicAdd(p.counters, 1u);
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
  __syncthreads();
  int img = ctl->next_img;
  uint8_t* scratch = p.scratch + (size_t)blockIdx.x * p.scratch_stride;
  // chunk k is issued by thread k (each arrives on its own barrier); chunks below `first` are already in flight
  auto issue_load = [&](const uint8_t* src, int first) {
    if (tid >= first && tid < n_chunks) {
      const uint32_t bytes = (uint32_t)min(RES_CHUNK, img_bytes - tid * RES_CHUNK);
      mbar_arrive_expect_tx(c.full0 + 8 * tid, bytes);
      bulk_load(c.img + (uint32_t)tid * RES_CHUNK, src + (size_t)tid * RES_CHUNK, bytes, c.full0 + 8 * tid);
    }
  };
  while (img < p.B) {
    c.par ^= 1u;
    issue_load(p.in + (size_t)img * img_bytes, ctl->n_prefetched);  // (a flat predecessor already issued some or all)
    __syncthreads();
    if (tid == 0) {
      ctl->n_prefetched = 0;
      ctl->n_claimed = (int)atomicAdd(p.counters, 1u);  // the next image: the round trip hides behind this one
    }
    // schedule decode + chain walk up to the first pass (the loads are in flight)
    for (int i = tid; i < STATE_VECS; i += RNT) reinterpret_cast<uint4*>(&ctl->st)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid < 32) decode_image(pl, &ctl->st, ctl->rnd, ctl->rndc, img, H, W, tid);
    reset_view(&ctl->st, tid, RNT);
    __syncthreads();
    advance(&ctl->st, &ctl->st, pl, C, H, W, ctl->hmap, ctl->etab, tid, RNT, [] { __syncthreads(); });
    uint8_t* out_img = p.out + (size_t)img * img_bytes;
    for (;;) {
      const TileState& t = ctl->st.t;
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
      c.hshift = 31u - (uint32_t)__clz(ncopy * 4);
      c.hcopy = c.aux + (uint32_t)(c.lane & (ncopy - 1)) * 4u;
      __syncthreads();
      if (pass_kind == PASS_COUNT) {
        res_run_pass<C, true>(c, 0);
        if (tid == 0) ctl->st.hist_valid = 1;
        __syncthreads();
        advance(&ctl->st, &ctl->st, pl, C, H, W, ctl->hmap, ctl->etab, tid, RNT, [] { __syncthreads(); });
        continue;
      }
      // WRITE_OUT, or WRITE_SCRATCH: point-wise views (and CutOut rectangles, painted over them)
      // materialise in the resident image itself, the rare neighbourhood-of-neighbourhood chains go
      // through this CTA's scratch image and come back through the TMA
      const bool last = pass_kind == PASS_WRITE_OUT;
      bool any_geom = false;
      for (int k = 0; k < t.n_sp; ++k) any_geom = any_geom || (t.sp[k].type == SP_GEOM);
      const bool in_place = !last && (t.kmode == K_NONE || t.kmode == K_COLOR) && !any_geom;
      c.dst = last ? out_img : scratch;
      res_run_pass<C, false>(c, last ? 1 : 0);
      if (last) break;
      if (!in_place) {
        __threadfence();
        fence_proxy_async_all();
        __syncthreads();
        c.par ^= 1u;
        issue_load(scratch, 0);
      }
      reset_view(&ctl->st, tid, RNT);
      __syncthreads();
      advance(&ctl->st, &ctl->st, pl, C, H, W, ctl->hmap, ctl->etab, tid, RNT, [] { __syncthreads(); });
    }
    // every thread is done with the resident source; bulk stores out of it have read their bytes;
    // in-place writes (generic proxy) are ordered before the TMA refills the buffer
    fence_proxy_async();
    if (tid == 0) bulk_wait_read0();
    __syncthreads();
    img = ctl->n_claimed;
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
