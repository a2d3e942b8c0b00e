// Host side of libchambers_aug.so: the C ABI of include/chambers_aug.h.
//
// Everything the reference computes in Python at layer construction (magnitude -> kwargs,
// augmentation_schemes.py:42-128) or per call on the host (float32 conversion of the signed
// parameters, the projective coefficients, image_augmentations.py:135-146, :334-341, :420-427) is
// resolved here into a DevOp table; the pixels are only ever touched by chb_kernels.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "chb_internal.h"

using chb::DevOp;

struct PolicyEntry {
  std::vector<DevOp> host;
  DevOp* dev = nullptr;
  uint8_t* optab = nullptr;  // [n][256] value maps of the point-wise ops (built on device)
};

// Per-stream working memory: calls on one stream are ordered, calls on different streams are not.
struct Workspace {
  cudaStream_t stream = nullptr;
  chb::ImgState* states = nullptr;
  int* lists = nullptr;
  unsigned int* counters = nullptr;
  size_t cap_images = 0;
  bool clean = false;                // counters + continuation list are zero (the pass kernel leaves them so)
  uint8_t* scratch = nullptr;
  size_t scratch_bytes = 0;
  uint8_t* inplace = nullptr;
  size_t inplace_bytes = 0;
  uint8_t* norm_tmp = nullptr;       // uint8 policy output in front of an unfused normalisation
  size_t norm_tmp_bytes = 0;
};

static const int kPipeMax = 8;      // e2e pipeline: at most this many streams / staging buffers
static int g_pipe = 4;              // streams in use (CHB_E2E_STREAMS)
static long long g_nf_first_images = 6;  // a batch is also "small" below this many images per CTA (CHB_NF_FIRST_IMAGES)
static int g_self_clean = 1;         // CHB_SELF_CLEAN=0: zero the counters with a memset in front of every call instead
static int g_chunk_kb = 3840;       // target chunk size of the host path in KiB (CHB_E2E_CHUNK_KB)
static int g_lpt = 1;               // resident engine: cost-sorted claim order for small batches (CHB_LPT=0 disables)
static int g_split_pct = 80;        // resident engine: split an item that would outlast this % of the average SM's load (CHB_SPLIT_PCT)
static int g_split = 1;             // resident engine: small batches cut expensive last passes into row ranges (CHB_SPLIT=0 disables)

struct chb_ctx {
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  std::string err;
  int64_t launches = 0;
  int force_generic = 0;  // CHB_FORCE_GENERIC=1: every tile through the scalar executor (debugging)
  int engine = 0;         // CHB_ENGINE_*: 0 = resident where eligible, else tiles
  int res_smem = 0;       // dynamic shared memory of a resident CTA (0: engine unavailable)
  int last_engine = 0;    // engine the last device call ran on (CHB_ENGINE_TILES / CHB_ENGINE_RESIDENT)
  unsigned long long* timeline = nullptr;  // debug builds only (chb_debug_timeline)
  std::vector<Workspace*> workspaces;
  std::vector<PolicyEntry*> cache;
  // e2e pipeline
  cudaStream_t streams[kPipeMax] = {};
  uint8_t* st_in[kPipeMax] = {};
  uint8_t* st_out[kPipeMax] = {};
  int32_t* st_replay[kPipeMax] = {};
  int32_t* st_record[kPipeMax] = {};
  size_t st_img_bytes = 0, st_sched_bytes = 0;
};

static std::string g_init_error;

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so that the library
// does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;

static const int kBoxRows = 96;    // 64 x 64 tile rotated by 45 degrees + margins
static const int kBoxBytes = 304;  // 76 words: consecutive staged rows start 12 banks apart

// Tensor map over `images` images of H rows of W*C bytes (uint32 elements), `stride` bytes apart,
// with a box of at most kBoxRows x kBoxBytes.  Returns the box actually encoded.
static bool encode_image_map(chb::TMap* out, const void* base, int H, int rowbytes, size_t stride, size_t images,
                             int* box_rows, int* box_bytes) {
  static_assert(sizeof(CUtensorMap) == sizeof(chb::TMap), "CUtensorMap size");
  if (!g_encode_tiled || !base || images == 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)(rowbytes / 4), (cuuint64_t)H, (cuuint64_t)images};
  const cuuint64_t strides[2] = {(cuuint64_t)rowbytes, (cuuint64_t)stride};
  const int bw = kBoxBytes < rowbytes ? kBoxBytes : rowbytes;
  const int bh = kBoxRows < H ? kBoxRows : H;
  const cuuint32_t box[3] = {(cuuint32_t)(bw / 4), (cuuint32_t)bh, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult r = g_encode_tiled(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_UINT32, 3,
                                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  *box_rows = bh;
  *box_bytes = bw;
  return r == CUDA_SUCCESS;
}

// Every entry point that touches the device makes ctx->device current for its own duration and puts
// the caller's current device back on every exit path (a caller with several GPUs in one process
// must not find its current device changed by a chb_* call).
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    if (prev == device) return cudaSuccess;
    e = cudaSetDevice(device);
    switched = (e == cudaSuccess);
    return e;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

static int fail(chb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_init_error = msg;
  return code;
}
static int cuda_fail(chb_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, CHB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CHB_CUDA(ctx, call)                                   \
  do {                                                        \
    cudaError_t e_ = (call);                                  \
    if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);  \
  } while (0)

// ---------------------------------------------------------------------------- table builders
static chb_op make_op(int kind) {
  chb_op op;
  memset(&op, 0, sizeof(op));
  op.kind = kind;
  op.probability = -1.0;
  if (kind == CHB_OP_SHEAR_X || kind == CHB_OP_SHEAR_Y || kind == CHB_OP_TRANSLATE_X ||
      kind == CHB_OP_TRANSLATE_Y || kind == CHB_OP_ROTATE) {
    op.interpolation = CHB_INTERP_NEAREST;  // augmentation_schemes.py:7
    op.fill_mode = CHB_FILL_CONSTANT;       // :8
    op.fill_value = 128.0f;                 // :9
  }
  return op;
}

// _get_transform(name, magnitude), augmentation_schemes.py:105-128 with the maps of :42-102.
// Same IEEE double arithmetic, same left-to-right order, int() truncation.
static chb_op op_from_magnitude(int kind, double m) {
  const double kMax = 10.0;  // _MAX_MAGNITUDE :10
  chb_op op = make_op(kind);
  switch (kind) {
    case CHB_OP_BRIGHTNESS: case CHB_OP_CONTRAST: case CHB_OP_COLOR: case CHB_OP_SHARPNESS:
      op.value = m / kMax * 1.8 + 0.1;  // :43
      break;
    case CHB_OP_SHEAR_X: case CHB_OP_SHEAR_Y:
      op.value = m / kMax * 0.3;  // :49
      break;
    case CHB_OP_TRANSLATE_X: case CHB_OP_TRANSLATE_Y:
      op.value = m / kMax * 100;  // :60
      break;
    case CHB_OP_POSTERIZE:
      op.ivalue[0] = (int)(m / kMax * 4);  // :71
      break;
    case CHB_OP_SOLARIZE:
      op.ivalue[0] = (int)(m / kMax * 256);  // :77
      break;
    case CHB_OP_SOLARIZE_ADD:
      op.ivalue[0] = (int)(m / kMax * 110);  // :83
      op.ivalue[1] = 128;                    // SolarizeAdd default threshold, image_augmentations.py:206
      break;
    case CHB_OP_ROTATE:
      op.value = m / kMax * 30.0;  // :89
      break;
    case CHB_OP_CUTOUT:
      op.ivalue[0] = (int)(m / kMax * 80);  // :100
      op.ivalue[1] = 128;                   // :101
      break;
    default:
      break;
  }
  return op;
}

extern "C" int chb_randaugment_table(double magnitude, chb_transform* out16) {
  if (!out16) return CHB_ERR_INVALID;
  for (int k = 0; k < CHB_OP_COUNT; ++k) {  // fixed order, augmentation_schemes.py:181-198
    memset(&out16[k], 0, sizeof(chb_transform));
    out16[k].n_ops = 1;
    out16[k].ops[0] = op_from_magnitude(k, magnitude);
  }
  return CHB_OK;
}

extern "C" int chb_autoaugment_table(chb_transform* out25) {
  if (!out25) return CHB_ERR_INVALID;
  struct E { int k1; double p1; double m1; int k2; double p2; double m2; };
  // _AUTO_AUGMENT_POLICY_V0, augmentation_schemes.py:12-39 (magnitude None -> 0, unused).
  static const E v0[25] = {
      {CHB_OP_EQUALIZE, 0.8, 0, CHB_OP_SHEAR_Y, 0.8, 4},
      {CHB_OP_COLOR, 0.4, 9, CHB_OP_EQUALIZE, 0.6, 0},
      {CHB_OP_COLOR, 0.4, 1, CHB_OP_ROTATE, 0.6, 8},
      {CHB_OP_SOLARIZE, 0.8, 3, CHB_OP_EQUALIZE, 0.4, 7},
      {CHB_OP_SOLARIZE, 0.4, 2, CHB_OP_SOLARIZE, 0.6, 2},
      {CHB_OP_COLOR, 0.2, 0, CHB_OP_EQUALIZE, 0.8, 0},
      {CHB_OP_EQUALIZE, 0.4, 0, CHB_OP_SOLARIZE_ADD, 0.8, 3},
      {CHB_OP_SHEAR_X, 0.2, 9, CHB_OP_ROTATE, 0.6, 8},
      {CHB_OP_COLOR, 0.6, 1, CHB_OP_EQUALIZE, 1.0, 0},
      {CHB_OP_INVERT, 0.4, 0, CHB_OP_ROTATE, 0.6, 0},
      {CHB_OP_EQUALIZE, 1.0, 0, CHB_OP_SHEAR_Y, 0.6, 3},
      {CHB_OP_COLOR, 0.4, 7, CHB_OP_EQUALIZE, 0.6, 0},
      {CHB_OP_POSTERIZE, 0.4, 6, CHB_OP_AUTOCONTRAST, 0.4, 0},
      {CHB_OP_SOLARIZE, 0.6, 8, CHB_OP_COLOR, 0.6, 9},
      {CHB_OP_SOLARIZE, 0.2, 4, CHB_OP_ROTATE, 0.8, 9},
      {CHB_OP_ROTATE, 1.0, 7, CHB_OP_TRANSLATE_Y, 0.8, 9},
      {CHB_OP_SHEAR_X, 0.0, 0, CHB_OP_SOLARIZE, 0.8, 4},
      {CHB_OP_SHEAR_Y, 0.8, 0, CHB_OP_COLOR, 0.6, 4},
      {CHB_OP_COLOR, 1.0, 0, CHB_OP_ROTATE, 0.6, 2},
      {CHB_OP_EQUALIZE, 0.8, 0, CHB_OP_EQUALIZE, 0.0, 0},
      {CHB_OP_EQUALIZE, 1.0, 0, CHB_OP_AUTOCONTRAST, 0.6, 0},
      {CHB_OP_SHEAR_Y, 0.4, 7, CHB_OP_SOLARIZE_ADD, 0.6, 7},
      {CHB_OP_POSTERIZE, 0.8, 2, CHB_OP_SOLARIZE, 0.6, 10},
      {CHB_OP_SOLARIZE, 0.6, 8, CHB_OP_EQUALIZE, 0.6, 1},
      {CHB_OP_COLOR, 0.8, 6, CHB_OP_ROTATE, 0.4, 5},
  };
  for (int i = 0; i < 25; ++i) {
    memset(&out25[i], 0, sizeof(chb_transform));
    out25[i].n_ops = 2;
    out25[i].ops[0] = op_from_magnitude(v0[i].k1, v0[i].m1);
    out25[i].ops[0].probability = v0[i].p1;  // RandomChance, :141-142
    out25[i].ops[1] = op_from_magnitude(v0[i].k2, v0[i].m2);
    out25[i].ops[1].probability = v0[i].p2;
  }
  return CHB_OK;
}

// --------------------------------------------------------------------- chb_op -> DevOp
static int wrap_threshold(int t) {  // oracle/switches.py SOLARIZE_THRESHOLD_OVERFLOW = "wrap"
  if (t >= 0 && t <= 255) return t;
  return ((t % 256) + 256) % 256;
}

static int blend_mode_of(double factor) {  // image_augmentations.py:28-31, :43
  if (factor == 0.0) return chb::BLEND_IMAGE1;
  if (factor == 1.0) return chb::BLEND_IMAGE2;
  if (factor > 0.0 && factor < 1.0) return chb::BLEND_INTERP;
  return chb::BLEND_EXTRAP;
}

static void affine_identity(float* t) {
  const float id[8] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
  memcpy(t, id, sizeof(id));
}

static int build_devop(const chb_op& op, int H, int W, int C, long long contrast_pixels, DevOp& d,
                       std::string& err) {
  memset(&d, 0, sizeof(d));
  d.kind = op.kind;
  d.interp = op.interpolation;
  d.fill_mode = op.fill_mode;
  if (op.probability < 0.0) {
    d.thr24 = 1 << 24;
  } else {
    double t = ceil(op.probability * 16777216.0);  // oracle/philox.py prob_threshold24
    if (t < 0) t = 0;
    if (t > 16777216.0) t = 16777216.0;
    d.thr24 = (int)t;
  }
  affine_identity(d.coef[0]);
  affine_identity(d.coef[1]);
  switch (op.kind) {
    case CHB_OP_AUTOCONTRAST: case CHB_OP_EQUALIZE: case CHB_OP_INVERT:
      break;
    case CHB_OP_BRIGHTNESS: case CHB_OP_CONTRAST: case CHB_OP_COLOR: case CHB_OP_SHARPNESS: {
      d.factor = (float)op.value;
      d.blend_mode = blend_mode_of(op.value);
      if ((op.kind == CHB_OP_CONTRAST || op.kind == CHB_OP_COLOR) && C != 3) {
        err = "Color / Contrast need 3 channels (tf.image.rgb_to_grayscale)";
        return CHB_ERR_INVALID;
      }
      if (op.kind == CHB_OP_CONTRAST) {
        // image_augmentations.py:260-264: sum(hist)/256 = (#pixels passed in)/256, clip, trunc.
        float mean = (float)contrast_pixels / 256.0f;
        mean = fminf(fmaxf(mean, 0.0f), 255.0f);
        d.ip0 = (int)mean;
      }
    } break;
    case CHB_OP_POSTERIZE: {
      int shift = 8 - op.ivalue[0];  // :168
      if (shift < 0) shift = 0;
      if (shift > 7) shift = 7;  // oracle/switches.py POSTERIZE_SHIFT8 = "clamp7"
      d.ip0 = shift;
    } break;
    case CHB_OP_SOLARIZE:
      d.ip0 = wrap_threshold(op.ivalue[0]);
      break;
    case CHB_OP_SOLARIZE_ADD:
      d.ip0 = op.ivalue[0];
      d.ip1 = wrap_threshold(op.ivalue[1]);
      break;
    case CHB_OP_CUTOUT:
      if (op.ivalue[0] < 0 || (op.ivalue[0] & 1)) {
        err = "CutOut mask_size should be a non-negative multiple of 2 (tfa.image.cutout)";
        return CHB_ERR_INVALID;
      }
      d.ip0 = op.ivalue[0] / 2;
      d.ip1 = op.ivalue[1] & 0xFF;
      break;
    case CHB_OP_SHEAR_X: case CHB_OP_SHEAR_Y: case CHB_OP_TRANSLATE_X: case CHB_OP_TRANSLATE_Y:
    case CHB_OP_ROTATE: {
      if (op.interpolation != CHB_INTERP_NEAREST && op.interpolation != CHB_INTERP_BILINEAR) {
        err = "interpolation must be nearest or bilinear";
        return CHB_ERR_INVALID;
      }
      if (op.fill_mode < CHB_FILL_CONSTANT || op.fill_mode > CHB_FILL_NEAREST) {
        err = "fill_mode must be constant, reflect, wrap or nearest";
        return CHB_ERR_INVALID;
      }
      d.fill_u8 = ((int)op.fill_value) & 0xFF;
      for (int neg = 0; neg < 2; ++neg) {
        float* t = d.coef[neg];
        // _randomly_negate_value, image_augmentations.py:52-56: +-value as a float32 tensor.
        const float v = (float)(neg ? -op.value : op.value);
        if (op.kind == CHB_OP_SHEAR_X) {
          t[1] = v;  // :337
        } else if (op.kind == CHB_OP_SHEAR_Y) {
          t[3] = v;  // :380
        } else if (op.kind == CHB_OP_TRANSLATE_X) {
          t[2] = v;  // :421-427 translate([-px, 0]) -> [1,0,px, 0,1,-0, 0,0]
          t[5] = -0.0f;
        } else if (op.kind == CHB_OP_TRANSLATE_Y) {
          t[2] = -0.0f;
          t[5] = v;  // :464-470
        } else {
          // Rotate: radians in double (:135), sign flip to float32 (:139), then tfa's
          // angles_to_projective_transforms in stepwise float32 (oracle/ops.py rotate_coeffs).
          const double radians = op.value * M_PI / 180.0;
          const float ang = (float)(neg ? -radians : radians);
          const float c = (float)cos((double)ang);
          const float s = (float)sin((double)ang);
          const float w1 = (float)W - 1.0f, h1 = (float)H - 1.0f;
          volatile float cw = c * w1, sh = s * h1, sw = s * w1, ch = c * h1;
          volatile float a = cw - sh, b = sw + ch;
          volatile float xo = w1 - a, yo = h1 - b;
          t[0] = c; t[1] = -s; t[2] = xo / 2.0f;
          t[3] = s; t[4] = c;  t[5] = yo / 2.0f;
        }
      }
    } break;
    default:
      err = "unknown op kind";
      return CHB_ERR_INVALID;
  }
  return CHB_OK;
}

// ------------------------------------------------------------------------------------ context
extern "C" int chb_version(void) { return CHB_VERSION_MAJOR * 1000 + CHB_VERSION_MINOR; }

extern "C" const char* chb_last_error(const chb_ctx* ctx) {
  return ctx ? ctx->err.c_str() : g_init_error.c_str();
}

extern "C" int64_t chb_kernel_launches(const chb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int chb_init(int device, chb_ctx** out) {
  if (!out) return fail(nullptr, CHB_ERR_INVALID, "chb_init: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, CHB_ERR_NO_DEVICE,
                std::string("chb_init: no CUDA device (") + cudaGetErrorString(e) +
                    "); libchambers_aug has no CPU fallback");
  if (device < 0 || device >= n) return fail(nullptr, CHB_ERR_INVALID, "chb_init: bad device index");
  DeviceGuard guard;
  CHB_CUDA(nullptr, guard.enter(device));
  cudaDeviceProp prop;
  CHB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, CHB_ERR_NO_DEVICE,
                "chb_init: this library is built for sm_100a (B200) only; found sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor));
  chb_ctx* ctx = new chb_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  e = chb::configure_kernels();
  if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "configure_kernels"); delete ctx; return r; }
  ctx->res_smem = (int)prop.sharedMemPerBlockOptin;  // the resident kernels have no static shared memory
  e = chb::configure_resident(ctx->res_smem);
  if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "configure_resident"); delete ctx; return r; }
  if (const char* en = getenv("CHB_ENGINE")) {
    if (!strcmp(en, "tiles")) ctx->engine = CHB_ENGINE_TILES;
    else if (!strcmp(en, "resident")) ctx->engine = CHB_ENGINE_RESIDENT;
  }
  if (!g_encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode_tiled = (EncodeTiledFn)fn;
  }
  if (const char* e = getenv("CHB_NF_FIRST_IMAGES")) g_nf_first_images = atoll(e);
  if (const char* e = getenv("CHB_SELF_CLEAN")) g_self_clean = (e[0] != '0') ? 1 : 0;
  if (const char* e = getenv("CHB_E2E_STREAMS")) { int v = atoi(e); if (v >= 1 && v <= kPipeMax) g_pipe = v; }
  if (const char* e = getenv("CHB_E2E_CHUNK_KB")) { int v = atoi(e); if (v >= 64) g_chunk_kb = v; }
  if (const char* e = getenv("CHB_LPT")) g_lpt = (e[0] != '0') ? 1 : 0;
  if (const char* e = getenv("CHB_SPLIT")) g_split = (e[0] != '0') ? 1 : 0;
  if (const char* e = getenv("CHB_SPLIT_PCT")) { int v = atoi(e); if (v >= 10 && v <= 1000) g_split_pct = v; }
  const char* fg = getenv("CHB_FORCE_GENERIC");
  ctx->force_generic = (fg && fg[0] && fg[0] != '0') ? 1 : 0;
  *out = ctx;
  return CHB_OK;
}

extern "C" void chb_destroy(chb_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard guard;
  guard.enter(ctx->device);
  cudaDeviceSynchronize();
  for (PolicyEntry* pe : ctx->cache) { cudaFree(pe->dev); cudaFree(pe->optab); delete pe; }
  for (Workspace* ws : ctx->workspaces) {
    cudaFree(ws->states); cudaFree(ws->lists); cudaFree(ws->counters); cudaFree(ws->scratch); cudaFree(ws->inplace);
    cudaFree(ws->norm_tmp);
    delete ws;
  }
  for (int i = 0; i < kPipeMax; ++i) {
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    cudaFree(ctx->st_in[i]); cudaFree(ctx->st_out[i]);
    cudaFree(ctx->st_replay[i]); cudaFree(ctx->st_record[i]);
  }
  cudaFree(ctx->timeline);
  delete ctx;
}

extern "C" int chb_set_debug(chb_ctx* ctx, int force_generic) {
  if (!ctx) return CHB_ERR_INVALID;
  ctx->force_generic = force_generic ? 1 : 0;
  return CHB_OK;
}

extern "C" int chb_set_engine(chb_ctx* ctx, int engine) {
  if (!ctx) return CHB_ERR_INVALID;
  if (engine < CHB_ENGINE_AUTO || engine > CHB_ENGINE_RESIDENT) return fail(ctx, CHB_ERR_INVALID, "chb_set_engine: unknown engine");
  ctx->engine = engine;
  return CHB_OK;
}

extern "C" int chb_last_engine(const chb_ctx* ctx) { return ctx ? ctx->last_engine : 0; }

extern "C" int chb_debug_timeline(chb_ctx* ctx, uint64_t* host_out, int max_words) {
  if (!ctx) return CHB_ERR_INVALID;
#ifdef CHB_TIMELINE
  const size_t kTimelineWords = (size_t)1024 * 2 * 16;
  DeviceGuard guard;
  CHB_CUDA(ctx, guard.enter(ctx->device));
  if (!ctx->timeline) {
    CHB_CUDA(ctx, cudaMalloc(&ctx->timeline, kTimelineWords * 8));
    CHB_CUDA(ctx, cudaMemset(ctx->timeline, 0, kTimelineWords * 8));
    CHB_CUDA(ctx, cudaMemset(ctx->timeline + 1, 0xFF, 8));  // resident engine: earliest kernel entry (atomicMin)
  }
  if (!host_out) return 0;
  CHB_CUDA(ctx, cudaDeviceSynchronize());
  size_t n = max_words < 0 ? 0 : (size_t)max_words;
  if (n > kTimelineWords) n = kTimelineWords;
  CHB_CUDA(ctx, cudaMemcpy(host_out, ctx->timeline, n * 8, cudaMemcpyDeviceToHost));
  CHB_CUDA(ctx, cudaMemset(ctx->timeline, 0, kTimelineWords * 8));
  CHB_CUDA(ctx, cudaMemset(ctx->timeline + 1, 0xFF, 8));
  return (int)n;
#else
  (void)host_out; (void)max_words;
  return fail(ctx, CHB_ERR_UNSUPPORTED, "chb_debug_timeline: the library was built without -DCHB_TIMELINE");
#endif
}

extern "C" int chb_tile_plan(int H, int W, int* tiles_x, int* tiles_y, int* tw, int* th) {
  const chb::TilePlan t = chb::plan_tiles(H, W);
  if (tiles_x) *tiles_x = t.tiles_x;
  if (tiles_y) *tiles_y = t.tiles_y;
  if (tw) *tw = t.tw;
  if (th) *th = t.th;
  return t.n_tiles;
}

// ------------------------------------------------------------------------------------- launch
static int validate_policy(chb_ctx* ctx, const chb_policy* pol, int& K) {
  if (!pol || !pol->table) return fail(ctx, CHB_ERR_INVALID, "policy is NULL");
  if (pol->n_table < 1) return fail(ctx, CHB_ERR_INVALID, "policy has no transforms");
  if (pol->n_draws < 0) return fail(ctx, CHB_ERR_INVALID, "n_transforms must be >= 0");
  K = 1;
  for (int t = 0; t < pol->n_table; ++t) {
    const int n = pol->table[t].n_ops;
    if (n < 0 || n > CHB_MAX_SUBOPS)
      return fail(ctx, CHB_ERR_UNSUPPORTED, "a transform holds more than CHB_MAX_SUBOPS ops");
    if (n > K) K = n;
  }
  if ((long long)pol->n_draws * K > CHB_MAX_CHAIN)
    return fail(ctx, CHB_ERR_UNSUPPORTED, "n_transforms * ops-per-transform exceeds CHB_MAX_CHAIN");
  return CHB_OK;
}

static int get_policy(chb_ctx* ctx, const chb_policy* pol, int K, int H, int W, int C,
                      int64_t batch_total, cudaStream_t stream, PolicyEntry** out) {
  std::vector<DevOp> host((size_t)pol->n_table * K);
  const long long contrast_pixels = (long long)(pol->elementwise ? 1 : batch_total) * H * W;
  for (int t = 0; t < pol->n_table; ++t)
    for (int j = 0; j < K; ++j) {
      DevOp& d = host[(size_t)t * K + j];
      if (j < pol->table[t].n_ops) {
        std::string err;
        const int r = build_devop(pol->table[t].ops[j], H, W, C, contrast_pixels, d, err);
        if (r != CHB_OK) return fail(ctx, r, err);
      } else {
        memset(&d, 0, sizeof(d));
        d.kind = -1;
      }
    }
  PolicyEntry* hit = nullptr;
  for (PolicyEntry* pe : ctx->cache)
    if (pe->host.size() == host.size() && memcmp(pe->host.data(), host.data(), host.size() * sizeof(DevOp)) == 0) {
      hit = pe;
      break;
    }
  if (!hit) {
    if (ctx->cache.size() >= 64) {  // bounded cache: drop everything once nothing can be in flight
      CHB_CUDA(ctx, cudaDeviceSynchronize());
      for (PolicyEntry* pe : ctx->cache) { cudaFree(pe->dev); cudaFree(pe->optab); delete pe; }
      ctx->cache.clear();
    }
    hit = new PolicyEntry();
    hit->host = host;
    cudaError_t e = cudaMalloc(&hit->dev, host.size() * sizeof(DevOp));
    if (e != cudaSuccess) { delete hit; return cuda_fail(ctx, e, "policy table alloc"); }
    // pageable source: the runtime stages the bytes before returning, so `host` may go away.
    e = cudaMalloc(&hit->optab, host.size() * 256);
    if (e != cudaSuccess) { cudaFree(hit->dev); delete hit; return cuda_fail(ctx, e, "policy value-map alloc"); }
    e = cudaMemcpyAsync(hit->dev, hit->host.data(), host.size() * sizeof(DevOp), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
      e = chb::launch_optab(hit->dev, hit->optab, (int)host.size(), stream);
      ctx->launches += 1;
    }
    // other streams may use this entry later: make the upload visible to them too.
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { cudaFree(hit->dev); cudaFree(hit->optab); delete hit; return cuda_fail(ctx, e, "policy table upload"); }
    ctx->cache.push_back(hit);
  }
  *out = hit;
  return CHB_OK;
}

static int grow(chb_ctx* ctx, void** ptr, size_t* have, size_t want) {
  if (want <= *have) return CHB_OK;
  if (*ptr) CHB_CUDA(ctx, cudaFree(*ptr));  // cudaFree synchronises the device
  *ptr = nullptr;
  *have = 0;
  CHB_CUDA(ctx, cudaMalloc(ptr, want));
  *have = want;
  return CHB_OK;
}

static int get_workspace(chb_ctx* ctx, cudaStream_t stream, int B, size_t scratch_bytes,
                         size_t inplace_bytes, Workspace** out) {
  Workspace* ws = nullptr;
  for (Workspace* w : ctx->workspaces)
    if (w->stream == stream) ws = w;
  if (!ws) {
    ws = new Workspace();
    ws->stream = stream;
    ctx->workspaces.push_back(ws);
  }
  if ((size_t)B > ws->cap_images) {
    const size_t nb = (size_t)B;
    if (ws->states) CHB_CUDA(ctx, cudaFree(ws->states));
    if (ws->lists) CHB_CUDA(ctx, cudaFree(ws->lists));
    if (ws->counters) CHB_CUDA(ctx, cudaFree(ws->counters));
    ws->states = nullptr; ws->lists = nullptr; ws->counters = nullptr;
    ws->cap_images = 0;
    CHB_CUDA(ctx, cudaMalloc(&ws->states, nb * sizeof(chb::ImgState)));
    CHB_CUDA(ctx, cudaMalloc(&ws->lists, nb * chb::NBINS * sizeof(int)));
    // counters and the continuation list share one allocation (one memset per call): 32 counter
    // words, then one entry per possible pass after an image's first plus slack for void tickets
    CHB_CUDA(ctx, cudaMalloc(&ws->counters, (32 + nb * CHB_MAX_CHAIN + 2048) * sizeof(unsigned int)));
    ws->cap_images = nb;
    ws->clean = false;
  }
  int r = grow(ctx, (void**)&ws->scratch, &ws->scratch_bytes, scratch_bytes);
  if (r != CHB_OK) return r;
  r = grow(ctx, (void**)&ws->inplace, &ws->inplace_bytes, inplace_bytes);
  if (r != CHB_OK) return r;
  *out = ws;
  return CHB_OK;
}

// Can this call run on the image-resident engine (chb_resident.cuh)?  The image, the control block and
// a working region (histogram copies / the halo band of a gathered Sharpness) must fit one SM's shared
// memory; rows and images are whole 16-byte units at 16-byte aligned addresses (bulk copies, vector
// loads); every geometric op of the table is nearest / constant-fill (what the policies build,
// augmentation_schemes.py:7-9) -- bilinear warps and the other fill modes stay on the tile engine.
static bool resident_eligible(const chb_ctx* ctx, const PolicyEntry* pe, const uint8_t* d_in, const uint8_t* d_out,
                              int H, int W, int C) {
  if (ctx->engine == CHB_ENGINE_TILES || ctx->force_generic || ctx->res_smem <= 0) return false;
  const size_t img_bytes = (size_t)H * W * C;
  const size_t row = (size_t)W * C;
  if ((row & 15) != 0 || img_bytes == 0) return false;
  if ((((uintptr_t)d_in) & 15) != 0 || (((uintptr_t)d_out) & 15) != 0) return false;
  if (img_bytes > (size_t)chb::resident_max_chunk_bytes()) return false;
  const size_t fixed = chb::resident_ctl_bytes() + (pe->host.size() * (sizeof(DevOp) + 256) + 127) / 128 * 128 +
                       (img_bytes + 127) / 128 * 128;
  if (fixed > (size_t)ctx->res_smem) return false;
  const size_t aux = (size_t)ctx->res_smem - fixed;
  const size_t hist_copy = (size_t)C * 1024 + 16;
  if (aux < hist_copy + 4 * (row + 16) || aux < 10 * 1024) return false;  // (plan_order scratch 8.4 KB, 4 histogram copies <= 8 KB)
  for (const DevOp& d : pe->host) {
    const bool geo = d.kind == CHB_OP_SHEAR_X || d.kind == CHB_OP_SHEAR_Y || d.kind == CHB_OP_TRANSLATE_X ||
                     d.kind == CHB_OP_TRANSLATE_Y || d.kind == CHB_OP_ROTATE;
    if (geo && (d.interp != CHB_INTERP_NEAREST || d.fill_mode != CHB_FILL_CONSTANT)) return false;
  }
  return true;
}

static int launch_device(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                         const chb_policy* pol, int64_t batch_total, int64_t image_index_base,
                         uint64_t seed, uint32_t call_counter, const int32_t* d_replay,
                         int32_t* d_record, cudaStream_t stream, float* d_outf = nullptr, int norm_mode = -1) {
  if (!ctx) return CHB_ERR_INVALID;
  if (B < 0 || H < 0 || W < 0) return fail(ctx, CHB_ERR_INVALID, "negative shape");
  if (C < 1 || C > 4) return fail(ctx, CHB_ERR_UNSUPPORTED, "channels must be 1..4");
  int K = 1;
  int r = validate_policy(ctx, pol, K);
  if (r != CHB_OK) return r;
  if ((long long)H * W * C > 0x3FFFFFFFll) return fail(ctx, CHB_ERR_UNSUPPORTED, "image larger than 1 GiB");
  if (H >= (1 << 22) || W >= (1 << 22)) return fail(ctx, CHB_ERR_UNSUPPORTED, "image side of 4M pixels or more");
  DeviceGuard guard;
  CHB_CUDA(ctx, guard.enter(ctx->device));
  PolicyEntry* pe = nullptr;
  r = get_policy(ctx, pol, K, H, W, C, batch_total, stream, &pe);  // validates ops even when B == 0
  if (r != CHB_OK) return r;
  if (B == 0 || H == 0 || W == 0) return CHB_OK;
  if (!d_in || (!d_out && !d_outf)) return fail(ctx, CHB_ERR_INVALID, "NULL image pointer");
  const size_t img_bytes = (size_t)H * W * C;
  const chb::TilePlan tp = chb::plan_tiles(H, W);
  if ((unsigned long long)B * 2ull * (unsigned long long)tp.n_tiles >= 0xFFFF0000ull)
    return fail(ctx, CHB_ERR_UNSUPPORTED, "batch * tiles exceeds the 32-bit work counter");
  const int chain = pol->n_draws * K;
  // An op triggers an extra pass over the pixels only if it is a histogram op (COUNT) or finds the
  // kernel slot K occupied / frozen by an earlier op (WRITE_SCRATCH).  The CTA that completes such a
  // pass publishes the image's next pass in the continuation list, where every CTA of the same
  // launch holds a ticket (chb_kernels.cuh), so a call is one plan launch and one pass launch
  // whatever the chain length.
  // tiles read neighbours of their own region: an in-place call goes through a temporary.
  // Fused ImageNetNormalization epilogue (d_outf): the resident engine writes float32 from its last pass (modes tf /
  // torch).  Anything else -- caffe's channel reversal, calls only the tile engine can take -- runs the policy into a
  // library-owned uint8 buffer and the standalone normalisation kernel behind it (same bytes, one more pass).
  if (d_outf) {
    const bool fusable = (norm_mode == CHB_NORM_TF || norm_mode == CHB_NORM_TORCH) && (((uintptr_t)d_outf) & 15) == 0 &&
                         resident_eligible(ctx, pe, d_in, d_in, H, W, C);
    if (!fusable) {
      Workspace* wt = nullptr;
      r = get_workspace(ctx, stream, 1, 0, 0, &wt);
      if (r != CHB_OK) return r;
      r = grow(ctx, (void**)&wt->norm_tmp, &wt->norm_tmp_bytes, (size_t)B * img_bytes);
      if (r != CHB_OK) return r;
      r = launch_device(ctx, d_in, wt->norm_tmp, B, H, W, C, pol, batch_total, image_index_base, seed, call_counter, d_replay, d_record, stream);
      if (r != CHB_OK) return r;
      cudaError_t e = chb::launch_normalize(wt->norm_tmp, 0, d_outf, (unsigned long long)B * img_bytes, C, norm_mode, ctx->num_sms, stream);
      if (e != cudaSuccess) return cuda_fail(ctx, e, "normalize kernel launch");
      ctx->launches += 1;
      return CHB_OK;
    }
    d_out = nullptr;
  }
  const bool overlap = d_out && (d_in < d_out + (size_t)B * img_bytes) && (d_out < d_in + (size_t)B * img_bytes);
  const size_t stride = (img_bytes + 255) / 256 * 256;
  Workspace* ws = nullptr;
  // The resident engine reads an image completely before it writes any byte of it and never reads
  // another image's bytes: d_in == d_out needs no temporary there (a partial overlap still does).
  const bool resident = resident_eligible(ctx, pe, d_in, d_out ? d_out : d_in, H, W, C);
  if (ctx->engine == CHB_ENGINE_RESIDENT && !resident)
    return fail(ctx, CHB_ERR_UNSUPPORTED, "CHB_ENGINE_RESIDENT: this call is not eligible for the image-resident engine");
  if (resident) {
    const bool temp = overlap && d_in != d_out;
    // one CTA per SM; the smallest batches may cut expensive images into up to four parts (plan_order)
    const bool may_split = g_split && !overlap && pol->elementwise && B <= 512;
    const long long max_items = (long long)B * (may_split ? 4 : 1);
    int grid = (long long)ctx->num_sms < max_items ? ctx->num_sms : (int)max_items;
    r = get_workspace(ctx, stream, 1, chain >= 2 ? (size_t)grid * stride : 0, temp ? (size_t)B * img_bytes : 0, &ws);
    if (r != CHB_OK) return r;
    chb::KParams p;
    memset(&p, 0, sizeof(p));
    p.in = d_in; p.out = temp ? ws->inplace : d_out; p.B = B; p.H = H; p.W = W;
    p.ops = pe->dev; p.optab = pe->optab; p.T = pol->n_table; p.n_draws = pol->n_draws; p.K = K;
    p.elementwise = pol->elementwise ? 1 : 0;
    p.seed = seed; p.call_counter = call_counter; p.image_index_base = (unsigned long long)image_index_base;
    p.replay = d_replay; p.record = d_record;
    p.scratch = ws->scratch; p.scratch_stride = stride;
    p.counters = ws->counters;
    p.res_smem_bytes = ctx->res_smem;
    p.outf = d_outf; p.norm_mode = norm_mode;
    p.res_lpt = g_lpt;
    p.res_rules = 1;
    p.res_split_pct = g_split_pct;
    p.res_split = may_split ? 1 : 0;  // (parts of one image run on different CTAs: they must not write what another still reads)
    p.timeline = ctx->timeline;
    {
      const size_t fixed = chb::resident_ctl_bytes() + (pe->host.size() * (sizeof(DevOp) + 256) + 127) / 128 * 128 +
                           (img_bytes + 127) / 128 * 128;
      chb::resident_splits(p, C, (int)((size_t)ctx->res_smem - fixed));
    }
    cudaError_t e = cudaSuccess;
    if (!ws->clean || !g_self_clean) {
      e = cudaMemsetAsync(ws->counters, 0, (32 + ws->cap_images * CHB_MAX_CHAIN + 2048) * sizeof(unsigned int), stream);
      if (e != cudaSuccess) return cuda_fail(ctx, e, "counter reset");
    }
    ws->clean = false;
    e = chb::launch_resident(p, C, grid, stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "resident kernel launch");
    ctx->launches += 1;
    ctx->last_engine = CHB_ENGINE_RESIDENT;
    ws->clean = true;
    if (temp) CHB_CUDA(ctx, cudaMemcpyAsync(d_out, ws->inplace, (size_t)B * img_bytes, cudaMemcpyDeviceToDevice, stream));
    return CHB_OK;
  }
  ctx->last_engine = CHB_ENGINE_TILES;
  r = get_workspace(ctx, stream, B, chain >= 2 ? (size_t)B * 2 * stride : 0,
                    overlap ? (size_t)B * img_bytes : 0, &ws);
  if (r != CHB_OK) return r;

  chb::KParams p;
  memset(&p, 0, sizeof(p));
  p.in = d_in; p.out = overlap ? ws->inplace : d_out; p.B = B; p.H = H; p.W = W;
  p.ops = pe->dev; p.optab = pe->optab; p.T = pol->n_table; p.n_draws = pol->n_draws; p.K = K;
  p.elementwise = pol->elementwise ? 1 : 0;
  p.seed = seed; p.call_counter = call_counter; p.image_index_base = (unsigned long long)image_index_base;
  p.replay = d_replay; p.record = d_record;
  p.scratch = ws->scratch; p.scratch_stride = stride;
  p.states = ws->states; p.lists = ws->lists; p.counters = ws->counters;
  p.tiles_x = tp.tiles_x; p.tiles_y = tp.tiles_y; p.tw = tp.tw; p.th = tp.th; p.n_tiles = tp.n_tiles;
  p.force_generic = ctx->force_generic;
  p.timeline = ctx->timeline;
  {
    p.strip_rows = (H + tp.n_tiles - 1) / tp.n_tiles;
    const int ub = (C == 3) ? 48 : 16;
    p.flat_units = (int)(img_bytes / ub);
    p.n_flat_tiles = tp.n_tiles;
    p.flat_upt = (p.flat_units + tp.n_tiles - 1) / tp.n_tiles;
    p.flags = ((((uintptr_t)d_in & 15) == 0) ? 1 : 0) | ((((uintptr_t)p.out & 15) == 0) ? 2 : 0) |
              (((img_bytes & 15) == 0) ? 4 : 0) | (((((size_t)W * C) & 15) == 0) ? 8 : 0);
    if (!(p.flags & 4)) p.flags &= ~3;  // images that are not whole 16-byte units are not aligned beyond the first
    // When every pass without a spatial op runs as a flat run (aligned batch, fast executors), flat
    // runs get their own, smaller tile count: tiles of up to 256 units keep all 256 consumer threads
    // busy (a 224 x 224 x 3 image is 13 runs of 242 units instead of 16 of 196).
    if ((p.flags & 7) == 7 && !p.force_generic && p.flat_units > 0) {
      int nf = (p.flat_units + 255) / 256;
      if (nf < 1) nf = 1;
      if (nf < tp.n_tiles) {
        p.n_flat_tiles = nf;
        p.flat_upt = (p.flat_units + nf - 1) / nf;
      }
    }
  }
  // gather tiles fetch their source bounding box as one 3-D tensor-map box (rows must be whole 16-byte units)
  chb::TMap tm_in, tm_scr;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_scr, 0, sizeof(tm_scr));
  if ((((size_t)W * C) & 15) == 0 && ((uintptr_t)d_in & 15) == 0) {
    int br = 0, bb = 0;
    bool ok = encode_image_map(&tm_in, d_in, H, W * C, img_bytes, (size_t)B, &br, &bb);
    if (ok && ws->scratch) ok = encode_image_map(&tm_scr, ws->scratch, H, W * C, stride, (size_t)B * 2, &br, &bb);
    if (ok) { p.use_tmap = 1; p.box_rows = br; p.box_bytes = bb; }
  }
  long long grid = (long long)ctx->num_sms * chb::pass_ctas_per_sm(C);
  if ((long long)B * tp.n_tiles < grid) grid = (long long)B * tp.n_tiles;
  p.nf_first = ((long long)B * tp.n_tiles < 96 * grid) ? 1 : 0;  // small batch (< 96 tiles per CTA): see bin_of
  if (g_nf_first_images > 0 && (long long)B < g_nf_first_images * grid) p.nf_first = 1;
  p.cont = reinterpret_cast<int*>(ws->counters + 32);
  // The pass kernel's last CTA leaves counters and continuation list zeroed, so only a fresh (or
  // possibly dirty) workspace needs a memset.
  cudaError_t e = cudaSuccess;
  if (!ws->clean || !g_self_clean) {
    e = cudaMemsetAsync(ws->counters, 0, (32 + ws->cap_images * CHB_MAX_CHAIN + 2048) * sizeof(unsigned int), stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "counter reset");
  }
  ws->clean = false;
  e = chb::launch_plan(p, C, stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "plan kernel launch");
  ctx->launches += 1;
  e = chb::launch_pass(p, tm_in, tm_scr, C, (int)grid, stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "pass kernel launch");
  ctx->launches += 1;
  ws->clean = true;
  if (overlap) CHB_CUDA(ctx, cudaMemcpyAsync(d_out, ws->inplace, (size_t)B * img_bytes, cudaMemcpyDeviceToDevice, stream));
  return CHB_OK;
}

extern "C" int chb_policy_apply(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W,
                                int C, const chb_policy* policy, int64_t batch_total,
                                int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                                const int32_t* d_replay, int32_t* d_record, void* stream) {
  return launch_device(ctx, d_in, d_out, B, H, W, C, policy, batch_total, image_index_base, seed,
                       call_counter, d_replay, d_record, (cudaStream_t)stream);
}

extern "C" int chb_policy_apply_normalized(chb_ctx* ctx, const uint8_t* d_in, float* d_out, int B, int H, int W, int C,
                                           const chb_policy* policy, int norm_mode, int64_t batch_total,
                                           int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                                           const int32_t* d_replay, int32_t* d_record, void* stream) {
  if (!ctx) return CHB_ERR_INVALID;
  if (norm_mode != CHB_NORM_CAFFE && norm_mode != CHB_NORM_TF && norm_mode != CHB_NORM_TORCH)
    return fail(ctx, CHB_ERR_INVALID, "Unknown mode");
  if (norm_mode != CHB_NORM_TF && C != 3)
    return fail(ctx, CHB_ERR_INVALID, "caffe / torch normalisation broadcast 3-element constants: C must be 3");
  if (!d_out && (long long)B * H * W > 0) return fail(ctx, CHB_ERR_INVALID, "NULL image pointer");
  return launch_device(ctx, d_in, nullptr, B, H, W, C, policy, batch_total, image_index_base, seed, call_counter,
                       d_replay, d_record, (cudaStream_t)stream, d_out, norm_mode);
}

extern "C" int chb_randaugment(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W,
                               int C, int n_transforms, double magnitude, int elementwise,
                               int64_t batch_total, int64_t image_index_base, uint64_t seed,
                               uint32_t call_counter, const int32_t* d_replay, int32_t* d_record,
                               void* stream) {
  chb_transform table[CHB_OP_COUNT];
  chb_randaugment_table(magnitude, table);
  chb_policy pol = {CHB_OP_COUNT, n_transforms, elementwise, 0, table};
  return launch_device(ctx, d_in, d_out, B, H, W, C, &pol, batch_total, image_index_base, seed,
                       call_counter, d_replay, d_record, (cudaStream_t)stream);
}

extern "C" int chb_autoaugment(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W,
                               int C, int elementwise, int64_t batch_total, int64_t image_index_base,
                               uint64_t seed, uint32_t call_counter, const int32_t* d_replay,
                               int32_t* d_record, void* stream) {
  chb_transform table[25];
  chb_autoaugment_table(table);
  chb_policy pol = {25, 1, elementwise, 0, table};
  return launch_device(ctx, d_in, d_out, B, H, W, C, &pol, batch_total, image_index_base, seed,
                       call_counter, d_replay, d_record, (cudaStream_t)stream);
}

extern "C" int chb_apply_op(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                            const chb_op* op, int64_t batch_total, int64_t image_index_base,
                            uint64_t seed, uint32_t call_counter, const int32_t* d_replay,
                            int32_t* d_record, void* stream) {
  if (!op) return fail(ctx, CHB_ERR_INVALID, "op is NULL");
  chb_transform t;
  memset(&t, 0, sizeof(t));
  t.n_ops = 1;
  t.ops[0] = *op;
  chb_policy pol = {1, 1, 0, 0, &t};
  return launch_device(ctx, d_in, d_out, B, H, W, C, &pol, batch_total, image_index_base, seed,
                       call_counter, d_replay, d_record, (cudaStream_t)stream);
}

// ------------------------------------------------------------------- neighbours of the policy path
extern "C" int chb_imagenet_normalize(chb_ctx* ctx, const void* d_in, int in_is_f32, float* d_out, int64_t n_values,
                                      int C, int mode, void* stream) {
  if (!ctx) return CHB_ERR_INVALID;
  if (mode != CHB_NORM_CAFFE && mode != CHB_NORM_TF && mode != CHB_NORM_TORCH)
    return fail(ctx, CHB_ERR_INVALID, "Unknown mode");  // image_augmentations.py:624-625
  if (n_values < 0 || C < 1) return fail(ctx, CHB_ERR_INVALID, "negative size");
  if (mode != CHB_NORM_TF && C != 3)
    return fail(ctx, CHB_ERR_INVALID, "caffe / torch normalisation broadcast 3-element constants: C must be 3");
  if (n_values % C) return fail(ctx, CHB_ERR_INVALID, "n_values is not a whole number of pixels");
  if (n_values == 0) return CHB_OK;
  if (!d_in || !d_out) return fail(ctx, CHB_ERR_INVALID, "NULL pointer");
  DeviceGuard guard;
  CHB_CUDA(ctx, guard.enter(ctx->device));
  cudaError_t e = chb::launch_normalize(d_in, in_is_f32, d_out, (unsigned long long)n_values, C, mode, ctx->num_sms,
                                        (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "normalize kernel launch");
  ctx->launches += 1;
  return CHB_OK;
}

extern "C" int chb_resize_min_max_shape(int H, int W, int min_side, int max_side, int* out_h, int* out_w) {
  if (min_side <= 0 && max_side <= 0) return CHB_ERR_INVALID;  // :704-705
  const float height = (float)H, width = (float)W;              // :711-712
  float scale;
  if (min_side > 0 && max_side > 0) {                           // :714-719
    const float a = (float)max_side / fmaxf(width, height), b = (float)min_side / fminf(width, height);
    scale = fminf(a, b);
  } else if (min_side > 0) {
    scale = (float)min_side / fminf(width, height);             // :720-723
  } else {
    scale = (float)max_side / fmaxf(width, height);             // :724-727
  }
  volatile float nh = height * scale, nw = width * scale;       // :729-730, float32 products, truncating casts
  if (out_h) *out_h = (int)nh;
  if (out_w) *out_w = (int)nw;
  return CHB_OK;
}

extern "C" int chb_resize(chb_ctx* ctx, const void* d_in, int in_is_f32, void* d_out, int B, int H, int W, int C,
                          int out_h, int out_w, int nearest, void* stream) {
  if (!ctx) return CHB_ERR_INVALID;
  if (B < 0 || H < 0 || W < 0 || C < 1 || out_h < 0 || out_w < 0) return fail(ctx, CHB_ERR_INVALID, "negative shape");
  if ((long long)B * out_h * out_w * C == 0) return CHB_OK;
  if (H == 0 || W == 0) return fail(ctx, CHB_ERR_INVALID, "cannot resize an empty image");
  if (!d_in || !d_out) return fail(ctx, CHB_ERR_INVALID, "NULL pointer");
  DeviceGuard guard;
  CHB_CUDA(ctx, guard.enter(ctx->device));
  cudaError_t e = chb::launch_resize(d_in, in_is_f32, d_out, B, H, W, C, out_h, out_w, nearest, ctx->num_sms, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "resize kernel launch");
  ctx->launches += 1;
  return CHB_OK;
}

// ------------------------------------------------------------------------------ e2e pipeline
static int ensure_staging(chb_ctx* ctx, size_t img_bytes, size_t sched_bytes) {
  for (int i = 0; i < g_pipe; ++i)
    if (!ctx->streams[i]) CHB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking));
  if (img_bytes > ctx->st_img_bytes) {
    for (int i = 0; i < g_pipe; ++i) {
      if (ctx->st_in[i]) CHB_CUDA(ctx, cudaFree(ctx->st_in[i]));
      if (ctx->st_out[i]) CHB_CUDA(ctx, cudaFree(ctx->st_out[i]));
      ctx->st_in[i] = ctx->st_out[i] = nullptr;
    }
    ctx->st_img_bytes = 0;
    for (int i = 0; i < g_pipe; ++i) {
      CHB_CUDA(ctx, cudaMalloc(&ctx->st_in[i], img_bytes));
      CHB_CUDA(ctx, cudaMalloc(&ctx->st_out[i], img_bytes));
    }
    ctx->st_img_bytes = img_bytes;
  }
  if (sched_bytes > ctx->st_sched_bytes) {
    for (int i = 0; i < g_pipe; ++i) {
      if (ctx->st_replay[i]) CHB_CUDA(ctx, cudaFree(ctx->st_replay[i]));
      if (ctx->st_record[i]) CHB_CUDA(ctx, cudaFree(ctx->st_record[i]));
      ctx->st_replay[i] = ctx->st_record[i] = nullptr;
    }
    ctx->st_sched_bytes = 0;
    for (int i = 0; i < g_pipe; ++i) {
      CHB_CUDA(ctx, cudaMalloc(&ctx->st_replay[i], sched_bytes));
      CHB_CUDA(ctx, cudaMalloc(&ctx->st_record[i], sched_bytes));
    }
    ctx->st_sched_bytes = sched_bytes;
  }
  return CHB_OK;
}

extern "C" int chb_policy_apply_host(chb_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, int B, int H,
                                     int W, int C, const chb_policy* policy, int64_t batch_total,
                                     int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                                     const int32_t* h_replay, int32_t* h_record) {
  if (!ctx) return CHB_ERR_INVALID;
  if (B < 0 || H < 0 || W < 0) return fail(ctx, CHB_ERR_INVALID, "negative shape");
  if (C < 1 || C > 4) return fail(ctx, CHB_ERR_UNSUPPORTED, "channels must be 1..4");
  int K = 1;
  int r = validate_policy(ctx, policy, K);
  if (r != CHB_OK) return r;
  DeviceGuard guard;
  CHB_CUDA(ctx, guard.enter(ctx->device));
  const size_t img = (size_t)H * W * C;
  if (B == 0 || img == 0) {
    // still validate the ops the way a real call would
    return launch_device(ctx, nullptr, nullptr, 0, H, W, C, policy, batch_total, image_index_base,
                         seed, call_counter, nullptr, nullptr, 0);
  }
  if (!h_in || !h_out) return fail(ctx, CHB_ERR_INVALID, "NULL image pointer");
  // (Zero copy was measured and rejected: the resident kernel bulk-loading pinned, device-mapped host buffers
  // and storing straight back moved 9.6 GB/s each way against 35.8 GB/s for the staged pipeline below,
  // profiles/r02_ab_notes.md 9.)
  // chunk so that copy-in, kernel and copy-out of neighbouring chunks overlap: small enough that the
  // first copy-in and the last copy-out (which overlap nothing) are short, large enough that a chunk's
  // copies dwarf its launch overheads.
  long long n_chunks_ll = (long long)(((size_t)B * img + (size_t)g_chunk_kb * 1024 - 1) / ((size_t)g_chunk_kb * 1024));
  if (n_chunks_ll < 1) n_chunks_ll = 1;
  if (n_chunks_ll > 64) n_chunks_ll = 64;
  int n_chunks = (int)n_chunks_ll;
  if (B < n_chunks) n_chunks = B;
  const int per = (B + n_chunks - 1) / n_chunks;
  const size_t sched_per_img = (size_t)policy->n_draws * K * CHB_SCHED_FIELDS * sizeof(int32_t);
  r = ensure_staging(ctx, (size_t)per * img, (size_t)per * sched_per_img + 16);
  if (r != CHB_OK) return r;
  for (int c = 0, b0 = 0; b0 < B; ++c, b0 += per) {
    const int nb = (B - b0 < per) ? (B - b0) : per;
    const int sl = c % g_pipe;
    cudaStream_t st = ctx->streams[sl];
    // staging slot reuse is ordered by the slot's own stream.
    CHB_CUDA(ctx, cudaMemcpyAsync(ctx->st_in[sl], h_in + (size_t)b0 * img, (size_t)nb * img, cudaMemcpyHostToDevice, st));
    if (h_replay && sched_per_img)
      CHB_CUDA(ctx, cudaMemcpyAsync(ctx->st_replay[sl], (const uint8_t*)h_replay + (size_t)b0 * sched_per_img,
                                    (size_t)nb * sched_per_img, cudaMemcpyHostToDevice, st));
    r = launch_device(ctx, ctx->st_in[sl], ctx->st_out[sl], nb, H, W, C, policy, batch_total,
                      image_index_base + b0, seed, call_counter,
                      (h_replay && sched_per_img) ? ctx->st_replay[sl] : nullptr,
                      (h_record && sched_per_img) ? ctx->st_record[sl] : nullptr, st);
    if (r != CHB_OK) return r;
    CHB_CUDA(ctx, cudaMemcpyAsync(h_out + (size_t)b0 * img, ctx->st_out[sl], (size_t)nb * img, cudaMemcpyDeviceToHost, st));
    if (h_record && sched_per_img)
      CHB_CUDA(ctx, cudaMemcpyAsync((uint8_t*)h_record + (size_t)b0 * sched_per_img, ctx->st_record[sl],
                                    (size_t)nb * sched_per_img, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < g_pipe; ++i) CHB_CUDA(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return CHB_OK;
}
