/*
 * chambers_aug.h -- C ABI of libchambers_aug.so: the B200-native (sm_100a) replacement for the
 * image-augmentation hot path of chjort/chambers (chambers.augmentations RandAugment /
 * AutoAugment policy ops on uint8 NHWC batches).
 *
 * The reference exposes NO FFI for this path: its boundary is the Keras Layer protocol
 * (SURVEY.md section 8b).  Each entry point below therefore names the reference *Python* interface
 * it replaces (file:line into /root/reference/chambers/augmentations); INTEGRATION.md shows the
 * ctypes stub a chambers maintainer would add.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures; `stream` is a cudaStream_t passed as void*.
 *   - images: contiguous NHWC uint8, device pointers unless the name ends in _host.
 *   - every call returns 0 on success, a CHB_ERR_* code otherwise; chb_last_error(ctx) has the text.
 *     Nothing throws across the ABI.  Device calls are stream-ordered and never synchronise.
 *   - one chb_ctx per (host thread, device); a ctx is not re-entrant.
 *   - there is NO CPU fallback: without a CUDA device chb_init fails.
 *   - working memory (per-image pass state, scratch images) is owned by the ctx, keyed by stream and
 *     grown on demand; a call that grows it synchronises the device once.
 */
#ifndef CHAMBERS_AUG_H_
#define CHAMBERS_AUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHB_VERSION_MAJOR 0
#define CHB_VERSION_MINOR 1

#define CHB_MAX_SUBOPS 4      /* ops per transform (a Sequential of RandomChance layers)      */
#define CHB_MAX_CHAIN 8       /* n_draws * (ops per transform): longest per-image op chain    */
#define CHB_MAX_TABLE_OPS 64  /* transforms * (ops per transform) in one policy               */

/* Op kinds.  0..15 are the RandAugment transform order, augmentation_schemes.py:181-198. */
enum chb_op_kind {
  CHB_OP_AUTOCONTRAST = 0, /* image_augmentations.py:62-90   */
  CHB_OP_EQUALIZE = 1,     /* :93-103  (tfa.image.equalize)  */
  CHB_OP_INVERT = 2,       /* :106-116 */
  CHB_OP_BRIGHTNESS = 3,   /* :276-293 */
  CHB_OP_CONTRAST = 4,     /* :246-273 */
  CHB_OP_COLOR = 5,        /* :226-243 */
  CHB_OP_SHARPNESS = 6,    /* :296-312 (tfa.image.sharpness) */
  CHB_OP_SHEAR_X = 7,      /* :315-355 */
  CHB_OP_SHEAR_Y = 8,      /* :358-398 */
  CHB_OP_TRANSLATE_X = 9,  /* :401-441 */
  CHB_OP_TRANSLATE_Y = 10, /* :444-484 */
  CHB_OP_POSTERIZE = 11,   /* :163-182 */
  CHB_OP_SOLARIZE = 12,    /* :185-201 */
  CHB_OP_SOLARIZE_ADD = 13,/* :204-223 */
  CHB_OP_CUTOUT = 14,      /* :487-507 (tfa.image.random_cutout) */
  CHB_OP_ROTATE = 15,      /* :119-160 (tfa.image.rotate) */
  CHB_OP_COUNT = 16
};

enum chb_interpolation { CHB_INTERP_NEAREST = 0, CHB_INTERP_BILINEAR = 1 };
enum chb_fill_mode { CHB_FILL_CONSTANT = 0, CHB_FILL_REFLECT = 1, CHB_FILL_WRAP = 2, CHB_FILL_NEAREST = 3 };

enum chb_error {
  CHB_OK = 0,
  CHB_ERR_INVALID = 1,     /* bad argument (the reference raises ValueError)              */
  CHB_ERR_CUDA = 2,        /* CUDA runtime error; text in chb_last_error                   */
  CHB_ERR_UNSUPPORTED = 3, /* legal in the reference but not built         */
  CHB_ERR_NO_DEVICE = 4
};

/* One op layer as CONSTRUCTED in the reference: the constructor kwargs, nothing derived.
 * `value` is the Python float given to the constructor (factor / level / pixels / degrees);
 * ivalue: Posterize {bits,-}; Solarize {threshold,-}; SolarizeAdd {addition, threshold};
 *         CutOut {mask_size, constant_values}.
 * probability < 0: the op is used directly (no coin);  >= 0: wrapped in RandomChance(op, p)
 * (image_augmentations.py:513-545). */
typedef struct chb_op {
  int32_t kind;          /* enum chb_op_kind */
  int32_t interpolation; /* enum chb_interpolation (geometric ops) */
  int32_t fill_mode;     /* enum chb_fill_mode     (geometric ops) */
  int32_t ivalue[2];
  float fill_value;      /* geometric ops; cast to uint8 like the TF kernel does */
  double probability;    /* < 0: no coin */
  double value;
} chb_op;                /* 40 bytes */

/* One entry of RandomChoice.transforms (image_augmentations.py:548-561): an op layer, a
 * RandomChance, or a Sequential of up to CHB_MAX_SUBOPS of them. */
typedef struct chb_transform {
  int32_t n_ops;
  int32_t _pad;
  chb_op ops[CHB_MAX_SUBOPS];
} chb_transform;

/* RandomChoice(transforms, n_transforms, elementwise), image_augmentations.py:550-561. */
typedef struct chb_policy {
  int32_t n_table;     /* len(transforms)                                                  */
  int32_t n_draws;     /* n_transforms: draws WITH replacement, :606-617                   */
  int32_t elementwise; /* 1: per-image schedule (:565-567); 0: one schedule per batch (:569) */
  int32_t _pad;
  const chb_transform* table;
} chb_policy;

/* Explicit schedule entry, int32 x 5, layout [B][n_draws][K][5] with K = max n_ops over the
 * table: (choice, applied, negate, cy, cx).  Used to record what the device decoded and to replay
 * a given schedule (parity against the oracle is defined on identical schedules). */
#define CHB_SCHED_FIELDS 5

typedef struct chb_ctx chb_ctx;

int chb_version(void); /* CHB_VERSION_MAJOR * 1000 + CHB_VERSION_MINOR */

/* Context: owns per-device scratch, the uploaded policy tables and the e2e staging buffers. */
int chb_init(int device, chb_ctx** out);
void chb_destroy(chb_ctx* ctx);
const char* chb_last_error(const chb_ctx* ctx); /* ctx may be NULL: last chb_init error */

/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t chb_kernel_launches(const chb_ctx* ctx);

/* Table builders -- the C twins of RandAugment.__init__ (augmentation_schemes.py:176-201, fills 16
 * transforms) and AutoAugment.__init__ (:135-149, fills 25) including the magnitude maps :42-102. */
int chb_randaugment_table(double magnitude, chb_transform* out16);
int chb_autoaugment_table(chb_transform* out25);

/* RandomChoice.call on the device (image_augmentations.py:563-617), i.e. what
 * RandAugment.call / AutoAugment.call run when training is true
 * (augmentation_schemes.py:204-213, :152-161).
 *   batch_total       number of images in the WHOLE batch (all shards); Contrast's constant
 *                     depends on it when elementwise == 0 (image_augmentations.py:260-264).
 *   image_index_base  global index of d_in's image 0; the RNG is keyed by global image index so a
 *                     batch sharded over 1/2/4/8 GPUs yields identical pixels.
 *   seed, call_counter  Philox4x32-10 key / call index (see oracle/philox.py for the layout).
 *   d_replay          NULL, or device int32 [B][n_draws][K][5] schedule to use instead of the RNG.
 *   d_record          NULL, or device int32 buffer of the same shape receiving the schedule used.
 * d_in may equal d_out (or overlap it): tiles read neighbours of their own region, so such a call
 * is routed through a library-owned temporary and one extra device-to-device copy. */
int chb_policy_apply(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                     const chb_policy* policy, int64_t batch_total, int64_t image_index_base,
                     uint64_t seed, uint32_t call_counter, const int32_t* d_replay,
                     int32_t* d_record, void* stream);

/* RandAugment(n_transforms, magnitude, elementwise)(x, training=True),
 * augmentation_schemes.py:174-213. */
int chb_randaugment(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                    int n_transforms, double magnitude, int elementwise, int64_t batch_total,
                    int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                    const int32_t* d_replay, int32_t* d_record, void* stream);

/* AutoAugment(elementwise)(x, training=True), augmentation_schemes.py:131-161. */
int chb_autoaugment(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                    int elementwise, int64_t batch_total, int64_t image_index_base, uint64_t seed,
                    uint32_t call_counter, const int32_t* d_replay, int32_t* d_record, void* stream);

/* One op layer called directly, e.g. Rotate(30.)(x) (image_augmentations.py:138-147): one sign
 * flip per batch, CutOut centres per image.  Equivalent to a 1-entry policy with elementwise=0. */
int chb_apply_op(chb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int B, int H, int W, int C,
                 const chb_op* op, int64_t batch_total, int64_t image_index_base, uint64_t seed,
                 uint32_t call_counter, const int32_t* d_replay, int32_t* d_record, void* stream);

/* End-to-end form with HOST buffers (h_in / h_out ideally pinned): the batch is cut into chunks
 * that are copied in, augmented and copied out on rotating streams so the three overlap.
 * Synchronous: returns when h_out (and h_record) are complete.  h_replay / h_record are host
 * pointers or NULL. */
int chb_policy_apply_host(chb_ctx* ctx, const uint8_t* h_in, uint8_t* h_out, int B, int H, int W,
                          int C, const chb_policy* policy, int64_t batch_total,
                          int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                          const int32_t* h_replay, int32_t* h_record);

/* Debugging switch: force_generic != 0 routes every tile through the scalar executor instead of
 * the vectorised ones (same results, used by the tests to cross-check the two on the device). */
int chb_set_debug(chb_ctx* ctx, int force_generic);

/* ImageNetNormalization(mode)(x), image_augmentations.py:620-682 -- the layer every chambers backbone
 * applies right after the augmentation policy (vision_transformer.py:655).  d_in: NHWC uint8
 * (in_is_f32 == 0) or float32 (in_is_f32 != 0) with n_values = B*H*W*C elements; d_out: float32 of the
 * same shape.  tf: x / 127.5 - 1 (:660-665); torch: (x / 255 - mean) / std per channel (:652-657);
 * caffe: channels reversed (RGB -> BGR), minus the BGR means (:647-650).  torch / caffe need C == 3
 * (the reference broadcasts 3-element constants).  Checked against the reference's own golden
 * vectors (test_units/augmentations/test_image_augmentations.py:21-64). */
enum chb_norm_mode { CHB_NORM_CAFFE = 0, CHB_NORM_TF = 1, CHB_NORM_TORCH = 2 };
int chb_imagenet_normalize(chb_ctx* ctx, const void* d_in, int in_is_f32, float* d_out, int64_t n_values, int C,
                           int mode, void* stream);

/* The policy followed by ImageNetNormalization(mode) -- the reference's call order in front of every backbone
 * (RandAugment / AutoAugment, then preprocess_input = ImageNetNormalization(mode="tf"), vision_transformer.py:655) --
 * as ONE call: d_out is float32 NHWC of the input's shape and holds exactly what chb_imagenet_normalize would
 * produce from chb_policy_apply's uint8 result.  On the image-resident engine (modes tf / torch) the
 * normalisation is the write epilogue of the policy's last pass (no uint8 image is written or re-read);
 * otherwise the library runs the two kernels back to back through a buffer of its own. */
int chb_policy_apply_normalized(chb_ctx* ctx, const uint8_t* d_in, float* d_out, int B, int H, int W, int C,
                                const chb_policy* policy, int norm_mode, int64_t batch_total,
                                int64_t image_index_base, uint64_t seed, uint32_t call_counter,
                                const int32_t* d_replay, int32_t* d_record, void* stream);

/* ResizingMinMax(min_side, max_side, interpolation)(x), image_augmentations.py:685-748.
 * chb_resize_min_max_shape is the size arithmetic of :711-730 (float32 scale, truncating casts);
 * min_side / max_side <= 0 mean None; returns CHB_ERR_INVALID if both are.  chb_resize is the
 * tf.image.resize call behind Keras' Resizing (:732-735): bilinear (float32 output whatever the
 * input) or nearest (output of the input's type), half-pixel centres, no antialiasing. */
int chb_resize_min_max_shape(int H, int W, int min_side, int max_side, int* out_h, int* out_w);
int chb_resize(chb_ctx* ctx, const void* d_in, int in_is_f32, void* d_out, int B, int H, int W, int C, int out_h,
               int out_w, int nearest, void* stream);

/* Engine selection.  The library holds two engines with identical results:
 *   CHB_ENGINE_RESIDENT  one CTA per SM keeps a whole image in shared memory for its entire op chain
 *                        (images of up to ~180 KB whose rows are whole 16-byte units, nearest /
 *                        constant-fill geometric ops: every BASELINE 224 x 224 x 3 configuration);
 *   CHB_ENGINE_TILES     the tile-parallel engine: any shape, any fill mode, bilinear warps.
 * CHB_ENGINE_AUTO (default) takes the resident engine whenever a call is eligible.  Forcing
 * CHB_ENGINE_RESIDENT makes ineligible calls fail with CHB_ERR_UNSUPPORTED (tests use it to be sure
 * which engine they compare).  Environment: CHB_ENGINE=tiles|resident at chb_init. */
enum chb_engine { CHB_ENGINE_AUTO = 0, CHB_ENGINE_TILES = 1, CHB_ENGINE_RESIDENT = 2 };
int chb_set_engine(chb_ctx* ctx, int engine);
int chb_last_engine(const chb_ctx* ctx); /* engine of the last device call: CHB_ENGINE_TILES / _RESIDENT */

/* Profiling hook of debug builds (-DCHB_TIMELINE; the production library returns
 * CHB_ERR_UNSUPPORTED).  host_out == NULL: start recording per-CTA timestamps of the pass kernels.
 * Otherwise: synchronise the device, copy up to max_words 64-bit words of the records to host_out
 * ([1024 CTAs][producer | consumer][16]) and clear them; returns the words copied. */
int chb_debug_timeline(chb_ctx* ctx, uint64_t* host_out, int max_words);

/* How an H x W image is cut into work items: tiles_x * tiles_y tiles of at most tw x th pixels
 * (the same count cuts the image into flat runs / row strips for passes without a spatial op).
 * Pure function of the shape; any of the out pointers may be NULL.  Returns the tile count. */
int chb_tile_plan(int H, int W, int* tiles_x, int* tiles_y, int* tw, int* th);

#ifdef __cplusplus
}
#endif
#endif /* CHAMBERS_AUG_H_ */
