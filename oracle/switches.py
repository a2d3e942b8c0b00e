"""Named switches for upstream (TensorFlow 2.6 / tensorflow-addons) behaviours the
oracle could not verify against a real install (see ``oracle/__init__.py``).

Each switch records the recalled upstream behaviour and the alternative.  The
CUDA path implements the DEFAULT of every switch; flipping one here without
changing ``chambers_b200/csrc`` makes the parity tests fail, by design.
"""

# Solarize(threshold) with threshold > 255 (M=10 -> 256, M=15 -> 384).
# ``inputs < self.threshold`` (image_augmentations.py:193) converts the Python
# int to the tensor dtype (uint8).  TF 2.x eager int->uint8 constant conversion
# casts through int32 and wraps silently: 256 -> 0, 384 -> 128.
# "wrap": threshold % 256.   "saturate": min(threshold, 256) i.e. nothing inverted at 256.
SOLARIZE_THRESHOLD_OVERFLOW = "wrap"

# Posterize(bits=0) -> shift = 8 (image_augmentations.py:168).  TF's
# bitwise shift functors clamp the shift to bitwidth-1 (cwise_ops.h) so the
# result is (x >> 7) << 7 in {0, 128}.   "clamp7" | "zero".
POSTERIZE_SHIFT8 = "clamp7"

# tfa.image.sharpness blends with TFA's own compose_ops.blend, which rounds
# (tf.round, half-to-even) after clipping; chambers' blend truncates.
# "round_half_even" | "truncate".
TFA_BLEND_ROUNDING = "round_half_even"

# Depthwise 3x3 tap accumulation order for tfa.image.sharpness: row-major,
# accumulator starting at 0, multiply and add unfused (AVX-only TF wheels).
SHARPNESS_TAP_ORDER = "row_major_unfused"

# cos / sin of the (float32) rotation angle: correctly rounded float32 of the
# float32 argument (scalar libm path).
ROTATE_TRIG = "correctly_rounded_f32"

# tf.image.rgb_to_grayscale: f32(x) * f32(1/255), weights accumulated r,g,b with
# unfused multiply-add, then trunc(g * 255.5).
GRAYSCALE_ACCUMULATION = "rgb_unfused"
