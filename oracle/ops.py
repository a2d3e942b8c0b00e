"""Numpy restatement of the 16 policy ops of chambers.augmentations.

Every function takes a uint8 NHWC batch ``[B, H, W, C]`` and returns a new uint8
batch of the same shape.  All randomness is EXPLICIT: the sign flip of the
geometric ops (``negate``) and the CutOut centres are arguments, so the oracle
is RNG-agnostic (SURVEY.md section 8a "RNG").  Float arithmetic is done in
numpy float32 one ufunc at a time, which reproduces the reference's op-by-op
float32 temporaries (no fused multiply-add, one rounding per op).

Citations are ``file:line`` into ``/root/reference/chambers/augmentations``.
See ``oracle/__init__.py``: test infrastructure only; parity unpinned.
"""

import math

import numpy as np

from . import switches

f32 = np.float32

__all__ = [
    "OP_NAMES", "OP_INDEX", "blend", "autocontrast", "equalize", "invert", "brightness",
    "contrast", "contrast_constant", "color", "rgb_to_grayscale", "sharpness", "shear_x",
    "shear_y", "translate_x", "translate_y", "posterize", "solarize", "solarize_add",
    "cutout", "rotate", "projective_transform", "rotate_coeffs", "shear_x_coeffs",
    "shear_y_coeffs", "translate_x_coeffs", "translate_y_coeffs", "apply_op", "round_half_away",
    "tfa_blend", "equalize_lut", "wrap_threshold", "posterize_shift",
]

# Fixed op order of RandAugment.transforms (augmentation_schemes.py:181-198); the
# index doubles as the op-kind id of the C ABI (include/chambers_aug.h).
OP_NAMES = [
    "AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness",
    "ShearX", "ShearY", "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd",
    "CutOut", "Rotate",
]
OP_INDEX = {n: i for i, n in enumerate(OP_NAMES)}


def _check(images):
    images = np.asarray(images)
    if images.ndim != 4:
        raise ValueError("expected a 4-D NHWC batch, got shape %r" % (images.shape,))
    if images.dtype != np.uint8:
        raise ValueError("expected uint8, got %s" % images.dtype)
    return images


def _trunc_u8(x):
    """tf.cast(float32 -> uint8): truncation toward zero (values already in range)."""
    return x.astype(np.uint8)


# --------------------------------------------------------------------------- blend
def blend(image1, image2, factor):
    """chambers' own blend, image_augmentations.py:10-49."""
    if factor == 0.0:  # :28-29
        return np.array(image1, dtype=np.uint8, copy=True)
    if factor == 1.0:  # :30-31
        return np.array(image2, dtype=np.uint8, copy=True)
    i1 = np.asarray(image1).astype(f32)  # :33
    i2 = np.asarray(image2).astype(f32)  # :34
    difference = i2 - i1  # :36
    scaled = f32(factor) * difference  # :37 (python float -> float32 constant)
    temp = i1 + scaled  # :40
    if factor > 0.0 and factor < 1.0:  # :43-45 truncating cast, no clip
        return _trunc_u8(temp)
    return _trunc_u8(np.clip(temp, f32(0.0), f32(255.0)))  # :49


def tfa_blend(image1, image2, factor):
    """tensorflow_addons.image.compose_ops.blend as used by tfa.image.sharpness
    (recalled; switch TFA_BLEND_ROUNDING)."""
    if factor == 0.0:
        return np.array(image1, dtype=np.uint8, copy=True)
    if factor == 1.0:
        return np.array(image2, dtype=np.uint8, copy=True)
    i1 = np.asarray(image1).astype(f32)
    i2 = np.asarray(image2).astype(f32)
    temp = i1 + f32(factor) * (i2 - i1)
    temp = np.clip(temp, f32(0.0), f32(255.0))
    if switches.TFA_BLEND_ROUNDING == "round_half_even":
        temp = np.rint(temp)
    return _trunc_u8(temp)


# -------------------------------------------------------------------- colour ops
def autocontrast(images):
    """AutoContrast.call, image_augmentations.py:68-87."""
    images = _check(images)
    lo = images.min(axis=(1, 2)).astype(f32)  # :69
    hi = images.max(axis=(1, 2)).astype(f32)  # :70
    rng = hi - lo
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(rng == 0, f32(0.0), f32(255.0) / rng).astype(f32)  # :72 divide_no_nan
    offset = (-lo) * scale  # :73
    mask = (hi > lo).astype(f32)  # :76
    scale = scale * mask + (f32(1.0) - mask)  # :77
    offset = offset * mask  # :78
    x = images.astype(f32) * scale[:, None, None, :] + offset[:, None, None, :]  # :84 (mul, then add)
    x = np.clip(x, f32(0.0), f32(255.0))  # :85
    return _trunc_u8(x)  # :86


def equalize_lut(histo):
    """LUT of tfa.image.equalize's _scale_channel for one 256-bin histogram
    (all int arithmetic); returns None when step == 0 (channel unchanged)."""
    histo = np.asarray(histo, dtype=np.int64)
    nz = histo[histo != 0]
    step = (int(nz.sum()) - int(nz[-1])) // 255
    if step == 0:
        return None
    excl = np.cumsum(histo) - histo
    lut = (excl + step // 2) // step
    return np.clip(lut, 0, 255).astype(np.uint8)


def equalize(images):
    """Equalize.call -> tfa.image.equalize, image_augmentations.py:99-100:
    per image, per channel 256-bin histogram -> CDF LUT -> gather."""
    images = _check(images)
    out = images.copy()
    B, _, _, C = images.shape
    for b in range(B):
        for c in range(C):
            ch = images[b, :, :, c]
            histo = np.bincount(ch.ravel(), minlength=256)
            lut = equalize_lut(histo)
            if lut is not None:
                out[b, :, :, c] = lut[ch]
    return out


def invert(images):
    """Invert.call, image_augmentations.py:112-113."""
    return (255 - _check(images)).astype(np.uint8)


def brightness(images, factor):
    """Brightness.call, image_augmentations.py:283-285."""
    images = _check(images)
    return blend(np.zeros_like(images), images, factor)


def rgb_to_grayscale(images):
    """tf.image.rgb_to_grayscale on uint8 (recalled, switch GRAYSCALE_ACCUMULATION):
    f32(x) * f32(1/255); dot with [0.2989, 0.5870, 0.1140] accumulated r, g, b;
    convert back with trunc(g * 255.5).  Returns [B, H, W, 1] uint8."""
    images = _check(images)
    if images.shape[-1] != 3:
        raise ValueError("rgb_to_grayscale needs 3 channels, got %d" % images.shape[-1])
    flt = images.astype(f32) * f32(1.0 / 255.0)
    g = flt[..., 0] * f32(0.2989)
    g = g + flt[..., 1] * f32(0.5870)
    g = g + flt[..., 2] * f32(0.1140)
    return _trunc_u8(g * f32(255.5))[..., None]


def color(images, factor):
    """Color.call, image_augmentations.py:233-235."""
    images = _check(images)
    degenerate = np.repeat(rgb_to_grayscale(images), 3, axis=-1)  # :234
    return blend(degenerate, images, factor)  # :235


def contrast_constant(n_pixels):
    """The constant Contrast blends against (image_augmentations.py:260-264):
    sum(hist)/256 == (number of pixels in the tensor passed in)/256, clipped to
    255 and truncated to uint8 -- NOT the mean intensity (inherited upstream bug,
    reproduced on purpose)."""
    mean = f32(n_pixels) / f32(256.0)
    return int(np.clip(mean, f32(0.0), f32(255.0)))


def contrast(images, factor):
    """Contrast.call, image_augmentations.py:253-265.  The degenerate image
    depends on how many pixels were passed in, i.e. on whether RandomChoice runs
    batch-wise (B images) or elementwise (1 image)."""
    images = _check(images)
    if images.shape[-1] != 3:
        raise ValueError("Contrast needs 3 channels (rgb_to_grayscale)")
    B, H, W, _ = images.shape
    c = contrast_constant(B * H * W)
    degenerate = np.full_like(images, c)
    return blend(degenerate, images, factor)


def sharpness(images, factor):
    """Sharpness.call -> tfa.image.sharpness, image_augmentations.py:303-304
    (recalled TFA algorithm; switches SHARPNESS_TAP_ORDER, TFA_BLEND_ROUNDING)."""
    images = _check(images)
    B, H, W, C = images.shape
    if H < 3 or W < 3:
        # VALID conv output is empty; every pixel is border -> result == original.
        return tfa_blend(images, images, factor)
    x = images.astype(f32)
    k1 = f32(1.0) / f32(13.0)
    k5 = f32(5.0) / f32(13.0)
    acc = np.zeros((B, H - 2, W - 2, C), dtype=f32)
    for dy in range(3):
        for dx in range(3):
            k = k5 if (dy == 1 and dx == 1) else k1
            acc = acc + x[:, dy:dy + H - 2, dx:dx + W - 2, :] * k
    degenerate = images.copy()
    degenerate[:, 1:-1, 1:-1, :] = _trunc_u8(acc)
    return tfa_blend(degenerate, images, factor)


def posterize_shift(bits):
    """shift = 8 - bits (image_augmentations.py:168) as TF's shift functors see
    it (switch POSTERIZE_SHIFT8): clamped to [0, 7]."""
    shift = 8 - int(bits)
    if shift < 0:
        shift = 0
    if shift > 7:
        shift = 7 if switches.POSTERIZE_SHIFT8 == "clamp7" else 8
    return shift


def posterize(images, bits):
    """Posterize.call, image_augmentations.py:171-174."""
    images = _check(images)
    shift = posterize_shift(bits)
    if shift >= 8:
        return np.zeros_like(images)
    return ((images >> shift) << shift).astype(np.uint8)


def wrap_threshold(threshold):
    """Python int threshold as the uint8 comparison sees it (switch
    SOLARIZE_THRESHOLD_OVERFLOW).  Returns an int in [0, 256]."""
    t = int(threshold)
    if 0 <= t <= 255:
        return t
    if switches.SOLARIZE_THRESHOLD_OVERFLOW == "wrap":
        return t % 256
    return max(0, min(t, 256))


def solarize(images, threshold=128):
    """Solarize.call, image_augmentations.py:192-193."""
    images = _check(images)
    t = wrap_threshold(threshold)
    return np.where(images.astype(np.int32) < t, images, 255 - images).astype(np.uint8)


def solarize_add(images, addition=0, threshold=128):
    """SolarizeAdd.call, image_augmentations.py:212-215."""
    images = _check(images)
    t = wrap_threshold(threshold)
    x = np.clip(images.astype(np.int64) + int(addition), 0, 255).astype(np.uint8)
    return np.where(images.astype(np.int32) < t, x, images).astype(np.uint8)


def cutout(images, mask_size, constant_values=0, centers=None):
    """CutOut.call -> tfa.image.random_cutout, image_augmentations.py:495-499
    with the per-image centres ``centers[b] = (cy, cx)`` given explicitly."""
    images = _check(images)
    B, H, W, _ = images.shape
    if int(mask_size) % 2 != 0:
        raise ValueError("mask_size should be divisible by 2 (tfa.image.cutout)")
    centers = np.asarray(centers, dtype=np.int64).reshape(B, 2)
    h = int(mask_size) // 2
    out = images.copy()
    fill = np.uint8(int(constant_values) & 0xFF)
    for b in range(B):
        cy, cx = int(centers[b, 0]), int(centers[b, 1])
        y0, y1 = max(0, cy - h), min(H, cy + h)
        x0, x1 = max(0, cx - h), min(W, cx + h)
        if y1 > y0 and x1 > x0:
            out[b, y0:y1, x0:x1, :] = fill
    return out


# ------------------------------------------------------------------ geometric ops
def round_half_away(v):
    """std::round on float32 arrays (half away from zero), exactly."""
    v = np.asarray(v, dtype=f32)
    r = np.trunc(v)
    d = v - r  # exact
    return (r + (d >= f32(0.5)).astype(f32) - (d <= f32(-0.5)).astype(f32)).astype(f32)


def _map_coordinate(coord, length, fill_mode):
    """tensorflow/core/kernels/image/image_ops.h MapCoordinate (recalled)."""
    coord = np.asarray(coord, dtype=f32)
    if fill_mode == "constant":
        return coord
    n = int(length)
    if fill_mode == "nearest":
        return np.clip(coord, f32(0.0), f32(n - 1))
    out = coord.copy()
    if fill_mode == "reflect":
        if n <= 1:
            out[:] = 0
        else:
            sz2 = f32(2 * n)
            neg = coord < 0
            v = coord[neg]
            lt = v < sz2  # always true for negatives; kept as in the kernel
            q = np.trunc(-v / sz2)  # static_cast<DenseIndex>(-in/sz2)
            v = np.where(lt, sz2 * q + v, v).astype(f32)
            v = np.where(v < f32(-n), v + sz2, -v - f32(1.0)).astype(f32)
            out[neg] = v
            big = coord > f32(n - 1)
            w = coord[big]
            w = (w - sz2 * np.trunc(w / sz2)).astype(f32)
            w = np.where(w >= f32(n), sz2 - w - f32(1.0), w).astype(f32)
            out[big] = w
    elif fill_mode == "wrap":
        if n <= 1:
            out[:] = 0
        else:
            sz = f32(n - 1)
            neg = coord < 0
            v = coord[neg]
            out[neg] = (v + f32(n) * (np.trunc(-v / sz) + f32(1.0))).astype(f32)
            big = coord > f32(n - 1)
            w = coord[big]
            out[big] = (w - f32(n) * np.trunc(w / sz)).astype(f32)
    else:
        raise ValueError("unknown fill_mode %r" % (fill_mode,))
    return np.clip(out, f32(0.0), f32(n - 1)).astype(f32)


def projective_transform(images, coeffs, interpolation="nearest", fill_mode="constant",
                         fill_value=0.0):
    """tf.raw_ops.ImageProjectiveTransformV3 as reached through
    tfa.image.transform (image_augmentations.py:335-341 etc.; kernel semantics
    recalled from image_ops.h ProjectiveGenerator).

    ``coeffs``: 8 float32 per image ``[B, 8]`` or shared ``[8]``; output pixel
    (x, y) reads input ((t0 x + t1 y + t2)/k, (t3 x + t4 y + t5)/k), k = t6 x + t7 y + 1.
    """
    images = _check(images)
    B, H, W, C = images.shape
    coeffs = np.asarray(coeffs, dtype=f32)
    if coeffs.ndim == 1:
        coeffs = np.broadcast_to(coeffs, (B, 8))
    fill = np.uint8(int(np.float32(fill_value)) & 0xFF)  # static_cast<T>(fill_value)
    xs = np.arange(W, dtype=f32)[None, :]
    ys = np.arange(H, dtype=f32)[:, None]
    out = np.empty_like(images)

    def read(img, iy, ix):
        inside = (iy >= 0) & (iy < H) & (ix >= 0) & (ix < W)
        iyc = np.clip(iy, 0, H - 1)
        ixc = np.clip(ix, 0, W - 1)
        v = img[iyc, ixc, :]
        return np.where(inside[..., None], v, fill)

    for b in range(B):
        t = coeffs[b]
        proj = (t[6] * xs + t[7] * ys) + f32(1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            sx = ((t[0] * xs + t[1] * ys) + t[2]) / proj
            sy = ((t[3] * xs + t[4] * ys) + t[5]) / proj
        sx = _map_coordinate(sx.astype(f32), W, fill_mode)
        sy = _map_coordinate(sy.astype(f32), H, fill_mode)
        img = images[b]
        if interpolation == "nearest":
            ix = round_half_away(sx).astype(np.int64)
            iy = round_half_away(sy).astype(np.int64)
            res = read(img, iy, ix)
        elif interpolation == "bilinear":
            xf = np.floor(sx)
            yf = np.floor(sy)
            xc = xf + f32(1.0)
            yc = yf + f32(1.0)
            ixf, iyf = xf.astype(np.int64), yf.astype(np.int64)
            ixc, iyc = xc.astype(np.int64), yc.astype(np.int64)
            wxf = (xc - sx)[..., None]
            wxc = (sx - xf)[..., None]
            v_yf = wxf * read(img, iyf, ixf).astype(f32) + wxc * read(img, iyf, ixc).astype(f32)
            v_yc = wxf * read(img, iyc, ixf).astype(f32) + wxc * read(img, iyc, ixc).astype(f32)
            val = (yc - sy)[..., None] * v_yf + (sy - yf)[..., None] * v_yc
            res = _trunc_u8(val.astype(f32))
        else:
            raise ValueError("unknown interpolation %r" % (interpolation,))
        res = np.where((proj == 0)[..., None], fill, res)
        out[b] = res
    return out


def _signed(value, negate):
    """_randomly_negate_value (image_augmentations.py:52-56) with the coin given;
    the result is a float32 tensor in the reference."""
    return f32(-value) if negate else f32(value)


def shear_x_coeffs(level, negate):
    lv = _signed(level, negate)  # image_augmentations.py:334
    return np.array([1.0, lv, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0], dtype=f32)  # :337


def shear_y_coeffs(level, negate):
    lv = _signed(level, negate)  # :377
    return np.array([1.0, 0.0, 0.0, lv, 1.0, 0.0, 0.0, 0.0], dtype=f32)  # :380


def translate_x_coeffs(pixels, negate):
    px = _signed(pixels, negate)  # :420
    # tfa.image.translate(x, [-px, 0]) -> transform [1, 0, -dx, 0, 1, -dy, 0, 0], dx = -px
    return np.array([1.0, 0.0, -(-px), 0.0, 1.0, -f32(0.0), 0.0, 0.0], dtype=f32)  # :421-427


def translate_y_coeffs(pixels, negate):
    px = _signed(pixels, negate)  # :463
    return np.array([1.0, 0.0, -f32(0.0), 0.0, 1.0, -(-px), 0.0, 0.0], dtype=f32)  # :464-470


def rotate_coeffs(degrees, negate, H, W):
    """Rotate: radians in Python double (image_augmentations.py:135), sign flip to
    float32 (:139), then tfa angles_to_projective_transforms in stepwise float32."""
    radians = degrees * math.pi / 180.0
    ang = _signed(radians, negate)
    c = f32(math.cos(float(ang)))  # switch ROTATE_TRIG: correctly rounded f32 of f32 arg
    s = f32(math.sin(float(ang)))
    w1 = f32(W) - f32(1.0)
    h1 = f32(H) - f32(1.0)
    x_off = (w1 - (c * w1 - s * h1)) / f32(2.0)
    y_off = (h1 - (s * w1 + c * h1)) / f32(2.0)
    return np.array([c, -s, x_off, s, c, y_off, 0.0, 0.0], dtype=f32)


def _geo(images, coeff_fn, negate, interpolation, fill_mode, fill_value):
    images = _check(images)
    B = images.shape[0]
    neg = np.broadcast_to(np.asarray(negate, dtype=bool), (B,))
    coeffs = np.stack([coeff_fn(bool(n)) for n in neg])
    return projective_transform(images, coeffs, interpolation, fill_mode, fill_value)


def shear_x(images, level, negate=False, interpolation="nearest", fill_mode="constant", fill_value=0.0):
    """ShearX.call, image_augmentations.py:333-342."""
    return _geo(images, lambda n: shear_x_coeffs(level, n), negate, interpolation, fill_mode, fill_value)


def shear_y(images, level, negate=False, interpolation="nearest", fill_mode="constant", fill_value=0.0):
    """ShearY.call, image_augmentations.py:376-385."""
    return _geo(images, lambda n: shear_y_coeffs(level, n), negate, interpolation, fill_mode, fill_value)


def translate_x(images, pixels, negate=False, interpolation="nearest", fill_mode="constant", fill_value=0.0):
    """TranslateX.call, image_augmentations.py:419-428."""
    return _geo(images, lambda n: translate_x_coeffs(pixels, n), negate, interpolation, fill_mode, fill_value)


def translate_y(images, pixels, negate=False, interpolation="nearest", fill_mode="constant", fill_value=0.0):
    """TranslateY.call, image_augmentations.py:462-471."""
    return _geo(images, lambda n: translate_y_coeffs(pixels, n), negate, interpolation, fill_mode, fill_value)


def rotate(images, degrees, negate=False, interpolation="nearest", fill_mode="constant", fill_value=0.0):
    """Rotate.call, image_augmentations.py:138-147."""
    images = _check(images)
    H, W = images.shape[1:3]
    return _geo(images, lambda n: rotate_coeffs(degrees, n, H, W), negate, interpolation, fill_mode, fill_value)


# ------------------------------------------------------------------- dispatcher
def apply_op(images, name, params, negate=False, centers=None, batch_pixels=None):
    """Apply one op by layer name with constructor kwargs ``params``.

    ``negate``/``centers`` carry the op's own random draws.  ``batch_pixels`` is
    unused here (Contrast derives it from ``images``); it is accepted so callers
    can document what they intend."""
    p = dict(params)
    geo = {k: p[k] for k in ("interpolation", "fill_mode", "fill_value") if k in p}
    if name == "AutoContrast":
        return autocontrast(images)
    if name == "Equalize":
        return equalize(images)
    if name == "Invert":
        return invert(images)
    if name == "Brightness":
        return brightness(images, p["factor"])
    if name == "Contrast":
        return contrast(images, p["factor"])
    if name == "Color":
        return color(images, p["factor"])
    if name == "Sharpness":
        return sharpness(images, p["factor"])
    if name == "ShearX":
        return shear_x(images, p["level"], negate, **geo)
    if name == "ShearY":
        return shear_y(images, p["level"], negate, **geo)
    if name == "TranslateX":
        return translate_x(images, p["pixels"], negate, **geo)
    if name == "TranslateY":
        return translate_y(images, p["pixels"], negate, **geo)
    if name == "Posterize":
        return posterize(images, p["bits"])
    if name == "Solarize":
        return solarize(images, p.get("threshold", 128))
    if name == "SolarizeAdd":
        return solarize_add(images, p.get("addition", 0), p.get("threshold", 128))
    if name == "CutOut":
        return cutout(images, p["mask_size"], p.get("constant_values", 0), centers)
    if name == "Rotate":
        return rotate(images, p["degrees"], negate, **geo)
    raise ValueError("unknown op %r" % (name,))
