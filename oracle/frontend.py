"""CPU restatement of the two chambers layers either side of the policy path (TEST INFRASTRUCTURE
ONLY, like the rest of ``oracle/``).

``imagenet_normalization`` IS PINNED: the reference's own tests hold exact float32 golden vectors for
all three modes (``/root/reference/test_units/augmentations/test_image_augmentations.py:21-64``,
committed as ``tests/golden/imagenet_norm_ref.json`` by ``tests/golden/make_imagenet_norm_fixture.py``)
and this restatement reproduces them bit for bit (``tests/test_oracle.py``).

``resizing_min_max_shape`` is pinned by the reference's four shape tests (``:66-80``).  The resize
arithmetic itself (``tf.image.resize`` behind Keras' ``Resizing``) lives in TensorFlow, which is not
installable here: ``resize_bilinear`` / ``resize_nearest`` restate its published half-pixel-centre
algorithm and are UNPINNED.
"""

import numpy as np

f32 = np.float32


def imagenet_normalization(images, mode="caffe"):
    """ImageNetNormalization.call, image_augmentations.py:629-682.  One float32 rounding per op."""
    if mode not in ("caffe", "tf", "torch"):
        raise ValueError("Unknown mode " + str(mode))  # :624-625
    x = np.asarray(images)
    if mode == "tf":  # :660-665
        x = x.astype(f32)
        return (x / f32(127.5)) - f32(1.0)
    if mode == "torch":  # :652-657 -> _normalize :668-682
        x = x.astype(f32) / f32(255.0)
        mean = np.array([0.485, 0.456, 0.406], dtype=f32)
        std = np.array([0.229, 0.224, 0.225], dtype=f32)
        return (x - mean) / std
    x = x[..., ::-1].astype(f32)  # :648 RGB -> BGR, then _normalize's cast
    return x - np.array([103.939, 116.779, 123.68], dtype=f32)  # :649


def resizing_min_max_shape(height, width, min_side=None, max_side=None):
    """ResizingMinMax.call :711-730: float32 scale, truncating int casts."""
    if min_side is None and max_side is None:
        raise ValueError("Must specify either 'min_side' or 'max_side'.")  # :704-705
    h, w = f32(height), f32(width)
    if min_side is not None and max_side is not None:
        scale = min(f32(max_side) / max(w, h), f32(min_side) / min(w, h))
    elif min_side is not None:
        scale = f32(min_side) / min(w, h)
    else:
        scale = f32(max_side) / max(w, h)
    return int(f32(h * scale)), int(f32(w * scale))


def _taps(out_size, in_size):
    scale = f32(in_size) / f32(out_size)
    src = (np.arange(out_size, dtype=f32) + f32(0.5)) * scale - f32(0.5)
    fl = np.floor(src)
    lo = np.maximum(fl.astype(np.int64), 0)
    hi = np.minimum(np.ceil(src).astype(np.int64), in_size - 1)
    return lo, hi, (src - fl).astype(f32)


def resize_bilinear(images, out_h, out_w):
    """tf.image.resize(..., method="bilinear") with half-pixel centres, no antialiasing: float32 out."""
    x = np.asarray(images).astype(f32)
    y0, y1, yl = _taps(out_h, x.shape[1])
    x0, x1, xl = _taps(out_w, x.shape[2])
    xl = xl[None, None, :, None]
    yl = yl[None, :, None, None]
    tl, tr = x[:, y0][:, :, x0], x[:, y0][:, :, x1]
    bl, br = x[:, y1][:, :, x0], x[:, y1][:, :, x1]
    top = tl + (tr - tl) * xl
    bot = bl + (br - bl) * xl
    return (top + (bot - top) * yl).astype(f32)


def resize_nearest(images, out_h, out_w):
    x = np.asarray(images)
    hs, ws = f32(x.shape[1]) / f32(out_h), f32(x.shape[2]) / f32(out_w)
    iy = np.minimum(np.floor((np.arange(out_h, dtype=f32) + f32(0.5)) * hs).astype(np.int64), x.shape[1] - 1)
    ix = np.minimum(np.floor((np.arange(out_w, dtype=f32) + f32(0.5)) * ws).astype(np.int64), x.shape[2] - 1)
    return x[:, iy][:, :, ix]


def resizing_min_max(images, min_side=None, max_side=None, interpolation="bilinear"):
    x = np.asarray(images)
    oh, ow = resizing_min_max_shape(x.shape[1], x.shape[2], min_side, max_side)
    if interpolation == "nearest":
        return resize_nearest(x, oh, ow)
    if interpolation != "bilinear":
        raise ValueError("interpolation must be bilinear or nearest")
    return resize_bilinear(x, oh, ow)
