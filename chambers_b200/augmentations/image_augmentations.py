"""The 16 policy op layers plus ``RandomChance`` / ``RandomChoice`` with the reference's names,
constructor arguments, ``get_config`` keys and error behaviour
(/root/reference/chambers/augmentations/image_augmentations.py, cited per class), backed by the
fused sm_100a kernel instead of TensorFlow ops.

A layer here holds only its constructor arguments; ``_fill_op`` writes them into the ``chb_op``
struct of the C ABI and every pixel is produced by ``libchambers_aug.so``.
"""

from .. import _lib
from .base import build_policy, Layer, register, serialize, deserialize, run_policy, check_uint8

_INTERP = _lib.INTERPOLATIONS
_FILL = _lib.FILL_MODES


class _OpLayer(Layer):
    """An op layer called directly behaves like RandomChoice([op], 1, elementwise=False): one
    sign flip per call for the whole batch (_randomly_negate_value, :52-56), CutOut centres per
    image (tfa.image.random_cutout draws [B] centres)."""

    _kind = None
    _keys = ()

    def call(self, inputs, seed=None, call_counter=None, replay=None, record=False, batch_total=None,
             image_index_base=0, **kwargs):
        seed, call_counter = self._stream(seed, call_counter)
        out, sched = run_policy(inputs, [[(self, None)]], 1, False, seed, call_counter,
                                batch_total=batch_total, image_index_base=image_index_base,
                                replay=replay, record=record)
        self.last_schedule = sched
        return out

    def get_config(self):
        config = {k: getattr(self, k) for k in self._keys}
        return dict(list(super().get_config().items()) + list(config.items()))

    def _fill_op(self, op):
        op.kind = _lib.OP_KIND[self._kind]
        op.interpolation = 0
        op.fill_mode = 0
        op.ivalue[0] = 0
        op.ivalue[1] = 0
        op.fill_value = 0.0
        op.value = 0.0


class _GeometricLayer(_OpLayer):
    _value_key = None

    def __init__(self, value, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        setattr(self, self._value_key, value)
        self.interpolation = interpolation
        self.fill_mode = fill_mode
        self.fill_value = fill_value

    def _fill_op(self, op):
        super()._fill_op(op)
        if str(self.interpolation).lower() not in _INTERP:
            raise ValueError("Unknown interpolation %r (nearest | bilinear)" % (self.interpolation,))
        if str(self.fill_mode).lower() not in _FILL:
            raise ValueError("Unknown fill_mode %r (constant | reflect | wrap | nearest)" % (self.fill_mode,))
        op.interpolation = _INTERP[str(self.interpolation).lower()]
        op.fill_mode = _FILL[str(self.fill_mode).lower()]
        op.fill_value = float(self.fill_value)
        op.value = float(getattr(self, self._value_key))


class _FactorLayer(_OpLayer):
    _keys = ("factor",)

    def __init__(self, factor, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.factor = factor

    def _fill_op(self, op):
        super()._fill_op(op)
        op.value = float(self.factor)


# ---------------------------------------------------------------- ops without parameters
@register
class AutoContrast(_OpLayer):
    """image_augmentations.py:62-90."""
    _kind = "AutoContrast"

    def __init__(self, name=None, **kwargs):
        super().__init__(name=name, **kwargs)


@register
class Equalize(_OpLayer):
    """image_augmentations.py:93-103 (tfa.image.equalize)."""
    _kind = "Equalize"

    def __init__(self, name=None, **kwargs):
        super().__init__(name=name, **kwargs)


@register
class Invert(_OpLayer):
    """image_augmentations.py:106-116."""
    _kind = "Invert"

    def __init__(self, name=None, **kwargs):
        super().__init__(name=name, **kwargs)


# ------------------------------------------------------------------------- geometric ops
@register
class Rotate(_GeometricLayer):
    """image_augmentations.py:119-160 (tfa.image.rotate about the image centre)."""
    _kind = "Rotate"
    _value_key = "degrees"
    _keys = ("degrees", "interpolation", "fill_mode", "fill_value")

    def __init__(self, degrees, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(degrees, interpolation, fill_mode, fill_value, name=name, **kwargs)


@register
class ShearX(_GeometricLayer):
    """image_augmentations.py:315-355."""
    _kind = "ShearX"
    _value_key = "level"
    _keys = ("level", "interpolation", "fill_mode", "fill_value")

    def __init__(self, level, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(level, interpolation, fill_mode, fill_value, name=name, **kwargs)


@register
class ShearY(_GeometricLayer):
    """image_augmentations.py:358-398."""
    _kind = "ShearY"
    _value_key = "level"
    _keys = ("level", "interpolation", "fill_mode", "fill_value")

    def __init__(self, level, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(level, interpolation, fill_mode, fill_value, name=name, **kwargs)


@register
class TranslateX(_GeometricLayer):
    """image_augmentations.py:401-441."""
    _kind = "TranslateX"
    _value_key = "pixels"
    _keys = ("pixels", "interpolation", "fill_mode", "fill_value")

    def __init__(self, pixels, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(pixels, interpolation, fill_mode, fill_value, name=name, **kwargs)


@register
class TranslateY(_GeometricLayer):
    """image_augmentations.py:444-484."""
    _kind = "TranslateY"
    _value_key = "pixels"
    _keys = ("pixels", "interpolation", "fill_mode", "fill_value")

    def __init__(self, pixels, interpolation="nearest", fill_mode="constant", fill_value=0.0, name=None, **kwargs):
        super().__init__(pixels, interpolation, fill_mode, fill_value, name=name, **kwargs)


# ----------------------------------------------------------------------- integer colour ops
@register
class Posterize(_OpLayer):
    """image_augmentations.py:163-182."""
    _kind = "Posterize"
    _keys = ("bits",)

    def __init__(self, bits, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.bits = bits
        self._shift = 8 - bits

    def _fill_op(self, op):
        super()._fill_op(op)
        op.ivalue[0] = int(self.bits)


@register
class Solarize(_OpLayer):
    """image_augmentations.py:185-201."""
    _kind = "Solarize"
    _keys = ("threshold",)

    def __init__(self, threshold=128, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.threshold = threshold

    def _fill_op(self, op):
        super()._fill_op(op)
        op.ivalue[0] = int(self.threshold)


@register
class SolarizeAdd(_OpLayer):
    """image_augmentations.py:204-223."""
    _kind = "SolarizeAdd"
    _keys = ("addition", "threshold")

    def __init__(self, addition=0, threshold=128, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.addition = addition
        self.threshold = threshold

    def _fill_op(self, op):
        super()._fill_op(op)
        op.ivalue[0] = int(self.addition)
        op.ivalue[1] = int(self.threshold)


# ----------------------------------------------------------------------------- blend ops
@register
class Color(_FactorLayer):
    """image_augmentations.py:226-243."""
    _kind = "Color"


@register
class Contrast(_FactorLayer):
    """image_augmentations.py:246-273 (including the sum(hist)/256 quirk)."""
    _kind = "Contrast"


@register
class Brightness(_FactorLayer):
    """image_augmentations.py:276-293."""
    _kind = "Brightness"


@register
class Sharpness(_FactorLayer):
    """image_augmentations.py:296-312 (tfa.image.sharpness)."""
    _kind = "Sharpness"


@register
class CutOut(_OpLayer):
    """image_augmentations.py:487-507 (tfa.image.random_cutout)."""
    _kind = "CutOut"
    _keys = ("mask_size", "constant_values")

    def __init__(self, mask_size, constant_values=0, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.mask_size = mask_size
        self.constant_values = constant_values

    def _fill_op(self, op):
        super()._fill_op(op)
        if int(self.mask_size) % 2 != 0:
            raise ValueError("mask_size should be divisible by 2")
        op.ivalue[0] = int(self.mask_size)
        op.ivalue[1] = int(self.constant_values)


# ------------------------------------------------------------------------ control layers
class Sequential(Layer):
    """Minimal tf.keras.Sequential stand-in for composing RandomChance / op layers inside a
    RandomChoice (AutoAugment's sub-policies, augmentation_schemes.py:139-146)."""

    def __init__(self, layers=None, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.layers = list(layers or [])

    def call(self, inputs, **kwargs):
        x = inputs
        for layer in self.layers:
            x = layer(x)
        return x

    def get_config(self):
        config = {"layers": [serialize(l) for l in self.layers]}
        return dict(list(super().get_config().items()) + list(config.items()))

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["layers"] = [deserialize(l) for l in config["layers"]]
        return cls(**config)


register(Sequential)


def _flatten_transform(t):
    """A RandomChoice entry -> list of (op layer, probability or None)."""
    if isinstance(t, _OpLayer):
        return [(t, None)]
    if isinstance(t, RandomChance):
        if not isinstance(t.transform, _OpLayer):
            raise ValueError("RandomChance inside a fused policy must wrap one of the 16 op layers, got %r"
                             % type(t.transform).__name__)
        return [(t.transform, t.probability)]
    if isinstance(t, Sequential):
        out = []
        for l in t.layers:
            out.extend(_flatten_transform(l))
        return out
    raise ValueError("Unsupported transform %r: expected an op layer, RandomChance or Sequential of those"
                     % type(t).__name__)


@register
class RandomChance(Layer):
    """image_augmentations.py:513-545: apply ``transform`` with the given probability (one coin
    per call, i.e. per batch, exactly like tf.random.uniform([]) < p at :523)."""

    def __init__(self, transform, probability, name=None, **kwargs):
        if name is None and getattr(transform, "name", None) is not None:
            name = "random_chance_" + transform.name
        super().__init__(name=name, **kwargs)
        self.transform = transform
        self.probability = probability

    def call(self, inputs, seed=None, call_counter=None, replay=None, record=False, batch_total=None,
             image_index_base=0, **kwargs):
        seed, call_counter = self._stream(seed, call_counter)
        out, sched = run_policy(inputs, [_flatten_transform(self)], 1, False, seed, call_counter,
                                batch_total=batch_total, image_index_base=image_index_base,
                                replay=replay, record=record)
        self.last_schedule = sched
        return out

    def compute_output_shape(self, input_shape):
        return self.transform.compute_output_shape(input_shape)

    def get_config(self):
        config = {"transform": serialize(self.transform), "probability": self.probability}
        return dict(list(super().get_config().items()) + list(config.items()))

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["transform"] = deserialize(config["transform"])
        return cls(**config)


@register
class RandomChoice(Layer):
    """image_augmentations.py:548-617: ``n_transforms`` draws with replacement from ``transforms``;
    one schedule for the whole batch (elementwise=False, :569) or one per image (:565-567)."""

    def __init__(self, transforms, n_transforms, elementwise=False, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.transforms = list(transforms)
        self.n_transforms = n_transforms
        self.elementwise = elementwise

    def call(self, inputs, seed=None, call_counter=None, replay=None, record=False, batch_total=None,
             image_index_base=0, out=None, normalize=None, **kwargs):
        seed, call_counter = self._stream(seed, call_counter)
        # the C structs of the policy are built once per (transforms, n, elementwise): per call this
        # costs tens of microseconds of Python, more than the kernels of a small batch
        key = (tuple(id(t) for t in self.transforms), self.n_transforms, bool(self.elementwise))
        cached = getattr(self, "_built_policy", None)
        if cached is None or cached[0] != key:
            table = [_flatten_transform(t) for t in self.transforms]
            cached = (key, table, build_policy(table, self.n_transforms, self.elementwise))
            self._built_policy = cached
        res, sched = run_policy(inputs, cached[1], self.n_transforms, self.elementwise, seed, call_counter,
                                batch_total=batch_total, image_index_base=image_index_base,
                                replay=replay, record=record, out=out, built=cached[2], normalize=normalize)
        self.last_schedule = sched
        return res

    def compute_output_shape(self, input_shape):
        # every transform preserves the shape; the reference's np.float-based merge (:572-586) would
        # return the same list for shape-preserving transforms (and crashes on numpy >= 1.24).
        return input_shape

    def get_config(self):
        config = {
            "transforms": [serialize(t) for t in self.transforms],
            "n_transforms": self.n_transforms,
            "elementwise": self.elementwise,
        }
        return dict(list(super().get_config().items()) + list(config.items()))

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["transforms"] = [deserialize(t) for t in config["transforms"]]
        return cls(**config)


# ------------------------------------------------------------------ neighbours of the policy path
def _device_call(inputs, out_dtype_of, out_shape, launch):
    """Shared plumbing of the two front-end layers: torch CUDA tensors run stream-ordered on their
    device; numpy / CPU torch inputs are copied to the current device and back (synchronous).  There
    is no CPU implementation."""
    import numpy as np
    from .base import torch
    if torch is None:
        raise _lib.ChambersAugError(4, "torch is required as the device-memory provider")
    is_torch = isinstance(inputs, torch.Tensor)
    x = inputs if is_torch else torch.from_numpy(np.ascontiguousarray(inputs))
    if x.dtype not in (torch.uint8, torch.float32):
        x = x.to(torch.float32)  # the reference casts whatever it is given to float32
    on_device = x.is_cuda
    if not on_device:
        if not torch.cuda.is_available():
            _lib.context(0)  # raises with the real reason (no device / wrong arch)
        x = x.cuda(non_blocking=True)
    x = x.contiguous()
    dev = x.device.index if x.device.index is not None else torch.cuda.current_device()
    out = torch.empty(out_shape, dtype=out_dtype_of(x), device=x.device)
    with torch.cuda.device(dev):
        ctx = _lib.context(dev)
        _lib.check(ctx, launch(ctx, x, out, torch.cuda.current_stream(dev).cuda_stream))
    if on_device:
        return out
    res = out.cpu()
    return res if is_torch else res.numpy()


@register
class ImageNetNormalization(Layer):
    """image_augmentations.py:620-682: float32 ImageNet input normalisation, modes "caffe" (RGB->BGR,
    minus the BGR means), "tf" (x / 127.5 - 1) and "torch" ((x / 255 - mean) / std)."""

    def __init__(self, mode="caffe", name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        if mode not in {"caffe", "tf", "torch"}:
            raise ValueError("Unknown mode " + str(mode))
        self.mode = mode

    def call(self, inputs, **kwargs):
        from .base import torch
        shape = tuple(int(d) for d in inputs.shape)
        n = 1
        for d in shape:
            n *= d
        lib = _lib.load()

        def launch(ctx, x, out, stream):
            return lib.chb_imagenet_normalize(ctx, x.data_ptr(), 1 if x.dtype == torch.float32 else 0, out.data_ptr(),
                                              n, shape[3], _lib.NORM_MODES[self.mode], stream)
        return _device_call(inputs, lambda x: torch.float32, shape, launch)

    def get_config(self):
        return dict(list(super().get_config().items()) + [("mode", self.mode)])


@register
class ResizingMinMax(Layer):
    """image_augmentations.py:685-748: resize so that the smallest side equals ``min_side`` or the
    largest equals ``max_side`` (whichever scales down more when both are given), keeping the aspect
    ratio.  Bilinear output is float32 (tf.image.resize), nearest keeps the input type."""

    def __init__(self, min_side=None, max_side=None, interpolation="bilinear", name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        if min_side is None and max_side is None:
            raise ValueError("Must specify either 'min_side' or 'max_side'.")
        self.min_side = min_side
        self.max_side = max_side
        self.interpolation = interpolation

    def call(self, inputs, **kwargs):
        from .base import torch
        if str(self.interpolation).lower() not in ("bilinear", "nearest"):
            raise ValueError("Unsupported interpolation %r (bilinear | nearest)" % (self.interpolation,))
        nearest = str(self.interpolation).lower() == "nearest"
        B, H, W, C = (int(d) for d in inputs.shape)
        oh, ow = _lib.resize_min_max_shape(H, W, self.min_side, self.max_side)
        lib = _lib.load()

        def launch(ctx, x, out, stream):
            return lib.chb_resize(ctx, x.data_ptr(), 1 if x.dtype == torch.float32 else 0, out.data_ptr(), B, H, W, C,
                                  oh, ow, 1 if nearest else 0, stream)
        return _device_call(inputs, (lambda x: x.dtype) if nearest else (lambda x: torch.float32), (B, oh, ow, C), launch)

    def compute_output_shape(self, input_shape):
        return [input_shape[0], self.min_side, self.max_side, input_shape[3]]  # as the reference states it (:737-738)

    def get_config(self):
        config = [("min_side", self.min_side), ("max_side", self.max_side), ("interpolation", self.interpolation)]
        return dict(list(super().get_config().items()) + config)
