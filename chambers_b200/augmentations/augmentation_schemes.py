"""``RandAugment`` and ``AutoAugment`` with the reference's constructors, magnitude semantics and
``training`` gate (/root/reference/chambers/augmentations/augmentation_schemes.py), running as ONE
fused kernel launch per call on the B200.
"""

from . import image_augmentations as ops
from .base import Layer, register, check_uint8, learning_phase

# augmentation_schemes.py:7-10
_INTERPOLATION_MODE = "nearest"
_FILL_MODE = "constant"
_FILL_VALUE = 128
_MAX_MAGNITUDE = 10.0

# (Transform, Probability, Magnitude) pairs -- the V0 policy, augmentation_schemes.py:12-39.
_AUTO_AUGMENT_POLICY_V0 = (
    (("Equalize", 0.8, None), ("ShearY", 0.8, 4)),
    (("Color", 0.4, 9), ("Equalize", 0.6, None)),
    (("Color", 0.4, 1), ("Rotate", 0.6, 8)),
    (("Solarize", 0.8, 3), ("Equalize", 0.4, 7)),
    (("Solarize", 0.4, 2), ("Solarize", 0.6, 2)),
    (("Color", 0.2, 0), ("Equalize", 0.8, None)),
    (("Equalize", 0.4, None), ("SolarizeAdd", 0.8, 3)),
    (("ShearX", 0.2, 9), ("Rotate", 0.6, 8)),
    (("Color", 0.6, 1), ("Equalize", 1.0, None)),
    (("Invert", 0.4, None), ("Rotate", 0.6, 0)),
    (("Equalize", 1.0, None), ("ShearY", 0.6, 3)),
    (("Color", 0.4, 7), ("Equalize", 0.6, None)),
    (("Posterize", 0.4, 6), ("AutoContrast", 0.4, None)),
    (("Solarize", 0.6, 8), ("Color", 0.6, 9)),
    (("Solarize", 0.2, 4), ("Rotate", 0.8, 9)),
    (("Rotate", 1.0, 7), ("TranslateY", 0.8, 9)),
    (("ShearX", 0.0, 0), ("Solarize", 0.8, 4)),
    (("ShearY", 0.8, 0), ("Color", 0.6, 4)),
    (("Color", 1.0, 0), ("Rotate", 0.6, 2)),
    (("Equalize", 0.8, None), ("Equalize", 0.0, None)),
    (("Equalize", 1.0, None), ("AutoContrast", 0.6, None)),
    (("ShearY", 0.4, 7), ("SolarizeAdd", 0.6, 7)),
    (("Posterize", 0.8, 2), ("Solarize", 0.6, 10)),
    (("Solarize", 0.6, 8), ("Equalize", 0.6, 1)),
    (("Color", 0.8, 6), ("Rotate", 0.4, 5)),
)

_RAND_AUGMENT_ORDER = (  # augmentation_schemes.py:181-198
    "AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness", "ShearX",
    "ShearY", "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd", "CutOut", "Rotate",
)

_GEOMETRIC = dict(interpolation=_INTERPOLATION_MODE, fill_mode=_FILL_MODE, fill_value=_FILL_VALUE)


def _scaled(magnitude, top):
    return magnitude / _MAX_MAGNITUDE * top


# magnitude -> constructor kwargs, augmentation_schemes.py:42-102 (same Python arithmetic, no clamp)
_MAGNITUDE_KWARGS = {
    "AutoContrast": lambda m: {},
    "Equalize": lambda m: {},
    "Invert": lambda m: {},
    "Brightness": lambda m: {"factor": _scaled(m, 1.8) + 0.1},
    "Contrast": lambda m: {"factor": _scaled(m, 1.8) + 0.1},
    "Color": lambda m: {"factor": _scaled(m, 1.8) + 0.1},
    "Sharpness": lambda m: {"factor": _scaled(m, 1.8) + 0.1},
    "ShearX": lambda m: dict(level=_scaled(m, 0.3), **_GEOMETRIC),
    "ShearY": lambda m: dict(level=_scaled(m, 0.3), **_GEOMETRIC),
    "TranslateX": lambda m: dict(pixels=_scaled(m, 100), **_GEOMETRIC),
    "TranslateY": lambda m: dict(pixels=_scaled(m, 100), **_GEOMETRIC),
    "Posterize": lambda m: {"bits": int(_scaled(m, 4))},
    "Solarize": lambda m: {"threshold": int(_scaled(m, 256))},
    "SolarizeAdd": lambda m: {"addition": int(_scaled(m, 110))},
    "Rotate": lambda m: dict(degrees=_scaled(m, 30.0), **_GEOMETRIC),
    "CutOut": lambda m: {"mask_size": int(_scaled(m, 80)), "constant_values": _FILL_VALUE},
}


def _get_transform(transform_name, magnitude):
    """augmentation_schemes.py:105-128."""
    return getattr(ops, transform_name)(**_MAGNITUDE_KWARGS[transform_name](magnitude))


class _Policy(Layer):
    """training gate shared by both policies (smart_cond, augmentation_schemes.py:152-161, :204-213):
    a falsy ``training`` returns the input object itself."""

    def call(self, inputs, training=None, **kwargs):
        if training is None:
            training = learning_phase()
        if not training:
            return inputs
        check_uint8(inputs, type(self).__name__)
        out = self._transform(inputs, **kwargs)
        self.last_schedule = self._transform.last_schedule
        return out

    def compute_output_shape(self, input_shape):
        return self._transform.compute_output_shape(input_shape)


@register
class AutoAugment(_Policy):
    """Applies a random (op, op) sub-policy of the V0 table -- to the batch, or per image with
    elementwise=True.  augmentation_schemes.py:131-171."""

    def __init__(self, elementwise=False, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.elementwise = elementwise
        self.transforms = [
            ops.Sequential([ops.RandomChance(_get_transform(t1, m1), p1),
                            ops.RandomChance(_get_transform(t2, m2), p2)])
            for (t1, p1, m1), (t2, p2, m2) in _AUTO_AUGMENT_POLICY_V0
        ]
        self._transform = ops.RandomChoice(self.transforms, n_transforms=1, elementwise=elementwise)

    def get_config(self):
        return dict(list(super().get_config().items()) + [("elementwise", self.elementwise)])


@register
class RandAugment(_Policy):
    """``n_transforms`` ops drawn uniformly with replacement from the 16, all at ``magnitude``.
    augmentation_schemes.py:174-225."""

    def __init__(self, n_transforms, magnitude, elementwise=False, name=None, **kwargs):
        super().__init__(name=name, **kwargs)
        self.n_transforms = n_transforms
        self.magnitude = magnitude
        self.elementwise = elementwise
        self.transforms = [_get_transform(t, magnitude) for t in _RAND_AUGMENT_ORDER]
        self._transform = ops.RandomChoice(self.transforms, n_transforms=n_transforms, elementwise=elementwise)

    def get_config(self):
        config = [("n_transforms", self.n_transforms), ("magnitude", self.magnitude),
                  ("elementwise", self.elementwise)]
        return dict(list(super().get_config().items()) + config)
