"""Drop-in surface of ``chambers.augmentations`` for the RandAugment / AutoAugment hot path
(export list of /root/reference/chambers/augmentations/__init__.py:14-39, including the two chambers
layers either side of the path, ``ImageNetNormalization`` and ``ResizingMinMax``).

The stock Keras preprocessing re-exports of the reference (``RandomRotation`` ... ``CenterCrop``,
``__init__.py:1-13``) are Keras code, not chambers code, and are out of scope.
"""

from .base import (  # noqa: F401
    Layer, serialize, deserialize, set_random_seed, set_learning_phase, learning_phase,
)
from .image_augmentations import (  # noqa: F401
    RandomChoice,
    RandomChance,
    Sequential,
    AutoContrast,
    Equalize,
    Invert,
    Rotate,
    Posterize,
    Solarize,
    SolarizeAdd,
    Color,
    Contrast,
    Brightness,
    Sharpness,
    ShearX,
    ShearY,
    TranslateX,
    TranslateY,
    CutOut,
    ImageNetNormalization,
    ResizingMinMax,
)
from .augmentation_schemes import (  # noqa: F401
    AutoAugment,
    RandAugment,
)
