"""chambers_b200 -- B200-native (sm_100a) drop-in for the image-augmentation hot path of
chjort/chambers: ``chambers_b200.augmentations`` mirrors ``chambers.augmentations`` for
RandAugment / AutoAugment and their 16 op layers; the pixels are produced by
``libchambers_aug.so`` (hand-written CUDA behind the C ABI of ``include/chambers_aug.h``).
"""

from . import augmentations  # noqa: F401
from .augmentations import set_random_seed  # noqa: F401

__version__ = "0.1.0"
