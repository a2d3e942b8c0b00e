"""CPU oracle for the chambers.augmentations RandAugment / AutoAugment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``chambers_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or as the CPU baseline that is timed *beside* the product), never as the
product.

PARITY UNPINNED.  The reference (chjort/chambers) holds no test, golden vector
or fixture for this path (its only augmentation tests cover
``ImageNetNormalization`` and ``ResizingMinMax``,
``test_units/augmentations/test_image_augmentations.py:21-80``), and the pixel
arithmetic of most ops lives in third-party code that is neither vendored in
``/root/reference`` nor installable here:

* ``tensorflow==2.6.0``            (``requirements.txt:2``)
* ``tensorflow-addons`` (unpinned) (``requirements.txt:3``; 0.14.x-0.16.x match TF 2.6)
* ``keras`` (unpinned)             (``requirements.txt:1``)

The in-repo arithmetic (``blend``, AutoContrast, Invert, Posterize, Solarize,
SolarizeAdd, Brightness, the Contrast quirk, the magnitude maps and the policy
tables) is restated from the reference source line by line.  The TF / TFA
delegated arithmetic (equalize, sharpness, the projective transform kernel,
random_cutout, rgb_to_grayscale) is restated from those libraries' published
algorithms; every behaviour that hinges on an upstream detail is a NAMED SWITCH
in :mod:`oracle.switches` so it can be flipped in one place once a real
TensorFlow is at hand.  What *is* pinned: the DERIVED known-answer vectors in
``tests/golden/`` (computed by hand from the in-repo formulas), and bit-exact
cross-checks of the integer ops against PIL ``ImageOps`` and torchvision, which
implement the same published algorithms.
"""

from . import switches  # noqa: F401
from .ops import *  # noqa: F401,F403
from .policy import *  # noqa: F401,F403
from .philox import *  # noqa: F401,F403
from .frontend import (  # noqa: F401
    imagenet_normalization, resizing_min_max, resizing_min_max_shape, resize_bilinear, resize_nearest,
)
