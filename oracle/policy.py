"""Policy layer of the oracle: magnitude -> constructor kwargs, the RandAugment /
AutoAugment transform tables, and schedule replay.

Restates ``/root/reference/chambers/augmentations/augmentation_schemes.py``
(constants :7-10, V0 table :12-39, magnitude maps :42-102, _get_transform
:105-128, AutoAugment :131-171, RandAugment :174-225) and the control layers
``RandomChance`` / ``RandomChoice`` of ``image_augmentations.py:513-617``.

A *policy* here is ``(transforms, n_transforms)`` where ``transforms`` is a list
of T transforms and each transform is a list of ``(op_name, kwargs, probability)``
sub-ops; ``probability is None`` means "always, no coin drawn" (an op layer used
directly), a float means ``RandomChance`` (a coin ``u < p`` is drawn).

A *schedule* is the explicit outcome of every random draw, so that the oracle
and the CUDA path can be compared on identical choices (SURVEY.md section 8a
"RNG"): ``int32 [B, n_transforms, K, 5]`` with fields
``(choice, applied, negate, cy, cx)``, K = max sub-ops per transform.

Test infrastructure only; parity unpinned (see ``oracle/__init__.py``).
"""

import numpy as np

from . import ops

__all__ = [
    "INTERPOLATION_MODE", "FILL_MODE", "FILL_VALUE", "MAX_MAGNITUDE", "AUTO_AUGMENT_POLICY_V0",
    "magnitude_kwargs", "randaugment_policy", "autoaugment_policy", "single_op_policy",
    "apply_schedule", "policy_k", "SCHED_FIELDS",
]

INTERPOLATION_MODE = "nearest"  # augmentation_schemes.py:7
FILL_MODE = "constant"  # :8
FILL_VALUE = 128  # :9
MAX_MAGNITUDE = 10.0  # :10

# augmentation_schemes.py:12-39 -- (op, probability, magnitude) pairs.
AUTO_AUGMENT_POLICY_V0 = [
    [("Equalize", 0.8, None), ("ShearY", 0.8, 4)],
    [("Color", 0.4, 9), ("Equalize", 0.6, None)],
    [("Color", 0.4, 1), ("Rotate", 0.6, 8)],
    [("Solarize", 0.8, 3), ("Equalize", 0.4, 7)],
    [("Solarize", 0.4, 2), ("Solarize", 0.6, 2)],
    [("Color", 0.2, 0), ("Equalize", 0.8, None)],
    [("Equalize", 0.4, None), ("SolarizeAdd", 0.8, 3)],
    [("ShearX", 0.2, 9), ("Rotate", 0.6, 8)],
    [("Color", 0.6, 1), ("Equalize", 1.0, None)],
    [("Invert", 0.4, None), ("Rotate", 0.6, 0)],
    [("Equalize", 1.0, None), ("ShearY", 0.6, 3)],
    [("Color", 0.4, 7), ("Equalize", 0.6, None)],
    [("Posterize", 0.4, 6), ("AutoContrast", 0.4, None)],
    [("Solarize", 0.6, 8), ("Color", 0.6, 9)],
    [("Solarize", 0.2, 4), ("Rotate", 0.8, 9)],
    [("Rotate", 1.0, 7), ("TranslateY", 0.8, 9)],
    [("ShearX", 0.0, 0), ("Solarize", 0.8, 4)],
    [("ShearY", 0.8, 0), ("Color", 0.6, 4)],
    [("Color", 1.0, 0), ("Rotate", 0.6, 2)],
    [("Equalize", 0.8, None), ("Equalize", 0.0, None)],
    [("Equalize", 1.0, None), ("AutoContrast", 0.6, None)],
    [("ShearY", 0.4, 7), ("SolarizeAdd", 0.6, 7)],
    [("Posterize", 0.8, 2), ("Solarize", 0.6, 10)],
    [("Solarize", 0.6, 8), ("Equalize", 0.6, 1)],
    [("Color", 0.8, 6), ("Rotate", 0.4, 5)],
]

_GEO = {"interpolation": INTERPOLATION_MODE, "fill_mode": FILL_MODE, "fill_value": FILL_VALUE}


def magnitude_kwargs(name, magnitude):
    """_get_transform's magnitude -> kwargs table (augmentation_schemes.py:42-128).
    Plain Python float/int arithmetic, no clamping (M=15 is legal)."""
    m = magnitude
    if name in ("AutoContrast", "Equalize", "Invert"):
        return {}
    if name in ("Brightness", "Contrast", "Color", "Sharpness"):
        return {"factor": m / MAX_MAGNITUDE * 1.8 + 0.1}  # :43
    if name in ("ShearX", "ShearY"):
        return dict(level=m / MAX_MAGNITUDE * 0.3, **_GEO)  # :49
    if name in ("TranslateX", "TranslateY"):
        return dict(pixels=m / MAX_MAGNITUDE * 100, **_GEO)  # :60
    if name == "Posterize":
        return {"bits": int(m / MAX_MAGNITUDE * 4)}  # :71
    if name == "Solarize":
        return {"threshold": int(m / MAX_MAGNITUDE * 256)}  # :77
    if name == "SolarizeAdd":
        return {"addition": int(m / MAX_MAGNITUDE * 110)}  # :83
    if name == "Rotate":
        return dict(degrees=m / MAX_MAGNITUDE * 30.0, **_GEO)  # :89
    if name == "CutOut":
        return {"mask_size": int(m / MAX_MAGNITUDE * 80), "constant_values": FILL_VALUE}  # :100-101
    raise ValueError("unknown transform %r" % (name,))


def randaugment_policy(n_transforms, magnitude):
    """RandAugment.__init__ (augmentation_schemes.py:176-201): 16 ops in fixed
    order, each applied unconditionally once chosen."""
    transforms = [[(name, magnitude_kwargs(name, magnitude), None)] for name in ops.OP_NAMES]
    return transforms, int(n_transforms)


def autoaugment_policy():
    """AutoAugment.__init__ (augmentation_schemes.py:135-149): 25 sub-policies of
    two RandomChance-wrapped ops, one sub-policy drawn per call."""
    transforms = [
        [(t1, magnitude_kwargs(t1, m1), p1), (t2, magnitude_kwargs(t2, m2), p2)]
        for (t1, p1, m1), (t2, p2, m2) in AUTO_AUGMENT_POLICY_V0
    ]
    return transforms, 1


def single_op_policy(name, kwargs, probability=None):
    """An op layer called directly (or wrapped in one RandomChance)."""
    return [[(name, dict(kwargs), probability)]], 1


def policy_k(transforms):
    return max(len(t) for t in transforms)


SCHED_FIELDS = ("choice", "applied", "negate", "cy", "cx")


def apply_schedule(images, policy, schedule, elementwise):
    """Replay ``schedule`` through ``policy`` on ``images``.

    elementwise=True mirrors RandomChoice.call's tf.map_fn branch
    (image_augmentations.py:565-567): every op sees a ``[1, H, W, C]`` tensor.
    elementwise=False mirrors :569: every op sees the whole batch, the op
    sequence / coins / sign flips of image 0's schedule row apply to all images
    and only the CutOut centres are per image."""
    transforms, n_transforms = policy
    images = np.asarray(images)
    schedule = np.asarray(schedule)
    B = images.shape[0]
    K = policy_k(transforms)
    assert schedule.shape == (B, n_transforms, K, 5), (schedule.shape, (B, n_transforms, K, 5))
    if elementwise:
        out = np.empty_like(images)
        for b in range(B):
            x = images[b:b + 1]
            for i in range(n_transforms):
                t = transforms[int(schedule[b, i, 0, 0])]
                for j, (name, kwargs, _p) in enumerate(t):
                    _, applied, negate, cy, cx = (int(v) for v in schedule[b, i, j])
                    if applied:
                        x = ops.apply_op(x, name, kwargs, negate=bool(negate), centers=[[cy, cx]])
            out[b] = x[0]
        return out
    x = images
    for i in range(n_transforms):
        t = transforms[int(schedule[0, i, 0, 0])]
        for j, (name, kwargs, _p) in enumerate(t):
            _, applied, negate, _, _ = (int(v) for v in schedule[0, i, j])
            if applied:
                x = ops.apply_op(x, name, kwargs, negate=bool(negate), centers=schedule[:, i, j, 3:5])
    return np.array(x, copy=True)
