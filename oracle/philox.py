"""Counter-based schedule generator: Philox4x32-10 (Salmon et al., SC'11 --
the Random123 reference algorithm) plus the draw layout the CUDA path decodes
on device (``chambers_b200/csrc/chb_device.cuh``: ``chb_philox4x32_10`` and
``decode_schedule``).

The reference draws from TensorFlow's stateful global-seed stream
(``tf.random.uniform`` at image_augmentations.py:54, :523, :608-610 and inside
``tfa.image.random_cutout``); that stream cannot be reproduced without
TensorFlow, so parity is defined on an explicit schedule (SURVEY.md section 8a
"RNG").  This module restates OUR counter scheme so a device-decoded schedule
can be checked bit for bit and replayed through the oracle.

Draw layout.  key = (seed_lo, seed_hi); counter = (image_lo, image_hi,
call_counter, slot).  For draw i of a policy with K sub-ops per transform:

* slot ``i*(K+1)``          word0 -> ``choice = mulhi(word0, T)``
* slot ``i*(K+1) + 1 + j``  words (coin, sign, cy, cx) of sub-op j:
  ``applied = (coin >> 8) < ceil(p * 2**24)`` (always 1 when p is None),
  ``negate = sign < 2**31``, ``cy = mulhi(w2, H)``, ``cx = mulhi(w3, W)``.

``image`` is the GLOBAL image index (``image_index_base + b``), so a batch
sharded over 1/2/4/8 GPUs decodes the same schedule.  With elementwise=False
the choice / coin / sign words come from the batch-level stream
``image = 2**64 - 1`` and only the CutOut centres are per image.

Test infrastructure only (see ``oracle/__init__.py``).
"""

import math

import numpy as np

from .policy import policy_k

__all__ = ["philox4x32_10", "decode_schedule", "prob_threshold24", "BATCH_STREAM"]

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
BATCH_STREAM = (1 << 64) - 1


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: uint32 [..., 2] -> uint32 [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k = np.asarray(key, dtype=np.uint64).copy()
    c0, c1, c2, c3 = (c[..., i].copy() for i in range(4))
    k0 = np.broadcast_to(k[..., 0], c0.shape).copy()
    k1 = np.broadcast_to(k[..., 1], c0.shape).copy()
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + np.uint64(_W0)) & _MASK
        k1 = (k1 + np.uint64(_W1)) & _MASK
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def prob_threshold24(p):
    """RandomChance(u < p) on a 24-bit grid: apply iff (coin >> 8) < threshold."""
    if p is None:
        return 1 << 24
    t = int(math.ceil(float(p) * float(1 << 24)))
    return max(0, min(t, 1 << 24))


def _mulhi(w, n):
    return ((w.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)


def decode_schedule(policy, seed, call_counter, image_index_base, B, H, W, elementwise):
    """-> int32 [B, n_transforms, K, 5] (choice, applied, negate, cy, cx)."""
    transforms, n_transforms = policy
    T = len(transforms)
    K = policy_k(transforms)
    seed = int(seed) & ((1 << 64) - 1)
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    img = (np.arange(B, dtype=np.uint64) + np.uint64(image_index_base)).astype(np.uint64)
    own = img
    if not elementwise:
        img = np.full(B, BATCH_STREAM, dtype=np.uint64)

    def block(image_idx, slot):
        ctr = np.stack([
            image_idx & _MASK, image_idx >> np.uint64(32),
            np.full(B, int(call_counter) & 0xFFFFFFFF, dtype=np.uint64),
            np.full(B, slot, dtype=np.uint64),
        ], axis=-1)
        return philox4x32_10(ctr, key)

    thr = np.array([[prob_threshold24(t[j][2]) if j < len(t) else 0 for j in range(K)]
                    for t in transforms], dtype=np.int64)
    nsub = np.array([len(t) for t in transforms], dtype=np.int64)
    sched = np.zeros((B, n_transforms, K, 5), dtype=np.int32)
    for i in range(n_transforms):
        choice = _mulhi(block(img, i * (K + 1))[:, 0], T)
        for j in range(K):
            slot = i * (K + 1) + 1 + j
            w = block(img, slot)
            wc = block(own, slot)
            applied = ((w[:, 0] >> np.uint32(8)).astype(np.int64) < thr[choice, j]) & (j < nsub[choice])
            sched[:, i, j, 0] = choice
            sched[:, i, j, 1] = applied
            sched[:, i, j, 2] = w[:, 1] < np.uint32(0x80000000)
            sched[:, i, j, 3] = _mulhi(wc[:, 2], H)
            sched[:, i, j, 4] = _mulhi(wc[:, 3], W)
    return sched
