"""GPU tests of the two chambers layers either side of the policy path.

ImageNetNormalization is the one op of this neighbourhood whose parity is PINNED: the kernel's output
on the reference's 4 x 4 test image equals the reference's own golden vectors with exact float32
equality (test_units/augmentations/test_image_augmentations.py:21-64 -- the same assertAllEqual), and
equals the oracle bit for bit on large random batches."""

import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, random_images

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def A():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from chambers_b200 import build, augmentations
    build.build_library()
    return augmentations


@pytest.fixture(scope="module")
def ref():
    return json.load(open(os.path.join(GOLDEN, "imagenet_norm_ref.json")))


@pytest.mark.parametrize("mode", ["caffe", "tf", "torch"])
def test_imagenet_normalization_reference_golden_vectors(A, ref, mode):
    img = np.array(ref["image"], dtype=np.uint8)
    x = np.stack([img, img, img], axis=-1)[None]
    target = np.array(ref["targets"][mode], dtype=np.float32)
    layer = A.ImageNetNormalization(mode=mode)
    y = layer(torch.from_numpy(x).cuda())
    assert y.dtype == torch.float32 and tuple(y.shape) == x.shape
    assert np.array_equal(y.cpu().numpy()[0, ..., 0], target)          # exact, as the reference asserts
    assert np.array_equal(layer(x)[0, ..., 0], target)                 # numpy in -> numpy out (host path)
    yf = layer(torch.from_numpy(x.astype(np.float32)).cuda())          # float32 input: the reference casts
    assert np.array_equal(yf.cpu().numpy()[0, ..., 0], target)


@pytest.mark.parametrize("mode", ["caffe", "tf", "torch"])
def test_imagenet_normalization_matches_oracle(A, mode):
    layer = A.ImageNetNormalization(mode=mode)
    for shape in ((256, 224, 224, 3), (3, 37, 53, 3), (1, 1, 1, 3), (2, 5, 7, 3), (0, 8, 8, 3)):
        x = random_images(*shape, seed=shape[1])
        y = layer(torch.from_numpy(x).cuda()).cpu().numpy()
        want = oracle.imagenet_normalization(x, mode)
        assert y.shape == want.shape and np.array_equal(y, want), (mode, shape)
    # every byte value in every channel position, unaligned input and output views
    x = np.tile(np.arange(256, dtype=np.uint8)[:, None], (1, 3)).reshape(1, 16, 16, 3)
    big = torch.zeros(x.size + 1, dtype=torch.uint8, device="cuda")
    big[1:] = torch.from_numpy(x.reshape(-1)).cuda()
    y = layer(big[1:].view(1, 16, 16, 3)).cpu().numpy()
    assert np.array_equal(y, oracle.imagenet_normalization(x, mode))
    if mode == "tf":  # any channel count
        x1 = random_images(4, 30, 31, 1, seed=2)
        assert np.array_equal(layer(torch.from_numpy(x1).cuda()).cpu().numpy(), oracle.imagenet_normalization(x1, mode))
    else:
        from chambers_b200._lib import ChambersAugError
        with pytest.raises(ChambersAugError):
            layer(torch.zeros((1, 4, 4, 1), dtype=torch.uint8, device="cuda"))


def test_policy_then_normalization_pipeline(A):
    """The reference's call order: RandAugment -> ImageNetNormalization(mode="tf") (vision_transformer.py:655)."""
    x = random_images(64, 224, 224, 3, seed=9)
    ra = A.RandAugment(2, 10, elementwise=True)
    aug = ra(torch.from_numpy(x).cuda(), training=True, seed=1, call_counter=0, record=True)
    y = A.ImageNetNormalization(mode="tf")(aug).cpu().numpy()
    from test_gpu_parity import policy_of
    want = oracle.imagenet_normalization(oracle.apply_schedule(x, policy_of(ra), ra.last_schedule, elementwise=True), "tf")
    assert np.array_equal(y, want)


def test_resizing_min_max(A, ref):
    for name, case in ref["shapes"].items():                      # the reference's four shape tests
        x = np.zeros(case["input_shape"], np.uint8)
        out = A.ResizingMinMax(**case["kwargs"])(torch.from_numpy(x).cuda())
        assert list(out.shape) == case["output_shape"] and out.dtype == torch.float32, name
    rng = np.random.default_rng(4)
    for shape, kw in (((3, 60, 90, 3), dict(min_side=224)), ((2, 403, 196, 3), dict(max_side=224)),
                      ((2, 300, 500, 3), dict(min_side=256, max_side=384)), ((1, 17, 23, 1), dict(min_side=5)),
                      ((2, 64, 64, 4), dict(max_side=64))):
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        for interp in ("bilinear", "nearest"):
            got = A.ResizingMinMax(interpolation=interp, **kw)(torch.from_numpy(x).cuda()).cpu().numpy()
            want = oracle.resizing_min_max(x, interpolation=interp, **kw)
            assert got.shape == want.shape and got.dtype == want.dtype, (shape, kw, interp)
            assert np.array_equal(got, want), (shape, kw, interp, float(np.abs(got.astype(np.float64) - want).max()))
    xf = rng.random(size=(2, 40, 30, 3)).astype(np.float32)
    got = A.ResizingMinMax(min_side=64)(xf)                      # float32 numpy in -> numpy out
    assert isinstance(got, np.ndarray) and np.array_equal(got, oracle.resizing_min_max(xf, min_side=64))


@pytest.mark.parametrize("mode", ["tf", "torch", "caffe"])
def test_fused_normalization_epilogue(A, mode):
    """policy(x, normalize=mode) == ImageNetNormalization(mode)(policy(x)) bit for bit: every ordered op pair at
    224 x 224 (flat, gather, row-shift and Sharpness last passes all write float32 from the resident engine for
    tf / torch; caffe and the tile engine take the two-kernel fallback inside the library), a small batch that
    splits images over several CTAs, AutoAugment, numpy in / out, and the oracle on a sample."""
    from test_gpu_parity import policy_of, _replay_all_pairs
    s, rng = _replay_all_pairs()
    s[..., 3] = rng.integers(0, 224, size=s.shape[:3])
    s[..., 4] = rng.integers(0, 224, size=s.shape[:3])
    x = random_images(256, 224, 224, 3, seed=31)
    xg = torch.from_numpy(x).cuda()
    ra = A.RandAugment(2, 10, elementwise=True)
    norm = A.ImageNetNormalization(mode=mode)
    for lo, hi in ((0, 256), (96, 112), (200, 203)):
        want = norm(ra(xg[lo:hi], training=True, replay=s[lo:hi]))
        got = ra(xg[lo:hi], training=True, replay=s[lo:hi], normalize=mode)
        assert got.dtype == torch.float32 and got.shape == want.shape
        same = (got == want).flatten(1).all(dim=1).cpu().numpy()
        bad = [(oracle.OP_NAMES[s[lo + b, 0, 0, 0]], oracle.OP_NAMES[s[lo + b, 1, 0, 0]]) for b in np.nonzero(~same)[0]]
        assert not bad, "fused != separate for pairs %r" % (bad[:12],)
    idx = list(range(3, 256, 29))
    ref = oracle.imagenet_normalization(oracle.apply_schedule(x[idx], policy_of(ra), s[idx], elementwise=True), mode)
    assert np.array_equal(ra(xg[idx], training=True, replay=s[idx], normalize=mode).cpu().numpy(), ref)
    aa = A.AutoAugment(elementwise=True)
    y = aa(xg, training=True, seed=4, call_counter=2, normalize=mode, record=True)
    assert torch.equal(y, norm(aa(xg, training=True, replay=aa.last_schedule)))
    yh = ra(x[:9], training=True, seed=1, call_counter=1, normalize=mode)          # numpy in -> numpy out
    assert isinstance(yh, np.ndarray) and np.array_equal(yh, norm(ra(xg[:9], training=True, seed=1, call_counter=1)).cpu().numpy())
    big = torch.from_numpy(random_images(3, 512, 512, 3, seed=2)).cuda()             # tile engine: fallback
    assert torch.equal(ra(big, training=True, seed=2, call_counter=0, normalize=mode), norm(ra(big, training=True, seed=2, call_counter=0)))
