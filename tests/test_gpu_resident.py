"""GPU parity tests of the image-resident engine (chb_resident.cuh), forced with chb_set_engine so a
silent fall-back to the tile engine cannot hide it: against the oracle on identical schedules, and bit
for bit against the tile engine of the same library (two independent executions of the same lazy
chain state)."""

import numpy as np
import pytest

import oracle
from conftest import random_images
from test_gpu_parity import assert_same, policy_of, to_gpu, _replay_all_pairs

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def A():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from chambers_b200 import build, augmentations
    build.build_library()
    return augmentations


class engine:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        from chambers_b200 import _lib
        self.dev = torch.cuda.current_device()
        _lib.set_engine(self.dev, self.name)

    def __exit__(self, *a):
        from chambers_b200 import _lib
        _lib.set_engine(self.dev, "auto")


def both_engines(layer, xg, **kw):
    from chambers_b200 import _lib
    with engine("resident"):
        r = layer(xg, **kw)
        assert _lib.last_engine(torch.cuda.current_device()) == "resident"
    with engine("tiles"):
        t = layer(xg, **kw)
        assert _lib.last_engine(torch.cuda.current_device()) == "tiles"
    torch.cuda.synchronize()
    return r, t


def test_auto_engine_choice(A):
    from chambers_b200 import _lib
    dev = torch.cuda.current_device()
    ra = A.RandAugment(2, 10, elementwise=True)
    ra(to_gpu(random_images(4, 224, 224, 3)), training=True)
    assert _lib.last_engine(dev) == "resident"          # every BASELINE 224 x 224 x 3 configuration
    ra(to_gpu(random_images(2, 512, 512, 3)), training=True)
    assert _lib.last_engine(dev) == "tiles"             # 768 KB does not fit an SM
    ra(to_gpu(random_images(2, 37, 53, 3)), training=True)
    assert _lib.last_engine(dev) == "tiles"             # rows are not whole 16-byte units
    A.Rotate(10.0, interpolation="bilinear")(to_gpu(random_images(2, 64, 64, 3)))
    assert _lib.last_engine(dev) == "tiles"
    A.Rotate(10.0, fill_mode="reflect")(to_gpu(random_images(2, 64, 64, 3)))
    assert _lib.last_engine(dev) == "tiles"
    with engine("resident"):
        with pytest.raises(_lib.ChambersAugError):
            ra(to_gpu(random_images(2, 37, 53, 3)), training=True)


@pytest.mark.parametrize("magnitude,shape", [(10, (224, 224)), (15, (224, 224)), (10, (64, 64)), (3, (96, 160)), (10, (16, 16)),
                                             (10, (250, 256)), (15, (2, 32)), (10, (1, 48))])
def test_all_256_op_pairs(A, magnitude, shape):
    """Every ordered pair of the 16 ops: resident == tiles everywhere, and both == oracle on a sample
    (all 256 for the small shapes)."""
    H, W = shape
    s, rng = _replay_all_pairs()
    s[..., 3] = rng.integers(0, H, size=s.shape[:3])
    s[..., 4] = rng.integers(0, W, size=s.shape[:3])
    x = random_images(256, H, W, 3, seed=H * 7 + W, kind="smooth" if magnitude == 3 else "uniform")
    xg = to_gpu(x)
    layer = A.RandAugment(2, magnitude, elementwise=True)
    r, t = both_engines(layer, xg, training=True, replay=s)
    same = (r == t).flatten(1).all(dim=1).cpu().numpy()
    bad = [(oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]) for b in np.nonzero(~same)[0]]
    assert not bad, "resident != tiles for pairs %r" % (bad[:16],)
    idx = list(range(256)) if H * W <= 64 * 64 else list(range(5, 256, 23))
    want = oracle.apply_schedule(x[idx], policy_of(layer), s[idx], elementwise=True)
    got = r[idx].cpu().numpy()
    for k, b in enumerate(idx):
        assert_same(got[k], want[k], "pair %s -> %s" % (oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]))


def test_triples_and_long_chains(A):
    for n, B, shape in ((3, 768, (48, 64)), (6, 192, (64, 80)), (4, 64, (224, 224))):
        x = random_images(B, shape[0], shape[1], 3, seed=n, kind="smooth" if n == 6 else "uniform")
        xg = to_gpu(x)
        layer = A.RandAugment(n, 15 if n == 3 else 10, elementwise=True)
        with engine("resident"):
            y = layer(xg, training=True, seed=3, call_counter=n, record=True)
        sched = layer.last_schedule
        with engine("tiles"):
            t = layer(xg, training=True, replay=sched)
        bad = np.nonzero(~(y == t).flatten(1).all(dim=1).cpu().numpy())[0]
        assert bad.size == 0, "resident != tiles for chains %r" % ([[oracle.OP_NAMES[i] for i in sched[b, :, 0, 0]] for b in bad[:6]],)
        idx = list(range(B)) if shape[0] < 100 else list(range(0, B, 7))
        want = oracle.apply_schedule(x[idx], policy_of(layer), sched[idx], elementwise=True)
        got = y[idx].cpu().numpy()
        for k, b in enumerate(idx):
            assert_same(got[k], want[k], "chain %r" % ([oracle.OP_NAMES[i] for i in sched[b, :, 0, 0]],))


def test_autoaugment_and_batch_mode(A, c1_batch):
    x = np.concatenate([c1_batch[:25], c1_batch[:25]])
    s = np.zeros((50, 1, 2, 5), np.int32)
    s[:, 0, :, 0] = np.tile(np.arange(25), 2)[:, None]
    s[..., 1] = 1
    s[:25, :, :, 2] = 1
    layer = A.AutoAugment(elementwise=True)
    with engine("resident"):
        y = layer(to_gpu(x), training=True, replay=s).cpu().numpy()
    want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
    for b in range(50):
        assert_same(y[b], want[b], "sub-policy %d" % s[b, 0, 0, 0])
    for batch_layer in (A.AutoAugment(), A.RandAugment(2, 10)):
        for call in range(6):
            with engine("resident"):
                y = batch_layer(to_gpu(c1_batch), training=True, seed=4, call_counter=call, record=True).cpu().numpy()
            want = oracle.apply_schedule(c1_batch, policy_of(batch_layer), batch_layer.last_schedule, False)
            assert_same(y, want, "%s batch mode call %d" % (type(batch_layer).__name__, call))


@pytest.mark.parametrize("C,W", [(1, 48), (2, 40), (4, 36)])
def test_other_channel_counts(A, C, W):
    x = random_images(80, 30, W, C, seed=C)
    names = [n for n in oracle.OP_NAMES if n not in ("Color", "Contrast")]
    layers = [getattr(A, n)(**oracle.magnitude_kwargs(n, 10)) for n in names]
    choice = A.RandomChoice(layers, 3, elementwise=True)
    with engine("resident"):
        y = choice(to_gpu(x), seed=1, call_counter=0, record=True).cpu().numpy()
    want = oracle.apply_schedule(x, policy_of(choice), choice.last_schedule, elementwise=True)
    for b in range(x.shape[0]):
        assert_same(y[b], want[b], "C=%d chain %r" % (C, [names[i] for i in choice.last_schedule[b, :, 0, 0]]))


def test_identity_branches_and_single_ops(A):
    """Constant images (Equalize step == 0, AutoContrast hi == lo), tiny images (fewer than 255 pixels),
    standalone op layers (batch mode: one sign flip per call, CutOut centres per image)."""
    for shape, kind in (((3, 224, 224, 3), "constant"), ((4, 9, 16, 3), "uniform"), ((3, 64, 48, 3), "lowentropy"), ((2, 224, 224, 3), "smooth")):
        x = random_images(*shape, seed=13, kind=kind)
        if kind == "constant":
            x[1] = 255 - x[0]
            x[2] = 0
        for name in oracle.OP_NAMES:
            layer = getattr(A, name)(**oracle.magnitude_kwargs(name, 10))
            for call in range(2):
                with engine("resident"):
                    y = layer(to_gpu(x), seed=99, call_counter=call, record=True).cpu().numpy()
                want = oracle.apply_schedule(x, policy_of(layer), layer.last_schedule, elementwise=False)
                assert_same(y, want, "%s %s call %d" % (name, kind, call))


def test_in_place_sharding_determinism(A):
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 256, (1500, 224, 224, 3), dtype=torch.uint8, generator=g)
    xg = x.cuda()
    for layer in (A.RandAugment(2, 10, elementwise=True), A.AutoAugment(elementwise=True)):
        with engine("resident"):
            y = layer(xg, training=True, seed=5, call_counter=1, record=True)
            sched = layer.last_schedule
            y2 = layer(xg, training=True, seed=5, call_counter=1)
            assert torch.equal(y, y2), "two runs of the same call differ"
            half = layer(xg[700:], training=True, seed=5, call_counter=1, batch_total=1500, image_index_base=700)
            assert torch.equal(half, y[700:]), "shard differs from the whole batch"
            t = xg.clone()
            out = layer._transform(t, seed=5, call_counter=1, out=t)   # d_in == d_out: no temporary on this engine
            assert out.data_ptr() == t.data_ptr() and torch.equal(t, y)
            big = torch.empty(1501 * 224 * 224 * 3, dtype=torch.uint8, device="cuda")
            src = big[: 1500 * 224 * 224 * 3].view(1500, 224, 224, 3)
            dst = big[224 * 224 * 3:].view(1500, 224, 224, 3)          # partial overlap, one image ahead
            src.copy_(xg)
            layer._transform(src, seed=5, call_counter=1, out=dst)
            assert torch.equal(dst, y)
        idx = list(range(0, 1500, 97))
        want = oracle.apply_schedule(x.numpy()[idx], policy_of(layer), sched[idx], elementwise=True)
        assert_same(y[idx].cpu().numpy(), want, type(layer).__name__ + " sample of 1500")


def test_concurrent_streams_and_repeatability(A):
    x = to_gpu(random_images(400, 224, 224, 3, seed=41))
    layer = A.RandAugment(3, 10, elementwise=True)
    with engine("resident"):
        ref = [layer(x, training=True, seed=2, call_counter=c) for c in range(4)]
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        outs = {}
        for rep in range(3):
            for c in range(4):
                with torch.cuda.stream(s1 if c % 2 == 0 else s2):
                    outs[c] = layer(x, training=True, seed=2, call_counter=c)
        torch.cuda.synchronize()
        for c in range(4):
            assert torch.equal(outs[c], ref[c]), "call %d differs when run concurrently" % c
        for rep in range(10):  # back-to-back calls on one stream (programmatic dependent launch, self-cleaned counter)
            ys = [layer(x, training=True, seed=2, call_counter=c) for c in range(4)]
            torch.cuda.synchronize()
            for c in range(4):
                assert torch.equal(ys[c], ref[c])


def test_row_shift_warps(A):
    """ShearX / TranslateX / TranslateY run as shifted copies (res_gather_rowshift) wherever four pixels
    share one exact shift: integral and half-integral shifts, shifts beyond the image, levels whose
    per-row offsets cross float32 binade boundaries, a LUT in front, a histogram op behind (the COUNT
    form), three image widths.  Oracle on identical schedules, and the tile engine."""
    kw = dict(interpolation="nearest", fill_mode="constant", fill_value=128.0)
    singles = ([A.ShearX(l, **kw) for l in (0.3, 0.44999999999999996, 0.07, 1.3, 3.0, 0.0)] +
               [A.TranslateX(p, **kw) for p in (100.0, 12.5, 0.5, 1.0, 223.0, 224.0, 300.0, 7.25)] +
               [A.TranslateY(p, **kw) for p in (17.0, 12.5, 250.0, 0.0, 1.0)])
    chains = [A.Sequential([A.Brightness(1.9), A.ShearX(0.3, **kw)]), A.Sequential([A.TranslateX(33.0, **kw), A.Equalize()]),
              A.Sequential([A.ShearX(0.21, **kw), A.AutoContrast(), A.Invert()]), A.Sequential([A.Solarize(77), A.TranslateY(9.0, **kw), A.Posterize(3)]),
              A.Sequential([A.TranslateX(40.0, **kw), A.CutOut(20, 7)]), A.Sequential([A.ShearX(0.3, **kw), A.Sharpness(1.9)]),
              A.Sequential([A.TranslateY(5.0, **kw), A.Color(1.9)])]
    transforms = singles + chains
    for shape in ((16, 16), (224, 224), (100, 256)):
        B = 2 * len(transforms)
        x = random_images(B, shape[0], shape[1], 3, seed=shape[1], kind="uniform")
        s = np.zeros((B, 1, 4, 5), np.int32)
        s[:, 0, :, 0] = np.repeat(np.arange(len(transforms)), 2)[:, None]
        s[..., 1] = 1
        s[1::2, :, :, 2] = 1                       # second copy of every transform: negated sign
        s[..., 3] = 3; s[..., 4] = 5
        layer = A.RandomChoice([t if isinstance(t, A.Sequential) else A.Sequential([t]) for t in transforms], 1, elementwise=True)
        # pad every transform to K = 4 sub-op slots: the schedule rows of the missing ops are ignored
        r, t = both_engines(layer, to_gpu(x), replay=_fit(s, layer))
        same = (r == t).flatten(1).all(dim=1).cpu().numpy()
        assert same.all(), "resident != tiles for transforms %r" % (np.nonzero(~same)[0] // 2,)
        sched = _fit(s, layer)
        want = oracle.apply_schedule(x, policy_of(layer), sched, elementwise=True)
        got = r.cpu().numpy()
        for b in range(B):
            assert_same(got[b], want[b], "transform %d (%s) negate=%d shape %r" % (b // 2, type(transforms[b // 2]).__name__, b % 2, shape))


def _fit(s, layer):
    """Schedule with as many sub-op slots as the layer's widest transform."""
    from chambers_b200.augmentations.image_augmentations import _flatten_transform
    K = max(len(_flatten_transform(t)) for t in layer.transforms)
    return np.ascontiguousarray(s[:, :, :K])


def test_small_batch_claim_order(A):
    """Batches of up to 2048 images are claimed most-expensive-chain-first (plan_order): the order is an
    internal schedule, so results must not depend on it -- device coins and replay, batch sizes around
    the CTA count and the 1024 / 2048 thresholds, against the tile engine and the oracle."""
    layer = A.RandAugment(2, 10, elementwise=True)
    for B in (149, 256, 300, 1023, 1025, 2048, 2049):
        x = random_images(B, 32, 48, 3, seed=B)
        xg = to_gpu(x)
        with engine("resident"):
            y = layer(xg, training=True, seed=B, call_counter=1, record=True)
            sched = layer.last_schedule
            y2 = layer(xg, training=True, replay=sched)
        assert torch.equal(y, y2), "replay differs from device coins at B=%d" % B
        with engine("tiles"):
            t = layer(xg, training=True, replay=sched)
        assert torch.equal(y, t), "resident != tiles at B=%d" % B
        idx = list(range(0, B, max(1, B // 40)))
        want = oracle.apply_schedule(x[idx], policy_of(layer), sched[idx], elementwise=True)
        assert_same(y[idx].cpu().numpy(), want, "B=%d" % B)


def test_split_last_pass_of_small_batches(A):
    """Batches of <= 512 images cut the last pass of an expensive image into row ranges that run on
    different CTAs (plan_order): every chain class, with its parts, must still produce the oracle's bytes --
    batch sizes from 1 image (one image on up to four SMs) to the 512 threshold, all 256 pairs at
    224 x 224 with few images per call so that nearly everything is split, AutoAugment, in place
    (no split: parts would race with the source), and the tile engine as a second witness."""
    s_all, rng = _replay_all_pairs()
    H = W = 224
    s_all[..., 3] = rng.integers(0, H, size=s_all.shape[:3])
    s_all[..., 4] = rng.integers(0, W, size=s_all.shape[:3])
    x_all = random_images(256, H, W, 3, seed=77)
    layer = A.RandAugment(2, 10, elementwise=True)
    for lo, hi in ((0, 1), (1, 3), (3, 40), (40, 104), (104, 256)):
        x, s = x_all[lo:hi], np.ascontiguousarray(s_all[lo:hi])
        r, t = both_engines(layer, to_gpu(x), training=True, replay=s)
        same = (r == t).flatten(1).all(dim=1).cpu().numpy()
        bad = [(oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]) for b in np.nonzero(~same)[0]]
        assert not bad, "resident != tiles for pairs %r (images %d..%d)" % (bad[:16], lo, hi)
    idx = list(range(96, 112)) + list(range(240, 256))   # Sharpness-first and Rotate-first chains
    with engine("resident"):
        y = layer(to_gpu(x_all[idx]), training=True, replay=np.ascontiguousarray(s_all[idx])).cpu().numpy()
        t = to_gpu(x_all[idx])
        layer._transform(t, replay=np.ascontiguousarray(s_all[idx]), out=t)     # d_in == d_out
    want = oracle.apply_schedule(x_all[idx], policy_of(layer), s_all[idx], elementwise=True)
    for k, b in enumerate(idx):
        assert_same(y[k], want[k], "pair %s -> %s" % (oracle.OP_NAMES[s_all[b, 0, 0, 0]], oracle.OP_NAMES[s_all[b, 1, 0, 0]]))
    assert (t.cpu().numpy() == y).all(), "in-place call differs"
    for B in (5, 37, 150, 511, 512, 513):
        x = random_images(B, 64, 64, 3, seed=B)
        for pol in (A.RandAugment(3, 10, elementwise=True), A.AutoAugment(elementwise=True)):
            with engine("resident"):
                y = pol(to_gpu(x), training=True, seed=B, call_counter=2, record=True).cpu().numpy()
            want = oracle.apply_schedule(x, policy_of(pol), pol.last_schedule, elementwise=True)
            assert_same(y, want, "%s B=%d" % (type(pol).__name__, B))


def test_closed_form_blend_and_presence_histogram(A):
    """l1 as one blend op is evaluated arithmetically by the flat executor (Brightness / Contrast at
    factors below and above 1, Contrast's constant 196 at 224 x 224 and 255 in batch mode), and
    AutoContrast behind a monotone table takes a min / max pass instead of a histogram -- including
    the cases that must NOT take the shortcuts: a second point-wise op on top of the blend, a
    non-monotone table (Solarize, SolarizeAdd) in front of AutoContrast, Equalize right after
    AutoContrast (the presence-only histogram may not be reused), narrow-range images."""
    seq = A.Sequential
    chains = [seq([A.Brightness(f)]) for f in (1.9, 0.28, 2.8, 0.999, 1.0, 0.0)] + \
             [seq([A.Contrast(f)]) for f in (1.9, 0.46, 2.8)] + \
             [seq([A.Brightness(1.9), A.Invert()]), seq([A.Contrast(1.9), A.Brightness(0.5)]), seq([A.Invert(), A.Brightness(1.3)]),
              seq([A.AutoContrast()]), seq([A.Brightness(0.4), A.AutoContrast()]), seq([A.Invert(), A.AutoContrast()]),
              seq([A.Posterize(3), A.AutoContrast()]), seq([A.Solarize(100), A.AutoContrast()]), seq([A.SolarizeAdd(60), A.AutoContrast()]),
              seq([A.AutoContrast(), A.Equalize()]), seq([A.AutoContrast(), A.Solarize(77), A.AutoContrast()]),
              seq([A.Brightness(0.3), A.AutoContrast(), A.AutoContrast()]), seq([A.Equalize(), A.AutoContrast()]),
              seq([A.CutOut(40, 9), A.AutoContrast()]), seq([A.Color(0.3), A.AutoContrast()]), seq([A.Contrast(0.2), A.AutoContrast(), A.Brightness(1.5)])]
    for ew in (True, False):
        layer = A.RandomChoice(chains, 1, elementwise=ew)
        n = len(chains)
        x = random_images(2 * n, 224, 224, 3, seed=5)
        x[n:] = (x[n:] // 3 + 40).astype(np.uint8)               # narrow range: AutoContrast really stretches
        s = np.zeros((2 * n, 1, 3, 5), np.int32)
        s[:, 0, :, 0] = np.tile(np.arange(n), 2)[:, None]
        s[..., 1] = 1; s[..., 3] = 100; s[..., 4] = 60
        if ew:
            r, t = both_engines(layer, to_gpu(x), replay=s)
            assert torch.equal(r, t), "resident != tiles"
            want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
            got = r.cpu().numpy()
            for b in range(2 * n):
                assert_same(got[b], want[b], "chain %d (%s)" % (b % n, "/".join(type(l).__name__ for l in chains[b % n].layers)))
        else:  # batch mode: one chain for the whole batch (Contrast's constant becomes 255)
            for k in range(n):
                sk = s.copy(); sk[:, 0, :, 0] = k
                with engine("resident"):
                    got = layer(to_gpu(x[:6]), replay=sk[:6]).cpu().numpy()
                want = oracle.apply_schedule(x[:6], policy_of(layer), sk[:6], elementwise=False)
                assert_same(got, want, "batch mode chain %d" % k)


def test_pinned_host_path(A):
    """Pinned host tensors through chb_policy_apply_host (staged, chunked copies on rotating streams): same
    bytes as the device path and as pageable numpy arrays; record on one, replay on the other; output tensor
    given by the caller; 512 x 512 images (tile engine) as well."""
    x = random_images(300, 224, 224, 3, seed=12)
    xp = torch.from_numpy(x).pin_memory()
    for layer in (A.RandAugment(2, 10, elementwise=True), A.AutoAugment(elementwise=True), A.RandAugment(2, 10, elementwise=False)):
        dev = layer(to_gpu(x), training=True, seed=3, call_counter=8, image_index_base=50, batch_total=900, record=True)
        dev_sched = layer.last_schedule
        hp = layer(xp, training=True, seed=3, call_counter=8, image_index_base=50, batch_total=900, record=True)
        assert isinstance(hp, torch.Tensor) and not hp.is_cuda
        assert (layer.last_schedule == dev_sched).all()
        assert torch.equal(hp, dev.cpu()), "pinned host path != device path"
        staged = layer(x, training=True, seed=3, call_counter=8, image_index_base=50, batch_total=900)   # pageable numpy
        assert (staged == hp.numpy()).all()
        out = torch.empty_like(xp).pin_memory()
        res = layer._transform(xp, replay=dev_sched, batch_total=900, out=out)
        assert res.data_ptr() == out.data_ptr() and torch.equal(out, hp)
    big = torch.from_numpy(random_images(3, 512, 512, 3, seed=1)).pin_memory()
    ra = A.RandAugment(2, 10, elementwise=True)
    assert torch.equal(ra(big, training=True, seed=1, call_counter=1), ra(big.cuda(), training=True, seed=1, call_counter=1).cpu())


def test_fuzz_resident_shapes(A):
    """Seeded fuzz over what the resident engine's arithmetic depends on: image shape (rows of one to a few hundred
    16-byte units, 1 to 300 rows, images just below the ~180 KB that fit beside the control block and 16 KB of working space), channel count, batch size (more or fewer
    images than SMs, split and unsplit), magnitude and chain length -- device coins, the tile engine as replay
    witness, the oracle on a sample."""
    rng = np.random.default_rng(77)
    cases = [(3, 1, 16, 3, 3, 10.0), (5, 2, 32, 3, 4, 15.0), (2, 234, 256, 3, 3, 10.0), (4, 280, 208, 3, 2, 5.0), (300, 16, 16, 3, 2, 10.0),
             (7, 58, 3072, 1, 3, 10.0), (6, 90, 480, 4, 4, 12.0), (9, 84, 1064, 2, 3, 10.0), (1, 224, 224, 3, 6, 10.0)]
    for _ in range(10):
        C = int(rng.choice([1, 2, 3, 4]))
        step = {1: 16, 2: 8, 3: 16, 4: 4}[C]
        W = step * int(rng.integers(1, 20))
        H = int(rng.integers(1, min(300, 175000 // (W * C)) + 1))
        cases.append((int(rng.integers(1, 200)), H, W, C, int(rng.integers(1, 6)), float(rng.integers(0, 16))))
    for (B, H, W, C, n, m) in cases:
        x = random_images(B, H, W, C, seed=B * H + W, kind="smooth" if (H + W) % 2 else "uniform")
        if C == 3:
            layer = A.RandAugment(n, m, elementwise=True)._transform
        else:
            names = [nm for nm in oracle.OP_NAMES if nm not in ("Color", "Contrast")]
            layer = A.RandomChoice([getattr(A, nm)(**oracle.magnitude_kwargs(nm, m)) for nm in names], n, elementwise=True)
        xg = to_gpu(x)
        with engine("resident"):
            y = layer(xg, seed=77, call_counter=3, record=True)
            sched = layer.last_schedule
        with engine("tiles"):
            t = layer(xg, replay=sched)
        assert torch.equal(y, t), "resident != tiles for case %r" % ((B, H, W, C, n, m),)
        idx = list(range(0, B, max(1, B // 12)))
        want = oracle.apply_schedule(x[idx], policy_of(layer), sched[idx], elementwise=True)
        got = y[idx].cpu().numpy()
        for k, b in enumerate(idx):
            assert_same(got[k], want[k], "case %r image %d chain %r" % ((B, H, W, C, n, m), b, sched[b, :, 0, 0].tolist()))
