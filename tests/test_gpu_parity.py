"""GPU parity tests (``-m gpu``): the CUDA path, called through the Python layers -> ctypes ->
C ABI, against the oracle on identical inputs and identical schedules.

Bar (BASELINE.json north_star): bit-exact for the integer ops (posterize, solarize, invert,
equalize, cutout, nearest warps given equal coordinates) and within +-1 LSB for the float-blended /
interpolated ops.  The kernel restates the oracle's float32 step order with unfused *_rn
arithmetic, so these tests assert 0 LSB everywhere; TOL documents the contractual bar.
"""

import numpy as np
import pytest

import oracle
from conftest import random_images

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL = {"Brightness": 1, "Contrast": 1, "Color": 1, "Sharpness": 1, "AutoContrast": 1}  # contractual, LSB
ASSERT_TOL = 0  # what we actually hold ourselves to


@pytest.fixture(scope="module")
def A():
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from chambers_b200 import build, augmentations
    build.build_library()
    return augmentations


def to_gpu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def run_layer(layer, x, **kw):
    y = layer(to_gpu(x), **kw)
    torch.cuda.synchronize()
    return y.cpu().numpy()


def assert_same(got, want, what=""):
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    bad = int((diff > ASSERT_TOL).sum())
    assert bad == 0, "%s: %d mismatching bytes of %d, max |diff| %d, first at %r" % (
        what, bad, diff.size, int(diff.max()), tuple(np.argwhere(diff > ASSERT_TOL)[0]))


def policy_of(layer):
    """oracle policy (transforms, n) equivalent to a chambers_b200 layer."""
    from chambers_b200.augmentations.image_augmentations import _flatten_transform, RandomChoice, _OpLayer
    def spec(l, p):
        cfg = {k: v for k, v in l.get_config().items() if k != "name"}
        return (type(l).__name__, cfg, p)
    if isinstance(layer, RandomChoice):
        return [[spec(l, p) for l, p in _flatten_transform(t)] for t in layer.transforms], layer.n_transforms
    if hasattr(layer, "_transform"):
        return policy_of(layer._transform)
    return [[spec(l, p) for l, p in _flatten_transform(layer)]], 1


SINGLE_OPS = [
    ("AutoContrast", {}), ("Equalize", {}), ("Invert", {}),
    ("Brightness", {"factor": 1.9000000000000001}), ("Brightness", {"factor": 0.28}),
    ("Brightness", {"factor": 1.0}), ("Brightness", {"factor": 0.0}),
    ("Contrast", {"factor": 1.9000000000000001}), ("Contrast", {"factor": 0.46}),
    ("Color", {"factor": 1.9000000000000001}), ("Color", {"factor": 0.1}), ("Color", {"factor": 0.0}),
    ("Sharpness", {"factor": 1.9000000000000001}), ("Sharpness", {"factor": 0.28}), ("Sharpness", {"factor": 0.0}),
    ("ShearX", {"level": 0.3, "fill_value": 128}), ("ShearY", {"level": 0.44999999999999996, "fill_value": 128}),
    ("TranslateX", {"pixels": 100.0, "fill_value": 128}), ("TranslateY", {"pixels": 17.0, "fill_value": 3}),
    ("Posterize", {"bits": 4}), ("Posterize", {"bits": 0}), ("Posterize", {"bits": 8}),
    ("Solarize", {"threshold": 256}), ("Solarize", {"threshold": 384}), ("Solarize", {"threshold": 77}),
    ("SolarizeAdd", {"addition": 110}), ("SolarizeAdd", {"addition": -40, "threshold": 200}),
    ("CutOut", {"mask_size": 80, "constant_values": 128}), ("CutOut", {"mask_size": 0}),
    ("Rotate", {"degrees": 30.0, "fill_value": 128}), ("Rotate", {"degrees": 45.0}), ("Rotate", {"degrees": 0.0}),
]


@pytest.mark.parametrize("name,kwargs", SINGLE_OPS, ids=lambda v: str(v) if isinstance(v, str) else "-".join("%s" % x for x in v.values()))
@pytest.mark.parametrize("shape,kind", [((5, 37, 53, 3), "uniform"), ((3, 224, 224, 3), "smooth"), ((2, 96, 64, 3), "lowentropy"),
                                        # one value everywhere: Equalize's step == 0 and AutoContrast's hi == lo identity
                                        # branches (image_augmentations.py:76-78; tfa equalize `step == 0`)
                                        ((2, 224, 224, 3), "constant"), ((3, 48, 32, 3), "constant"),
                                        # fewer than 255 pixels per channel: (sum - last) // 255 == 0 whatever the content
                                        ((4, 9, 16, 3), "uniform")])
def test_single_op_layers(A, name, kwargs, shape, kind):
    x = random_images(*shape, seed=11, kind=kind)
    if kind == "constant":
        x[1] = 255 - x[0]  # a second constant value (and one image per branch of the solarize / posterize maps)
    layer = getattr(A, name)(**kwargs)
    for call in range(2):  # two calls: both sign flips / different centres are likely to occur
        y = run_layer(layer, x, seed=99, call_counter=call, record=True)
        sched = layer.last_schedule
        want = oracle.apply_schedule(x, policy_of(layer), sched, elementwise=False)
        assert_same(y, want, "%s%r call %d" % (name, kwargs, call))


def test_device_schedule_matches_oracle_philox(A):
    x = random_images(300, 32, 40, 3, seed=1)
    for layer, ew in [(A.RandAugment(3, 15, elementwise=True), True), (A.RandAugment(2, 10), False),
                      (A.AutoAugment(elementwise=True), True), (A.AutoAugment(), False)]:
        run_layer(layer, x, training=True, seed=0xDEADBEEFCAFE1234, call_counter=5, image_index_base=1000,
                  batch_total=5000, record=True)
        want = oracle.decode_schedule(policy_of(layer), 0xDEADBEEFCAFE1234, 5, 1000, 300, 32, 40, ew)
        assert (layer.last_schedule == want).all(), type(layer).__name__


def _replay_all_pairs(n_ops=16):
    pairs = [(i, j) for i in range(n_ops) for j in range(n_ops)]
    B = len(pairs)
    rng = np.random.default_rng(7)
    s = np.zeros((B, 2, 1, 5), np.int32)
    for b, (i, j) in enumerate(pairs):
        s[b, 0, 0, 0], s[b, 1, 0, 0] = i, j
    s[..., 1] = 1
    s[..., 2] = rng.integers(0, 2, size=(B, 2, 1))
    return s, rng


@pytest.mark.parametrize("magnitude,shape", [(10, (64, 64)), (15, (50, 70)), (3, (224, 224))])
def test_all_256_op_pairs_fused(A, magnitude, shape):
    """Every ordered pair of the 16 ops through the fused chain kernel (replayed schedule)."""
    H, W = shape
    s, rng = _replay_all_pairs()
    s[..., 3] = rng.integers(0, H, size=s.shape[:3])
    s[..., 4] = rng.integers(0, W, size=s.shape[:3])
    x = random_images(256, H, W, 3, seed=21, kind="smooth" if magnitude == 3 else "uniform")
    layer = A.RandAugment(2, magnitude, elementwise=True)
    y = run_layer(layer, x, training=True, replay=s)
    want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
    for b in range(256):
        assert_same(y[b], want[b], "pair %s -> %s" % (oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]))


def test_random_triples_m15(A):
    x = random_images(768, 40, 56, 3, seed=5)
    layer = A.RandAugment(3, 15, elementwise=True)
    y = run_layer(layer, x, training=True, seed=3, call_counter=0, record=True)
    want = oracle.apply_schedule(x, policy_of(layer), layer.last_schedule, elementwise=True)
    for b in range(x.shape[0]):
        assert_same(y[b], want[b], "triple %r" % ([oracle.OP_NAMES[i] for i in layer.last_schedule[b, :, 0, 0]],))


@pytest.mark.parametrize("elementwise", [False, True])
def test_config0_randaugment_on_reference_sample_data(A, c1_batch, elementwise):
    """BASELINE.json configs[0]: RandAugment N=2 M=10 on the 32x224x224x3 sample_data batch."""
    layer = A.RandAugment(2, 10, elementwise=elementwise)
    for call in range(6 if not elementwise else 2):
        y = run_layer(layer, c1_batch, training=True, seed=0, call_counter=call, record=True)
        want = oracle.apply_schedule(c1_batch, policy_of(layer), layer.last_schedule, elementwise)
        assert_same(y, want, "call %d" % call)


def test_autoaugment_every_subpolicy(A, c1_batch):
    """All 25 sub-policies with both coins forced on (replay), then the device's own coins."""
    x = np.concatenate([c1_batch[:25], c1_batch[:25]])
    s = np.zeros((50, 1, 2, 5), np.int32)
    s[:, 0, :, 0] = np.tile(np.arange(25), 2)[:, None]
    s[..., 1] = 1
    s[:25, :, :, 2] = 1
    layer = A.AutoAugment(elementwise=True)
    y = run_layer(layer, x, training=True, replay=s)
    want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
    for b in range(50):
        assert_same(y[b], want[b], "sub-policy %d" % s[b, 0, 0, 0])
    y = run_layer(layer, c1_batch, training=True, seed=17, call_counter=0, record=True)
    want = oracle.apply_schedule(c1_batch, policy_of(layer), layer.last_schedule, elementwise=True)
    assert_same(y, want, "device coins")
    batch_layer = A.AutoAugment()
    for call in range(4):
        y = run_layer(batch_layer, c1_batch[:8], training=True, seed=4, call_counter=call, record=True)
        want = oracle.apply_schedule(c1_batch[:8], policy_of(batch_layer), batch_layer.last_schedule, False)
        assert_same(y, want, "batch mode call %d" % call)


@pytest.mark.parametrize("shape", [(6, 256, 256, 3), (3, 512, 512, 3), (4, 301, 299, 3)])
def test_large_images_many_tiles(A, shape):
    """Images cut into many tiles, aligned (config 4's 512x512) and ragged (301x299: scalar executor)."""
    from chambers_b200 import _lib
    tx, ty, _, _ = _lib.tile_plan(shape[1], shape[2])
    assert tx * ty >= 16
    x = random_images(*shape, seed=9, kind="smooth")
    layer = A.RandAugment(3, 15, elementwise=True)
    for call in range(3):
        y = run_layer(layer, x, training=True, seed=8, call_counter=call, record=True)
        want = oracle.apply_schedule(x, policy_of(layer), layer.last_schedule, elementwise=True)
        assert_same(y, want, "call %d %r" % (call, layer.last_schedule[:, :, 0, 0].tolist()))


def test_large_images_all_pairs(A):
    s, rng = _replay_all_pairs()
    H, W = 244, 260
    s[..., 3] = rng.integers(0, H, size=s.shape[:3])
    s[..., 4] = rng.integers(0, W, size=s.shape[:3])
    x = random_images(256, H, W, 3, seed=2)
    layer = A.RandAugment(2, 10, elementwise=True)
    y = run_layer(layer, x, training=True, replay=s)
    want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
    for b in range(256):
        assert_same(y[b], want[b], "pair %s -> %s" % (oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]))


@pytest.mark.parametrize("C", [1, 2, 4])
def test_other_channel_counts(A, C):
    x = random_images(40, 28, 28, C, seed=C)
    names = [n for n in oracle.OP_NAMES if n not in ("Color", "Contrast")]
    layers = [getattr(A, n)(**oracle.magnitude_kwargs(n, 10)) for n in names]
    choice = A.RandomChoice(layers, 2, elementwise=True)
    y = run_layer(choice, x, seed=1, call_counter=0, record=True)
    want = oracle.apply_schedule(x, policy_of(choice), choice.last_schedule, elementwise=True)
    assert_same(y, want, "C=%d" % C)
    from chambers_b200._lib import ChambersAugError
    with pytest.raises(ChambersAugError):
        run_layer(A.Color(1.5), x)


@pytest.mark.parametrize("interpolation", ["nearest", "bilinear"])
@pytest.mark.parametrize("fill_mode", ["constant", "reflect", "wrap", "nearest"])
def test_geometric_interpolation_and_fill_modes(A, interpolation, fill_mode):
    x = random_images(3, 45, 61, 3, seed=13, kind="smooth")
    kw = dict(interpolation=interpolation, fill_mode=fill_mode, fill_value=77.0)
    for layer in (A.Rotate(33.0, **kw), A.ShearX(0.7, **kw), A.ShearY(0.21, **kw), A.TranslateX(70.0, **kw),
                  A.TranslateY(12.5, **kw)):
        for call in range(2):
            y = run_layer(layer, x, seed=5, call_counter=call, record=True)
            want = oracle.apply_schedule(x, policy_of(layer), layer.last_schedule, elementwise=False)
            assert_same(y, want, "%s %s %s" % (type(layer).__name__, interpolation, fill_mode))
    # a bilinear warp inside a chain (materialisation path)
    chain = A.RandomChoice([A.Sequential([A.Equalize(), A.Rotate(20.0, **kw), A.Sharpness(1.5), A.ShearX(0.2, **kw)])], 1,
                           elementwise=True)
    y = run_layer(chain, x, seed=2, call_counter=0, record=True)
    want = oracle.apply_schedule(x, policy_of(chain), chain.last_schedule, elementwise=True)
    assert_same(y, want, "bilinear chain")


def test_shard_invariance_and_host_path(A):
    """RNG keyed by global image index: 1 shard == 4 shards; numpy host path == device path."""
    x = random_images(64, 64, 64, 3, seed=3)
    for layer in (A.RandAugment(2, 10, elementwise=True), A.RandAugment(2, 10, elementwise=False), A.AutoAugment(True)):
        whole = run_layer(layer, x, training=True, seed=77, call_counter=3)
        parts = [run_layer(layer, x[i:i + 16], training=True, seed=77, call_counter=3, image_index_base=i, batch_total=64)
                 for i in range(0, 64, 16)]
        assert (np.concatenate(parts) == whole).all()
        host = layer(x, training=True, seed=77, call_counter=3)  # numpy in -> chb_policy_apply_host
        assert isinstance(host, np.ndarray) and (host == whole).all()
        pinned = torch.from_numpy(x).pin_memory()
        host_t = layer(pinned, training=True, seed=77, call_counter=3)
        assert (host_t.numpy() == whole).all()


def test_in_place_and_empty_batches(A):
    x = random_images(9, 32, 32, 3, seed=4)
    layer = A.RandAugment(2, 10, elementwise=True)
    want = run_layer(layer, x, training=True, seed=1, call_counter=0)
    t = to_gpu(x)
    out = layer._transform(t, seed=1, call_counter=0, out=t)
    torch.cuda.synchronize()
    assert out.data_ptr() == t.data_ptr() and (t.cpu().numpy() == want).all()
    empty = layer(torch.empty((0, 32, 32, 3), dtype=torch.uint8, device="cuda"), training=True)
    assert tuple(empty.shape) == (0, 32, 32, 3)
    with pytest.raises(ValueError):
        run_layer(A.CutOut(3), x)


def test_full_size_batches_sampled_against_oracle(A):
    """BASELINE configs[1] (256x224x224x3) and a slice of configs[4]'s 8192 sweep: sampled images are
    checked against the oracle; whole-batch size-independent properties are checked on the device."""
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, generator=g)
    xg = x.cuda()
    layer = A.RandAugment(2, 10, elementwise=True)
    y = layer(xg, training=True, seed=0, call_counter=0, record=True)
    idx = [0, 1, 17, 100, 147, 148, 200, 255]
    want = oracle.apply_schedule(x.numpy()[idx], policy_of(layer), layer.last_schedule[idx], elementwise=True)
    assert_same(y[idx].cpu().numpy(), want, "config 1 sample")
    big = torch.randint(0, 256, (2048, 224, 224, 3), dtype=torch.uint8, generator=g).cuda()
    inv = A.Invert()
    assert torch.equal(inv(inv(big)), big)                                  # involution
    post = A.Posterize(4)
    p1 = post(big)
    assert torch.equal(post(p1), p1) and int((p1 & 15).max()) == 0         # idempotent, low bits cleared
    sol = A.Solarize(256)
    assert torch.equal(sol(big), inv(big))                                  # wrapped threshold == invert
    eq = A.Equalize()(big)
    assert int((eq.amax(dim=(1, 2)) - eq.amin(dim=(1, 2))).min()) >= 250    # equalised noise spans the range
    tx = A.TranslateX(100.0, fill_value=128)
    t = tx(big, seed=1, call_counter=0, record=True)
    if tx.last_schedule[0, 0, 0, 2]:   # negated: src_x = x - 100
        assert torch.equal(t[:, :, 100:], big[:, :, :124]) and int((t[:, :, :100] != 128).sum()) == 0
    else:
        assert torch.equal(t[:, :, :124], big[:, :, 100:]) and int((t[:, :, 124:] != 128).sum()) == 0
    cut = A.CutOut(80, 128)(big, seed=2, call_counter=0)
    changed = (cut != big).any(dim=3).sum(dim=(1, 2))
    assert int(changed.max()) <= 80 * 80 and int(((cut != big) & (cut != 128)).sum()) == 0


@pytest.mark.parametrize("shape", [(224, 224), (512, 512), (96, 160)])
def test_fast_executors_match_scalar_executor(A, shape):
    """Every ordered op pair at the BASELINE image sizes: the vectorised tile executors (flat /
    gather / sharpness) against the scalar executor of the same library, bit for bit, plus a sample
    of the pairs against the oracle."""
    from chambers_b200 import _lib
    H, W = shape
    s, rng = _replay_all_pairs()
    s[..., 3] = rng.integers(0, H, size=s.shape[:3])
    s[..., 4] = rng.integers(0, W, size=s.shape[:3])
    g = torch.Generator().manual_seed(H)
    x = torch.randint(0, 256, (256, H, W, 3), dtype=torch.uint8, generator=g)
    xg = x.cuda()
    dev = torch.cuda.current_device()
    for magnitude in (10, 15):
        layer = A.RandAugment(2, magnitude, elementwise=True)
        fast = layer(xg, training=True, replay=s)
        _lib.set_debug(dev, True)
        try:
            slow = layer(xg, training=True, replay=s)
        finally:
            _lib.set_debug(dev, False)
        torch.cuda.synchronize()
        same = (fast == slow).flatten(1).all(dim=1).cpu().numpy()
        bad = [(oracle.OP_NAMES[s[b, 0, 0, 0]], oracle.OP_NAMES[s[b, 1, 0, 0]]) for b in np.nonzero(~same)[0]]
        assert not bad, "fast != scalar executor for pairs %r" % (bad[:12],)
        if H <= 224:
            idx = list(range(3, 256, 37))
            want = oracle.apply_schedule(x.numpy()[idx], policy_of(layer), s[idx], elementwise=True)
            assert_same(fast[idx].cpu().numpy(), want, "oracle sample M=%d" % magnitude)


def test_long_chains_every_pass_kind(A):
    """Chains of six ops (up to five passes per image: COUNT and WRITE_SCRATCH passes published to
    the continuation list and picked up by other CTAs of the same launch) against the oracle."""
    x = random_images(192, 64, 80, 3, seed=31, kind="smooth")
    layer = A.RandAugment(6, 10, elementwise=True)
    for call in range(2):
        y = run_layer(layer, x, training=True, seed=9, call_counter=call, record=True)
        want = oracle.apply_schedule(x, policy_of(layer), layer.last_schedule, elementwise=True)
        for b in range(x.shape[0]):
            assert_same(y[b], want[b], "chain %r" % ([oracle.OP_NAMES[i] for i in layer.last_schedule[b, :, 0, 0]],))


def test_large_batch_scheduler_paths(A):
    """4096 x 224 x 224 x 3 (BASELINE configs[2] size): multi-tile claims and multi-tile continuation
    tickets.  Sampled images against the oracle, run-to-run determinism, and shard invariance."""
    g = torch.Generator().manual_seed(7)
    x = torch.randint(0, 256, (4096, 224, 224, 3), dtype=torch.uint8, generator=g)
    xg = x.cuda()
    for layer in (A.RandAugment(2, 10, elementwise=True), A.AutoAugment(elementwise=True)):
        y = layer(xg, training=True, seed=5, call_counter=1, record=True)
        sched = layer.last_schedule
        idx = list(range(0, 4096, 171))
        want = oracle.apply_schedule(x.numpy()[idx], policy_of(layer), sched[idx], elementwise=True)
        assert_same(y[idx].cpu().numpy(), want, type(layer).__name__ + " sample of 4096")
        y2 = layer(xg, training=True, seed=5, call_counter=1)
        assert torch.equal(y, y2), "two runs of the same call differ"
        # the second half as its own shard: same pixels
        half = layer(xg[2048:], training=True, seed=5, call_counter=1, batch_total=4096, image_index_base=2048)
        assert torch.equal(half, y[2048:]), "shard differs from the whole batch"


def test_concurrent_streams(A):
    """Two streams run policy calls at the same time (their persistent kernels share the SMs, so not
    every CTA of a launch is resident): nothing may be owned by a CTA that has not started.  Results
    equal the single-stream run."""
    x = random_images(512, 224, 224, 3, seed=41)
    xg = to_gpu(x)
    layer = A.RandAugment(3, 10, elementwise=True)
    ref = [layer(xg, training=True, seed=2, call_counter=c) for c in range(4)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = {}
    for rep in range(3):
        for c in range(4):
            with torch.cuda.stream(s1 if c % 2 == 0 else s2):
                outs[c] = layer(xg, training=True, seed=2, call_counter=c)
    torch.cuda.synchronize()
    for c in range(4):
        assert torch.equal(outs[c], ref[c]), "call %d differs when run concurrently" % c


def test_repeatability_stress(A):
    """Race detector of last resort (the pass kernel schedules dynamically: claims, tickets, elections):
    the same call must produce the same bytes every time, at a small and a large batch, on the default
    stream and on a side stream, with chains of three ops (COUNT and WRITE_SCRATCH passes included)."""
    side = torch.cuda.Stream()
    for B, reps in ((192, 12), (2048, 4)):
        x = to_gpu(random_images(B, 224, 224, 3, seed=B))
        for layer in (A.RandAugment(3, 10, elementwise=True), A.AutoAugment(elementwise=True)):
            for call in range(reps):
                y0 = layer(x, training=True, seed=11, call_counter=call)
                with torch.cuda.stream(side):
                    side.wait_stream(torch.cuda.current_stream())
                    y1 = layer(x, training=True, seed=11, call_counter=call)
                torch.cuda.current_stream().wait_stream(side)
                y2 = layer(x, training=True, seed=11, call_counter=call)
                torch.cuda.synchronize()
                assert torch.equal(y0, y1) and torch.equal(y0, y2), "%s B=%d call %d is not repeatable" % (type(layer).__name__, B, call)


def test_fuzz_shapes_magnitudes_chain_lengths(A):
    """Seeded fuzz over what the scheduler's arithmetic depends on: image shape (tile counts, ragged tiles,
    rows that are / are not whole 16-byte units), channel count, batch size (claims of 1 / 2 / 4 tiles,
    more or fewer tiles than CTAs), magnitude and chain length; device coins, oracle replay."""
    rng = np.random.default_rng(2026)
    cases = [(16, 224, 224, 3, 4, 10.0), (3, 130, 192, 3, 3, 7.0), (7, 64, 80, 4, 2, 13.0), (5, 100, 36, 1, 4, 4.0),
             (9, 48, 112, 2, 3, 10.0), (2, 257, 131, 3, 2, 10.0), (640, 32, 32, 3, 2, 10.0), (1, 384, 320, 3, 5, 10.0)]
    for _ in range(6):
        cases.append((int(rng.integers(1, 12)), int(rng.integers(17, 200)), int(rng.integers(17, 200)),
                      int(rng.choice([1, 2, 3, 4])), int(rng.integers(1, 6)), float(rng.integers(0, 16))))
    for (B, H, W, C, n, m) in cases:
        x = random_images(B, H, W, C, seed=B * H + W, kind="smooth" if (H + W) % 2 else "uniform")
        if C == 3:
            layer = A.RandAugment(n, m, elementwise=True)
        else:
            names = [nm for nm in oracle.OP_NAMES if nm not in ("Color", "Contrast")]
            layer = A.RandomChoice([getattr(A, nm)(**oracle.magnitude_kwargs(nm, m)) for nm in names], n, elementwise=True)
        y = run_layer(layer, x, training=True, seed=77, call_counter=3, record=True) if C == 3 else \
            run_layer(layer, x, seed=77, call_counter=3, record=True)
        inner = layer._transform if hasattr(layer, "_transform") else layer
        sched = layer.last_schedule if layer.last_schedule is not None else inner.last_schedule
        want = oracle.apply_schedule(x, policy_of(layer), sched, elementwise=True)
        for b in range(B):
            assert_same(y[b], want[b], "case %r image %d chain %r" % ((B, H, W, C, n, m), b, [oracle.OP_NAMES[i] if C == 3 else i for i in sched[b, :, 0, 0]]))


def test_c_abi_policy_entry_points(A):
    """chb_randaugment / chb_autoaugment / chb_apply_op called through ctypes exactly as INTEGRATION.md
    section 2 binds them, against chb_policy_apply (the entry point the Python layers use)."""
    import ctypes
    from chambers_b200 import _lib
    lib = _lib.load()
    x = to_gpu(random_images(48, 64, 80, 3, seed=5, kind="smooth"))
    dev = torch.cuda.current_device()
    ctx = _lib.context(dev)
    stream = torch.cuda.current_stream().cuda_stream
    n_sched = 48 * 3 * 2 * _lib.CHB_SCHED_FIELDS

    def call(fn, *args):
        out = torch.empty_like(x)
        rec = torch.zeros(n_sched, dtype=torch.int32, device=x.device)
        _lib.check(ctx, fn(ctx, x.data_ptr(), out.data_ptr(), 48, 64, 80, 3, *args, 48, 1000, 0xABCDEF, 7, None, rec.data_ptr(), stream))
        torch.cuda.synchronize()
        return out, rec.cpu().numpy()

    for ew in (0, 1):
        got, rec = call(lib.chb_randaugment, 3, ctypes.c_double(7.0), ew)
        layer = A.RandAugment(3, 7.0, elementwise=bool(ew))
        want = layer(x, training=True, seed=0xABCDEF, call_counter=7, image_index_base=1000, batch_total=48, record=True)
        assert torch.equal(got, want), "chb_randaugment elementwise=%d" % ew
        assert (rec[:48 * 3 * 5].reshape(48, 3, 1, 5) == layer.last_schedule).all()
        got, rec = call(lib.chb_autoaugment, ew)
        layer = A.AutoAugment(elementwise=bool(ew))
        want = layer(x, training=True, seed=0xABCDEF, call_counter=7, image_index_base=1000, batch_total=48, record=True)
        assert torch.equal(got, want), "chb_autoaugment elementwise=%d" % ew
        assert (rec[:48 * 2 * 5].reshape(48, 1, 2, 5) == layer.last_schedule).all()
    for layer in (A.Rotate(25.0, fill_value=9.0), A.Equalize(), A.CutOut(20, 7), A.Sharpness(1.7), A.TranslateX(11.0)):
        op = _lib.ChbOp()
        layer._fill_op(op)
        op.probability = -1.0
        got, _ = call(lib.chb_apply_op, ctypes.byref(op))
        want = layer(x, seed=0xABCDEF, call_counter=7, image_index_base=1000, batch_total=48)
        assert torch.equal(got, want), "chb_apply_op %s" % type(layer).__name__
    # the oracle on the direct entry point as well (not only through the layers)
    got, rec = call(lib.chb_randaugment, 2, ctypes.c_double(10.0), 1)
    pol = oracle.randaugment_policy(2, 10)
    want = oracle.apply_schedule(x.cpu().numpy(), pol, rec[:48 * 2 * 5].reshape(48, 2, 1, 5), elementwise=True)
    assert_same(got.cpu().numpy(), want, "chb_randaugment vs oracle")


def test_config3_full_shape_sampled(A):
    """BASELINE.json configs[3] as specified: RandAugment(N=3, M=15) on 512 x 512 x 512 x 3 (64 tiles per image
    in the tile engine, whole-pass claims); a sample of the images against the oracle, shard invariance on the
    second half, and run-to-run determinism."""
    g = torch.Generator().manual_seed(3)
    x = torch.randint(0, 256, (512, 512, 512, 3), dtype=torch.uint8, generator=g)
    xg = x.cuda()
    layer = A.RandAugment(3, 15, elementwise=True)
    y = layer(xg, training=True, seed=15, call_counter=2, record=True)
    sched = layer.last_schedule
    idx = [0, 63, 147, 148, 255, 296, 400, 511]
    want = oracle.apply_schedule(x.numpy()[idx], policy_of(layer), sched[idx], elementwise=True)
    got = y[idx].cpu().numpy()
    for k, b in enumerate(idx):
        assert_same(got[k], want[k], "image %d chain %r" % (b, [oracle.OP_NAMES[i] for i in sched[b, :, 0, 0]]))
    assert torch.equal(layer(xg, training=True, seed=15, call_counter=2), y)
    half = layer(xg[256:], training=True, seed=15, call_counter=2, batch_total=512, image_index_base=256)
    assert torch.equal(half, y[256:])


def test_host_path_chunk_pipeline(A):
    """chb_policy_apply_host with many chunks (> 2 * streams + 1: every staging slot is reused, the
    last chunk is ragged): bit-identical to the device path, elementwise and batch mode (Contrast's
    constant and the batch-level schedule must be the same in every chunk; CutOut centres are keyed by
    the global image index), record on the host path then replay on both."""
    B = 223  # 223 x 150528 B = 33.6 MB -> 9 chunks (3840 KiB target) of 25 images, the last one of 23: every one of the 4 staging slots is reused
    x = random_images(B, 224, 224, 3, seed=77, kind="smooth")
    xg = to_gpu(x)
    kw = dict(interpolation="nearest", fill_mode="constant", fill_value=128.0)
    for ew in (True, False):
        chain = A.Sequential([A.Contrast(1.5), A.CutOut(40, 128), A.Rotate(20.0, **kw)])
        for layer in (A.RandomChoice([chain, A.Equalize(), A.Sharpness(1.9)], 2, elementwise=ew),
                      A.RandAugment(2, 10, elementwise=ew)._transform):
            dev_out = layer(xg, seed=5, call_counter=9, image_index_base=300, batch_total=1000, record=True)
            dev_sched = layer.last_schedule
            host_out = layer(x, seed=5, call_counter=9, image_index_base=300, batch_total=1000, record=True)
            host_sched = layer.last_schedule
            assert (host_sched == dev_sched).all(), "recorded schedules differ (elementwise=%s)" % ew
            assert (host_out == dev_out.cpu().numpy()).all(), "host path != device path (elementwise=%s)" % ew
            # replay what was recorded, on the host path (rows sliced per chunk) and on the device
            rep_host = layer(x, replay=host_sched, batch_total=1000)
            rep_dev = layer(xg, replay=host_sched, batch_total=1000)
            assert (rep_host == host_out).all() and torch.equal(rep_dev, dev_out), "replay differs (elementwise=%s)" % ew


def test_multi_gpu_fixed_batch_is_shard_invariant(A):
    """BASELINE.json configs[2]: AutoAugment on a FIXED 4096 x 224 x 224 x 3 batch sharded over 2 / 4 / 8
    visible GPUs -- the concatenated shards equal the 1-GPU result bit for bit (RNG keyed by global image
    index, Contrast constant by batch_total).  One process driving several devices also exercises the
    current-device guard of the C ABI."""
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least 2 visible GPUs")
    from chambers_b200.sharding import shard_bounds, shard_kwargs
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, 256, (4096, 224, 224, 3), dtype=torch.uint8, generator=g)
    for layer in (A.AutoAugment(elementwise=True), A.RandAugment(2, 10, elementwise=False)):
        whole = layer(x.cuda(0), training=True, seed=21, call_counter=4).cpu()
        for world in [w for w in (2, 4, 8) if w <= n_dev]:
            parts = []
            for r in range(world):
                a, b = shard_bounds(4096, r, world)
                parts.append(layer(x[a:b].cuda(r), training=True, seed=21, call_counter=4, **shard_kwargs(4096, r, world)))
            assert torch.cuda.current_device() == 0, "a chb_* call changed the caller's current device"
            got = torch.cat([p.cpu() for p in parts])
            assert torch.equal(got, whole), "%s: %d shards differ from 1 GPU" % (type(layer).__name__, world)
