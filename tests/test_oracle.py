"""CPU tests of the oracle itself: it is pinned against (a) the hand-derived known-answer vectors
of tests/golden/kat_img4x4.json on the reference's own 4x4 fixture, (b) Random123's Philox4x32-10
known answers, (c) PIL / torchvision, which implement the same published integer algorithms
(equalize, posterize, solarize, invert), and (d) structural properties of each op.
The reference holds no test of this path, so nothing stronger exists.
"""

import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN, random_images


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "kat_img4x4.json")) as f:
        return json.load(f)


def _img(kat):
    ch = np.array(kat["img"], dtype=np.uint8)
    return np.stack([ch, ch, ch], axis=-1)[None]


def test_kat_integer_ops(kat):
    x = _img(kat)
    assert (oracle.invert(x)[0, ..., 0] == np.array(kat["Invert"])).all()
    assert (oracle.posterize(x, 4)[0, ..., 1] == np.array(kat["Posterize_bits4"])).all()
    assert (oracle.solarize(x, 128)[0, ..., 2] == np.array(kat["Solarize_128"])).all()
    assert (oracle.solarize_add(x, 110, 128)[0, ..., 0] == np.array(kat["SolarizeAdd_110_128"])).all()


def test_kat_float_ops(kat):
    x = _img(kat)
    f19 = 10 / 10.0 * 1.8 + 0.1  # 1.9000000000000001, augmentation_schemes.py:43
    assert (oracle.brightness(x, f19)[0, ..., 0] == np.array(kat["Brightness_1.9"])).all()
    assert (oracle.brightness(x, 0.28)[0, ..., 0] == np.array(kat["Brightness_0.28"])).all()
    assert (oracle.autocontrast(x)[0, ..., 0] == np.array(kat["AutoContrast"])).all()
    assert (oracle.equalize(x) == x).all()  # step = (16 - 1) // 255 = 0
    assert (oracle.color(x, f19) == x).all()  # grey input
    assert (oracle.contrast(x, f19) == oracle.brightness(x, f19)).all()  # constant 16/256 -> 0


def test_contrast_constant_quirk():
    # (number of pixels)/256, clipped
    assert oracle.contrast_constant(32 * 224 * 224) == 255
    assert oracle.contrast_constant(224 * 224) == 196
    assert oracle.contrast_constant(28 * 28) == 3
    assert oracle.contrast_constant(512 * 512) == 255
    assert oracle.contrast_constant(16) == 0


def test_philox_known_answers():
    f = oracle.philox4x32_10
    z = f(np.zeros(4, np.uint32), np.zeros(2, np.uint32))
    assert [int(v) for v in z] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    o = f(np.full(4, 0xFFFFFFFF, np.uint32), np.full(2, 0xFFFFFFFF, np.uint32))
    assert [int(v) for v in o] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    p = f(np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32),
          np.array([0xA4093822, 0x299F31D0], np.uint32))
    assert [int(v) for v in p] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


@pytest.mark.parametrize("kind", ["uniform", "smooth", "lowentropy", "constant"])
def test_integer_ops_match_pil(kind):
    from PIL import Image, ImageOps
    x = random_images(3, 37, 53, 3, seed=3, kind=kind)
    eq, po, so, inv = oracle.equalize(x), oracle.posterize(x, 3), oracle.solarize(x, 100), oracle.invert(x)
    for b in range(x.shape[0]):
        im = Image.fromarray(x[b])
        assert (np.asarray(ImageOps.equalize(im)) == eq[b]).all()
        assert (np.asarray(ImageOps.posterize(im, 3)) == po[b]).all()
        assert (np.asarray(ImageOps.solarize(im, 100)) == so[b]).all()
        assert (np.asarray(ImageOps.invert(im)) == inv[b]).all()


def test_integer_ops_match_torchvision():
    torch = pytest.importorskip("torch")
    F = pytest.importorskip("torchvision.transforms.functional")
    x = random_images(2, 64, 48, 3, seed=5, kind="smooth")
    t = torch.from_numpy(x).permute(0, 3, 1, 2).contiguous()
    back = lambda y: y.permute(0, 2, 3, 1).numpy()
    assert (back(F.equalize(t)) == oracle.equalize(x)).all()
    assert (back(F.posterize(t, 5)) == oracle.posterize(x, 5)).all()
    assert (back(F.solarize(t, 77)) == oracle.solarize(x, 77)).all()
    assert (back(F.invert(t)) == oracle.invert(x)).all()


def test_switch_edge_cases():
    x = random_images(1, 16, 16, 3, seed=1)
    # Solarize threshold 256 wraps to 0: full invert; 384 -> 128 (switch SOLARIZE_THRESHOLD_OVERFLOW)
    assert (oracle.solarize(x, 256) == oracle.invert(x)).all()
    assert (oracle.solarize(x, 384) == oracle.solarize(x, 128)).all()
    # Posterize bits=0: shift clamped to 7 (switch POSTERIZE_SHIFT8)
    assert set(np.unique(oracle.posterize(x, 0))) <= {0, 128}
    assert (oracle.posterize(x, 8) == x).all()


def test_blend_branches():
    a = random_images(1, 8, 8, 3, seed=2)
    b = random_images(1, 8, 8, 3, seed=3)
    assert (oracle.blend(a, b, 0.0) == a).all()
    assert (oracle.blend(a, b, 1.0) == b).all()
    mid = oracle.blend(a, b, 0.5)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    assert ((mid >= lo) & (mid <= hi)).all()
    ext = oracle.blend(a, b, 1.9).astype(int)
    ref = np.clip(a.astype(np.float64) + 1.9 * (b.astype(np.float64) - a), 0, 255)
    assert np.abs(ext - np.floor(ref)).max() <= 1


def test_geometric_identities():
    x = random_images(2, 20, 30, 3, seed=4)
    assert (oracle.rotate(x, 0.0, fill_value=128) == x).all()
    assert (oracle.shear_x(x, 0.0) == x).all()
    # TranslateX(+10): src_x = x + 10 -> content moves left, right edge filled
    t = oracle.translate_x(x, 10.0, negate=False, fill_value=128)
    assert (t[:, :, :20] == x[:, :, 10:]).all() and (t[:, :, 20:] == 128).all()
    t = oracle.translate_y(x, 5.0, negate=True, fill_value=7)
    assert (t[:, 5:] == x[:, :15]).all() and (t[:, :5] == 7).all()
    # 180 degree rotation about the centre is an exact flip for nearest sampling
    r = oracle.rotate(x, 180.0, fill_value=0)
    assert (r == x[:, ::-1, ::-1]).all()
    # ShearX shears about the top-left origin: row y reads x + level*y
    s = oracle.shear_x(x, 0.5, fill_value=9)
    assert (s[:, 0] == x[:, 0]).all()
    assert (s[:, 4, :28] == x[:, 4, 2:30]).all() and (s[:, 4, 28:] == 9).all()


def test_round_half_away():
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, 0.49999997, -0.49999997, 3.0, -3.0, 8388607.5], dtype=np.float32)
    assert oracle.round_half_away(v).tolist() == [1, 2, 3, -1, -2, 0, 0, 3, -3, 8388608]


def test_bilinear_and_fill_modes_are_consistent():
    x = random_images(1, 12, 12, 3, seed=6)
    # integral translation: bilinear == nearest
    a = oracle.translate_x(x, 3.0, interpolation="bilinear", fill_value=50)
    b = oracle.translate_x(x, 3.0, interpolation="nearest", fill_value=50)
    assert (a == b).all()
    # nearest fill mode replicates the edge
    c = oracle.translate_x(x, 3.0, fill_mode="nearest")
    assert (c[:, :, 9:] == x[:, :, 11:12]).all()
    # wrap: [abcd] -> [abcd|abcd]
    w = oracle.translate_x(x, 3.0, fill_mode="wrap")
    assert (w[:, :, :9] == x[:, :, 3:]).all() and (w[:, :, 9:] == x[:, :, :3]).all()
    # reflect: [abcd] -> [abcd|dcba]
    r = oracle.translate_x(x, 3.0, fill_mode="reflect")
    assert (r[:, :, 9:] == x[:, :, :8:-1]).all()


def test_cutout_geometry():
    x = np.zeros((2, 10, 12, 3), np.uint8)
    y = oracle.cutout(x, 4, 128, centers=[[0, 0], [5, 6]])
    assert (y[0, :2, :2] == 128).all() and y[0].sum() == 128 * 4 * 3
    assert (y[1, 3:7, 4:8] == 128).all() and y[1].sum() == 128 * 16 * 3
    with pytest.raises(ValueError):
        oracle.cutout(x, 3, 0, centers=[[0, 0], [0, 0]])


def test_sharpness_border_and_identity():
    x = random_images(1, 9, 11, 3, seed=8)
    assert (oracle.sharpness(x, 1.0) == x).all()
    y = oracle.sharpness(x, 1.9)
    assert (y[:, 0] == x[:, 0]).all() and (y[:, -1] == x[:, -1]).all()
    assert (y[:, :, 0] == x[:, :, 0]).all() and (y[:, :, -1] == x[:, :, -1]).all()
    flat = np.full((1, 6, 6, 3), 77, np.uint8)
    assert (oracle.sharpness(flat, 0.0) >= 76).all()  # float32 sum of 13 * 77/13 may truncate to 76


def test_magnitude_tables():
    kw = oracle.magnitude_kwargs
    assert kw("Brightness", 10)["factor"] == 1.9000000000000001
    assert kw("ShearX", 15)["level"] == 0.44999999999999996
    assert kw("Posterize", 10)["bits"] == 4 and kw("Posterize", 2)["bits"] == 0
    assert kw("Solarize", 10)["threshold"] == 256 and kw("Solarize", 15)["threshold"] == 384
    assert kw("SolarizeAdd", 10)["addition"] == 110
    assert kw("CutOut", 10) == {"mask_size": 80, "constant_values": 128}
    assert kw("Rotate", 10)["degrees"] == 30.0 and kw("Rotate", 10)["fill_value"] == 128
    t, n = oracle.autoaugment_policy()
    assert len(t) == 25 and n == 1 and t[22][0][0] == "Posterize" and t[22][0][1]["bits"] == 0
    assert t[22][1][1]["threshold"] == 256


def test_schedule_decode_properties():
    pol = oracle.randaugment_policy(2, 10)
    s = oracle.decode_schedule(pol, seed=1234, call_counter=7, image_index_base=0, B=4096, H=224, W=224, elementwise=True)
    assert s.shape == (4096, 2, 1, 5)
    counts = np.bincount(s[:, :, 0, 0].ravel(), minlength=16)
    assert counts.min() > 380 and counts.max() < 650  # uniform over 16, 8192 draws
    assert abs(s[:, :, 0, 2].mean() - 0.5) < 0.03
    assert s[..., 3].min() >= 0 and s[..., 3].max() <= 223 and s[..., 4].max() <= 223
    assert (s[..., 1] == 1).all()
    # sharding invariance: the schedule is keyed by GLOBAL image index
    a = oracle.decode_schedule(pol, 1234, 7, 1000, 96, 224, 224, True)
    assert (a == s[1000:1096]).all()
    # batch mode: one op sequence / sign for everybody, centres per image
    b = oracle.decode_schedule(pol, 1234, 7, 0, 64, 224, 224, False)
    assert (b[:, :, :, :3] == b[0:1, :, :, :3]).all()
    assert len(np.unique(b[:, 0, 0, 3])) > 10
    # AutoAugment coins follow the table's probabilities
    ap = oracle.autoaugment_policy()
    c = oracle.decode_schedule(ap, 5, 0, 0, 20000, 224, 224, True)
    sel = c[c[:, 0, 0, 0] == 0]  # sub-policy 0: Equalize .8 | ShearY .8
    assert abs(sel[:, 0, 0, 1].mean() - 0.8) < 0.05 and abs(sel[:, 0, 1, 1].mean() - 0.8) < 0.05
    sel = c[c[:, 0, 0, 0] == 16]  # ShearX 0.0 | Solarize .8
    assert sel[:, 0, 0, 1].sum() == 0
    sel = c[c[:, 0, 0, 0] == 15]  # Rotate 1.0
    assert (sel[:, 0, 0, 1] == 1).all()


def test_apply_schedule_modes(c1_batch):
    x = c1_batch[:4]
    pol = oracle.randaugment_policy(2, 10)
    s = oracle.decode_schedule(pol, 0, 0, 0, 4, 224, 224, True)
    y = oracle.apply_schedule(x, pol, s, elementwise=True)
    assert y.shape == x.shape and y.dtype == np.uint8
    # elementwise result of image b depends only on image b
    y1 = oracle.apply_schedule(x[2:3], pol, s[2:3], elementwise=True)
    assert (y1[0] == y[2]).all()


def test_imagenet_normalization_matches_the_reference_golden_vectors():
    """PINNED: the reference's own float32 golden vectors (test_units/augmentations/
    test_image_augmentations.py:21-64, extracted by tests/golden/make_imagenet_norm_fixture.py),
    asserted with exact equality exactly as the reference's tests do (assertAllEqual)."""
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "imagenet_norm_ref.json")))
    img = np.array(ref["image"], dtype=np.uint8)
    x = np.stack([img, img, img], axis=-1)[None]          # IMG, :14-15
    for mode in ("caffe", "tf", "torch"):
        got = oracle.imagenet_normalization(x, mode)
        assert got.dtype == np.float32 and got.shape == x.shape
        target = np.array(ref["targets"][mode], dtype=np.float32)
        assert np.array_equal(got[0, ..., 0], target), mode
    # caffe reverses the channels before subtracting the BGR means
    y = np.arange(2 * 3 * 5 * 3, dtype=np.uint8).reshape(2, 3, 5, 3)
    c = oracle.imagenet_normalization(y, "caffe")
    assert np.array_equal(c[..., 0], y[..., 2].astype(np.float32) - np.float32(103.939))
    assert np.array_equal(c[..., 2], y[..., 0].astype(np.float32) - np.float32(123.68))
    with pytest.raises(ValueError):
        oracle.imagenet_normalization(y, "keras")


def test_resizing_min_max_shapes_match_the_reference_tests():
    import json
    import os
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "imagenet_norm_ref.json")))
    for name, case in ref["shapes"].items():
        x = np.zeros(case["input_shape"], np.uint8)
        out = oracle.resizing_min_max(x, **case["kwargs"])
        assert list(out.shape) == case["output_shape"], name
        assert out.dtype == np.float32
    # identity-size bilinear resize reproduces the input; nearest picks existing pixels
    x = np.random.default_rng(1).integers(0, 256, size=(2, 7, 9, 3), dtype=np.uint8)
    assert np.array_equal(oracle.resize_bilinear(x, 7, 9), x.astype(np.float32))
    n = oracle.resize_nearest(x, 14, 18)
    assert n.dtype == np.uint8 and np.array_equal(n[:, ::2, ::2], x)


def check_tf_vectors(path):
    """Compares the oracle with vectors produced by the real reference on TensorFlow
    (tests/golden/make_tf_vectors.py).  Returns (report lines, number of mismatching cases).  For the
    behaviours behind oracle/switches.py it also tries the alternative setting and says which one
    the vector confirms."""
    from oracle import switches
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    lines, bad = ["vectors from tensorflow %s / tensorflow-addons %s" % (meta["tensorflow"], meta["tensorflow_addons"])], 0
    alternatives = {"Solarize": ("SOLARIZE_THRESHOLD_OVERFLOW", ("wrap", "saturate")),
                    "Posterize": ("POSTERIZE_SHIFT8", ("clamp7", "zero")),
                    "Sharpness": ("TFA_BLEND_ROUNDING", ("round_half_even", "truncate"))}
    for case in meta["cases"]:
        x = z["in_" + case["input"]]
        want = z[case["tag"]]
        centres = case["centres"] if case["centres"] is not None else [[0, 0]] * x.shape[0]
        got = oracle.apply_op(x, case["name"], case["kwargs"], negate=bool(case["negate"]), centers=centres)
        diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
        n_bad, worst = int((diff > 0).sum()), int(diff.max()) if diff.size else 0
        verdict = "exact" if n_bad == 0 else "%d bytes differ (max %d LSB)" % (n_bad, worst)
        if n_bad and case["name"] in alternatives:
            sw, values = alternatives[case["name"]]
            keep = getattr(switches, sw)
            for v in values:
                setattr(switches, sw, v)
                alt = oracle.apply_op(x, case["name"], case["kwargs"], negate=bool(case["negate"]), centers=centres)
                if np.array_equal(alt, want):
                    verdict += "; switch %s=%r reproduces it (default is %r)" % (sw, v, keep)
            setattr(switches, sw, keep)
        bad += 1 if n_bad else 0
        lines.append("%-12s %-60s on %-5s negate=%d: %s" % (case["name"], json.dumps(case["kwargs"])[:60], case["input"],
                                                           case["negate"], verdict))
    return lines, bad


def test_oracle_against_tensorflow_vectors_when_present():
    """Flips parity from "unpinned" to pinned the day tests/golden/tf_vectors.npz is generated (needs
    TensorFlow 2.6 + tensorflow-addons: tests/golden/make_tf_vectors.py)."""
    path = os.path.join(GOLDEN, "tf_vectors.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/tf_vectors.npz not generated yet (TensorFlow is not installable in this image)")
    lines, bad = check_tf_vectors(path)
    print("\n".join(lines))
    assert bad == 0, "%d cases differ from TensorFlow:\n%s" % (bad, "\n".join(l for l in lines if "differ" in l))


def test_tf_vector_file_format_roundtrip(tmp_path):
    """The consumer above on a file in the generator's format (written here FROM THE ORACLE, in a temp
    directory -- plumbing only, it pins nothing): every case must read back as exact, and a corrupted
    Solarize(256) vector must be reported together with the switch value that would reproduce it."""
    from oracle import switches
    x = random_images(2, 16, 16, 3, seed=3)
    cases, arrays = [], {"in_a": x}
    specs = [("Invert", {}, 0, None), ("Rotate", {"degrees": 30.0, "interpolation": "nearest", "fill_mode": "constant", "fill_value": 128}, 1, None),
             ("CutOut", {"mask_size": 8, "constant_values": 128}, 0, [[3, 4], [15, 0]]), ("Solarize", {"threshold": 256}, 0, None)]
    for i, (name, kw, neg, cen) in enumerate(specs):
        tag = "out_%03d_a_%d" % (i, neg)
        arrays[tag] = oracle.apply_op(x, name, kw, negate=bool(neg), centers=cen if cen is not None else [[0, 0]] * 2)
        cases.append({"tag": tag, "name": name, "kwargs": kw, "magnitude": None, "negate": neg, "input": "a", "centres": cen})
    arrays["meta"] = np.array(json.dumps({"tensorflow": "fake", "tensorflow_addons": "fake", "cases": cases}))
    path = str(tmp_path / "tf_vectors.npz")
    np.savez_compressed(path, **arrays)
    lines, bad = check_tf_vectors(path)
    assert bad == 0 and sum("exact" in l for l in lines) == len(specs)
    assert switches.SOLARIZE_THRESHOLD_OVERFLOW == "wrap"
    arrays["out_003_a_0"] = x.copy()          # what "saturate" (nothing inverted at 256) would give
    np.savez_compressed(path, **arrays)
    lines, bad = check_tf_vectors(path)
    assert bad == 1 and any("SOLARIZE_THRESHOLD_OVERFLOW='saturate' reproduces it" in l for l in lines)
    assert switches.SOLARIZE_THRESHOLD_OVERFLOW == "wrap"   # restored
