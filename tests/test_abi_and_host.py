"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/chambers_aug.h declares (no compute calls without a GPU), the C table builders agree with
the reference's magnitude maps as restated by the oracle, and the Python layers mirror the
reference's constructor / get_config / training-gate / error behaviour.
"""

import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from chambers_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from chambers_b200 import _lib
    header = open(os.path.join(ROOT, "include", "chambers_aug.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(chb_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 13
    for name in declared:
        assert hasattr(lib, name), "libchambers_aug.so does not export " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.chb_version() == 1


def test_struct_layout_matches_header():
    from chambers_b200 import _lib
    assert ctypes.sizeof(_lib.ChbOp) == 40
    assert ctypes.sizeof(_lib.ChbTransform) == 8 + 4 * 40
    assert _lib.ChbOp.probability.offset == 24 and _lib.ChbOp.value.offset == 32


def test_init_fails_loudly_without_gpu(lib):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    handle = ctypes.c_void_p()
    rc = lib.chb_init(0, ctypes.byref(handle))
    assert rc != 0 and not handle.value
    assert b"no CPU fallback" in lib.chb_last_error(None)
    from chambers_b200 import augmentations as A
    with pytest.raises(RuntimeError):
        A.Invert()(np.zeros((1, 4, 4, 3), np.uint8))


def _op_tuple(o):
    from chambers_b200 import _lib
    return (_lib.OP_KINDS[o.kind], o.value, tuple(o.ivalue), o.interpolation, o.fill_mode, o.fill_value)


@pytest.mark.parametrize("magnitude", [0, 1, 2, 3, 5, 7, 9, 10, 15, 4.5])
def test_c_randaugment_table_matches_reference_maps(lib, magnitude):
    from chambers_b200 import _lib
    t = (_lib.ChbTransform * 16)()
    assert lib.chb_randaugment_table(float(magnitude), t) == 0
    for i, name in enumerate(oracle.OP_NAMES):
        o = t[i].ops[0]
        kw = oracle.magnitude_kwargs(name, magnitude)
        assert t[i].n_ops == 1 and _lib.OP_KINDS[o.kind] == name and o.probability < 0
        for key in ("factor", "level", "pixels", "degrees"):
            if key in kw:
                assert o.value == kw[key], (name, o.value, kw[key])  # bit-identical doubles
        if name == "Posterize":
            assert o.ivalue[0] == kw["bits"]
        if name == "Solarize":
            assert o.ivalue[0] == kw["threshold"]
        if name == "SolarizeAdd":
            assert (o.ivalue[0], o.ivalue[1]) == (kw["addition"], 128)
        if name == "CutOut":
            assert (o.ivalue[0], o.ivalue[1]) == (kw["mask_size"], kw["constant_values"])
        if "fill_value" in kw:
            assert o.fill_value == 128.0 and o.interpolation == 0 and o.fill_mode == 0


def test_c_autoaugment_table_matches_python_layers(lib):
    from chambers_b200 import _lib, augmentations as A
    from chambers_b200.augmentations.image_augmentations import _flatten_transform
    t = (_lib.ChbTransform * 25)()
    assert lib.chb_autoaugment_table(t) == 0
    aa = A.AutoAugment()
    ref, _ = oracle.autoaugment_policy()
    for i in range(25):
        flat = _flatten_transform(aa.transforms[i])
        assert t[i].n_ops == 2 == len(flat)
        for j in range(2):
            py = _lib.ChbOp()
            flat[j][0]._fill_op(py)
            assert _op_tuple(py) == _op_tuple(t[i].ops[j])
            assert t[i].ops[j].probability == flat[j][1] == ref[i][j][2]
            assert _lib.OP_KINDS[py.kind] == ref[i][j][0]


def test_export_list_matches_reference():
    from chambers_b200 import augmentations as A
    # every chambers-owned name of chambers/augmentations/__init__.py:14-39 (the Keras re-exports of
    # :1-13 are Keras code); ImageNetNormalization and ResizingMinMax included
    names = ["RandomChoice", "RandomChance", "AutoContrast", "Equalize", "Invert", "Rotate", "Posterize",
             "Solarize", "SolarizeAdd", "Color", "Contrast", "Brightness", "Sharpness", "ShearX", "ShearY",
             "TranslateX", "TranslateY", "CutOut", "AutoAugment", "RandAugment", "ImageNetNormalization",
             "ResizingMinMax"]
    for n in names:
        assert hasattr(A, n), n
    import os
    ref_init = "/root/reference/chambers/augmentations/__init__.py"
    if os.path.exists(ref_init):  # in the build container: the list above is the reference's own
        import ast
        tree = ast.parse(open(ref_init).read())
        ours = set()
        for node in tree.body:
            if isinstance(node, ast.ImportFrom) and node.module in ("image_augmentations", "augmentation_schemes") and node.level == 1:
                ours.update(a.name for a in node.names)
        assert ours == set(names), ours ^ set(names)


def test_frontend_layers_host_logic():
    """ImageNetNormalization / ResizingMinMax: constructor errors, configs, and the resize size
    arithmetic of the C ABI (host code, no GPU) against the reference's four shape tests
    (test_units/augmentations/test_image_augmentations.py:66-80) and the oracle."""
    import json
    from chambers_b200 import augmentations as A, _lib
    with pytest.raises(ValueError):
        A.ImageNetNormalization(mode="keras")
    with pytest.raises(ValueError):
        A.ResizingMinMax()
    n = A.ImageNetNormalization()
    assert n.mode == "caffe" and n.get_config() == {"name": n.name, "mode": "caffe"}
    r = A.ResizingMinMax(min_side=100, max_side=50, interpolation="nearest")
    assert r.get_config() == {"name": r.name, "min_side": 100, "max_side": 50, "interpolation": "nearest"}
    assert A.deserialize(A.serialize(r)).get_config()["max_side"] == 50
    assert r.compute_output_shape([8, 4, 3, 3]) == [8, 100, 50, 3]
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "imagenet_norm_ref.json")))
    for name, case in ref["shapes"].items():
        _, H, W, C = case["input_shape"]
        kw = case["kwargs"]
        got = _lib.resize_min_max_shape(H, W, kw.get("min_side"), kw.get("max_side"))
        assert list(got) == case["output_shape"][1:3], (name, got)
        assert oracle.resizing_min_max_shape(H, W, **kw) == got
    rng = np.random.default_rng(0)
    for _ in range(300):
        H, W = (int(v) for v in rng.integers(1, 3000, size=2))
        mn = int(rng.integers(1, 1200)) if rng.random() < 0.7 else None
        mx = int(rng.integers(1, 1200)) if (mn is None or rng.random() < 0.6) else None
        assert _lib.resize_min_max_shape(H, W, mn, mx) == oracle.resizing_min_max_shape(H, W, mn, mx), (H, W, mn, mx)


def test_constructors_and_configs():
    from chambers_b200 import augmentations as A
    r = A.Rotate(30.0)
    assert r.get_config() == {"name": r.name, "degrees": 30.0, "interpolation": "nearest",
                              "fill_mode": "constant", "fill_value": 0.0}
    assert A.Solarize().threshold == 128 and A.SolarizeAdd().get_config()["addition"] == 0
    assert A.CutOut(80).constant_values == 0
    ra = A.RandAugment(2, 10)
    cfg = ra.get_config()
    assert {k: cfg[k] for k in ("n_transforms", "magnitude", "elementwise")} == \
        {"n_transforms": 2, "magnitude": 10, "elementwise": False}
    assert [type(t).__name__ for t in ra.transforms] == oracle.OP_NAMES
    assert ra.transforms[3].factor == 1.9000000000000001 and ra.transforms[15].fill_value == 128
    ra2 = A.RandAugment.from_config(cfg)
    assert ra2.get_config() == cfg
    aa = A.AutoAugment(elementwise=True)
    assert aa.get_config()["elementwise"] is True and len(aa.transforms) == 25
    rc = aa._transform
    rc2 = A.RandomChoice.from_config(rc.get_config())
    assert rc2.get_config() == rc.get_config()
    ch = A.RandomChance(A.Invert(name="inv"), 0.25)
    assert ch.name == "random_chance_inv"
    assert A.deserialize(A.serialize(ch)).probability == 0.25
    assert ra.compute_output_shape([8, 224, 224, 3]) == [8, 224, 224, 3]
    with pytest.raises(TypeError):
        A.Invert(bogus=1)


def test_training_gate_and_input_checks():
    from chambers_b200 import augmentations as A
    x = np.zeros((2, 8, 8, 3), np.uint8)
    ra = A.RandAugment(2, 10)
    assert ra(x) is x and ra(x, training=False) is x  # smart_cond false branch returns the input
    A.set_learning_phase(False)
    assert A.AutoAugment()(x, training=None) is x
    with pytest.raises(ValueError):
        ra(np.zeros((8, 8, 3), np.uint8), training=True)
    with pytest.raises(ValueError):
        ra(np.zeros((1, 8, 8, 3), np.float32), training=True)
    with pytest.raises(ValueError):
        A.RandomChoice([A.RandomChoice([A.Invert()], 1)], 1)(x)
    with pytest.raises(ValueError):
        A.RandomChoice([A.Invert()] , 9)(x)  # chain longer than the fused limit


def test_layer_streams_are_reproducible():
    from chambers_b200 import augmentations as A
    A.set_random_seed(42)
    a, b = A.Invert(), A.Rotate(10.0)
    sa, sb = a._stream(), b._stream()
    A.set_random_seed(42)
    a2, b2 = A.Invert(), A.Rotate(10.0)
    assert a2._stream() == sa and b2._stream() == sb and sa[0] != sb[0]
    assert a2._stream()[1] == 1  # call counter advances
