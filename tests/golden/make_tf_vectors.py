"""Generates tests/golden/tf_vectors.npz: outputs of the REAL reference -- chambers.augmentations on
TensorFlow 2.6 + tensorflow-addons -- for every op of the RandAugment / AutoAugment hot path, with the
op's own random draws (sign flip, CutOut centres) captured, so that tests/test_oracle.py can pin the
oracle (and through it the CUDA path) to TensorFlow bit for bit.

IT CANNOT RUN IN THE BUILD CONTAINER: TensorFlow 2.6 has no wheel for its Python 3.12 and there is no
network.  Run it anywhere the reference runs, e.g.

    python3.7 -m venv tfenv && tfenv/bin/pip install tensorflow==2.6.0 keras==2.6.0 tensorflow-addons==0.14.0 numpy
    tfenv/bin/python tests/golden/make_tf_vectors.py --reference /path/to/chjort-chambers

and commit the resulting tests/golden/tf_vectors.npz.  Until then the oracle stays "parity
unpinned" for everything but ImageNetNormalization (oracle/__init__.py).

What is recorded, per case:
    name, kwargs (JSON), magnitude, negate (0/1, what _randomly_negate_value returned), centres
    ([B, 2] = (cy, cx) handed to tfa.image.cutout by random_cutout), input key, output array.
Cases: the 16 ops built by augmentation_schemes._get_transform at M in {0, 2, 10, 15} (both sign
outcomes for the signed ops), plus the parameter edge cases the oracle's switches are about
(Solarize 256 / 384, Posterize bits 0, Sharpness / Brightness / Color at factor < 1 and > 1), each on
    img4    the reference's 4 x 4 test image, three equal channels (test_image_augmentations.py:5-15)
    c1      the first four 224 x 224 x 3 images of tests/golden/c1_batch.npz
    const   a constant image and a 9 x 16 image (Equalize step == 0, AutoContrast hi == lo)
Ops are called exactly as RandomChoice would call them: layer(x) with x uint8 [B, H, W, 3], eagerly.
"""
import argparse
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OPS = ["AutoContrast", "Equalize", "Invert", "Brightness", "Contrast", "Color", "Sharpness", "ShearX", "ShearY",
       "TranslateX", "TranslateY", "Posterize", "Solarize", "SolarizeAdd", "CutOut", "Rotate"]
SIGNED = {"ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"}
IMG4 = [[139, 186, 208, 200], [175, 201, 198, 200], [166, 191, 193, 195], [124, 155, 172, 151]]


def inputs():
    img4 = np.stack([np.array(IMG4, np.uint8)] * 3, axis=-1)[None]
    c1 = np.load(os.path.join(HERE, "c1_batch.npz"))["images"][:4]
    rng = np.random.default_rng(0)
    const = np.full((2, 32, 48, 3), 77, np.uint8)
    const[1] = 200
    small = rng.integers(0, 256, size=(2, 9, 16, 3), dtype=np.uint8)
    return {"img4": img4, "c1": c1, "const": const, "small": small}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference", help="checkout of chjort/chambers")
    ap.add_argument("--out", default=os.path.join(HERE, "tf_vectors.npz"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import tensorflow as tf
    import tensorflow_addons as tfa
    from chambers.augmentations import image_augmentations as ia
    from chambers.augmentations import augmentation_schemes as sch

    # --- capture the ops' own random draws -------------------------------------------------------
    forced = {"negate": 0}

    def fixed_negate(value):  # stands in for image_augmentations._randomly_negate_value (:52-56)
        v = tf.cast(value, tf.float32)
        return -v if forced["negate"] else v
    ia._randomly_negate_value = fixed_negate

    seen = {"centres": None}
    cutout_mod = sys.modules[tfa.image.cutout.__module__]
    real_cutout = cutout_mod.cutout

    def spy_cutout(images, mask_size, offset=(0, 0), constant_values=0, **kw):
        seen["centres"] = np.array(tf.convert_to_tensor(offset)).reshape(-1, 2)
        return real_cutout(images, mask_size, offset, constant_values, **kw)
    cutout_mod.cutout = spy_cutout  # random_cutout resolves `cutout` in its module at call time

    cases = []
    for name in OPS:
        for m in (0, 2, 10, 15):
            cases.append((name, sch._get_transform(name, m), m))
    extra = [("Solarize", dict(threshold=256)), ("Solarize", dict(threshold=384)), ("Solarize", dict(threshold=128)),
             ("Posterize", dict(bits=0)), ("Posterize", dict(bits=8)), ("SolarizeAdd", dict(addition=-40, threshold=200)),
             ("Sharpness", dict(factor=0.28)), ("Sharpness", dict(factor=1.9)), ("Brightness", dict(factor=0.28)),
             ("Color", dict(factor=0.1)), ("Contrast", dict(factor=0.46)),
             ("Rotate", dict(degrees=33.0, interpolation="bilinear", fill_mode="reflect", fill_value=77.0)),
             ("ShearX", dict(level=0.7, interpolation="nearest", fill_mode="wrap", fill_value=0.0))]
    for name, kw in extra:
        cases.append((name, getattr(ia, name)(**kw), None))

    out = {}
    meta = []
    data = inputs()
    for key, x in data.items():
        out["in_" + key] = x
    tf.random.set_seed(1234)
    for ci, (name, layer, m) in enumerate(cases):
        cfg = {k: v for k, v in layer.get_config().items() if k not in ("name", "trainable", "dtype")}
        for key, x in data.items():
            for negate in ((0, 1) if name in SIGNED else (0,)):
                forced["negate"] = negate
                seen["centres"] = None
                if name == "CutOut" and cfg.get("mask_size", 0) % 2:
                    continue
                y = np.array(layer(tf.constant(x)))
                tag = "out_%03d_%s_%d" % (ci, key, negate)
                out[tag] = y
                rec = {"tag": tag, "name": name, "kwargs": cfg, "magnitude": m, "negate": negate, "input": key,
                       "centres": None if seen["centres"] is None else seen["centres"].tolist()}
                meta.append(rec)
    out["meta"] = np.array(json.dumps({"tensorflow": tf.__version__, "tensorflow_addons": tfa.__version__,
                                       "cases": meta}))
    np.savez_compressed(args.out, **out)
    print("wrote %s: %d cases (tf %s, tfa %s)" % (args.out, len(meta), tf.__version__, tfa.__version__))


if __name__ == "__main__":
    main()
