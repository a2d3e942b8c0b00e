"""Builds tests/golden/c1_batch.npz: BASELINE.json configs[0]'s input -- the first 32 files of
sorted(triplets/train/*/*/*) + sorted(triplets/test/*/*/*) under the reference's
test_units/sample_data, PIL-decoded, .convert("RGB"), bilinear-resized to 224x224 (input
preparation, not under test).  Run in the build container only (needs /root/reference);
the GPU box uses the committed .npz.

    python tests/golden/make_c1_fixture.py
"""
import glob
import hashlib
import os

import numpy as np
from PIL import Image

ROOT = "/root/reference/test_units/sample_data/triplets"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    files = sorted(glob.glob(os.path.join(ROOT, "train", "*", "*", "*"))) + \
        sorted(glob.glob(os.path.join(ROOT, "test", "*", "*", "*")))
    files = [f for f in files if os.path.isfile(f)][:32]
    assert len(files) == 32, len(files)
    batch = np.stack([
        np.asarray(Image.open(f).convert("RGB").resize((224, 224), Image.BILINEAR), dtype=np.uint8)
        for f in files
    ])
    sha = hashlib.sha256(batch.tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, "c1_batch.npz"), images=batch,
                        files=np.array([os.path.relpath(f, ROOT) for f in files]), sha256=np.array(sha))
    print(batch.shape, batch.dtype, sha)


if __name__ == "__main__":
    main()
