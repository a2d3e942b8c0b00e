import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def c1_batch():
    """BASELINE.json configs[0] input: 32x224x224x3 uint8 from the reference's sample_data."""
    import hashlib
    z = np.load(os.path.join(GOLDEN, "c1_batch.npz"))
    imgs = z["images"]
    assert hashlib.sha256(imgs.tobytes()).hexdigest() == str(z["sha256"])
    return imgs


def random_images(B, H, W, C=3, seed=0, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, size=(B, H, W, C), dtype=np.uint8)
    if kind == "lowentropy":
        vals = rng.integers(0, 256, size=8, dtype=np.uint8)
        return vals[rng.integers(0, 8, size=(B, H, W, C))]
    if kind == "constant":
        return np.full((B, H, W, C), int(rng.integers(0, 256)), dtype=np.uint8)
    if kind == "smooth":
        y, x = np.mgrid[0:H, 0:W]
        base = (np.sin(x / 7.0)[None, :, :, None] * 60 + np.cos(y / 5.0)[None, :, :, None] * 60 + 128
                + rng.normal(0, 6, size=(B, H, W, C)))
        return np.clip(base, 0, 255).astype(np.uint8)
    raise ValueError(kind)
