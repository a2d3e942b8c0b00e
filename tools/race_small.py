"""Small resident-engine workload for compute-sanitizer --tool racecheck (shared-memory hazards): every chain class
once, 64 x 64 images, a few CTAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chambers_b200 import augmentations as A, _lib
rng = np.random.default_rng(0)
x = torch.from_numpy(rng.integers(0, 256, size=(24, 64, 64, 3), dtype=np.uint8)).cuda()
_lib.set_engine(0, "resident")
for layer in (A.RandAugment(3, 10, elementwise=True), A.AutoAugment(elementwise=True)):
    for call in range(3):
        y = layer(x, training=True, seed=1, call_counter=call)
torch.cuda.synchronize()
print("done", int(y.sum()))
