import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from chambers_b200 import augmentations as A, _lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import policy_of
names = oracle.OP_NAMES
pairs = [tuple(p.split(",")) for p in sys.argv[1:]] or [("Color", "Equalize"), ("Rotate", "Equalize"), ("CutOut", "Equalize"), ("Rotate", "AutoContrast")]
rng = np.random.default_rng(0)
x = rng.integers(0, 256, size=(len(pairs), 224, 224, 3), dtype=np.uint8)
s = np.zeros((len(pairs), 2, 1, 5), np.int32)
for b, (a, c) in enumerate(pairs):
    s[b, 0, 0, 0] = names.index(a); s[b, 1, 0, 0] = names.index(c)
s[..., 1] = 1; s[..., 3] = 100; s[..., 4] = 50
layer = A.RandAugment(2, 10, elementwise=True)
xg = torch.from_numpy(x).cuda()
_lib.set_engine(0, "resident"); r = layer(xg, training=True, replay=s).cpu().numpy()
_lib.set_engine(0, "tiles"); t = layer(xg, training=True, replay=s).cpu().numpy()
want = oracle.apply_schedule(x, policy_of(layer), s, elementwise=True)
for b, p in enumerate(pairs):
    dr = (r[b].astype(int) - want[b]); dt = (t[b].astype(int) - want[b])
    print(p, "resident vs oracle: %d bytes differ (max %d)" % ((dr != 0).sum(), np.abs(dr).max()), " tiles vs oracle: %d (max %d)" % ((dt != 0).sum(), np.abs(dt).max()))
    if (dr != 0).any():
        idx = np.argwhere(dr != 0)[:5]
        print("   first diffs (y,x,c): got/want", [(tuple(i), int(r[b][tuple(i)]), int(want[b][tuple(i)])) for i in idx])
