"""Write a synthetic scene (query + reference PNGs) for exercising crossscore_b200.predict (development tool)."""
import os, sys
import numpy as np
from PIL import Image
root, nq, nr = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
H, W = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (540, 960)
rng = np.random.default_rng(0)
for sub, n in (("m/ds/scene/test/ours/renders", nq), ("m/ds/scene/train/ours/gt", nr)):
    d = os.path.join(root, sub)
    os.makedirs(d, exist_ok=True)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n):
        a = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        a[..., 1] = (127 + 100 * np.sin(xx / (9.0 + i)) * np.cos(yy / 7.0)).astype(np.uint8)
        Image.fromarray(a).save(os.path.join(d, f"frame_{i:05}.png"))
print(os.path.join(root, "m/ds/scene/test/ours/renders"), os.path.join(root, "m/ds/scene/train/ours/gt"))
