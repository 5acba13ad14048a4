"""Decode the reference's shipped golden images into .npy fixtures.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The PNGs are the only artefacts of the reference that pin the result of the hot path
(SURVEY.md section 4.2); md5 of the sources is recorded in golden.json.
"""
import hashlib
import json
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scenes"

meta = {}
for name in ("sphere", "sphere2"):
    path = os.path.join(SRC, name + ".png")
    raw = open(path, "rb").read()
    img = np.array(Image.open(path).convert("RGB"), dtype=np.uint8)
    assert img.shape == (200, 200, 3)
    np.save(os.path.join(HERE, name + ".npy"), img)
    meta[name] = {"source": path, "md5": hashlib.md5(raw).hexdigest(), "shape": list(img.shape),
                  "blue_census": {str(int(v)): int(c) for v, c in zip(*np.unique(img[..., 2], return_counts=True))}}
json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1)
print(json.dumps(meta, indent=1))
