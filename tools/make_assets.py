"""Regenerates assets/earth_stb.png from the reference's textures/earth.jpg, decoded by the
reference's own vendored stb_image (through `oracle/_ref/ref_harness earth`, host only)."""
import subprocess
import sys
from pathlib import Path

import numpy as np
from PIL import Image

root = Path(__file__).resolve().parent.parent
subprocess.check_call(["make", "-C", str(root / "oracle"), "_ref/earth_stb.f32"])
raw = np.fromfile(root / "oracle/_ref/earth_stb.f32", dtype=np.uint8)
w, h = np.frombuffer(raw[:8], dtype=np.int32)
img = np.frombuffer(raw[8:], dtype=np.float32).reshape(h, w, 3)
b = np.rint(img * 255).astype(np.uint8)
assert np.array_equal(b.astype(np.float32) / np.float32(255), img), "stb floats are not byte/255.f"
(root / "assets").mkdir(exist_ok=True)
Image.fromarray(b).save(root / "assets/earth_stb.png", optimize=True)
print("wrote assets/earth_stb.png", w, h, file=sys.stderr)
