"""tests/golden/ref_render_earth_emitter_300x150.png: the reference's own committed render of its hard-coded scene
(/root/reference/renders/earth_emitter.jpg, README.md:17-18: 1200x600, 100 spp, JPEG quality 100), box-filtered 4x4 in
linear byte space to 300x150 and stored losslessly.  SURVEY.md 8c lists it as one of the two fixtures the reference
holds for this path (a loose, noise-limited one).  Run in the build container only (the GPU box has no /root/reference)."""
from pathlib import Path

import numpy as np
from PIL import Image

src = Path("/root/reference/renders/earth_emitter.jpg")
im = np.asarray(Image.open(src).convert("RGB"), dtype=np.float64)
assert im.shape == (600, 1200, 3)
small = im.reshape(150, 4, 300, 4, 3).mean(axis=(1, 3))
out = Path(__file__).resolve().parent / "ref_render_earth_emitter_300x150.png"
Image.fromarray(np.clip(np.rint(small), 0, 255).astype(np.uint8)).save(out, optimize=True)
print(out, out.stat().st_size, "bytes")
