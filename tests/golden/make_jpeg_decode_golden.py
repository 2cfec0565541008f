"""Generates tests/golden/jpeg_decode_golden.npz: small JPEG files (written by Pillow/libjpeg in every mode the reader
covers, and by this repo's stb-identical encoder) together with the bytes the REFERENCE's own vendored stb_image.h decodes
from them (oracle/_ref/libref_stb.so, `make -C oracle ref`; needs /root/reference).
    python tests/golden/make_jpeg_decode_golden.py"""
import io
import sys
from pathlib import Path

import numpy as np
from PIL import Image

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import oracle_api as oa  # noqa: E402

out, k = {}, 0
for (w, h) in [(1, 1), (8, 8), (17, 9), (33, 47), (100, 75)]:
    img = oa.jpeg_test_image("photo", w, h, seed=w)
    files = []
    for prog in (False, True):
        for sub in (0, 1, 2):  # 4:4:4, 4:2:2, 4:2:0
            b = io.BytesIO()
            Image.fromarray(img).save(b, "JPEG", quality=85, progressive=prog, subsampling=sub)
            files.append((f"pil {w}x{h} progressive={prog} subsampling={sub}", b.getvalue()))
        b = io.BytesIO()
        Image.fromarray(img).convert("L").save(b, "JPEG", quality=80, progressive=prog)
        files.append((f"pil grey {w}x{h} progressive={prog}", b.getvalue()))
    files.append((f"stb-identical encoder {w}x{h} q100", oa.oracle_jpeg(img, 100)))
    files.append((f"stb-identical encoder {w}x{h} q60 (4:2:0)", oa.oracle_jpeg(img, 60)))
    for name, data in files:
        out[f"name{k:03d}"] = np.array(name)
        out[f"file{k:03d}"] = np.frombuffer(data, dtype=np.uint8)
        out[f"pix{k:03d}"] = oa.ref_stb_load_jpeg(data)
        k += 1
np.savez_compressed(ROOT / "tests" / "golden" / "jpeg_decode_golden.npz", **out)
print("wrote", k, "cases")
