"""Generates tests/golden/jpeg_golden.npz from the REFERENCE's own JPEG writer: the vendored stb_image_write.h
compiled where it lies under /root/reference (oracle/_ref/libref_stb.so, `make -C oracle ref`).  Run in the build
container (the GPU box has no /root/reference):  python tests/golden/make_jpeg_golden.py
Stores, per case, the generator arguments of the input image (tests/oracle_api.jpeg_test_image) and the bytes stb wrote."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from tests import oracle_api as oa  # noqa: E402

CASES = [("noise", 8, 8, 100), ("noise", 1, 1, 100), ("noise", 7, 3, 100), ("noise", 37, 21, 100), ("noise", 37, 21, 90),
         ("noise", 37, 21, 50), ("noise", 33, 17, 1), ("smooth", 64, 48, 100), ("smooth", 64, 48, 75), ("flat", 40, 24, 100),
         ("flat", 40, 24, 10), ("sat", 48, 40, 100), ("sat", 50, 30, 35), ("photo", 120, 60, 100), ("photo", 120, 60, 91),
         ("photo", 130, 70, 90)]
out = {}
for k, (kind, w, h, q) in enumerate(CASES):
    img = oa.jpeg_test_image(kind, w, h, seed=k)
    out[f"case{k:02d}"] = np.array([kind, str(w), str(h), str(q), str(k)])
    out[f"jpg{k:02d}"] = np.frombuffer(oa.ref_stb_jpeg(img, q), dtype=np.uint8)
np.savez_compressed(ROOT / "tests" / "golden" / "jpeg_golden.npz", **out)
print("wrote", len(CASES), "cases,", sum(v.size for k, v in out.items() if k.startswith("jpg")), "bytes of JPEG")
