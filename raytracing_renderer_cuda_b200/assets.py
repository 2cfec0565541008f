"""Input assets of the benchmark scenes."""
from __future__ import annotations

from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
EARTH_PNG = ROOT / "assets" / "earth_stb.png"


def load_earth() -> np.ndarray:
    """The reference's textures/earth.jpg as stbi_loadf returns it with ldr_to_hdr gamma = scale = 1
    (main.cu:376-380): float32 [600, 1200, 3], row 0 = top, value = byte / 255.f.

    assets/earth_stb.png is a lossless copy of the bytes stb_image v2.26 decodes from that JPEG
    (libjpeg/Pillow decode 0.8 % of the bytes differently, by up to 3 levels), written by
    `oracle/_ref/ref_harness earth` + tools/make_assets.py.
    """
    from PIL import Image

    b = np.asarray(Image.open(EARTH_PNG).convert("RGB"), dtype=np.uint8)
    return (b.astype(np.float32) / np.float32(255.0)).copy()
