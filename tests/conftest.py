import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


def _have_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The shared libraries are build artefacts (git-ignored).  Build what is missing; never rebuild on
    the GPU box, which receives the prebuilt files."""
    lib = ROOT / "raytracing_renderer_cuda_b200" / "librt_b200.so"
    orc = ROOT / "oracle" / "liboracle.so"
    if not lib.exists():
        subprocess.check_call(["make", "-C", str(ROOT), "lib"])
    if not orc.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle.so"])
    yield


@pytest.fixture(scope="session")
def earth():
    from raytracing_renderer_cuda_b200.assets import load_earth

    return load_earth()


@pytest.fixture(scope="session")
def oracle():
    from tests.oracle_api import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def cpu_golden():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "ref_cpu_golden.npz")


SCENES = ("earth_emitter", "book1_final", "perlin_motion")


def record_parity(test: str, **values) -> None:
    """Appends the MEASURED value behind a parity gate (PSNR dB, differing-pixel fractions, mismatch counts) to
    gpurun_out/parity.jsonl (or $RT_PARITY_LOG), so that the numbers the assertions only bound are kept:
    tools/parity_report.py turns the file into profiles/rNN_parity.md."""
    import json

    path = Path(os.environ.get("RT_PARITY_LOG", ROOT / "gpurun_out" / "parity.jsonl"))
    try:
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "a") as f:
            f.write(json.dumps({"test": test, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass


def make_env_image(w: int = 250, h: int = 130):
    """Procedural environment image for the reference's second scene (populate_scene_hdr, main.cu:136-182; its
    textures/hdr.jpg is not shipped).  Deliberately NOT the render size and not a multiple of 8: the reference's
    image_texture takes any width x height (texture.h:116-132)."""
    import numpy as np

    v, u = np.mgrid[0:h, 0:w].astype(np.float32)
    u, v = u / (w - 1), v / (h - 1)
    sky = np.stack([0.3 + 0.5 * v, 0.45 + 0.4 * v, 0.9 - 0.3 * v], -1)
    sun = np.exp(-(((u - 0.3) / 0.03) ** 2 + ((v - 0.35) / 0.05) ** 2))[..., None] * np.array([4.0, 3.6, 2.5], np.float32)
    bands = (0.1 * np.sin(40 * u) * np.cos(25 * v))[..., None]
    return np.ascontiguousarray(np.clip(sky + bands, 0, 1) + sun, dtype=np.float32)


@pytest.fixture(scope="session")
def scene_descs(earth):
    import raytracing_renderer_cuda_b200 as rt

    return {"earth_emitter": rt.SceneDesc.builtin("earth_emitter", earth), "book1_final": rt.SceneDesc.builtin("book1_final"),
            "perlin_motion": rt.SceneDesc.builtin("perlin_motion"), "hdr_sphere": rt.SceneDesc.builtin("hdr_sphere", make_env_image())}
