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


@pytest.fixture(scope="session")
def scene_descs(earth):
    import raytracing_renderer_cuda_b200 as rt

    return {"earth_emitter": rt.SceneDesc.builtin("earth_emitter", earth), "book1_final": rt.SceneDesc.builtin("book1_final"),
            "perlin_motion": rt.SceneDesc.builtin("perlin_motion")}
