import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        if lib.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        lib.cuDeviceGetCount(ctypes.byref(n))
        return n.value
    except OSError:
        return 0


HAS_GPU = _cuda_device_count() > 0


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_port():
    """The C restatement (test infrastructure), compiled on demand."""
    from oracle import port
    port.build()
    return port


@pytest.fixture(scope="session")
def golden():
    from cases import CASES

    def load(name):
        z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        return CASES[name], z["frames"], z["pixels"]

    return load


@pytest.fixture(scope="session")
def renderer_lib():
    from swift3drenderer_b200 import renderer
    renderer.build_library()
    return renderer


@pytest.fixture(scope="session")
def gpu_renderer(renderer_lib):
    r = renderer_lib.Renderer(0)
    yield r
    r.close()
