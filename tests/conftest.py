import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def built():
    """libwpt.so + liboracle.so + the generated stand-in meshes."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def meshes(built):
    import wasm_pathtracer_b200 as W
    out = {}
    for sub in (3, 4):
        path = os.path.join(ROOT, "assets", "_gen", "standin_%d.obj" % sub)
        out[sub] = W.parse_obj(open(path).read(), True)
    return out


@pytest.fixture(scope="session")
def gpu_ok(built):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback to fall back to)")
    return True
