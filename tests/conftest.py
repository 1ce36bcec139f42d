import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_ids():
    return [os.path.basename(f)[:-4] for f in golden_files()]


def load_golden(path):
    d = np.load(path, allow_pickle=False)
    g = {k: d[k] for k in d.files}
    g["albedo"] = str(g["albedo"])
    g["shading"] = str(g["shading"])
    for k in ("num_vertices", "num_cameras", "width", "height"):
        g[k] = int(g[k])
    return g


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Builds the product library (nvcc cross-compiles without a GPU) and the CPU oracle once."""
    from gvv_differentiable_cuda_renderer_b200 import _native
    from oracle import cpu
    _native.build_library()
    cpu.build()
    yield


def rel_l2(x, y):
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    d = np.linalg.norm(y)
    return float(np.linalg.norm(x - y) / d) if d > 0 else float(np.linalg.norm(x))
