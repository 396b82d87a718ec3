import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native_built():
    """Build the native libraries in-tree if they are missing (nvcc cross-compiles without a GPU)."""
    from rayrs_b200 import build
    build.build_all()
    import oracle
    oracle.build()
    return True


@pytest.fixture(scope="session")
def hdri_small(native_built):
    from rayrs_b200 import scenes
    return scenes.synthetic_hdri(512, 256)


def relrmse(g, r):
    """relative RMSE over linear-radiance pixels."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    return float(np.sqrt(np.mean((g - r) ** 2 / (r ** 2 + 0.01))))
