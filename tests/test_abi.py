"""The C-ABI library loads, exports every symbol include/rayrs_b200.h declares, and — on a box
without a GPU — refuses to compute instead of falling back to the CPU.  CPU only."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from rayrs_b200 import _ffi

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module", autouse=True)
def _built(native_built):
    return native_built


def _declared_functions():
    text = (ROOT / "include" / "rayrs_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rrs_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_functions()
    assert len(names) >= 13
    lib = C.CDLL(str(_ffi.CUDA_LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in rayrs_b200.h but not exported"
    assert sorted(_ffi.CUDA_SYMBOLS) == names  # the Python binding covers exactly the header


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.RrsNode) == 64
    assert C.sizeof(_ffi.RrsNodeF64) == 128
    assert C.sizeof(_ffi.RrsPrim) == 88
    assert C.sizeof(_ffi.RrsRay) == 48
    assert C.sizeof(_ffi.RrsMaterial) == 72
    assert C.sizeof(_ffi.RrsStats) == 104


def test_abi_version_and_error_channel():
    lib = _ffi.cuda_lib()
    assert lib.rrs_abi_version() == _ffi.RRS_ABI_VERSION
    out = C.c_void_p()
    assert lib.rrs_scene_create(None, 0, C.byref(out)) == _ffi.RRS_ERR_INVALID
    assert b"null" in lib.rrs_last_error()
    desc = _ffi.RrsSceneDesc()
    desc.abi_version = 99
    assert lib.rrs_scene_create(C.byref(desc), 0, C.byref(out)) == _ffi.RRS_ERR_INVALID
    assert b"ABI" in lib.rrs_last_error()


def test_no_cpu_fallback_without_a_device():
    """With no usable sm_100 device the product must fail loudly (RRS_ERR_NO_DEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    from rayrs_b200 import scenes
    assert _ffi.cuda_lib().rrs_device_count() == 0
    hdri = scenes.synthetic_hdri(16, 8)
    spec = scenes.diffuse_single_sphere(16, 16)
    with pytest.raises(ValueError, match="no CUDA device|NO_DEVICE|no CPU fallback"):
        spec.scene(hdri)
    # the host half alone (BVH + flattening) needs no device
    sc = spec.scene(hdri, upload=False)
    with pytest.raises(_ffi.RayrsError):
        sc.intersect(np.zeros((1, 6)))


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py may touch oracle/ (scope rule 3)."""
    pkg = ROOT / "rayrs_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.hpp")):
        text = p.read_text()
        assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, p
        assert "oracle/" not in text or p.name == "scenes.py", p
