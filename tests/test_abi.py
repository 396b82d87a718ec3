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
    assert C.sizeof(_ffi.RrsStats) == 112
    assert C.sizeof(_ffi.RrsUniqueId) == 128
    assert C.sizeof(_ffi.RrsSceneDesc) == 112


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


def _tiny_desc(keep):
    """A valid 3-node description (virtual root -> inner node -> two leaf runs) the tests below damage."""
    prims = (_ffi.RrsPrim * 2)()
    for i, x in enumerate((-2.0, 2.0)):
        prims[i].type, prims[i].obj_id, prims[i].material, prims[i].emission = 0, i, 0, -1
        prims[i].v[0], prims[i].v[1] = 1.0, x
    nodes = (_ffi.RrsNode * 3)()
    inf = float("inf")
    for nd in nodes:
        nd.lo0[:] = nd.lo1[:] = [inf] * 3
        nd.hi0[:] = nd.hi1[:] = [-inf] * 3
        nd.ref0 = nd.ref1 = _ffi.RRS_REF_EMPTY
    nodes[0].ref0 = 1
    nodes[0].lo0[:], nodes[0].hi0[:] = [-3, -1, -1], [3, 1, 1]
    nodes[1].ref0, nodes[1].ref1 = _ffi.RRS_REF_LEAF | 0, _ffi.RRS_REF_LEAF | 1
    nodes[1].lo0[:], nodes[1].hi0[:] = [-3, -1, -1], [-1, 1, 1]
    nodes[1].lo1[:], nodes[1].hi1[:] = [1, -1, -1], [3, 1, 1]
    mats = (_ffi.RrsMaterial * 1)()
    mats[0].tag = 0
    mats[0].color[:] = [0.5, 0.5, 0.5]
    hdri = (C.c_float * 12)(*([1.0] * 12))
    d = _ffi.RrsSceneDesc()
    d.abi_version = _ffi.RRS_ABI_VERSION
    d.n_prims, d.prims = 2, prims
    d.n_nodes, d.nodes = 3, nodes
    d.max_depth = 2
    d.n_materials, d.materials = 1, mats
    d.hdri_width, d.hdri_height, d.hdri_rgb = 2, 2, hdri
    d.t_min, d.t_max = 1e-6, 1e6
    keep.extend([prims, nodes, mats, hdri])
    return d, nodes


def test_malformed_trees_are_rejected_not_trusted():
    """The traversal stack is sized from max_depth and release kernels do not bounds-check it: a cycle, a node
    reached twice, an understated depth or a wrapping max_depth must come back as an error (validated on the host,
    before any device is touched, so this runs on the CPU box too)."""
    lib = _ffi.cuda_lib()
    out = C.c_void_p()
    keep = []

    def status(d):
        rc = lib.rrs_scene_create(C.byref(d), 0, C.byref(out))
        if rc == _ffi.RRS_OK:  # a GPU box accepts the valid description
            lib.rrs_scene_destroy(out)
        return rc, lib.rrs_last_error().decode()

    d, nodes = _tiny_desc(keep)
    rc, msg = status(d)
    assert rc in (_ffi.RRS_OK, _ffi.RRS_ERR_NO_DEVICE), msg  # the undamaged description passes validation
    d, nodes = _tiny_desc(keep)
    nodes[1].ref1 = 1  # a node that is its own child: the persistent kernel would never finish
    rc, msg = status(d)
    assert rc == _ffi.RRS_ERR_INVALID and "twice" in msg
    d, nodes = _tiny_desc(keep)
    nodes[1].ref0, nodes[1].ref1 = 2, 2  # a shared subtree (DAG)
    nodes[2].ref0 = _ffi.RRS_REF_LEAF | 0
    rc, msg = status(d)
    assert rc == _ffi.RRS_ERR_INVALID and "twice" in msg
    d, nodes = _tiny_desc(keep)
    d.max_depth = 1  # understated: the real chain is 2 nodes long
    rc, msg = status(d)
    assert rc == _ffi.RRS_ERR_INVALID and "deeper" in msg
    for wrap in (0xFFFFFFFD, 0xFFFFFFFF, 118):
        d, nodes = _tiny_desc(keep)
        d.max_depth = wrap  # + 3 must not wrap around the 120-entry limit
        rc, msg = status(d)
        assert rc == _ffi.RRS_ERR_TOO_DEEP, (wrap, msg)
    d, nodes = _tiny_desc(keep)
    d.n_emissions = 1  # emissions == NULL
    rc, msg = status(d)
    assert rc == _ffi.RRS_ERR_INVALID and "emissions" in msg
    d, nodes = _tiny_desc(keep)
    d.refill_lanes = 33
    rc, msg = status(d)
    assert rc == _ffi.RRS_ERR_INVALID


def test_sample_range_is_a_partition_and_matches_the_python_mirror():
    from rayrs_b200 import api
    from rayrs_b200.multigpu import sample_range
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 5, 8, 4096, 4097):
            got = [api.sample_range(r, world, spp) for r in range(world)]
            assert got == [sample_range(r, world, spp) for r in range(world)]
            assert sum(c for _, c in got) == spp
            pos = 0
            for first, count in got:
                assert first == pos
                pos += count
    with pytest.raises(_ffi.RayrsError):
        api.sample_range(2, 2, 8)


def test_single_device_communicator_needs_no_nccl():
    lib = _ffi.cuda_lib()
    comm = C.c_void_p()
    devs = (C.c_int * 1)(0)
    assert lib.rrs_comm_init_all(devs, 1, C.byref(comm)) == _ffi.RRS_OK and comm.value
    lib.rrs_comm_destroy(comm)
    two = (C.c_int * 2)(0, 0)
    assert lib.rrs_comm_init_all(two, 2, C.byref(comm)) == _ffi.RRS_ERR_INVALID  # the same device twice


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
