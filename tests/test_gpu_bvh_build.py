"""The device BVH build (rrs_bvh_build, csrc/bvh_build.cu) against the host mirror's recursive build — which
tests/test_host_bvh.py pins to the oracle's literal restatement of rayrs-lib/src/bvh.rs:227-389: the SAME tree, i.e.
every array of the flattened scene equal byte for byte (RrsNode, RrsNodeF64, primitive order, the pre-order topology
dump and its f64 boxes).  Cases: the 46 of test_host_bvh.py (ties, lattice centres under six split counts, all three
heuristics), a 90k-triangle mesh, and the 1.0M- and 4.0M-triangle meshes of configurations 4 and 5."""
import numpy as np
import pytest

from rayrs_b200 import scenes
from rayrs_b200.api import BvhHeuristic, Image, Scene

from test_host_bvh import CASES, _lattice_spheres, _random_spheres

pytestmark = pytest.mark.gpu
HDRI = Image(2, 2, np.ones((2, 2, 3)))


def _same_tree(objects, heuristic):
    host = Scene(objects, 1e-6, 1e6, heuristic, HDRI, upload=False)
    dev = Scene(objects, 1e-6, 1e6, heuristic, HDRI, upload=False, device_build=True)
    assert (dev.n_nodes, dev.n_prims, dev.max_depth, dev.dead_nodes) == (host.n_nodes, host.n_prims, host.max_depth, host.dead_nodes)
    fh, fd = host.flat(), dev.flat()
    assert np.array_equal(fh[2], fd[2]), "primitive order differs"
    assert np.array_equal(fh[3], fd[3]), "topology differs"
    assert np.array_equal(fh[4], fd[4]), "f64 boxes differ"
    assert bytes(fh[0]) == bytes(fd[0]) and bytes(fh[1]) == bytes(fd[1]) and bytes(fh[5]) == bytes(fd[5])
    t = (host.build_timing["tree"], dev.build_timing["tree"], dev.build_timing["tree_device"])
    host.close()
    dev.close()
    return t


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("heuristic", [BvhHeuristic.Sah(1000), BvhHeuristic.Sah(7), BvhHeuristic.Midpoint()],
                         ids=["sah1000", "sah7", "midpoint"])
def test_device_tree_equals_host_tree(case, heuristic, native_built):
    _same_tree(CASES[case](), heuristic)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("splits", [2, 3, 10, 257, 1000, 5000])
def test_device_tree_on_lattice_centres(seed, splits, native_built):
    _same_tree(_lattice_spheres(seed) + _random_spheres(60, 100 + seed), BvhHeuristic.Sah(splits))


@pytest.mark.parametrize("shape", [(300, 150), (1000, 500), (2000, 1000)], ids=["90k", "1M", "4M"])
def test_device_tree_on_the_config_meshes(shape, native_built):
    nu, nv = shape
    spec = scenes.copper_torus(nu, nv, 32, 32) if nu < 2000 else scenes.mixed_scene(nu, nv, 32, 32)
    th, td, tdev = _same_tree(spec.objects, spec.heuristic)
    print(f"[{2 * nu * nv} triangles] tree build: host {th:.3f} s, device path {td:.3f} s (of which on the device {tdev:.3f} s)")
