"""The reference's own unit tests for the hot path, run through the CUDA backend (C ABI: rrs_intersect, both the
production fp32 traversal and the fp64 literal traversal).  Each case cites the reference test it restates; the same
cases pin the oracle in tests/test_oracle_kat.py.  Primitives are put behind a BVH exactly as `Scene::new` does, so the
answers also cross the flattener, the fp16-box nodes / brute-force list and the leaf filter (t > tmin && t < tmax)."""
import numpy as np
import pytest

from rayrs_b200.api import Axis, BvhHeuristic, Image, Material, Object, Scene

pytestmark = pytest.mark.gpu

HDRI = Image(2, 2, np.ones((2, 2, 3)))
M = Material.no_reflect()
# a second, far-away object: a lone Plane / flat triangle would sit in a zero-thickness Node box, which the reference
# slab test never accepts (geometry.rs:474,491,508) — the reference tests call the primitives directly
FAR = Object.sphere(0.5, (300.0, 300.0, 300.0), M)


def _hit(objects, ray, z_near=1e-6, z_far=1e6):
    sc = Scene(objects, z_near, z_far, BvhHeuristic.Sah(1000), HDRI)
    r = np.asarray([ray], dtype=np.float64)
    out = {}
    for prec in (32, 64):
        ids, t = sc.intersect(r, prec)
        out[prec] = (int(ids[0]), float(t[0]))
    sc.close()
    assert out[32][0] == out[64][0]
    if out[64][0] >= 0:
        assert abs(out[32][1] - out[64][1]) <= 1e-5 * abs(out[64][1])
    return out[64]


def test_bvh_intersect_node_leafnode(native_built):
    """bvh.rs:543-559: unit sphere at the origin, ray from (-5,0,0) along +x, (tmin, tmax) = (0.001, 1000): t == 4.0"""
    i, t = _hit([Object.sphere(1.0, (0, 0, 0), M)], [-5.0, 0, 0, 1, 0, 0], 0.001, 1000.0)
    assert i == 0 and t == 4.0


@pytest.mark.parametrize("ray,hits", [
    ([0, 0, 5, 0, 0, -1], True),            # geometry.rs:744-751 outside
    ([0, 0, 0, 0, 1, 0], True),             # :753-759 inside
    ([0, 5, 0, 0, 1, 0], False),            # :761-766 pointing away
    ([0.99999, -5, 0, 0, 1, 0], True),      # :768-774 glancing
])
def test_sphere_cases(native_built, ray, hits):
    i, t = _hit([Object.sphere(1.0, (0, 0, 0), M)], ray)
    assert (i == 0) is hits
    if hits:
        o, d = np.array(ray[:3], dtype=float), np.array(ray[3:], dtype=float)
        assert abs(np.linalg.norm(o + t * d) - 1.0) < 1e-9   # the hit point lies on the sphere (fp64 traversal)


@pytest.mark.parametrize("axis,ray", [
    (Axis.X, [5, 0, 0, -1, 0, 0]), (Axis.X, [-5, 0, 0, 1, 0, 0]),      # geometry.rs:782-828 front / back per axis
    (Axis.Y, [0, 5, 0, 0, -1, 0]), (Axis.Y, [0, -5, 0, 0, 1, 0]),
    (Axis.Z, [0, 0, 5, 0, 0, -1]), (Axis.Z, [0, 0, -5, 0, 0, 1]),
])
def test_plane_front_and_back(native_built, axis, ray):
    i, t = _hit([Object.plane(axis, -1, 1, -1, 1, 0.0, M), FAR], ray)
    assert i == 0 and t == 5.0


def test_plane_half_open_ranges_and_leaf_filter(native_built):
    plane = [Object.plane(Axis.Y, -1, 1, -1, 1, 0.0, M), FAR]
    assert _hit(plane, [-1.0, 5, 0, 0, -1, 0])[0] == 0     # Range::contains is [start, end): geometry.rs:229-271
    assert _hit(plane, [1.0, 5, 0, 0, -1, 0])[0] == -1
    assert _hit(plane, [0, 5, 0, 1, 0, 0])[0] == -1        # parallel
    assert _hit(plane, [0, 5, 0, 0, 1, 0])[0] == -1        # Plane::intersect returns t = -5; the leaf filter drops it (bvh.rs:404-413)
    assert _hit(plane, [0, 5, 0, 0, -1, 0], 1e-6, 4.0)[0] == -1   # beyond tmax
    assert _hit(plane, [0, 5, 0, 0, -1, 0], 1e-6, 6.0) == (0, 5.0)


def test_lone_flat_primitive_is_invisible(native_built):
    """a Node whose box has zero thickness never passes the reference's slab test, so a scene that
    holds only one axis-aligned plane shows nothing — and the backend reproduces that."""
    assert _hit([Object.plane(Axis.Y, -1, 1, -1, 1, 0.0, M)], [0, 5, 0, 0, -1, 0])[0] == -1


def test_triangle_cases(native_built):
    """Unpinned by the reference (no triangle test exists); closed-form cases of geometry.rs:359-375: two-sided,
    inclusive edges, nothing behind the origin."""
    tri = [Object.triangle((-1, 0, 0), (1, 0, 0), (0, 1, 0), M), FAR]
    assert _hit(tri, [0, 0.25, 5, 0, 0, -1]) == (0, 5.0)
    assert _hit(tri, [0, 0.25, -5, 0, 0, 1]) == (0, 5.0)
    assert _hit(tri, [0, 2.0, 5, 0, 0, -1])[0] == -1
    assert _hit(tri, [0, 0.25, 5, 0, 0, 1])[0] == -1
    assert _hit(tri, [0, 0.25, 5, 1, 0, 0])[0] == -1
    assert _hit(tri, [0.0, 0.0, 5, 0, 0, -1]) == (0, 5.0)   # on the edge p1-p2: inclusive
