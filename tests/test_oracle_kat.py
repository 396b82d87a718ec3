"""The oracle against every numeric known-answer test the reference holds for the hot path
(SURVEY.md 4 / 8c).  Each case cites the reference test it restates.  CPU only."""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module", autouse=True)
def _built(native_built):
    return native_built


# ---- RNG: Random123 known answers (kat_vectors of the Random123 distribution): Philox4x32 with 7 rounds — what the render
# stream draws — and with 10 rounds, the paper's default
PHILOX_KATS = [
    (7, [0, 0, 0, 0], [0, 0], [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
    (7, [0xffffffff] * 4, [0xffffffff] * 2, [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
    (7, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
    (10, [0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    (10, [0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    (10, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("rounds,ctr,key,expect", PHILOX_KATS)
def test_philox_known_answers(rounds, ctr, key, expect):
    assert list(oracle.philox(ctr, key, rounds)) == expect
    if rounds == 7:
        assert list(oracle.philox(ctr, key)) == expect  # the render stream's default


def test_render_stream_is_philox_7():
    """rng_uniforms(seed, pixel, sample, slot) = the top 24 bits of Philox4x32-7(counter (pixel, sample, slot, 0), key seed)"""
    seed, pixel, sample, slot = 0x299f31d0a4093822, 0x243f6a88, 0x85a308d3, 0x13198a2e
    w = oracle.philox([pixel, sample, slot, 0], [seed & 0xFFFFFFFF, seed >> 32], 7)
    u = oracle.rng_uniforms(seed, pixel, sample, slot, oracle.RNG_MATCHED)
    assert [float(x >> 8) / 16777216.0 for x in w] == list(u)


def test_rng_uniform_ranges():
    for mode in (oracle.RNG_WIDE, oracle.RNG_MATCHED):
        u = np.array([oracle.rng_uniforms(7, p, s, 1, mode) for p in range(50) for s in range(4)])
        assert (u >= 0).all() and (u < 1).all()
        assert 0.4 < u.mean() < 0.6


# ---- bvh.rs:543-559 test_bvh_intersect_node_leafnode: t == 4.0 exactly
def test_bvh_node_leafnode_t_equals_4():
    from rayrs_b200.api import Material, Object, build_tables
    tables = build_tables([Object.sphere(1.0, (0, 0, 0), Material.no_reflect())])
    s = oracle.OracleScene(tables, np.ones((2, 2, 3)), z_near=0.001, z_far=1000.0)
    ids, t = s.intersect(np.array([[-5.0, 0, 0, 1, 0, 0]]))
    assert ids[0] == 0 and t[0] == 4.0


# ---- geometry.rs:744-774 sphere tests
def test_sphere_outside():
    assert oracle.sphere_intersect(1.0, [0, 0, 0], [0, 0, 5, 0, 0, -1]) > 0


def test_sphere_inside():
    assert oracle.sphere_intersect(1.0, [0, 0, 0], [0, 0, 0, 0, 1, 0]) > 0


def test_sphere_miss():
    assert oracle.sphere_intersect(1.0, [0, 0, 0], [0, 5, 0, 0, 1, 0]) is None


def test_sphere_glancing():
    assert oracle.sphere_intersect(1.0, [0, 0, 0], [0.99999, -5, 0, 0, 1, 0]) > 0


# ---- geometry.rs:782-828 plane tests (front and back for each axis)
@pytest.mark.parametrize("axis,ray", [
    (0, [5, 0, 0, -1, 0, 0]), (0, [-5, 0, 0, 1, 0, 0]),
    (2, [0, 5, 0, 0, -1, 0]), (2, [0, -5, 0, 0, 1, 0]),
    (4, [0, 0, 5, 0, 0, -1]), (4, [0, 0, -5, 0, 0, 1]),
])
def test_plane_front_back(axis, ray):
    assert oracle.plane_intersect(axis, -1, 1, -1, 1, 0.0, ray) > 0


def test_plane_half_open_ranges():
    # Range::contains is [start, end): geometry.rs:229-271
    assert oracle.plane_intersect(2, -1, 1, -1, 1, 0.0, [-1.0, 5, 0, 0, -1, 0]) is not None
    assert oracle.plane_intersect(2, -1, 1, -1, 1, 0.0, [1.0, 5, 0, 0, -1, 0]) is None
    # parallel ray: d_k == 0 -> None
    assert oracle.plane_intersect(2, -1, 1, -1, 1, 0.0, [0, 5, 0, 1, 0, 0]) is None
    # any sign of t is returned
    assert oracle.plane_intersect(2, -1, 1, -1, 1, 0.0, [0, 5, 0, 0, 1, 0]) == -5.0


# ---- geometry.rs:847-887 AABB tests
@pytest.mark.parametrize("ray,tmin,expect", [
    ([-5, 0, 0, 1, 0, 0], 0.001, True),
    ([0, -5, 0, 0, 1, 0], 0.0001, True),
    ([0, 0, -5, 0, 0, 1], 0.001, True),
    ([0, 0, 0, 0, 0, 1], 0.001, True),
    ([1.1, 0, 0, 0, 1, 1], 0.001, False),
    ([2, 0, 0, -1, -2, 0], 0.001, False),
])
def test_aabb(ray, tmin, expect):
    assert oracle.aabb_intersect([-1, 1, -1, 1, -1, 1], ray, tmin, 1000.0) is expect


def test_aabb_zero_extent_never_hit():
    # geometry.rs:474,491,508: tmax <= tmin -> false; a flat box has tmin == tmax on its flat axis
    assert oracle.aabb_intersect([-1, 1, 0, 0, -1, 1], [0, 5, 0, 0.1, -1, 0.1], 1e-6, 1e6) is False


# ---- doctests lib.rs:150-151,172-173: pixel counts
def test_camera_pixel_counts():
    c = oracle.camera_new([1, 1, 1], [0, 1, 0], [0, 0, 0], 90.0, 20.0, 10.0, 90)
    assert c[15] == 4580 and c[16] == 2290
    assert c[14] == 229  # ppc = round(90 * 2.54)


# ---- doctests geometry.rs:540-542,575,607,638: box of two unit spheres at x = +-1
def test_bbox_two_spheres():
    from rayrs_b200.api import Material, Object, build_tables
    m = Material.no_reflect()
    tables = build_tables([Object.sphere(1.0, (1, 0, 0), m), Object.sphere(1.0, (-1, 0, 0), m)])
    s = oracle.OracleScene(tables, np.ones((2, 2, 3)))
    b = s.bbox()
    assert b[0] == -2.0 and b[1] == 2.0
    assert tuple(b[6:9]) == (0.0, 0.0, 0.0)
    assert b[9] == 16.0 and b[10] == 40.0


# ---- doctest vecmath.rs:336-339: right-handed basis
def test_orthonormal_basis_handedness():
    e1, e2 = oracle.orthonormal_basis([0, 0, 1])
    assert np.allclose(np.cross(e1, e2), [0, 0, 1], atol=0, rtol=0)
    for n in ([1, 0, 0], [0, 1, 0], [0.6, 0.0, 0.8], [-0.3, 0.9, np.sqrt(1 - 0.09 - 0.81)]):
        e1, e2 = oracle.orthonormal_basis(n)
        assert abs(np.dot(e1, n)) < 1e-15 and abs(np.dot(e2, n)) < 1e-15
        assert np.allclose(np.cross(e1, e2), n, atol=1e-15)


# ---- triangle (unpinned by the reference; closed-form cases)
def test_triangle_basic():
    tri = [-1, 0, 0, 1, 0, 0, 0, 1, 0]
    t, n = oracle.triangle_intersect(tri, [0, 0.25, 5, 0, 0, -1])
    assert t == 5.0 and tuple(n) == (0.0, 0.0, 1.0)
    t, _ = oracle.triangle_intersect(tri, [0, 0.25, -5, 0, 0, 1])  # two sided
    assert t == 5.0
    t, _ = oracle.triangle_intersect(tri, [0, 2.0, 5, 0, 0, -1])
    assert t is None
    t, _ = oracle.triangle_intersect(tri, [0, 0.25, 5, 0, 0, 1])  # behind
    assert t is None
    t, _ = oracle.triangle_intersect(tri, [0, 0.25, 5, 1, 0, 0])  # parallel: NaN/inf quotients
    assert t is None or not np.isfinite(t)


# ---- the conditioning probe the fp32 parity tests lean on
def test_intersect_sensitivity_is_the_stability_probe_plus_conditioning():
    from rayrs_b200 import scenes
    hdri = scenes.synthetic_hdri(64, 32)
    spec = scenes.copper_torus(24, 12, 64, 48)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
    cam = spec.camera()
    rng = np.random.default_rng(3)
    n = 4096
    rays = oracle.primary_rays(cam.derived17(), 64, 48, rng.integers(0, 48, n), rng.integers(0, 64, n), rng.integers(0, 64, n))
    oid, ot = osc.intersect(rays)
    stable = osc.intersect_stable(rays)
    st2, tchange = osc.intersect_sensitivity(rays)
    assert np.array_equal(stable, st2)
    assert np.all(tchange[~stable] == 0) and np.all(tchange[oid < 0] == 0)
    hit = stable & (oid >= 0)
    assert np.all(tchange[hit] <= 1e-3) and tchange[hit].max() > 0
    # a head-on hit of the floor moves by about the perturbation itself; the probe must see that scale
    floor = hit & (oid == 0)
    assert floor.any() and np.median(tchange[floor]) < 1e-4


# ---- bvh.rs:436-541: four construction tests the reference carries COMMENTED OUT.  Their expected values are the author's
# own `{:?}` dumps of `Bvh::build(BvhHeuristic::Midpoint, ..)`; they are not run by `cargo test`, but they are the only
# statement of a tree SHAPE the reference makes — and the code as it stands still produces them: the oracle's build (both
# of its build modes), the fast host build of the C++ mirror and its flattening reproduce every one.
def _dump(objects, mode):
    from rayrs_b200.api import build_tables
    s = oracle.OracleScene(build_tables(objects), np.ones((2, 2, 3)), heuristic=(0, 0), build_mode=mode)
    topo, boxes = s.tree_dump()
    s.close()
    return topo.tolist(), boxes


def _host_dump(objects):
    from rayrs_b200.api import BvhHeuristic, Image, Scene
    sc = Scene(objects, 1e-6, 1e6, BvhHeuristic.Midpoint(), Image(2, 2, np.ones((2, 2, 3))), upload=False)
    flat = sc.flat()
    topo, boxes = flat[3].tolist(), flat[4].copy()
    sc.close()
    return topo, boxes


@pytest.mark.parametrize("axis", [0, 1, 2], ids=["x", "y", "z"])
def test_commented_out_midpoint_construction_along_an_axis(axis):
    """bvh.rs:435-487: 8 unit spheres at -10.5 + 3 i along one axis -> Node(all, [Node(left four), Node(right four)]),
    boxes -11.5..11.5, -11.5..-0.5 and 0.5..11.5 on that axis, -1..1 on the others"""
    from rayrs_b200.api import Material, Object
    objects = []
    for i in range(8):
        c = [0.0, 0.0, 0.0]
        c[axis] = -10.5 + 3.0 * i
        objects.append(Object.sphere(1.0, tuple(c), Material.no_reflect()))
    want_topo = [-2, -4, 0, 1, 2, 3, -4, 4, 5, 6, 7]
    want_boxes = np.tile(np.array([-1.0, 1.0] * 3), (3, 1))
    want_boxes[:, 2 * axis:2 * axis + 2] = [[-11.5, 11.5], [-11.5, -0.5], [0.5, 11.5]]
    for topo, boxes in (_dump(objects, 0), _dump(objects, 1), _host_dump(objects)):
        assert topo == want_topo
        assert np.array_equal(boxes, want_boxes)


def test_commented_out_midpoint_construction_sphere_in_center():
    """bvh.rs:489-541: three walls, a light and a triangle inside a 5 x 5 x 5 box (all three extents tie: the x axis is
    taken) -> Node(box, [Node(box, [top, bottom]), Node(z -2.5..0.8, [back, light, triangle])]); the dump also spells out
    Triangle::new's derived fields: e1 (0.5, 0, -0.5), e2 (2, 2, -1), unit normal (2/3, -1/3, 2/3), area 0.75"""
    from rayrs_b200.api import Axis, Emission, Material, Object
    nr = Material.no_reflect()
    top = Object.plane(Axis.YRev, -2.5, 2.5, -2.5, 2.5, 2.5, nr)
    bottom = Object.plane(Axis.Y, -2.5, 2.5, -2.5, 2.5, -2.5, nr)
    back = Object.plane(Axis.Z, -2.5, 2.5, -2.5, 2.5, -2.5, nr)
    light = Object.plane(Axis.YRev, -0.8, 0.8, -0.8, 0.8, 2.4999, nr, Emission.new(5.0, (1, 1, 1)))
    tri = Object.triangle((-1, -1, -0.5), (-0.5, -1, -1), (1, 1, -1.5), nr)
    objects = [top, bottom, back, light, tri]
    want_boxes = np.array([[-2.5, 2.5, -2.5, 2.5, -2.5, 2.5], [-2.5, 2.5, -2.5, 2.5, -2.5, 2.5], [-2.5, 2.5, -2.5, 2.5, -2.5, 0.8]])
    for topo, boxes in (_dump(objects, 0), _dump(objects, 1), _host_dump(objects)):
        assert topo == [-2, -2, 0, 1, -3, 2, 3, 4]
        assert np.array_equal(boxes, want_boxes)
    _, normal = oracle.triangle_intersect([-1, -1, -0.5, -0.5, -1, -1, 1, 1, -1.5], [0, 0, 5, 0, 0, -1])
    assert list(normal) == [0.6666666666666666, -0.3333333333333333, 0.6666666666666666]  # the dump's digits
    area, _ = oracle.hittable_area_sample(tri.rows[0], [0.0, 0.0])
    assert area == 0.75


# ---- geometry.rs:829-845 test_plane_area / test_plane_sample (the two live reference tests of Hittable::area / sample,
# restated with the dormant next-event-estimation hook: tests/test_pdf_hook.py)
def test_plane_area_and_sample():
    from rayrs_b200.api import Axis, Material, Object
    row = Object.plane(Axis.Z, -1.0, 1.0, -1.0, 1.0, 0.0, Material.no_reflect()).rows[0]
    rng = np.random.default_rng(3)
    for u in rng.random((200, 2)):
        area, s = oracle.hittable_area_sample(row, u)
        assert area == 4.0
        assert -1.0 <= s[0] < 1.0 and -1.0 <= s[1] < 1.0 and s[2] == 0.0   # Range::contains is half-open


# ---- vecmath.rs:816-893: the reference's unit tests of the Vec3 operators (cross1 / cross2 fix the handedness everything
# from orthonormal_basis to the camera frame relies on); powf and clip are the two operators of the output stage
def test_vecmath_unit_tests():
    ops = oracle.vecmath_ops([1, 2, 3], [2, 4, 6], 3.0)
    assert list(ops["add"]) == [3, 6, 9]                                  # test_add
    assert list(oracle.vecmath_ops([4, 3, 2], [1, 1, 1], 1.0)["sub"]) == [3, 2, 1]      # test_sub
    assert list(oracle.vecmath_ops([1, 4, 8], [2, 2, 2], 1.0)["mul"]) == [2, 8, 16]     # test_mul
    assert list(ops["scalar_mul"]) == [3, 6, 9] and list(ops["mul_scalar"]) == [3, 6, 9]  # test_scalar_mul, test_mul_scalar
    assert oracle.vecmath_ops([1, 2, 3], [1, 2, 3], 1.0)["dot"] == 14.0     # test_dot
    assert list(oracle.vecmath_ops([1, 0, 0], [0, 1, 0], 1.0)["cross"]) == [0, 0, 1]    # test_cross1
    assert list(oracle.vecmath_ops([1, 0, 0], [0, 0, 1], 1.0)["cross"]) == [0, -1, 0]   # test_cross2
    assert ops["mag2"] == 14.0                                             # test_mag2
    # test_pow (1, 2, 3).powf(2) == (1, 4, 9) and test_clip (-1, 2, 0.5).clip(0, 1) == (0, 1, 0.5), through the restated
    # output stage that applies them (image.rs:193-222): clip then powf(gamma) then (255.99 x) as u8
    by, census = oracle.to_raw_bytes(np.array([[[-1.0, 2.0, 0.5]]]), gamma=1.0)
    assert by[0, 0].tolist() == [0, 255, 127] and census == {"clamped": 1, "nan": 0, "negative": 1}
    by2, _ = oracle.to_raw_bytes(np.array([[[0.1, 0.2, 0.3]]]), gamma=2.0)
    assert by2[0, 0].tolist() == [int(255.99 * 0.1 ** 2), int(255.99 * 0.2 ** 2), int(255.99 * 0.3 ** 2)]
