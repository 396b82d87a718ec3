"""Host BVH builder + flattening (C++ host mirror) against the oracle's literal restatement of
rayrs-lib/src/bvh.rs:227-389.  CPU only (Scene(upload=False) never touches a GPU)."""
import numpy as np
import pytest

import oracle
from rayrs_b200 import _ffi, scenes
from rayrs_b200.api import Axis, BvhHeuristic, Emission, Fresnel, Image, Material, Object, Scene, build_tables

HDRI = Image(2, 2, np.ones((2, 2, 3)))


@pytest.fixture(scope="module", autouse=True)
def _built(native_built):
    return native_built


def _both(objects, heuristic):
    tables = build_tables(objects)
    host = Scene(objects, 1e-6, 1e6, heuristic, HDRI, upload=False)
    orc = oracle.OracleScene(tables, HDRI.pixels, heuristic=(heuristic.kind, heuristic.splits), build_mode=0)
    return host, orc


def _random_spheres(n, seed):
    rng = np.random.default_rng(seed)
    m = Material.lambertian_diffuse((0.5, 0.5, 0.5))
    return [Object.sphere(float(rng.uniform(0.05, 0.5)), rng.uniform(-5, 5, 3), m) for _ in range(n)]


def _random_triangles(n, seed, scale=0.3):
    rng = np.random.default_rng(seed)
    base = rng.uniform(-4, 4, (n, 1, 3))
    tris = base + rng.uniform(-scale, scale, (n, 3, 3))
    return [Object.from_triangles(tris, Material.lambertian_diffuse((0.5, 0.5, 0.5)))]


CASES = {
    "single_sphere_scene": lambda: scenes.diffuse_single_sphere(32, 32).objects,
    "seven_spheres": lambda: scenes.cook_torrance_spheres_metallic(32, 32).objects,
    "material_test": lambda: scenes.material_test(32, 32).objects,
    "random_spheres_300": lambda: _random_spheres(300, 1),
    "random_triangles_3000": lambda: _random_triangles(3000, 2),
    "torus_40x20": lambda: scenes.copper_torus(40, 20, 32, 32).objects,
    "box_and_spheres": lambda: Object.box_geom((-1, 0, -1), (1, 2, 1), Material.no_reflect()) + _random_spheres(9, 3),
    # ties: many identical centres (stable sort order + median fallback bvh.rs:279-287)
    "coincident_centres": lambda: [Object.sphere(0.1 + 0.01 * i, (0.0, 0.0, 0.0), Material.no_reflect()) for i in range(11)],
    "two_clusters_equal_keys": lambda: [Object.sphere(0.2, (float(i % 2) * 3.0, 0.0, 0.0), Material.no_reflect()) for i in range(13)],
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("heuristic", [BvhHeuristic.Sah(1000), BvhHeuristic.Sah(7), BvhHeuristic.Midpoint()],
                         ids=["sah1000", "sah7", "midpoint"])
def test_host_tree_equals_reference_tree(case, heuristic):
    host, orc = _both(CASES[case](), heuristic)
    _, _, _, topo_h, boxes_h, _ = host.flat()
    topo_o, boxes_o = orc.tree_dump()
    assert np.array_equal(topo_h, topo_o)
    assert np.array_equal(boxes_h, boxes_o)  # bit-identical f64 boxes


def _lattice_spheres(seed, n=700):
    """centres on a coarse lattice: equal keys, thresholds that land exactly on centres, long runs between two
    thresholds — what the distinct-split walk of the host builder (few primitives, many thresholds) has to get right"""
    rng = np.random.default_rng(seed)
    m = Material.no_reflect()
    pts = rng.integers(0, 9, (n, 3)).astype(np.float64) * np.array([0.5, 0.125, 2.0])
    return [Object.sphere(float(rng.choice([0.05, 0.1, 0.25])), p, m) for p in pts]


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("splits", [2, 3, 10, 257, 1000, 5000])
def test_distinct_split_walk_equals_threshold_loop(seed, splits):
    objects = _lattice_spheres(seed) + _random_spheres(60, 100 + seed)
    host, orc = _both(objects, BvhHeuristic.Sah(splits))
    _, _, _, topo_h, boxes_h, _ = host.flat()
    topo_o, boxes_o = orc.tree_dump()  # build_mode=0: the literal loop over every threshold
    assert np.array_equal(topo_h, topo_o)
    assert np.array_equal(boxes_h, boxes_o)


def test_parallel_sort_path_builds_the_same_tree():
    """90k triangles: the root ranges are above the builder's 64K threshold for chunked sorting on spare threads +
    stable merges.  Against the oracle's fast build (itself pinned to the literal one below)."""
    spec = scenes.copper_torus(300, 150, 32, 32)
    host = Scene(spec.objects, 1e-6, 1e6, spec.heuristic, HDRI, upload=False)
    assert host.n_prims == 90001
    orc = oracle.OracleScene(spec.tables(), HDRI.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
    _, _, _, topo_h, boxes_h, _ = host.flat()
    topo_o, boxes_o = orc.tree_dump()
    assert np.array_equal(topo_h, topo_o)
    assert np.array_equal(boxes_h, boxes_o)


def test_oracle_fast_build_equals_literal():
    objects = _random_triangles(5000, 5)
    tables = build_tables(objects)
    a = oracle.OracleScene(tables, HDRI.pixels, build_mode=0).tree_dump()
    b = oracle.OracleScene(tables, HDRI.pixels, build_mode=1).tree_dump()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def _walk(nodes):
    """yield (node index, child slot, ref) over the reachable flat tree"""
    stack = [0]
    seen = set()
    while stack:
        f = stack.pop()
        assert f not in seen
        seen.add(f)
        for ch, ref in ((0, nodes[f].ref0), (1, nodes[f].ref1)):
            yield f, ch, ref
            if ref != _ffi.RRS_REF_EMPTY and not (ref & _ffi.RRS_REF_LEAF):
                stack.append(ref)


@pytest.mark.parametrize("case", ["seven_spheres", "random_triangles_3000", "torus_40x20", "box_and_spheres"])
def test_flatten_invariants(case):
    objects = CASES[case]()
    host = Scene(objects, 1e-6, 1e6, BvhHeuristic.Sah(1000), HDRI, upload=False)
    nodes, nodes64, order, topo, boxes, prims = host.flat()
    n_prims = host.n_prims
    assert sorted(order.tolist()) == list(range(n_prims))  # a permutation
    assert [p.obj_id for p in prims] == order.tolist()
    # DFS leaf order == order of leaves in the pre-order topology dump
    assert [int(x) for x in topo if x >= 0] == order.tolist()
    covered = []
    for f, ch, ref in _walk(nodes):
        n32, n64 = nodes[f], nodes64[f]
        lo32 = (n32.lo1 if ch else n32.lo0)[:]
        hi32 = (n32.hi1 if ch else n32.hi0)[:]
        lo64 = (n64.lo1 if ch else n64.lo0)[:]
        hi64 = (n64.hi1 if ch else n64.hi0)[:]
        assert (n64.ref1 if ch else n64.ref0) == ref
        if ref == _ffi.RRS_REF_EMPTY:
            assert all(l > h for l, h in zip(lo32, hi32))  # inverted box: never accepted
            continue
        bare = (n32.flags >> ch) & 1
        if not bare:
            # fp32 box rounded outward from the reference's f64 box
            assert all(np.float64(l) <= d for l, d in zip(lo32, lo64))
            assert all(np.float64(h) >= d for h, d in zip(hi32, hi64))
            assert all(np.float64(np.nextafter(np.float32(l), np.float32(np.inf))) > np.float64(d) for l, d in zip(lo32, lo64))  # tight
        if ref & _ffi.RRS_REF_LEAF:
            first, count = ref & 0x0FFFFFFF, ((ref >> 28) & 7) + 1
            assert 1 <= count <= 4 and first + count <= n_prims
            if bare:
                assert count == 1
            covered.extend(range(first, first + count))
    assert host.dead_nodes > 0 or sorted(covered) == list(range(n_prims))  # every primitive reachable exactly once
    assert host.max_depth >= 1


def test_zero_extent_group_is_dead():
    # a leaf group of coplanar axis-aligned planes has a flat Node box -> never entered
    m = Material.no_reflect()
    flat = [Object.plane(Axis.Y, -1 + i, i, -1, 1, 0.0, m) for i in range(3)]  # 3 planes, all at y = 0
    spheres = [Object.sphere(0.3, (10.0 + i, 2.0, 0.0), m) for i in range(6)]
    host = Scene(flat + spheres, 1e-6, 1e6, BvhHeuristic.Sah(1000), HDRI, upload=False)
    assert host.dead_nodes >= 1
    # and the oracle agrees that those planes are invisible
    orc = oracle.OracleScene(build_tables(flat + spheres), HDRI.pixels)
    ids, _ = orc.intersect(np.array([[0.5, 5.0, 0.0, 0.0, -1.0, 0.0]]))
    assert ids[0] == -1


def test_primitive_tables_number_materials_by_first_appearance():
    """Scene::init builds the primitive / material / emission tables on all host threads (chunk tables merged, then
    renumbered): the result must be what the one-thread walk gives — primitives in DFS leaf order, identical materials
    sharing one entry, entries numbered by first appearance along that order.  mixed_scene(400, 200) spans three
    65536-object chunks with eight materials."""
    for spec in (scenes.material_test(), scenes.emissive_room(), scenes.mixed_scene(400, 200, 64, 64)):
        host = spec.scene(HDRI, upload=False)
        t = build_tables(spec.objects)
        _, _, order, _, _, prims = host.flat()
        rec = np.frombuffer(prims, dtype=np.dtype([("type", "<u4"), ("obj", "<u4"), ("mat", "<u4"), ("emi", "<i4"), ("v", "<f8", 9)]))
        assert (rec["obj"] == order).all()
        assert (rec["type"] == t.objs[order, 0]).all()
        tri = rec["type"] == 2  # a triangle's nine values are its vertices as given; spheres / planes keep derived values
        assert (rec["v"][tri] == t.objs[order, 3:12][tri]).all()
        seen_m, seen_e = {}, {}
        for k, o in enumerate(order):
            key = tuple(t.mats[int(t.objs[o, 1])])
            assert rec["mat"][k] == seen_m.setdefault(key, len(seen_m))
            ei = int(t.objs[o, 2])
            assert rec["emi"][k] == (-1 if ei < 0 else seen_e.setdefault(tuple(t.emis[ei]), len(seen_e)))
        assert host.n_materials == len(seen_m)


def test_constructor_panics_mirror_reference():
    m = Material.no_reflect()
    with pytest.raises(ValueError):  # geometry.rs:97
        Scene([Object.sphere(-1.0, (0, 0, 0), m)], 1e-6, 1e6, BvhHeuristic.Midpoint(), HDRI, upload=False)
    with pytest.raises(ValueError):  # geometry.rs:205-212
        Scene([Object.plane(Axis.Z, 1, -1, 1, -1, 0, m)], 1e-6, 1e6, BvhHeuristic.Midpoint(), HDRI, upload=False)
    with pytest.raises(ValueError):  # lib.rs:235
        Scene([Object.sphere(1.0, (0, 0, 0), m)], 1.0, 0.5, BvhHeuristic.Midpoint(), HDRI, upload=False)
    with pytest.raises(ValueError):  # bvh.rs:229
        Scene([], 1e-6, 1e6, BvhHeuristic.Midpoint(), HDRI, upload=False)
    with pytest.raises(ValueError):  # material.rs:706
        Scene([Object.sphere(1.0, (0, 0, 0), Material.cook_torrance((1, 1, 1), 0.0, Fresnel.schlick_metallic((1, 1, 1))))],
              1e-6, 1e6, BvhHeuristic.Midpoint(), HDRI, upload=False)
    with pytest.raises(ValueError):  # material.rs:597
        Scene([Object.sphere(1.0, (0, 0, 0), Material.lambertian_diffuse((1.5, 0, 0)))], 1e-6, 1e6,
              BvhHeuristic.Midpoint(), HDRI, upload=False)


def test_camera_matches_oracle_and_reference_doctest():
    from rayrs_b200.api import Camera
    c = Camera((1, 1, 1), (0, 1, 0), (0, 0, 0), 90.0, 20.0, 10.0, 90)
    assert c.x_pixels() == 4580 and c.y_pixels() == 2290  # lib.rs:150-151,172-173
    assert np.array_equal(c.derived17(), oracle.camera_new([1, 1, 1], [0, 1, 0], [0, 0, 0], 90.0, 20.0, 10.0, 90))
    with pytest.raises(ValueError):  # lib.rs:108
        Camera((1, 1, 1), (0, 1, 0), (0, 0, 0), 180.0, 20.0, 10.0, 90)
    with pytest.raises(ValueError):  # lib.rs:111
        Camera((1, 1, 1), (0, 1, 0), (1, 1, 1), 90.0, 20.0, 10.0, 90)
    for (w, h) in [(512, 512), (1024, 1024), (1920, 1080), (3840, 2160), (975, 549), (37, 19)]:
        fw, fh = scenes.film(w, h)
        c = Camera((0, 5, 10), (0, 1, 0), (0, 1, 0), 50.0, fw, fh, scenes.PPI)
        assert (c.x_pixels(), c.y_pixels()) == (w, h)
