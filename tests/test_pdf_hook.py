"""The reference's dormant next-event-estimation hook: Material::evaluate with pdf = Some(Pdf::Hittable(light))
(material.rs:91-109 -> LambertianDiffuse::scatter :259-281 -> Pdf::Mix :951-959,1028-1034 -> Pdf::Hittable :943-950,1027 ->
Hittable::{intersect, area, sample} geometry.rs:106-152,229-299,359-387).  radiance() never passes a pdf (lib.rs:532), so no
image depends on it; the hook is held to the oracle directly.

CPU: the oracle's restatement against closed forms and against an independent numpy restatement written from the
reference text.  GPU: rrs_material_evaluate_pdf (f64, the reference's operation order) against the oracle."""
import numpy as np
import pytest

import oracle
from rayrs_b200.api import Axis, BvhHeuristic, Fresnel, Material, Object, Scene

ALBEDO = (0.8, 0.5, 0.25)
MATS = {
    "lambertian": Material.lambertian_diffuse(ALBEDO),
    "plastic": Material.plastic(ALBEDO, (1, 1, 1), 0.25, 1.45),
    "ct_copper": Material.cook_torrance((1, 1, 1), 0.05, Fresnel.schlick_metallic((0.722, 0.451, 0.2))),
    "glass": Material.glass((0.8, 0.9, 1.0), 1.45),
}
# the lights: one primitive of each kind above the shaded region (positions are drawn in [-1, 1]^2 x [-0.5, 0.5])
LIGHTS = {
    "sphere": lambda m: Object.sphere(0.75, (0.5, 3.25, -0.5), m),
    "plane": lambda m: Object.plane(Axis.YRev, -1.5, 1.0, -0.75, 1.25, 4.0, m),
    "plane_x": lambda m: Object.plane(Axis.X, 1.0, 4.0, -2.0, 2.0, -3.0, m),
    "triangle": lambda m: Object.triangle((-2.0, -3.0, -1.0), (2.5, -3.5, -1.5), (0.25, -2.5, 2.0), m),
}


def _tables():
    """one scene holding every material on a sphere row plus the four lights; returns (objects, material index by name,
    object index of each light)"""
    objs = [Object.sphere(0.25, (10.0 + i, 0.0, 0.0), m) for i, m in enumerate(MATS.values())]
    light_obj = {}
    for k, make in LIGHTS.items():
        light_obj[k] = len(objs)
        objs.append(make(MATS["lambertian"]))
    return objs, light_obj


def _light_record(sc, obj_id):
    prims = sc.flat()[5]
    for p in prims:
        if p.obj_id == obj_id:
            return p
    raise AssertionError("light not among the flattened primitives")


def _inputs(n, seed):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1.0, 1.0, (n, 3)) * np.array([1.0, 0.5, 1.0])
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    view = rng.normal(size=(n, 3))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    flip = np.sum(nrm * view, axis=1) < 0
    view[flip] *= -1
    # f32-representable inputs: the pass-through arms run the fp32 production shading
    q = np.concatenate([pos, nrm, view], axis=1).astype(np.float32).astype(np.float64)
    u = rng.integers(0, 1 << 24, (n, 4)) / float(1 << 24)
    return q, u


# ------------------------------------------------------------------------------------------------------------------
# an independent numpy restatement, written from the reference text (not from oracle.cpp)
# ------------------------------------------------------------------------------------------------------------------
def _np_basis(n):
    a = np.abs(n[:, 0]) > np.abs(n[:, 1])
    e1 = np.where(a[:, None], np.stack([n[:, 2], 0 * n[:, 0], -n[:, 0]], 1), np.stack([0 * n[:, 0], n[:, 2], -n[:, 1]], 1))
    e1 = e1 / np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(n, e1)
    e2 = e2 / np.linalg.norm(e2, axis=1, keepdims=True)
    return e1, e2


def _np_light(kind, row):
    """(area, sample(u, v) -> points, intersect(o, d) -> (hit, t)) of an object row"""
    if kind == "sphere":
        r, c = row[3], row[4:7]

        def sample(a, b):
            phi = 2 * np.pi * b
            s = 2 * np.sqrt(a * (1 - a))
            return np.stack([np.cos(phi) * s, np.sin(phi) * s, 1 - 2 * a], 1) * r + c

        def intersect(o, d):
            od = o - c
            A = np.sum(d * d, 1)
            B = 2 * np.sum(d * od, 1)
            Cc = np.sum(od * od, 1) - r * r
            disc = B * B - 4 * A * Cc
            with np.errstate(invalid="ignore"):
                sq = np.sqrt(np.where(disc > 0, disc, 0))
                t1, t2 = (-B - sq) / (2 * A), (-B + sq) / (2 * A)
            t = np.where(t1 < 0, t2, t1)
            return (disc > 0) & ~((t1 < 0) & (t2 < 0)), t
        return 4 * np.pi * r * r, sample, intersect
    if kind.startswith("plane"):
        axis, u0, u1, v0, v1, p = int(row[3]) >> 1, *row[4:9]
        ui, vi = [(1, 2), (0, 2), (0, 1)][axis]

        def sample(a, b):
            out = np.empty((a.size, 3))
            out[:, axis] = p
            out[:, ui] = a * (u1 - u0) + u0
            out[:, vi] = b * (v1 - v0) + v0
            return out

        def intersect(o, d):
            with np.errstate(divide="ignore", invalid="ignore"):
                t = (p - o[:, axis]) / d[:, axis]
                q = o + d * t[:, None]
                ok = (d[:, axis] != 0) & (u0 <= q[:, ui]) & (q[:, ui] < u1) & (v0 <= q[:, vi]) & (q[:, vi] < v1)
            return ok, t
        return (u1 - u0) * (v1 - v0), sample, intersect
    p1, p2, p3 = row[3:6], row[6:9], row[9:12]
    e1, e2 = p2 - p1, p3 - p1

    def sample(a, b):
        return np.zeros((a.size, 3))

    def intersect(o, d):
        T = o - p1
        P = np.cross(d, e2)
        Q = np.cross(T, e1)
        den = P @ e1
        with np.errstate(divide="ignore", invalid="ignore"):
            dist, uu, vv = (Q @ e2) / den, np.sum(P * T, 1) / den, np.sum(Q * d, 1) / den
        return ~((dist < 0) | (uu < 0) | (vv < 0) | (uu + vv > 1)), dist
    return np.linalg.norm(np.cross(e1, e2)) / 2, sample, intersect


def _np_lambert_with_pdf(albedo, kind, row, q, u3):
    pos, n = q[:, 0:3], q[:, 3:6]
    area, sample, intersect = _np_light(kind, row)
    side = u3[:, 0] < 0.5
    to_light = sample(u3[:, 1], u3[:, 2]) - pos
    to_light = to_light / np.linalg.norm(to_light, axis=1, keepdims=True)
    e1, e2 = _np_basis(n)
    phi = 2 * np.pi * u3[:, 2]
    cosl = (np.cos(phi) * np.sqrt(u3[:, 1]))[:, None] * e1 + (np.sin(phi) * np.sqrt(u3[:, 1]))[:, None] * e2 + np.sqrt(1 - u3[:, 1])[:, None] * n
    l = np.where(side[:, None], to_light, cosl)
    nl = np.sum(n * l, 1)
    hit, t = intersect(pos, l)
    with np.errstate(divide="ignore", invalid="ignore"):
        hv = np.where(hit, (t * t * np.sum(l * l, 1)) / (nl * area), 0.0)
        pdf = np.where(nl < 0, np.inf, 0.5 * hv + 0.5 * nl / np.pi)
        color = np.asarray(albedo)[None, :] / np.pi * nl[:, None] / pdf[:, None]
    return l, color, pdf, hit, side


# ------------------------------------------------------------------------------------------------------------------
# CPU: the oracle
# ------------------------------------------------------------------------------------------------------------------
def test_hittable_area_and_sample_closed_forms(native_built):
    objs, light_obj = _tables()
    rows = {k: objs[i].rows[0] for k, i in light_obj.items()}
    a, p = oracle.hittable_area_sample(rows["sphere"], [0.3, 0.6])
    assert abs(a - 4 * np.pi * 0.75 ** 2) < 1e-14
    assert abs(np.linalg.norm(p - np.array([0.5, 3.25, -0.5])) - 0.75) < 1e-14          # on the sphere
    assert abs(p[2] - (-0.5 + 0.75 * (1 - 2 * 0.3))) < 1e-15                             # z = 1 - 2u
    a, p = oracle.hittable_area_sample(rows["plane"], [0.25, 0.5])
    assert a == 2.5 * 2.0 and np.array_equal(p, [-1.5 + 0.25 * 2.5, 4.0, -0.75 + 0.5 * 2.0])  # Axis::Y: (u, pos, v)
    a, p = oracle.hittable_area_sample(rows["plane_x"], [0.5, 0.25])
    assert a == 3.0 * 4.0 and np.array_equal(p, [-3.0, 2.5, -1.0])                       # Axis::X: (pos, u, v)
    a, p = oracle.hittable_area_sample(rows["triangle"], [0.9, 0.1])
    assert np.array_equal(p, [0.0, 0.0, 0.0])                                            # geometry.rs:385-387, as written
    a2, _ = oracle.hittable_area_sample(Object.triangle((-1, 0, 0), (1, 0, 0), (0, 1, 0), MATS["glass"]).rows[0], [0, 0])
    assert a2 == 1.0 and a > 0


@pytest.mark.parametrize("kind", sorted(LIGHTS))
def test_oracle_lambertian_with_pdf_matches_numpy_restatement(native_built, kind):
    objs, light_obj = _tables()
    row = objs[light_obj[kind]].rows[0]
    mat_row = _tables_of(objs).mats[list(MATS).index("lambertian")]
    q, u = _inputs(50_000, 7)
    o = oracle.material_evaluate_pdf(mat_row, row, q, u)
    l, color, pdf, hit, side = _np_lambert_with_pdf(ALBEDO, kind, row, q, u[:, :3])
    assert (o[:, 0] == 1).all()                                   # LambertianDiffuse::scatter always scatters
    # the numpy restatement sums dot products in another order than vecmath.rs: well away from the decision
    # boundaries (n.l = 0, the light's silhouette) the two agree to rounding
    assert np.abs(o[:, 4:7] - l).max() < 1e-12
    stable = np.abs(np.sum(q[:, 3:6] * l, 1)) > 1e-9
    if kind != "triangle":
        assert hit[side & stable].mean() > 0.99                    # a direction sampled on the light meets the light
    err = np.abs(o[:, 1:4] - color) / np.maximum(np.abs(color), 1.0)
    err = err[stable & np.isfinite(color).all(axis=1)]
    assert np.quantile(err, 0.999) < 1e-9 and np.median(err) < 1e-14, (np.quantile(err, 0.999), np.median(err))
    below = stable & (np.sum(q[:, 3:6] * l, 1) < 0)
    assert below.any() and np.all(o[below, 1:4] == 0.0)            # pdf = INFINITY below the surface -> weight 0
    # the cosine side missing the light: pdf = n.l / 2 pi, so the weight is exactly twice the albedo
    miss = stable & ~side & ~hit & ~below
    assert miss.sum() > 1000
    assert np.abs(o[miss, 1:4] / np.asarray(ALBEDO) - 2.0).max() < 1e-12
    # a direction that meets the light carries the extra density: the weight is strictly below 2 x albedo
    lit = stable & hit & ~below
    assert lit.sum() > 1000 and np.all(o[lit, 1] < 2.0 * ALBEDO[0])


def test_oracle_reproduces_pdf_hook_golden(native_built):
    """tests/golden/pdf_hook_golden.npz (tests/golden/make_golden.py): the restatement frozen against silent drift"""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "pdf_hook_golden.npz")
    objs, light_obj = _tables()
    mats = _tables_of(objs).mats
    q, u = _inputs(512, 5)
    assert np.array_equal(q, g["q"]) and np.array_equal(u, g["u"])
    for kind in sorted(LIGHTS):
        for mname in ("lambertian", "plastic"):
            got = oracle.material_evaluate_pdf(mats[list(MATS).index(mname)], objs[light_obj[kind]].rows[0], q, u)
            assert np.array_equal(got, g[f"{kind}/{mname}"], equal_nan=True), (kind, mname)


def _tables_of(objs):
    from rayrs_b200.api import build_tables
    return build_tables(objs)


def test_oracle_materials_that_ignore_the_pdf(native_built):
    """every arm but the diffuse lobe names its parameter `_pdf`: same event as evaluate(..., None)"""
    objs, light_obj = _tables()
    t = _tables_of(objs)
    q, u = _inputs(20_000, 11)
    row = objs[light_obj["sphere"]].rows[0]
    for name in ("ct_copper", "glass"):
        m = t.mats[list(MATS).index(name)]
        a = oracle.material_evaluate_pdf(m, row, q, u)
        b = oracle.material_evaluate(m, q[:, 3:9], u[:, :3])
        assert np.array_equal(a, b, equal_nan=True), name
    # Plastic: the specular lobe ignores it, the diffuse lobe is the Lambertian arm one draw later (material.rs:575-591)
    mp = t.mats[list(MATS).index("plastic")]
    ml = t.mats[list(MATS).index("lambertian")]
    a = oracle.material_evaluate_pdf(mp, row, q, u)
    b = oracle.material_evaluate(mp, q[:, 3:9], u[:, :3])
    r0 = ((1 - 1.45) / (1 + 1.45)) ** 2
    fres = r0 + (1 - r0) * (1 - np.sum(q[:, 3:6] * q[:, 6:9], 1)) ** 5
    spec = u[:, 0] < fres
    assert 0.02 < spec.mean() < 0.9
    assert np.array_equal(a[spec], b[spec], equal_nan=True)
    shifted = np.concatenate([u[:, 1:], np.zeros((u.shape[0], 1))], axis=1)
    c = oracle.material_evaluate_pdf(ml, row, q, shifted)
    assert np.array_equal(a[~spec], c[~spec], equal_nan=True)


# ------------------------------------------------------------------------------------------------------------------
# CPU: the device code itself.  csrc/nee_f64.cuh is __host__ __device__; tests/native/nee_host_check.cu compiles it for
# the host (nvcc, no GPU needed), so the code k_material_evaluate_pdf runs is held to the oracle here as well.
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def host_check(native_built, tmp_path_factory):
    import ctypes as C
    import shutil
    import subprocess
    from pathlib import Path
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    src = Path(__file__).resolve().parent / "native" / "nee_host_check.cu"
    so = tmp_path_factory.mktemp("nee") / "nee_host_check.so"
    subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
                    "-o", str(so), str(src)], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    lib = C.CDLL(str(so))
    lib.nee_host_lambert.argtypes = [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p]
    lib.nee_host_area_sample.argtypes = [C.c_void_p] * 3
    return lib


@pytest.fixture(scope="module")
def host_scene(hdri_small):
    objs, light_obj = _tables()
    sc = Scene(objs, 1e-6, 1e6, BvhHeuristic.Sah(1000), hdri_small, upload=False)
    yield sc, objs, light_obj
    sc.close()


@pytest.mark.parametrize("kind", sorted(LIGHTS))
def test_device_f64_code_on_the_host_matches_oracle(host_check, host_scene, kind):
    """the light's record comes out of the host mirror's flattening (RrsPrim), as a C-ABI caller would pass it"""
    import ctypes as C
    sc, objs, light_obj = host_scene
    light = _light_record(sc, light_obj[kind])
    row = objs[light_obj[kind]].rows[0]
    a4 = np.zeros(4)
    u2 = np.array([0.3, 0.6])
    host_check.nee_host_area_sample(C.byref(light), u2.ctypes.data, a4.ctypes.data)
    area, point = oracle.hittable_area_sample(row, u2)
    assert a4[0] == area and np.array_equal(a4[1:], point)
    if kind == "plane":
        # the reference's own two tests of this code, geometry.rs:829-845 (test_plane_area, test_plane_sample), on the device's
        # functions: Plane::new(Axis::Z, -1, 1, -1, 1, 0) has area 4 and samples inside its ranges at z = 0
        from rayrs_b200 import _ffi
        kat = _ffi.RrsPrim()
        kat.type = 1
        kat.v[0], kat.v[1], kat.v[2], kat.v[3], kat.v[4], kat.v[5] = 4.0, -1.0, 1.0, -1.0, 1.0, 0.0   # RrsAxis Z = 4
        for uu in np.random.default_rng(3).random((100, 2)):
            host_check.nee_host_area_sample(C.byref(kat), uu.ctypes.data, a4.ctypes.data)
            assert a4[0] == 4.0 and -1.0 <= a4[1] < 1.0 and -1.0 <= a4[2] < 1.0 and a4[3] == 0.0
    q, u = _inputs(100_000, 21)
    u3 = np.ascontiguousarray(u[:, :3])
    got = np.zeros((q.shape[0], 7))
    albedo = np.asarray(ALBEDO, dtype=np.float64)
    host_check.nee_host_lambert(albedo.ctypes.data, C.byref(light), q.ctypes.data, u3.ctypes.data, q.shape[0], got.ctypes.data)
    mat_row = sc.tables.mats[list(MATS).index("lambertian")]
    want = oracle.material_evaluate_pdf(mat_row, row, q, np.concatenate([u3, np.zeros((q.shape[0], 1))], axis=1))
    # same operations, same libm, no contraction on either side: bit for bit, NaNs included
    assert np.array_equal(got, want, equal_nan=True), float(np.nanmax(np.abs(got - want)))


# ------------------------------------------------------------------------------------------------------------------
# GPU: rrs_material_evaluate_pdf against the oracle
# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def scene(hdri_small):
    objs, light_obj = _tables()
    sc = Scene(objs, 1e-6, 1e6, BvhHeuristic.Sah(1000), hdri_small)
    yield sc, objs, light_obj
    sc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(LIGHTS))
def test_device_lambertian_with_pdf_matches_oracle(scene, kind):
    sc, objs, light_obj = scene
    idx = list(MATS).index("lambertian")
    light = _light_record(sc, light_obj[kind])
    q, u = _inputs(100_000, 21)
    g = sc.material_evaluate_pdf(idx, light, q, u)
    o = oracle.material_evaluate_pdf(sc.tables.mats[idx], objs[light_obj[kind]].rows[0], q, u)
    assert np.array_equal(g[:, 0], o[:, 0])
    # directions: the same f64 operations; device sin / cos / sqrt differ from libm in the last place at most
    assert np.abs(g[:, 4:7] - o[:, 4:7]).max() < 1e-12
    # weights: decisions (n.l < 0, hit / miss of the light) can only flip within rounding of their boundaries
    nl = np.sum(q[:, 3:6] * o[:, 4:7], 1)
    stable = np.abs(nl) > 1e-9
    err = np.abs(g[:, 1:4] - o[:, 1:4]) / np.maximum(np.abs(o[:, 1:4]), 1.0)
    fin = np.isfinite(o[:, 1:4]).all(axis=1) & np.isfinite(g[:, 1:4]).all(axis=1)
    assert (fin | ~stable).mean() > 0.999
    # the device keeps the albedo in fp32 (DMat): 6e-8 relative; everything else is f64
    assert np.quantile(err[stable & fin], 0.9999) < 1e-6, np.quantile(err[stable & fin], 0.9999)
    assert np.median(err[stable & fin]) < 1e-7
    below = stable & (nl < 0)
    assert np.all(g[below, 1:4] == 0.0)
    print(f"[pdf hook, {kind}] direction max diff {np.abs(g[:, 4:7] - o[:, 4:7]).max():.1e}, weight rel err median "
          f"{np.median(err[stable & fin]):.1e} p99.99 {np.quantile(err[stable & fin], 0.9999):.1e}, below surface {int(below.sum())}")


@pytest.mark.gpu
def test_device_plastic_and_pass_through_arms(scene):
    sc, objs, light_obj = scene
    light = _light_record(sc, light_obj["plane"])
    row = objs[light_obj["plane"]].rows[0]
    q, u = _inputs(100_000, 33)
    # arms that ignore the pdf return what rrs_material_evaluate returns (fp32 production shading)
    for name in ("ct_copper", "glass"):
        idx = list(MATS).index(name)
        g = sc.material_evaluate_pdf(idx, light, q, u)
        p = sc.material_evaluate(idx, q[:, 3:9], u[:, :3]).astype(np.float64)
        both = (g[:, 0] == 1) & (p[:, 0] == 1)
        assert (g[:, 0] != p[:, 0]).mean() < 2e-4
        d = np.abs(g[both, 1:7] - p[both, 1:7])
        d = d[np.isfinite(d).all(axis=1)]
        # two translation units (FMA contraction differs): equal to fp32 rounding, not bit for bit
        assert np.quantile(d.max(axis=1), 0.999) < 1e-3, name
    # Plastic: lobe choice by the first draw; the diffuse lobe consumes the pdf with draws 1..3
    idx = list(MATS).index("plastic")
    g = sc.material_evaluate_pdf(idx, light, q, u)
    o = oracle.material_evaluate_pdf(sc.tables.mats[idx], row, q, u)
    r0 = ((1 - 1.45) / (1 + 1.45)) ** 2
    fres = r0 + (1 - r0) * (1 - np.sum(q[:, 3:6] * q[:, 6:9], 1)) ** 5
    clear = np.abs(u[:, 0] - fres) > 1e-5           # the fp32 Schlick term decides within rounding of the boundary
    diffuse = clear & ~(u[:, 0] < fres)
    assert 0.1 < diffuse.mean() < 0.98
    assert np.abs(g[diffuse, 4:7] - o[diffuse, 4:7]).max() < 1e-12
    nl = np.sum(q[:, 3:6] * o[:, 4:7], 1)
    ok = diffuse & (np.abs(nl) > 1e-9) & np.isfinite(o[:, 1:4]).all(axis=1) & np.isfinite(g[:, 1:4]).all(axis=1)
    err = np.abs(g[ok, 1:4] - o[ok, 1:4]) / np.maximum(np.abs(o[ok, 1:4]), 1.0)
    assert np.quantile(err, 0.9999) < 1e-6
    spec = clear & (u[:, 0] < fres) & (g[:, 0] == 1) & (o[:, 0] == 1)
    ddir = np.linalg.norm(g[spec, 4:7] - o[spec, 4:7], axis=1)
    assert np.quantile(ddir, 0.999) < 5e-4          # the bound tests/test_gpu_shading.py holds the fp32 arms to


@pytest.mark.gpu
def test_device_pdf_hook_argument_checks(scene):
    from rayrs_b200 import _ffi
    sc, objs, light_obj = scene
    light = _light_record(sc, light_obj["sphere"])
    q, u = _inputs(4, 1)
    with pytest.raises(_ffi.RayrsError):
        sc.material_evaluate_pdf(999, light, q, u)
    bad = _ffi.RrsPrim()
    bad.type = 7
    with pytest.raises(_ffi.RayrsError):
        sc.material_evaluate_pdf(0, bad, q, u)
    assert sc.material_evaluate_pdf(0, light, q[:0], u[:0]).shape == (0, 7)
