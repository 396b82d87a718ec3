"""A second, independent restatement (numpy, written from the reference text) of the pieces of the path that no
reference test pins — triangle intersection, camera rays, background lookup, the Lambertian and Cook-Torrance
arms of Material::evaluate — against the C++ oracle.  Two restatements written separately and agreeing to rounding
is the strongest pin available for them without a Rust toolchain.  CPU only."""
import numpy as np
import pytest

import oracle
from rayrs_b200 import scenes
from rayrs_b200.api import Fresnel, Material, Object, build_tables


@pytest.fixture(scope="module", autouse=True)
def _built(native_built):
    return native_built


def unit(v):
    return v / np.sqrt(v @ v)


# ---- Triangle::new + Triangle::intersect, geometry.rs:341-375
def tri_ref(p1, p2, p3, o, d):
    e1, e2 = p2 - p1, p3 - p1
    t = o - p1
    p = np.cross(d, e2)
    q = np.cross(t, e1)
    den = p @ e1
    with np.errstate(divide="ignore", invalid="ignore"):
        dist, u, v = (q @ e2) / den, (p @ t) / den, (q @ d) / den
    if dist < 0.0 or u < 0.0 or v < 0.0 or u + v > 1.0:
        return None
    return dist


def test_triangle_restatements_agree():
    rng = np.random.default_rng(1)
    hits = 0
    for _ in range(4000):
        p = rng.uniform(-1, 1, (3, 3))
        o = rng.uniform(-3, 3, 3)
        target = p[0] + rng.uniform(-0.2, 1.2) * (p[1] - p[0]) + rng.uniform(-0.2, 1.2) * (p[2] - p[0])
        d = (target - o) * rng.uniform(0.2, 3.0) * rng.choice([1.0, 1.0, 1.0, -1.0])
        want = tri_ref(p[0], p[1], p[2], o, d)
        got, _ = oracle.triangle_intersect(p.ravel(), np.concatenate([o, d]))
        if want is None or got is None:
            if (want is None) != (got is None):   # only within rounding of an edge
                e1, e2 = p[1] - p[0], p[2] - p[0]
                den = np.cross(d, e2) @ e1
                u = (np.cross(d, e2) @ (o - p[0])) / den
                v = (np.cross(o - p[0], e1) @ d) / den
                assert min(abs(u), abs(v), abs(u + v - 1)) < 1e-9
            continue
        hits += 1
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
    assert hits > 500


# ---- Camera::new + generate_primary_ray, lib.rs:99-133,202-210; index mapping of rayrs/src/main.rs:71-76
def test_camera_and_primary_rays_restatement():
    origin, up, lookat = np.array([0.3, 5.0, 10.0]), np.array([0.0, 1.0, 0.0]), np.array([0.0, 1.0, -0.5])
    fov, W, H, ppi = 50.0, 96, 64, 100
    width, height = W / 254.0, H / 254.0
    c = oracle.camera_new(origin, up, lookat, fov, width, height, ppi)
    z = unit(lookat - origin)
    x = unit(np.cross(up, z))
    y = unit(np.cross(z, x))
    ppc = round(ppi * 2.54)
    assert np.allclose(c[3:6], x, rtol=0, atol=1e-15) and np.allclose(c[6:9], y, rtol=0, atol=1e-15)
    zs = (width / np.tan(np.radians(fov) / 2.0)) * z                     # the FOV quirk: width, not width / 2
    assert np.allclose(c[9:12], zs, rtol=1e-15, atol=0) and c[14] == ppc and (c[15], c[16]) == (W, H)
    rng = np.random.default_rng(2)
    rows, cols, smp = rng.integers(0, H, 200), rng.integers(0, W, 200), rng.integers(0, 64, 200)
    seed = 0x5EEDB200
    rays = oracle.primary_rays(c, W, H, rows, cols, smp, seed, oracle.RNG_WIDE)
    for k in range(200):
        u = oracle.rng_uniforms(seed, int(rows[k]) * W + int(cols[k]), int(smp[k]), 0, oracle.RNG_WIDE)
        i, j = float(H - rows[k]), float(W - cols[k])                    # main.rs:71-76
        xx = (j + u[0]) / ppc - width / 2.0
        yy = (i + u[1]) / ppc - height / 2.0
        assert np.allclose(rays[k, :3], origin, rtol=0, atol=0)
        assert np.allclose(rays[k, 3:], zs + xx * x + yy * y, rtol=1e-14, atol=1e-16)


# ---- Scene::background, lib.rs:254-285 (+ Image::pixel image.rs:183-186)
def test_background_restatement():
    hdri = scenes.synthetic_hdri(64, 32)
    spec = scenes.diffuse_single_sphere(16, 16)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    rng = np.random.default_rng(3)
    d = rng.normal(size=(3000, 3)) * rng.uniform(0.1, 5.0, (3000, 1))
    got = osc.background(d)
    px = np.asarray(hdri.pixels, dtype=np.float64).reshape(hdri.height, hdri.width, 3)
    for k in range(0, 3000, 7):
        u = unit(d[k])
        phi = np.arctan2(u[2], u[0]) + np.pi
        theta = np.arccos(u[1])
        x = phi / (2.0 * np.pi) * (hdri.width - 1)
        y = theta / np.pi * (hdri.height - 1)
        xf, xc, yf, yc = np.floor(x), np.ceil(x), np.floor(y), np.ceil(y)
        i, j = int(yf), int(xf)
        f = [px[i, j], px[min(i + 1, hdri.height - 1), j], px[i, min(j + 1, hdri.width - 1)],
             px[min(i + 1, hdri.height - 1), min(j + 1, hdri.width - 1)]]
        want = f[0] * (xc - x) * (yc - y) + f[1] * (xc - x) * (y - yf) + f[2] * (x - xf) * (yc - y) + f[3] * (x - xf) * (y - yf)
        assert np.allclose(got[k], want, rtol=1e-11, atol=1e-13)
    osc.close()


# ---- Material::evaluate: LambertianDiffuse (material.rs:259-281,982-993) and CookTorrance (:403-424,721-758,915-941,
#      1006-1020,1276-1322,1472-1496), literally: brdf * cos / pdf * dwh/dwi
def basis(n):
    e1 = unit(np.array([n[2], 0.0, -n[0]])) if abs(n[0]) > abs(n[1]) else unit(np.array([0.0, n[2], -n[1]]))
    return e1, unit(np.cross(n, e1))


def lambertian_ref(color, n, v, u):
    e1, e2 = basis(n)
    phi = 2.0 * np.pi * u[1]
    l = np.cos(phi) * np.sqrt(u[0]) * e1 + np.sin(phi) * np.sqrt(u[0]) * e2 + np.sqrt(1.0 - u[0]) * n
    brdf = color / np.pi
    return 1.0, brdf * (n @ l) / ((n @ l) / np.pi), l


def cook_torrance_ref(color, alpha, r0, n, v, u):
    a2 = alpha * alpha
    e1, e2 = basis(n)
    phi = 2.0 * np.pi * u[0]
    tan2 = -a2 * np.log(1.0 - u[1])
    cost = 1.0 / np.sqrt(1.0 + tan2)
    sint = np.sqrt(1.0 - cost * cost)
    h = np.cos(phi) * sint * e1 + np.sin(phi) * sint * e2 + cost * n
    l = 2.0 * (v @ h) * h - v                                        # reflect(halfway, view)
    # Pdf::Beckmann value (the cos(theta_h) factor is missing in the reference: SURVEY.md F2)
    hv_vec = l + v
    if not hv_vec.any():
        pdf = 1.0
    else:
        hh = unit(hv_vec)
        nh = abs(n @ hh)
        tan_t = np.tan(np.arccos(nh))
        pdf = 1.0 if np.isinf(tan_t) else np.exp(-tan_t * tan_t / a2) / (np.pi * a2 * nh ** 4)
    if h @ v < 0.0:
        return 0.0, np.zeros(3), l
    nl_s = n @ l
    if nl_s < 0.0:
        return 0.0, np.zeros(3), l
    # brdf
    nv, nl = abs(n @ v), abs(n @ l)
    if nv == 0.0 or nl == 0.0 or not hv_vec.any():
        return 0.0, np.zeros(3), l
    hh = unit(hv_vec)
    nh = n @ hh
    tan_t = np.tan(np.arccos(nh))
    if np.isinf(tan_t):
        return 0.0, np.zeros(3), l
    beck = np.exp(-tan_t * tan_t / a2) / (np.pi * a2 * nh ** 4)
    hv = hh @ v
    g = min(2.0 * nh * nv / hv, min(2.0 * nh * nl / hv, 1.0))
    fres = r0 + (1.0 - r0) * (1.0 - hh @ v) ** 5                      # schlick_vec(r0, halfway, view)
    brdf = color * fres * beck * g / (4.0 * nv * nl)
    col = brdf * nl_s / pdf * (4.0 * (h @ l))
    if not col.any():
        return 0.0, np.zeros(3), l
    return 1.0, col, l


@pytest.mark.parametrize("alpha", [0.01, 0.05, 0.25, 0.5])
def test_cook_torrance_and_lambertian_restatements(alpha):
    rng = np.random.default_rng(int(alpha * 1000))
    r0 = np.array([0.722, 0.451, 0.2])
    ct = Material.cook_torrance((1.0, 0.9, 0.8), alpha, Fresnel.schlick_metallic(tuple(r0)))
    lam = Material.lambertian_diffuse((0.8, 0.7, 0.6))
    rows = build_tables([Object.sphere(1.0, (0, 0, 0), ct), Object.sphere(1.0, (3, 0, 0), lam)]).mats
    n_s = 1500
    nrm = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    view = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    flip = np.einsum("ij,ij->i", nrm, view) < 0
    view[flip] *= -1.0                                                # views on the outside (the shipped scenes' case)
    u = rng.random((n_s, 3))
    nv = np.concatenate([nrm, view], axis=1)
    got_ct = oracle.material_evaluate(rows[0], nv, u)
    got_l = oracle.material_evaluate(rows[1], nv, u)
    scattered = 0
    for k in range(n_s):
        f, c, l = cook_torrance_ref(np.array([1.0, 0.9, 0.8]), alpha, r0, nrm[k], view[k], u[k])
        assert got_ct[k, 0] == f, k
        if f:
            scattered += 1
            assert np.allclose(got_ct[k, 4:7], l, rtol=0, atol=1e-12)
            assert np.allclose(got_ct[k, 1:4], c, rtol=1e-9, atol=1e-14), (k, got_ct[k, 1:4], c)
        f, c, l = lambertian_ref(np.array([0.8, 0.7, 0.6]), nrm[k], view[k], u[k])
        assert got_l[k, 0] == 1.0 and np.allclose(got_l[k, 4:7], l, rtol=0, atol=1e-12)
        assert np.allclose(got_l[k, 1:4], c, rtol=1e-12, atol=0)
    assert scattered > 0.3 * n_s


# ---- Glass (material.rs:339-401 with Reflect::brdf :1254-1266, Refract::btdf :1333-1352, refract :1502-1518) and
#      Plastic (:567-593, Plastic::new :887-900): Fresnel-selected branches
def schlick(ior_curr, ior_new, n, v):
    r0 = ((ior_curr - ior_new) / (ior_curr + ior_new)) ** 2
    return r0 + (1.0 - r0) * (1.0 - n @ v) ** 5


def refract_ref(n, v, ratio):
    cos_t = v @ n
    sin_t = np.sqrt(1.0 - cos_t * cos_t)
    if ratio * sin_t > 1.0:
        return None
    par = ratio * (cos_t * n - v)
    return -np.sqrt(1.0 - par @ par) * n + par


def glass_ref(color, ior, n, v, u):
    cos_t = n @ v
    entering = cos_t > 0.0
    nn = n if entering else -n
    ratio = 1.0 / ior if entering else ior
    sin2 = 1.0 - cos_t * cos_t

    def reflect_branch():
        l = 2.0 * (v @ nn) * nn - v
        return 1.0, (color / abs(nn @ l)) * (nn @ l), l

    if ratio * ratio * sin2 >= 1.0:
        return reflect_branch()
    fres = schlick(1.0, ior, nn, v) if entering else schlick(ior, 1.0, nn, v)
    if u[0] < fres:
        return reflect_branch()
    l = refract_ref(nn, v, ratio)
    assert l is not None
    btdf = np.zeros(3) if l @ v > 0.0 else color / abs(nn @ l)
    return 1.0, btdf * abs(nn @ l), l


def plastic_ref(color, spec, alpha, ior, n, v, u):
    fres = schlick(1.0, ior, n, v)
    if u[0] < fres:
        f, c, l = cook_torrance_dielectric_ref(spec, alpha, ior, n, v, u[1:])
        return f, (c / fres if f else c), l
    return lambertian_ref(color, n, v, u[1:])


def cook_torrance_dielectric_ref(color, alpha, ior, n, v, u):
    """cook_torrance_ref with Fresnel::SchlickDielectric(ior).value(halfway, view, Entering) (material.rs:1457-1468)"""
    f, c, l = cook_torrance_ref(color, alpha, np.ones(3), n, v, u)   # r0 = 1 -> Fresnel factor 1: divide it back in below
    if not f:
        return f, c, l
    hh = unit(v + l)
    return f, c * schlick(1.0, ior, hh, v), l


def test_glass_and_plastic_restatements():
    rng = np.random.default_rng(9)
    glass = Material.glass((0.9, 1.0, 0.8), 1.45)
    plastic = Material.plastic((0.2, 0.5, 0.8), (1.0, 0.9, 0.7), 0.1, 1.45)
    rows = build_tables([Object.sphere(1.0, (0, 0, 0), glass), Object.sphere(1.0, (3, 0, 0), plastic)]).mats
    n_s = 3000
    nrm = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    view = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])      # both sides of the surface: entering and exiting glass
    u = rng.random((n_s, 3))
    nv = np.concatenate([nrm, view], axis=1)
    got_g = oracle.material_evaluate(rows[0], nv, u)
    outside = np.einsum("ij,ij->i", nrm, view) > 0
    got_p = oracle.material_evaluate(rows[1], nv, u)
    branches = set()
    for k in range(n_s):
        f, c, l = glass_ref(np.array([0.9, 1.0, 0.8]), 1.45, nrm[k], view[k], u[k])
        assert got_g[k, 0] == f
        assert np.allclose(got_g[k, 4:7], l, rtol=0, atol=1e-12) and np.allclose(got_g[k, 1:4], c, rtol=1e-12, atol=1e-15)
        branches.add(("refl" if l @ nrm[k] * (1 if outside[k] else -1) > 0 else "refr", bool(outside[k])))
        if outside[k]:   # Plastic is an opaque material: the reference only ever sees it from outside
            f, c, l = plastic_ref(np.array([0.2, 0.5, 0.8]), np.array([1.0, 0.9, 0.7]), 0.1, 1.45, nrm[k], view[k], u[k])
            assert got_p[k, 0] == f, k
            if f:
                assert np.allclose(got_p[k, 4:7], l, rtol=0, atol=1e-12)
                assert np.allclose(got_p[k, 1:4], c, rtol=1e-9, atol=1e-14), (k, got_p[k, 1:4], c)
    assert len(branches) == 4   # reflection and refraction, entering and exiting


# ---- CookTorranceGlass (material.rs:469-565): MicrofacetDistribution::Beckmann generate (:1137-1161), reflection through
#      CookTorrance::evaluate_reflection / brdf with the dielectric Fresnel, refraction through evaluate_refraction
#      (:764-812) / CookTorrance::btdf (:1362-1442)
def ct_brdf_dielectric(color, a2, ior, n, l, v):
    nv, nl = abs(n @ v), abs(n @ l)
    hvec = v + l
    if nv == 0.0 or nl == 0.0 or not hvec.any():
        return np.zeros(3)
    hh = unit(hvec)
    nh = n @ hh
    tan_t = np.tan(np.arccos(nh))
    if np.isinf(tan_t):
        return np.zeros(3)
    beck = np.exp(-tan_t * tan_t / a2) / (np.pi * a2 * nh ** 4)
    hv = hh @ v
    g = min(2.0 * nh * nv / hv, min(2.0 * nh * nl / hv, 1.0))
    return color * schlick(1.0, ior, hh, v) * beck * g / (4.0 * nv * nl)     # brdf always asks the Fresnel as "Entering"


def ct_btdf(color, a2, ior, n, l, v, entering):
    nv, nl = abs(n @ v), abs(n @ l)
    ratio = 1.0 / ior if entering else ior
    hvec = l + ratio * v if ratio > 1.0 else (-ratio) * v - l
    if nv == 0.0 or nl == 0.0 or not hvec.any():
        return np.zeros(3)
    hh = unit(hvec)
    nh = n @ hh
    tan_t = np.tan(np.arccos(nh))
    if np.isinf(tan_t):
        return np.zeros(3)
    beck = np.exp(-tan_t * tan_t / a2) / (np.pi * a2 * nh ** 4)
    hl, hv = abs(hh @ l), abs(hh @ v)
    g = min(2.0 * nh * nv / hv, min(2.0 * nh * nl / hv, 1.0))
    denom = (ratio * hv + hl) ** 2
    fres = schlick(1.0, ior, hh, v) if entering else schlick(ior, 1.0, hh, v)
    return color * (1.0 - fres) * beck * g * (hv * hl / (nv * nl)) * ratio * ratio / denom


def ct_glass_ref(color, alpha, ior, n, v, u):
    a2 = alpha * alpha
    e1, e2 = basis(n)                                               # sampled about the UNflipped normal
    phi = 2.0 * np.pi * u[0]
    tan2 = -a2 * np.log(1.0 - u[1])
    cost = 1.0 / np.sqrt(1.0 + tan2)
    sint = np.sqrt(1.0 - cost * cost)
    h = np.cos(phi) * sint * e1 + np.sin(phi) * sint * e2 + cost * n
    pdf = np.exp(-tan2 / a2) / (np.pi * a2 * (n @ h) ** 4)
    entering = n @ v > 0.0
    if not entering:
        h, n = -h, -n
    cos_t = h @ v
    ratio = 1.0 / ior if entering else ior
    sin2 = 1.0 - cos_t * cos_t

    def reflection(div):
        l = 2.0 * (v @ h) * h - v
        if h @ v < 0.0 or n @ l < 0.0:
            return 0.0, np.zeros(3), l
        col = ct_brdf_dielectric(color, a2, ior, n, l, v) * (n @ l) / pdf * (4.0 * (h @ l))
        return (1.0, col / div, l) if col.any() else (0.0, np.zeros(3), l)

    if ratio * ratio * sin2 >= 1.0:
        return reflection(1.0)
    fres = schlick(1.0, ior, h, v) if entering else schlick(ior, 1.0, h, v)
    if u[2] < fres:
        return reflection(fres)
    l = refract_ref(h, v, ratio)
    assert l is not None
    nl = n @ l
    if h @ v < 0.0 or nl > 0.0:
        return 0.0, np.zeros(3), l
    hl, hv = abs(h @ l), abs(h @ v)
    dwh = hl / (ratio * hv + hl) ** 2
    col = ct_btdf(color, a2, ior, n, l, v, entering) * abs(nl) / (ratio * ratio) / (pdf * dwh)
    return (1.0, col / (1.0 - fres), l) if col.any() else (0.0, np.zeros(3), l)


@pytest.mark.parametrize("alpha", [0.05, 0.25])
def test_cook_torrance_glass_restatement(alpha):
    rng = np.random.default_rng(int(alpha * 100) + 7)
    color = np.array([1.0, 0.95, 0.9])
    mat = Material.cook_torrance_glass(tuple(color), alpha, 1.45)
    row = build_tables([Object.sphere(1.0, (0, 0, 0), mat)]).mats[0]
    n_s = 4000
    nrm = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    view = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])      # entering and exiting
    u = rng.random((n_s, 3))
    got = oracle.material_evaluate(row, np.concatenate([nrm, view], axis=1), u)
    kinds = set()
    for k in range(n_s):
        f, c, l = ct_glass_ref(color, alpha, 1.45, nrm[k], view[k], u[k])
        assert got[k, 0] == f, (k, got[k], f, c)
        if f:
            side = np.sign((nrm[k] @ view[k]) * (nrm[k] @ l))
            kinds.add((side > 0, nrm[k] @ view[k] > 0))
            assert np.allclose(got[k, 4:7], l, rtol=0, atol=1e-11)
            assert np.allclose(got[k, 1:4], c, rtol=1e-8, atol=1e-13), (k, got[k, 1:4], c)
    assert len(kinds) == 4   # reflected and transmitted, from outside and from inside


# ---- the whole path: main.rs:61-94 (pixel loop, F8 index mapping) -> generate_primary_ray -> radiance (lib.rs:521-560:
#      closest hit with the leaf filter t > tmin && t < tmax, emission only in the Scatter arm, Russian roulette from
#      bounce 0 with true division) -> Sphere / Plane intersect (geometry.rs:106-132, 229-271) -> background,
#      as a pure-Python path tracer on the diffuse_single_sphere scene, consuming the oracle's counter-based uniforms
def background_ref(px, W, H, d):
    u = unit(d)
    phi = np.arctan2(u[2], u[0]) + np.pi
    theta = np.arccos(u[1])
    x, y = phi / (2.0 * np.pi) * (W - 1), theta / np.pi * (H - 1)
    xf, xc, yf, yc = np.floor(x), np.ceil(x), np.floor(y), np.ceil(y)
    i, j = int(yf), int(xf)
    return (px[i, j] * (xc - x) * (yc - y) + px[i + 1, j] * (xc - x) * (y - yf) + px[i, j + 1] * (x - xf) * (yc - y)
            + px[i + 1, j + 1] * (x - xf) * (y - yf))


def sphere_ref(radius2, c, o, d):
    od = o - c
    a, b, cc = d @ d, 2.0 * (d @ od), od @ od - radius2
    desc = b * b - 4.0 * a * cc
    if not desc > 0.0:
        return None
    t1, t2 = (-b - np.sqrt(desc)) / (2.0 * a), (-b + np.sqrt(desc)) / (2.0 * a)
    if t1 < 0.0:
        return None if t2 < 0.0 else t2
    return t1


def floor_ref(o, d):   # Plane(Axis::Y, -25..25, -25..25, pos 0): half-open ranges, any sign of t
    if d[1] == 0.0:
        return None
    t = (0.0 - o[1]) / d[1]
    p = o + t * d
    return t if (-25.0 <= p[0] < 25.0 and -25.0 <= p[2] < 25.0) else None


def test_whole_path_restatement():
    W, H, spp, max_bounces, seed = 12, 8, 3, 8, 0x5EEDB200
    hdri = scenes.synthetic_hdri(64, 32)
    px = np.asarray(hdri.pixels, dtype=np.float64).reshape(32, 64, 3)
    spec = scenes.diffuse_single_sphere(W, H)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    cam = spec.camera().derived17()
    want, st = osc.render(cam, W, H, spp, max_bounces=max_bounces, seed=seed, nthreads=1)
    origin, e_x, e_y, zs, width, height, ppc = cam[0:3], cam[3:6], cam[6:9], cam[9:12], cam[12], cam[13], cam[14]
    tmin, tmax = 1e-6, 1e6
    got = np.zeros((H, W, 3))
    rays = 0
    for row in range(H):
        for col in range(W):
            pixel = row * W + col
            acc = np.zeros(3)
            for s in range(spp):
                u = oracle.rng_uniforms(seed, pixel, s, 0)
                x = (float(W - col) + u[0]) / ppc - width / 2.0
                y = (float(H - row) + u[1]) / ppc - height / 2.0
                o, d = origin.copy(), zs + x * e_x + y * e_y
                thr, light = np.ones(3), np.zeros(3)
                for b in range(max_bounces):
                    rays += 1
                    best, which = None, None
                    for name, t in (("floor", floor_ref(o, d)), ("sphere", sphere_ref(1.0, np.array([0.0, 1.0, 0.0]), o, d))):
                        if t is not None and t > tmin and t < tmax and (best is None or t < best):
                            best, which = t, name
                    if best is None:
                        light = light + thr * background_ref(px, 64, 32, d)
                        break
                    pos = o + best * d
                    view = unit(-1.0 * d)
                    uu = oracle.rng_uniforms(seed, pixel, s, b + 1)
                    if which == "sphere":
                        f, color, l = lambertian_ref(np.array([0.8, 0.8, 0.8]), unit(pos - np.array([0.0, 1.0, 0.0])), view, uu)
                    else:
                        f, color, l = cook_torrance_ref(np.ones(3), 0.5, np.array([0.8, 0.8, 0.8]), np.array([0.0, 1.0, 0.0]), view, uu)
                    if not f:
                        break
                    thr = thr * color                       # Emission::Dark: nothing to gather
                    p = max(thr)
                    if uu[3] > p:
                        break
                    thr = thr / p
                    o, d = pos, l
                acc += light
            got[row, col] = acc / spp
    assert rays == st["rays"]
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12), np.abs(got - want).max()
    osc.close()


# ---- the three "next" arms: Reflect (material.rs:283-303), Refract (:305-337), CookTorranceRefract (:426-467)
def test_reflect_refract_ct_refract_restatements():
    rng = np.random.default_rng(21)
    col = np.array([0.9, 0.8, 0.7])
    mats = [Material.reflect(tuple(col)), Material.refract(tuple(col), 1.45), Material.cook_torrance_refract(tuple(col), 0.2, 1.45)]
    rows = build_tables([Object.sphere(1.0, (3.0 * k, 0, 0), m) for k, m in enumerate(mats)]).mats
    n_s = 3000
    nrm = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    view = np.array([unit(x) for x in rng.normal(size=(n_s, 3))])
    u = rng.random((n_s, 3))
    nv = np.concatenate([nrm, view], axis=1)
    got = [oracle.material_evaluate(r, nv, u) for r in rows]
    scat = 0
    for k in range(n_s):
        n, v = nrm[k], view[k]
        # Reflect: Dirac pdf (value 1), brdf = color / |n.l|
        l = 2.0 * (v @ n) * n - v
        assert got[0][k, 0] == 1.0 and np.allclose(got[0][k, 4:7], l, atol=1e-13)
        assert np.allclose(got[0][k, 1:4], col / abs(n @ l) * (n @ l), rtol=1e-12)
        # Refract
        entering = n @ v > 0.0
        nn = n if entering else -n
        ratio = 1.0 / 1.45 if entering else 1.45
        l = refract_ref(nn, v, ratio)
        if l is None:
            assert got[1][k, 0] == 0.0
        else:
            btdf = np.zeros(3) if l @ v > 0.0 else col / abs(nn @ l)
            assert got[1][k, 0] == 1.0 and np.allclose(got[1][k, 4:7], l, atol=1e-13)
            assert np.allclose(got[1][k, 1:4], btdf * abs(nn @ l), rtol=1e-12)
        # CookTorranceRefract: half vector about the FLIPPED normal, flipped back for the exiting case
        a2 = 0.2 * 0.2
        e1, e2 = basis(nn)
        phi = 2.0 * np.pi * u[k, 0]
        tan2 = -a2 * np.log(1.0 - u[k, 1])
        cost = 1.0 / np.sqrt(1.0 + tan2)
        sint = np.sqrt(1.0 - cost * cost)
        h = np.cos(phi) * sint * e1 + np.sin(phi) * sint * e2 + cost * nn
        pdf = np.exp(-tan2 / a2) / (np.pi * a2 * (nn @ h) ** 4)
        if not entering:
            h = -h
        l = refract_ref(h, v, ratio)
        if l is None:
            assert got[2][k, 0] == 0.0, k
            continue
        nl = nn @ l
        if h @ v < 0.0 or nl > 0.0:
            assert got[2][k, 0] == 0.0, k
            continue
        hl, hv = abs(h @ l), abs(h @ v)
        dwh = hl / (ratio * hv + hl) ** 2
        c = ct_btdf(col, a2, 1.45, nn, l, v, entering) * abs(nl) / (ratio * ratio) / (pdf * dwh)
        if not c.any():
            assert got[2][k, 0] == 0.0, k
            continue
        scat += 1
        assert got[2][k, 0] == 1.0 and np.allclose(got[2][k, 4:7], l, atol=1e-11)
        assert np.allclose(got[2][k, 1:4], c, rtol=1e-8, atol=1e-13), (k, got[2][k, 1:4], c)
    assert scat > 200
