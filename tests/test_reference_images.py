"""A pin on the reference's OWN OUTPUT: the six renders rayrs ships (examples/*.png, rendered by the reference itself).

Their colours pin nothing (unshipped HDRI, unknown spp) — their geometry does.  Where the spheres' silhouettes fall in the
image depends only on Camera::new (lib.rs:99-133: the F9 field-of-view quirk `z = width / tan(fov / 2)`, the ppc rounding),
Camera::x_pixels / y_pixels, generate_primary_ray (lib.rs:202-210), the mirrored pixel mapping of the tile loop
(main.rs:71-76, F8), the constants of test_scenes.rs:14-44,163-256 and Sphere::intersect — the part of the path SURVEY.md 8c
lists as pinned by no reference test.  tests/golden/reference_image_pins.npz holds the edge maps measured from the PNGs
(tests/golden/make_reference_image_pins.py; the images themselves are not in the repository).

Held to them: the oracle's primary-ray hit mask (CPU) and the CUDA path's (rrs_intersect, GPU), with the reference's exact
camera arguments.  Each test also shows its own power: the same mask shifted by 3 pixels, or built with the textbook
field of view instead of the reference's quirk, or with the sphere row mirrored, does NOT fit.
"""
from pathlib import Path

import numpy as np
import pytest
from scipy.ndimage import binary_erosion, distance_transform_edt

import oracle
from rayrs_b200 import scenes
from rayrs_b200.api import Camera

PINS = np.load(Path(__file__).resolve().parent / "golden" / "reference_image_pins.npz")

# Camera::new arguments exactly as test_scenes.rs:27-35 (single sphere) and :194-202 (sphere rows) write them
SINGLE = dict(origin=(0.0, 5.0, 10.0), up=(0.0, 1.0, 0.0), lookat=(0.0, 1.0, 0.0), fov=50.0, width=1920.0 / 500.0,
              height=1080.0 / 500.0, ppi=100)
ROW = dict(origin=(0.0, 10.0, 20.0), up=(0.0, 1.0, 0.0), lookat=(0.0, 1.0, 0.0), fov=72.0, width=1920.0 / 500.0,
           height=400.0 / 500.0, ppi=125)
CASES = {
    "diffuse_single_sphere": (scenes.diffuse_single_sphere, SINGLE),
    "copper_sphere": (scenes.copper_single_sphere, SINGLE),
    "cook_torrance_glass_sphere": (scenes.cook_torrance_glass_single_sphere, SINGLE),
    "spheres_metallic": (scenes.cook_torrance_spheres_metallic, ROW),
    "spheres_plastic": (scenes.cook_torrance_spheres_plastic, ROW),
    "cook_torrance_spheres_frosted_glass": (scenes.cook_torrance_spheres_frosted_glass, ROW),
}
NEAR = 1.5  # pixels


def _primary_rays(cam17, W, H):
    rows, cols = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return oracle.primary_rays(cam17, W, H, rows.ravel(), cols.ravel(), np.zeros(rows.size))


def _edge_distance(name):
    W, H = (int(v) for v in PINS[f"{name}/size"])
    e = np.zeros((H, W), dtype=bool)
    xy = PINS[f"{name}/edge_xy"].astype(np.int64)
    e[xy[:, 1], xy[:, 0]] = True
    return distance_transform_edt(~e), W, H


def _fit(dist, mask, shift=(0, 0)):
    """fraction of the silhouette's pixels that lie within NEAR pixels of an edge of the reference image"""
    b = mask & ~binary_erosion(mask)
    b = np.roll(np.roll(b, shift[0], 0), shift[1], 1)
    return float((dist[b] <= NEAR).mean())


def _check_silhouettes(name, ids_of):
    builder, cam_args = CASES[name]
    dist, W, H = _edge_distance(name)
    cam = Camera(**cam_args)
    # Camera::x_pixels / y_pixels with the reference's film constants give the size of the shipped render
    assert (cam.x_pixels(), cam.y_pixels()) == (W, H)
    ids = ids_of(builder(W, H), cam.derived17(), W, H)
    spheres = ids >= 1  # object 0 is the floor (test_scenes.rs:22-24,190-191)
    assert 0.05 < spheres.mean() < 0.6
    fit = _fit(dist, spheres)
    assert fit >= 0.80, (name, fit)
    for shift in ((0, 3), (0, -3), (3, 0), (-3, 0), (3, 3), (-3, -3)):
        other = _fit(dist, spheres, shift)
        assert other <= fit - 0.08, (name, shift, other, fit)
    return ids, fit, dist, cam


def _oracle_ids(spec, cam17, W, H):
    osc = oracle.OracleScene(spec.tables(), scenes.synthetic_hdri(64, 32).pixels)
    ids, _ = osc.intersect(_primary_rays(cam17, W, H))
    osc.close()
    return ids.reshape(H, W)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_silhouettes_fit_the_reference_renders(native_built, name):
    ids, fit, dist, cam = _check_silhouettes(name, _oracle_ids)
    # the textbook field of view, z = (width / 2) / tan(fov / 2), instead of the reference's z = width / tan(fov / 2)
    # (lib.rs:131): the silhouettes land elsewhere
    builder, cam_args = CASES[name]
    W, H = cam.x_pixels(), cam.y_pixels()
    c17 = cam.derived17().copy()
    c17[9:12] *= 0.5  # z_scaled
    wrong = _oracle_ids(builder(W, H), c17, W, H) >= 1
    assert _fit(dist, wrong) <= fit - 0.25, (name, _fit(dist, wrong), fit)
    print(f"[{name}] {fit * 100:.1f} % of the silhouette within {NEAR} px of an edge of the reference render; "
          f"textbook fov {_fit(dist, wrong) * 100:.1f} %")


def _sharpness_per_sphere(name, ids):
    sh = PINS[f"{name}/sharpness_4x4"].astype(np.float64)
    hb, wb = sh.shape
    out = []
    for k in range(1, 8):
        inner = binary_erosion(ids == k, iterations=10)
        blocks = inner[:hb * 4, :wb * 4].reshape(hb, 4, wb, 4).all(axis=(1, 3))
        assert blocks.sum() > 300
        out.append(float(sh[blocks].mean()))
    return out


def test_sphere_row_orientation_and_roughness_order(native_built):
    """test_scenes.rs:178-189,213-224: sphere i sits at x = 2.2 (i - 3) with alpha = 0.01 (4 i + 1).  In the reference's
    render the mirror-like end is on the LEFT; the interior of each sphere is less sharp than its smoother neighbour's.
    A mapping mirrored about the vertical axis (a `width - j` dropped from main.rs:75, or e_x flipped) reverses it."""
    cam = Camera(**ROW)
    W, H = cam.x_pixels(), cam.y_pixels()
    ids = _oracle_ids(scenes.cook_torrance_spheres_metallic(W, H), cam.derived17(), W, H)
    centres = [float(np.nonzero((ids == k).any(axis=0))[0].mean()) for k in range(1, 8)]
    assert centres == sorted(centres) and centres[0] < W / 4 and centres[-1] > 3 * W / 4  # sphere 0 (alpha 0.01) on the left
    sharp = _sharpness_per_sphere("spheres_metallic", ids)
    assert all(a > b for a, b in zip(sharp, sharp[1:])), sharp          # strictly decreasing with alpha
    assert sharp[0] > 1.8 * sharp[-1]
    mirrored = _sharpness_per_sphere("spheres_metallic", ids[:, ::-1])
    assert all(a < b for a, b in zip(mirrored, mirrored[1:]))           # the mirrored hypothesis contradicts the image
    plastic = _sharpness_per_sphere("spheres_plastic", _oracle_ids(scenes.cook_torrance_spheres_plastic(W, H), cam.derived17(), W, H))
    assert plastic[0] == max(plastic) and plastic[0] > 1.15 * plastic[-1], plastic
    print("[spheres_metallic] interior sharpness, sphere 0..6:", [round(s, 2) for s in sharp])


def test_mirror_reflections_inside_the_smoothest_sphere(native_built):
    """The leftmost sphere of the metallic row is almost a mirror (alpha = 0.01, test_scenes.rs:213-224): what it shows —
    its neighbour, the floor's horizon, the sky — is drawn by Sphere::normal (geometry.rs:134-136), CookTorrance::scatter at
    the centre of its lobe (half vector = normal: material.rs:403-424,1006-1020 -> reflect, :1492-1496) and the second
    closest-hit query.  The contours of the oracle's second-hit mask INSIDE that sphere (4 px away from its own silhouette)
    coincide with edges of the reference's render; shifted by 2 or 3 pixels they do not."""
    name = "spheres_metallic"
    dist, W, H = _edge_distance(name)
    cam = Camera(**ROW)
    spec = scenes.cook_torrance_spheres_metallic(W, H)
    tables = spec.tables()
    osc = oracle.OracleScene(tables, scenes.synthetic_hdri(64, 32).pixels)
    rays = _primary_rays(cam.derived17(), W, H)
    ids, t = osc.intersect(rays)
    first = np.nonzero(ids == 1)[0]                      # object 1 = sphere 0 at x = -6.6
    o, d = rays[first, :3], rays[first, 3:]
    p = o + d * t[first, None]                           # Ray::point lib.rs:41-43
    c = tables.objs[1, 4:7]
    n = (p - c) / np.linalg.norm(p - c, axis=1, keepdims=True)
    v = -d / np.linalg.norm(d, axis=1, keepdims=True)
    # the oracle's own CookTorrance::scatter with the half-vector draw at the centre of the lobe (tan^2 = -alpha^2 ln(1 - 0) = 0)
    u = np.tile(np.array([0.3, 0.0, 0.5]), (first.size, 1))
    ev = oracle.material_evaluate(tables.mats[int(tables.objs[1, 1])], np.concatenate([n, v], axis=1), u)
    assert (ev[:, 0] == 1).all()
    l = ev[:, 4:7]
    assert np.abs(l - (2 * np.sum(v * n, 1, keepdims=True) * n - v)).max() < 1e-9   # = reflect(n, v)
    second, _ = osc.intersect(np.concatenate([p + 1e-9 * l, l], axis=1))
    osc.close()
    seen = np.full(H * W, -2)
    seen[first] = second
    seen = seen.reshape(H, W)
    assert (seen == 2).sum() > 1000 and (seen == 0).sum() > 5000 and (seen == -1).sum() > 5000  # neighbour, floor, sky
    inner = binary_erosion(ids.reshape(H, W) == 1, iterations=4)
    contour = np.zeros((H, W), dtype=bool)
    for what in (2, 0, -1):
        m = seen == what
        contour |= m & ~binary_erosion(m)
    contour &= inner
    assert contour.sum() > 400

    def fit(shift):
        b = np.roll(np.roll(contour, shift[0], 0), shift[1], 1)
        return float((dist[b] <= NEAR).mean())
    aligned = fit((0, 0))
    assert aligned >= 0.82, aligned
    worst = max(fit(s) for s in ((0, 2), (0, -2), (2, 0), (-2, 0), (0, 3), (0, -3), (3, 0), (-3, 0), (3, 3), (-3, -3), (3, -3), (-3, 3)))
    assert worst <= aligned - 0.08, (aligned, worst)
    print(f"[spheres_metallic, reflections in sphere 0] {aligned * 100:.1f} % of {int(contour.sum())} contour pixels on an edge of the "
          f"reference render; best shifted hypothesis {worst * 100:.1f} %")


def _through_the_glass_sphere(osc, tables, mat_row, rays, ids, t, W, H, straight_through):
    """contours of what the leftmost glass sphere (object 1) shows by transmission, 4 px inside its own silhouette.
    Both refractions are the oracle's own CookTorranceGlass::scatter (material.rs:469-565) evaluated at the centre of its
    lobe (half vector = normal) with the Fresnel draw on the refraction side — i.e. `refract` (material.rs:1502-1518) with
    the direction / ior bookkeeping of :1191-1231.  straight_through: the far hit of the sphere is lost and the ray goes
    on in the ONCE-refracted direction — what the reference's re-entry quirk (geometry.rs:106-132 +
    bvh.rs:404-413) does to part of the rays."""
    first = np.nonzero(ids == 1)[0]
    o, d = rays[first, :3], rays[first, 3:]
    p = o + d * t[first, None]
    c = tables.objs[1, 4:7]
    n = (p - c) / np.linalg.norm(p - c, axis=1, keepdims=True)
    v = -d / np.linalg.norm(d, axis=1, keepdims=True)
    u = np.tile(np.array([0.3, 0.0, 0.999999]), (first.size, 1))
    l1 = oracle.material_evaluate(mat_row, np.concatenate([n, v], axis=1), u)[:, 4:7]
    assert np.all(np.sum(l1 * n, 1) < 0)                        # refracted into the sphere
    t2 = -2 * np.sum(l1 * (p - c), 1) / np.sum(l1 * l1, 1)      # far root from a point on the sphere
    p2 = p + l1 * t2[:, None]
    if straight_through:
        second, _ = osc.intersect(np.concatenate([p2 + 1e-6 * l1, l1], axis=1))
    else:
        n2 = (p2 - c) / np.linalg.norm(p2 - c, axis=1, keepdims=True)
        v2 = -l1 / np.linalg.norm(l1, axis=1, keepdims=True)
        l2 = oracle.material_evaluate(mat_row, np.concatenate([n2, v2], axis=1), u)[:, 4:7]
        assert np.all(np.sum(l2 * n2, 1) > 0)                   # and out again
        second, _ = osc.intersect(np.concatenate([p2 + 1e-9 * l2, l2], axis=1))
    seen = np.full(H * W, -2)
    seen[first] = second
    seen = seen.reshape(H, W)
    contour = np.zeros((H, W), dtype=bool)
    for what in np.unique(second):
        m = seen == what
        contour |= m & ~binary_erosion(m)
    return contour & binary_erosion(ids.reshape(H, W) == 1, iterations=4)


def test_refraction_contours_inside_the_clearest_glass_sphere(native_built):
    """The leftmost sphere of the frosted-glass row is almost clear (alpha = 0.01, ior 1.45, test_scenes.rs:241-256): through
    it the floor's horizon appears upside down.  Where that contour lies is decided by two refractions at ior 1.45 — and the
    reference's render shows a SECOND horizon 26 rows lower: the rays whose far hit the reference loses (F7) and which go on
    once-refracted.  Both contours of the oracle lie on edges of the reference's render; moved vertically by 2-4 pixels, or
    computed with another index of refraction, they do not."""
    from rayrs_b200.api import Material, Object, build_tables
    name = "cook_torrance_spheres_frosted_glass"
    dist, W, H = _edge_distance(name)
    cam = Camera(**ROW)
    spec = scenes.cook_torrance_spheres_frosted_glass(W, H)
    tables = spec.tables()
    osc = oracle.OracleScene(tables, scenes.synthetic_hdri(64, 32).pixels)
    rays = _primary_rays(cam.derived17(), W, H)
    ids, t = osc.intersect(rays)
    mat = tables.mats[int(tables.objs[1, 1])]
    rows_of = {}
    for through in (False, True):
        contour = _through_the_glass_sphere(osc, tables, mat, rays, ids, t, W, H, through)
        assert contour.sum() > 150

        def fit(dy, ct=contour):
            return float((dist[np.roll(ct, dy, 0)] <= NEAR).mean())
        aligned = fit(0)
        assert aligned >= 0.90, (through, aligned)
        moved = max(fit(dy) for dy in (2, -2, 3, -3, 4, -4))
        assert moved <= aligned - 0.25, (through, aligned, moved)
        rows_of[through] = float(np.nonzero(contour)[0].mean())
        # another glass: the same construction with a different index of refraction misses the reference's edges
        for ior in (1.33, 1.6):
            other = build_tables([Object.sphere(1.0, (0, 0, 0), Material.cook_torrance_glass((1, 1, 1), 0.01, ior))]).mats[0]
            wrong = _through_the_glass_sphere(osc, tables, other, rays, ids, t, W, H, through)
            assert float((dist[wrong] <= NEAR).mean()) <= aligned - 0.4, (through, ior)
        print(f"[frosted glass row, sphere 0, {'once-refracted (F7)' if through else 'twice refracted'}] {aligned * 100:.1f} % of "
              f"{int(contour.sum())} contour pixels on an edge of the reference render (mean row {rows_of[through]:.1f}); "
              f"moved 2-4 px vertically: at most {moved * 100:.1f} %")
    osc.close()
    assert rows_of[True] - rows_of[False] > 15   # two distinct horizons


def test_refracted_horizon_in_the_single_glass_sphere(native_built):
    """cook_torrance_glass_single_sphere (test_scenes.rs:65-68: alpha 0.05, ior 1.45) with the single-sphere camera: the lobe
    is five times wider and the sphere 2.5 times larger on screen than in the row, so the refracted horizon is a soft edge — the
    fit is lower, but it still peaks where the oracle puts the contour and falls off within 4 rows either way."""
    name = "cook_torrance_glass_sphere"
    dist, W, H = _edge_distance(name)
    cam = Camera(**SINGLE)
    tables = scenes.cook_torrance_glass_single_sphere(W, H).tables()
    osc = oracle.OracleScene(tables, scenes.synthetic_hdri(64, 32).pixels)
    rays = _primary_rays(cam.derived17(), W, H)
    ids, t = osc.intersect(rays)
    contour = _through_the_glass_sphere(osc, tables, tables.mats[int(tables.objs[1, 1])], rays, ids, t, W, H, False)
    osc.close()
    assert contour.sum() > 500
    fit = {dy: float((dist[np.roll(contour, dy, 0)] <= NEAR).mean()) for dy in (0, 2, -2, 4, -4, 8, -8)}
    assert fit[0] == max(fit.values()) and fit[0] >= 0.45, fit
    assert max(fit[4], fit[-4]) <= fit[0] - 0.2 and max(fit[8], fit[-8]) <= fit[0] - 0.3, fit
    print("[cook_torrance_glass_sphere, twice-refracted horizon] fit by vertical offset:", {k: round(v, 3) for k, v in fit.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["diffuse_single_sphere", "spheres_metallic"])
def test_gpu_silhouettes_fit_the_reference_renders(hdri_small, name):
    """the production fp32 traversal on the same primary rays (rounded to fp32, as the device consumes them)"""
    def gpu_ids(spec, cam17, W, H):
        sc = spec.scene(hdri_small)
        rays = _primary_rays(cam17, W, H).astype(np.float32).astype(np.float64)
        ids, _ = sc.intersect(rays, 32)
        sc.close()
        return ids.reshape(H, W)

    ids, fit, _, cam = _check_silhouettes(name, gpu_ids)
    builder, _ = CASES[name]
    W, H = cam.x_pixels(), cam.y_pixels()
    same = (ids == _oracle_ids(builder(W, H), cam.derived17(), W, H)).mean()
    assert same > 0.999, same  # a handful of silhouette pixels may flip under the fp32 rounding of the ray
    print(f"[{name}] GPU mask: {fit * 100:.1f} % of the silhouette on an edge of the reference render; == oracle on {same * 100:.3f} % of the pixels")
