"""Checks whose expected values come from physics, not from a restatement by the same author (task 8):

 * white furnace — an albedo-1 Lambertian sphere under a uniform environment of radiance 1 returns radiance exactly 1
   in every pixel: each bounce multiplies the throughput by the albedo (cosine sampling, weight = colour,
   material.rs:259-281), Russian roulette survives with probability max(thr) = 1 (lib.rs:543-546), a ray leaving a
   convex sphere escapes, and the environment lookup interpolates a constant (lib.rs:254-285).  Exercises radiance(),
   the Lambertian arm, the roulette, the background and the accumulator end to end;
 * Schlick's approximation at normal incidence is the exact Fresnel reflectance ((n1 - n2) / (n1 + n2))^2
   (material.rs:1457-1518): Glass reflects with that probability, found here by bisection on the uniform that drives
   the choice (material.rs:339-401);
 * the reference's gamma known answer 0.5^(1/2.2) = 0.7297400528407231 (vecmath.rs:360-366).
The oracle half runs on CPU; the CUDA half is marked gpu."""
import numpy as np
import pytest

import oracle
from rayrs_b200 import scenes
from rayrs_b200.api import BvhHeuristic, Emission, Image, Material, Object


def _furnace_spec(W=96, H=64):
    w, h = scenes.film(W, H)
    objects = [Object.sphere(1.0, (0.0, 0.0, 0.0), Material.lambertian_diffuse((1.0, 1.0, 1.0)), Emission.Dark())]
    cam = dict(origin=(0.0, 1.0, 4.0), up=(0.0, 1.0, 0.0), lookat=(0.0, 0.0, 0.0), fov=50.0, width=w, height=h, ppi=scenes.PPI)
    return scenes.SceneSpec("white_furnace", cam, objects, BvhHeuristic.Sah(1000))


def _uniform_env():
    return Image(16, 8, np.ones((8, 16, 3)))


def test_white_furnace_oracle(native_built):
    spec = _furnace_spec()
    env = _uniform_env()
    osc = oracle.OracleScene(spec.tables(), env.pixels)
    cam = spec.camera()
    img, st = osc.render(cam.derived17(), cam.x_pixels(), cam.y_pixels(), 16)
    assert st["scatters"] > 0.2 * img.shape[0] * img.shape[1] * 16   # the sphere is actually hit
    assert np.max(np.abs(img - 1.0)) <= 1e-12


@pytest.mark.gpu
def test_white_furnace_gpu(native_built):
    from rayrs_b200 import api
    spec = _furnace_spec()
    sc = spec.scene(_uniform_env(), with_f64=False)
    img = api.render_gpu(spec.camera(), sc, 64, 50)
    st = sc.stats()
    assert st["rays"] > 1.2 * st["paths"] and st["census_mismatch_pixels"] == 0
    assert np.max(np.abs(img.astype(np.float64) - 1.0)) <= 2e-6
    sc.close()


def _reflect_threshold(evaluate, lo=0.0, hi=1.0, iters=60):
    """largest uniform for which Glass reflects at normal incidence (reflection: the new direction is the normal)"""
    nv = np.array([[0.0, 0.0, 1.0, 0.0, 0.0, 1.0]])
    def reflects(u):
        out = evaluate(nv, np.array([[u, 0.5, 0.5]]))
        return out[0, 6] > 0.0   # direction z: +1 reflected, -1 refracted
    assert reflects(lo) and not reflects(hi - 1e-9)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        if reflects(mid):
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


def test_schlick_normal_incidence_oracle(native_built):
    ior = 1.45
    row = Material.glass((1, 1, 1), ior).row
    thr = _reflect_threshold(lambda nv, u: oracle.material_evaluate(row, nv, u))
    assert abs(thr - ((1.0 - ior) / (1.0 + ior)) ** 2) <= 1e-12


@pytest.mark.gpu
def test_schlick_normal_incidence_gpu(native_built, hdri_small):
    from rayrs_b200.api import Scene
    ior = 1.45
    sc = Scene([Object.sphere(1.0, (0, 0, 0), Material.glass((1, 1, 1), ior))], 1e-6, 1e6, BvhHeuristic.Sah(1000), hdri_small)
    for cases_form in (False, True):
        thr = _reflect_threshold(lambda nv, u: sc.material_evaluate(0, nv, u, cases_form=cases_form), iters=30)
        assert abs(thr - ((1.0 - ior) / (1.0 + ior)) ** 2) <= 2e-7
    sc.close()


def test_reference_gamma_known_answer():
    """vecmath.rs:360-366 doctest: Vec3(0.5).powf(1 / 2.2) == 0.7297400528407231 — the pow Image::to_raw_bytes applies"""
    assert np.power(np.float64(0.5), 1.0 / 2.2) == 0.7297400528407231
    out, _ = oracle.to_raw_bytes(np.full((1, 1, 3), 0.5), 1.0 / 2.2)
    assert out.tolist() == [[[int(255.99 * 0.7297400528407231)] * 3]] == [[[186, 186, 186]]]
