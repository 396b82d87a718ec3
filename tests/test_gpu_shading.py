"""Shading parity probes through the C ABI: Material::evaluate, Scene::background and the RNG,
GPU fp32 closed forms against the oracle's literal f64 evaluation on the same inputs."""
import numpy as np
import pytest

import oracle
from rayrs_b200 import scenes
from rayrs_b200.api import Fresnel, Material, Object, Scene, BvhHeuristic

pytestmark = pytest.mark.gpu

MATERIALS = {
    "lambertian": Material.lambertian_diffuse((0.8, 0.7, 0.6)),
    "reflect": Material.reflect((0.8, 0.8, 0.8)),
    "refract": Material.refract((1, 1, 1), 1.45),
    "glass": Material.glass((0.8, 0.9, 1.0), 1.45),
    "ct_metal_rough": Material.cook_torrance((1, 1, 1), 0.5, Fresnel.schlick_metallic((0.8, 0.8, 0.8))),
    "ct_copper": Material.cook_torrance((1, 1, 1), 0.05, Fresnel.schlick_metallic((0.722, 0.451, 0.2))),
    "ct_metal_sharp": Material.cook_torrance((1, 1, 1), 0.01, Fresnel.schlick_metallic((0.8, 0.8, 0.8))),
    "ct_dielectric": Material.cook_torrance((0.9, 0.9, 0.9), 0.2, Fresnel.schlick_dielectric(1.45)),
    "ct_refract": Material.cook_torrance_refract((1, 1, 1), 0.09, 1.45),
    "ct_glass_005": Material.cook_torrance_glass((1, 1, 1), 0.05, 1.45),
    "ct_glass_025": Material.cook_torrance_glass((0.8, 0.8, 0.8), 0.25, 1.45),
    "plastic_005": Material.plastic((0.8, 0.8, 0.8), (1, 1, 1), 0.05, 1.45),
    "plastic_025": Material.plastic((0.8, 0.2, 0.1), (1, 1, 1), 0.25, 1.45),
    "no_reflect": Material.no_reflect(),
}


@pytest.fixture(scope="module")
def scene(hdri_small):
    objs = [Object.sphere(1.0, (3.0 * i, 0.0, 0.0), m) for i, m in enumerate(MATERIALS.values())]
    sc = Scene(objs, 1e-6, 1e6, BvhHeuristic.Sah(1000), hdri_small)
    yield sc
    sc.close()


def _inputs(n, seed, both_sides):
    rng = np.random.default_rng(seed)
    nrm = rng.normal(size=(n, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    view = rng.normal(size=(n, 3))
    view /= np.linalg.norm(view, axis=1, keepdims=True)
    if not both_sides:
        flip = np.sum(nrm * view, axis=1) < 0
        view[flip] *= -1
    # both sides consume f32-rounded unit vectors
    nv = np.concatenate([nrm, view], axis=1).astype(np.float32).astype(np.float64)
    u = (rng.integers(0, 1 << 24, (n, 3)) / float(1 << 24))
    return nv, u


@pytest.mark.parametrize("cases_form", [False, True], ids=["staged", "cases"])
@pytest.mark.parametrize("name", sorted(MATERIALS))
def test_material_evaluate_matches_oracle(name, scene, cases_form):
    mat = MATERIALS[name]
    idx = list(MATERIALS).index(name)
    assert scene.tables.mats.shape[0] == len(MATERIALS)
    n = 200_000
    # view on both sides of the surface: transmissive materials see it from inside, and the
    # reflective ones must reproduce the reference's NoScatter / sign behaviour there
    nv, u = _inputs(n, 100 + idx, both_sides=True)
    g = scene.material_evaluate(idx, nv, u, cases_form=cases_form).astype(np.float64)
    o = oracle.material_evaluate(scene.tables.mats[idx], nv, u)
    flag_diff = g[:, 0] != o[:, 0]
    # scatter/no-scatter decisions differ only within fp32 noise of a decision boundary
    assert flag_diff.mean() < 2e-4, flag_diff.mean()
    both = (g[:, 0] == 1) & (o[:, 0] == 1)
    if name == "no_reflect":
        assert not g[:, 0].any() and not o[:, 0].any()
        return
    assert both.mean() > 0.05
    # branch choice (reflect vs refract vs diffuse lobe) agrees except within noise of xi == F
    ddir = np.linalg.norm(g[both, 4:7] - o[both, 4:7], axis=1)
    branch_diff = ddir > 1e-2
    assert branch_diff.mean() < 2e-4, branch_diff.mean()
    same = np.where(both)[0][~branch_diff]
    assert np.quantile(ddir[~branch_diff], 0.999) < 5e-4
    # the reference itself yields NaN / inf weights at singular configurations (0/0 in G at h.v = 0);
    # they must appear on both sides or on neither
    fin_g, fin_o = np.isfinite(g[same, 1:4]).all(axis=1), np.isfinite(o[same, 1:4]).all(axis=1)
    assert (fin_g != fin_o).mean() < 5e-4, (name, int((~fin_g).sum()), int((~fin_o).sum()))
    same = same[fin_g & fin_o]
    scale = np.maximum(np.abs(o[same, 1:4]).max(axis=1), 1e-3)
    cerr = np.abs(g[same, 1:4] - o[same, 1:4]).max(axis=1) / scale
    # fp32 closed form vs the literal f64 brdf*cos/pdf: equal up to conditioning of 1/(n.v), G
    assert np.median(cerr) < 3e-5, (name, np.median(cerr))
    # (h.v -> 0 makes unit(v + l) ill-conditioned; those samples carry weight ~ h.v -> 0)
    assert np.quantile(cerr, 0.999) < 2e-2, (name, np.quantile(cerr, 0.999))
    # no systematic offset: the mean weight agrees to 2e-4 relative (heavy-tailed weights at grazing
    # view angles, where fp32 conditioning is worst, dominate this difference)
    mw_g, mw_o = g[same, 1:4].mean(), o[same, 1:4].mean()
    assert abs(mw_g - mw_o) <= 2e-4 * abs(mw_o) + 1e-7, (name, mw_g, mw_o)
    nv_cos = np.abs(np.sum(nv[same, :3] * nv[same, 3:], axis=1))
    core = nv_cos > 0.2
    mc_g, mc_o = g[same][core, 1:4].mean(), o[same][core, 1:4].mean()
    assert abs(mc_g - mc_o) <= 1e-4 * abs(mc_o) + 1e-7, (name, mc_g, mc_o)
    print(f"[{name}] scatter {both.mean() * 100:.1f}%  flag diff {flag_diff.sum()}  branch diff {branch_diff.sum()}  "
          f"color err median {np.median(cerr):.1e} p99.9 {np.quantile(cerr, 0.999):.1e}  mean weight rel diff "
          f"{(mw_g - mw_o) / mw_o:+.1e} (|n.v|>0.2: {(mc_g - mc_o) / mc_o:+.1e})  non-finite gpu/oracle {int((~fin_g).sum())}/{int((~fin_o).sum())}")


def test_background_matches_oracle(scene, hdri_small):
    osc = oracle.OracleScene(scene.tables, hdri_small.pixels)
    rng = np.random.default_rng(5)
    d = rng.normal(size=(200_000, 3)) * rng.uniform(0.1, 8.0, (200_000, 1))  # not normalised, like primary rays
    d = d.astype(np.float32).astype(np.float64)
    g = scene.background(d).astype(np.float64)
    o = osc.background(d)
    err = np.abs(g - o).max(axis=1)
    assert np.isfinite(g).all()
    assert err.mean() < 2e-4
    assert np.quantile(err, 0.999) < 2e-2  # the Gaussian sun is steep: 3e-4 texel of fp32 jitter
    assert err.max() < 0.1
    assert abs(g.mean() - o.mean()) < 1e-5
    # directions whose texel coordinate is exactly integral (straight up/down, the phi seam): the
    # reference's weights ceil(x)-x and x-floor(x) both vanish and it returns BLACK (lib.rs:268-284);
    # the kernel uses (1-fx, fx) — what f64 gives for the non-integral x that an fp32-integral x
    # stands for — and must at least stay finite and in range there, clamping the indices the
    # reference would panic on
    axes = np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [-1, 0, 0], [0, 0, 1], [0, 0, -1], [-1, 0, 1e-9], [-1, 0, -1e-9]], dtype=np.float64)
    ga = scene.background(axes)
    assert np.isfinite(ga).all() and (ga >= 0).all() and (ga <= 3.0001).all()
    assert not osc.background(axes[:2]).any()


def test_device_rng_reproduces_the_published_philox_7_vector(scene):
    """Random123 kat_vectors, philox4x32 7, counter 0 key 0 -> 5f6fb709 0d893f64 4f121f81 4f730a48: the device stream's
    uniforms are the top 24 bits of those words (the counter's 4th word is always 0 in the render stream)."""
    u = scene.rng_uniforms(0, 0, 0, 0)
    expect = [float(w >> 8) / 16777216.0 for w in (0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48)]
    assert list(u.astype(np.float64)) == expect


def test_rng_is_bit_identical_to_oracle(scene):
    rng = np.random.default_rng(3)
    for _ in range(64):
        seed = int(rng.integers(0, 1 << 62))
        pixel, sample, slot = int(rng.integers(0, 1 << 23)), int(rng.integers(0, 4096)), int(rng.integers(0, 51))
        g = scene.rng_uniforms(seed, pixel, sample, slot)
        o = oracle.rng_uniforms(seed, pixel, sample, slot, oracle.RNG_MATCHED)
        assert np.array_equal(g.astype(np.float64), o)
