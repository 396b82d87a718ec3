"""Committed golden vectors (tests/golden/oracle_golden.npz, made by tests/golden/make_golden.py from the pinned
oracle): the oracle must keep reproducing them bit for bit (CPU), and the CUDA path must match them through the
C ABI (GPU) — closest hits (ids exact, t to 1e-5), a sample-matched render, and the output-stage bytes."""
from pathlib import Path

import numpy as np
import pytest

import oracle
from rayrs_b200 import scenes

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "oracle_golden.npz")
SCENES = {
    "diffuse_single_sphere": lambda: scenes.diffuse_single_sphere(96, 64),
    "material_test": lambda: scenes.material_test(160, 32),
    "copper_torus_3200": lambda: scenes.copper_torus(40, 40, 96, 64),
    "mixed_1800": lambda: scenes.mixed_scene(30, 30, 160, 90),
}


@pytest.mark.parametrize("name", sorted(SCENES))
def test_oracle_reproduces_golden_hits(name, native_built):
    hdri = scenes.synthetic_hdri(128, 64)
    spec = SCENES[name]()
    osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
    rays = GOLD[f"{name}/rays"].astype(np.float64)
    ids, t = osc.intersect(rays)
    assert np.array_equal(ids, GOLD[f"{name}/ids"])
    assert np.array_equal(t, GOLD[f"{name}/t"])            # same machine arithmetic: bit for bit
    assert (ids >= 0).mean() > 0.2
    osc.close()


def test_oracle_reproduces_golden_render_and_bytes(native_built):
    hdri = scenes.synthetic_hdri(128, 64)
    spec = scenes.cook_torrance_spheres_plastic(48, 24)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    img, st = osc.render(spec.camera().derived17(), 48, 24, 16, nthreads=3)   # the image does not depend on the thread count
    assert st["rays"] == int(GOLD["render_plastic_48x24_spp16/rays"][0])
    assert np.allclose(img, GOLD["render_plastic_48x24_spp16/image"], rtol=1e-13, atol=0)
    b, census = oracle.to_raw_bytes(GOLD["to_raw_bytes/input"])
    assert np.array_equal(b, GOLD["to_raw_bytes/bytes"])
    assert [census["clamped"], census["nan"], census["negative"]] == GOLD["to_raw_bytes/census"].tolist()
    osc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENES))
def test_gpu_matches_golden_hits(name, native_built):
    hdri = scenes.synthetic_hdri(128, 64)
    spec = SCENES[name]()
    sc = spec.scene(hdri)
    rays = GOLD[f"{name}/rays"].astype(np.float64)
    gid, gt = sc.intersect(rays, 32)
    did, dt = sc.intersect(rays, 64)
    oid, ot, stable = GOLD[f"{name}/ids"], GOLD[f"{name}/t"], GOLD[f"{name}/stable"]
    assert np.array_equal(did, oid)
    hit = oid >= 0
    assert np.max(np.abs(dt[hit] - ot[hit]) / ot[hit]) <= 1e-12
    assert np.array_equal(gid[stable], oid[stable])
    ok = stable & hit
    assert np.max(np.abs(gt[ok] - ot[ok]) / ot[ok]) <= 1e-5
    sc.close()


@pytest.mark.gpu
def test_gpu_matches_golden_render(native_built):
    from rayrs_b200 import api
    from conftest import relrmse
    hdri = scenes.synthetic_hdri(128, 64)
    spec = scenes.cook_torrance_spheres_plastic(48, 24)
    sc = spec.scene(hdri, with_f64=False)
    img = api.render_gpu(spec.camera(), sc, 16, 50).astype(np.float64)
    assert relrmse(img, GOLD["render_plastic_48x24_spp16/image"]) < 5e-3     # sample-matched: same paths, fp32 vs f64
    assert abs(sc.stats()["rays"] - int(GOLD["render_plastic_48x24_spp16/rays"][0])) <= 20
    sc.close()
