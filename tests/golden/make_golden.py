"""Regenerates tests/golden/*.npz from the CPU oracle (run in the build container; needs no GPU):

    python tests/golden/make_golden.py

The reference itself cannot run here (Rust, no toolchain), so these are outputs of the pinned restatement
(oracle/oracle.cpp, checked against the reference's own KATs in tests/test_oracle_kat.py) on seeded inputs.
They freeze the oracle against silent drift and give the GPU tests a checker that does not depend on
rebuilding it: closest hits on fixed ray sets, one sample-matched render, one output-stage byte image, and the
Pdf::Hittable hook (pdf_hook_golden.npz; `--pdf-hook-only` regenerates that file alone).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from rayrs_b200 import scenes  # noqa: E402

OUT = Path(__file__).resolve().parent
N = 4096

SCENES = {
    "diffuse_single_sphere": lambda: scenes.diffuse_single_sphere(96, 64),
    "material_test": lambda: scenes.material_test(160, 32),
    "copper_torus_3200": lambda: scenes.copper_torus(40, 40, 96, 64),
    "mixed_1800": lambda: scenes.mixed_scene(30, 30, 160, 90),
}


def ray_set(spec, osc, n, seed):
    from test_gpu_intersect import fixed_ray_set
    return fixed_ray_set(spec, osc, n, seed)


def main():
    oracle.build()
    hdri = scenes.synthetic_hdri(128, 64)
    out = {}
    for name, builder in SCENES.items():
        spec = builder()
        osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
        rays = ray_set(spec, osc, N, seed=21)
        ids, t = osc.intersect(rays)
        stable = osc.intersect_stable(rays)
        out[f"{name}/rays"] = rays.astype(np.float32)   # exactly representable: the set was rounded to f32
        out[f"{name}/ids"] = ids.astype(np.int32)
        out[f"{name}/t"] = t
        out[f"{name}/stable"] = stable
        osc.close()
    spec = scenes.cook_torrance_spheres_plastic(48, 24)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    img, st = osc.render(spec.camera().derived17(), 48, 24, 16)
    out["render_plastic_48x24_spp16/image"] = img
    out["render_plastic_48x24_spp16/rays"] = np.array([st["rays"]], dtype=np.int64)
    bytes_, census = oracle.to_raw_bytes(np.concatenate([img, img * 3.0 - 0.2], axis=0))
    out["to_raw_bytes/input"] = np.concatenate([img, img * 3.0 - 0.2], axis=0)
    out["to_raw_bytes/bytes"] = bytes_
    out["to_raw_bytes/census"] = np.array([census["clamped"], census["nan"], census["negative"]], dtype=np.int64)
    np.savez_compressed(OUT / "oracle_golden.npz", **out)
    print("wrote", OUT / "oracle_golden.npz", sum(v.nbytes for v in out.values()), "bytes uncompressed")


def pdf_hook():
    """tests/golden/pdf_hook_golden.npz: Material::evaluate with Some(Pdf::Hittable(light)) (the dormant next-event-estimation
    hook, tests/test_pdf_hook.py) on a seeded input set, for every light kind, Lambertian and Plastic."""
    from test_pdf_hook import LIGHTS, MATS, _inputs, _tables
    from rayrs_b200.api import build_tables
    objs, light_obj = _tables()
    mats = build_tables(objs).mats
    q, u = _inputs(512, 5)
    out = {"q": q, "u": u}
    for kind in sorted(LIGHTS):
        row = objs[light_obj[kind]].rows[0]
        for mname in ("lambertian", "plastic"):
            out[f"{kind}/{mname}"] = oracle.material_evaluate_pdf(mats[list(MATS).index(mname)], row, q, u)
    np.savez_compressed(OUT / "pdf_hook_golden.npz", **out)
    print("wrote", OUT / "pdf_hook_golden.npz")


if __name__ == "__main__":
    if "--pdf-hook-only" not in sys.argv:
        main()
    pdf_hook()
