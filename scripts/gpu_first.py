"""First GPU contact: parity spot checks + rough timings (development aid, not a test)."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import oracle
from rayrs_b200 import scenes, api, _ffi

def relrmse(g, r):
    return float(np.sqrt(np.mean((g - r) ** 2 / (r ** 2 + 0.01))))

print("devices", _ffi.cuda_lib().rrs_device_count())
hdri = scenes.synthetic_hdri(512, 256)
for builder, W, H, spp, mb in [(scenes.diffuse_single_sphere, 128, 96, 32, 8),
                               (scenes.cook_torrance_spheres_metallic, 160, 64, 32, 50),
                               (scenes.cook_torrance_spheres_plastic, 160, 64, 32, 50),
                               (scenes.cook_torrance_spheres_frosted_glass, 160, 64, 32, 50),
                               (scenes.glass_single_sphere, 128, 96, 32, 50),
                               (scenes.material_test, 200, 40, 32, 50),
                               (lambda w, h: scenes.copper_torus(60, 30, w, h), 128, 96, 16, 50)]:
    spec = builder(W, H)
    sc = spec.scene(hdri)
    cam = spec.camera()
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    c17 = cam.derived17()
    rng = np.random.default_rng(1)
    n = 20000
    rows = rng.integers(0, H, n); cols = rng.integers(0, W, n); smp = rng.integers(0, 64, n)
    rays = oracle.primary_rays(c17, W, H, rows, cols, smp)
    rays = rays.astype(np.float32).astype(np.float64)
    oid, ot = osc.intersect(rays)
    gid, gt = sc.intersect(rays, 32)
    did, dt = sc.intersect(rays, 64)
    hit = oid >= 0
    print(spec.name, "fp32 id mismatch", int((gid != oid).sum()), "f64 id mismatch", int((did != oid).sum()),
          "max rel t32", float(np.max(np.abs(gt[hit & (gid == oid)] - ot[hit & (gid == oid)]) / ot[hit & (gid == oid)])) if hit.any() else None,
          "max rel t64", float(np.max(np.abs(dt[hit] - ot[hit]) / ot[hit])) if hit.any() else None,
          "t64 bitexact", float(np.mean(dt[hit] == ot[hit])) if hit.any() else None)
    img = api.render_gpu(cam, sc, spp, mb)
    st = sc.stats()
    ref, ost = osc.render(c17, W, H, spp, max_bounces=mb)
    ref2, _ = osc.render(c17, W, H, spp, max_bounces=mb, seed=12345)
    print("   render relRMSE gpu-vs-oracle(matched)", relrmse(img.astype(np.float64), ref), "oracle-vs-oracle(noise)", relrmse(ref2, ref),
          "rays gpu", st["rays"], "oracle", ost["rays"], "nan", st["nan_pixels"], "mean", img.mean(), ref.mean())
    sc.close()

# timings at full size
hdri = scenes.synthetic_hdri(2048, 1024)
for key in ["c1", "c2"]:
    cfg = scenes.CONFIGS[key]
    for spec in cfg.specs():
        sc = spec.scene(hdri)
        cam = spec.camera()
        for it in range(3):
            t0 = time.time()
            img = api.render_gpu(cam, sc, cfg.spp, cfg.max_bounces, flags=_ffi.RRS_FLAG_TIME_PHASES if it == 2 else 0)
            dt = time.time() - t0
            st = sc.stats()
            print(key, spec.name, "wall %.3f s" % dt, "device %.1f ms" % st["device_ms"], "rays %.3e" % st["rays"],
                  "Mrays/s %.1f" % (st["rays"] / st["device_ms"] / 1e3), "iters", st["iterations"],
                  "gen/ext/shade ms", st["generate_ms"], st["extend_ms"], st["shade_ms"])
        sc.close()
