"""Tiny renders through every kernel form, for compute-sanitizer (one tool per gpurun call):

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py

No torch import (ctypes only), so the tool instruments little besides our own kernels."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402

from rayrs_b200 import _ffi, api, scenes  # noqa: E402

F = _ffi
hdri = scenes.synthetic_hdri(64, 32)
cases = [
    ("plastic / path loop", scenes.cook_torrance_spheres_plastic(37, 19), [0, F.RRS_FLAG_FORCE_QUEUES, F.RRS_FLAG_SPLIT_KERNELS]),
    ("frosted glass / queued f64", scenes.cook_torrance_spheres_frosted_glass(40, 16), [0, F.RRS_FLAG_FORCE_PATHLOOP]),
    ("material_test / BVH f64", scenes.material_test(56, 16), [0, F.RRS_FLAG_SPLIT_KERNELS | F.RRS_FLAG_COUNT_TRAVERSAL]),
    ("torus / BVH", scenes.copper_torus(16, 8, 40, 24), [0, F.RRS_FLAG_COUNT_TRAVERSAL]),
    ("mixed / BVH f64", scenes.mixed_scene(12, 8, 48, 27), [0]),
]
rng = np.random.default_rng(3)
for name, spec, flag_list in cases:
    sc = spec.scene(hdri)
    cam = spec.camera()
    for flags in flag_list:
        for q in (0, 2048):
            img = api.render_gpu(cam, sc, 3, 50, queue_capacity=q, flags=flags)
            st = sc.stats()
            assert np.isfinite(img).all() and st["paths"] == cam.x_pixels() * cam.y_pixels() * 3
    rays = np.concatenate([rng.uniform(-3, 3, (512, 3)) + [0, 2, 6], rng.normal(size=(512, 3))], axis=1)
    sc.intersect(rays, 32)
    sc.intersect(rays, 64)
    print("ok", name, st["rays"], "rays, kernel form", st["kernel_form"], flush=True)
    sc.close()
print("sanitize_small: all renders finished")
