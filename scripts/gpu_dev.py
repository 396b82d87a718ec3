"""Development timing probe (not a test): per-config device time, phase split, traversal counters, queue / flag sweeps.

    python scripts/gpu_dev.py <configs> <queues> <spp> <render flags,...> [scene_flags]
"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from rayrs_b200 import scenes, api, _ffi

keys = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2"]
queues = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
spp_over = int(sys.argv[3]) if len(sys.argv) > 3 else 0
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]  # render flags: 1 count, 4 split kernels, 32 no L2 window
scene_flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
refill = int(sys.argv[6]) if len(sys.argv) > 6 else 0
hdri = scenes.synthetic_hdri(2048, 1024)
for key in keys:
    cfg = scenes.CONFIGS[key]
    for spec in cfg.specs():
        t0 = time.time()
        sc = spec.scene(hdri, with_f64=False, scene_flags=scene_flags, refill_lanes=refill, device_build=True, topology=False)
        print(key, spec.name, "scene build+upload %.2f s (bvh %.2f s) nodes %d depth %d" % (time.time() - t0, sc.build_seconds, sc.n_nodes, sc.max_depth), flush=True)
        cam = spec.camera()
        spp = spp_over or cfg.spp
        for mode in modes:
          for q in queues:
            best = None
            for it in range(6):  # the fastest of 6 renders (clock / thermal noise between renders is +-5 %)
                t0 = time.time()
                img = api.render_gpu(cam, sc, spp, cfg.max_bounces, queue_capacity=q, flags=mode)
                dt = time.time() - t0
                st_i = sc.stats()
                if best is None or st_i["device_ms"] < best["device_ms"]:
                    best = st_i
            st = best
            print("   flags %d queue %9d: wall %.1f ms device %.2f ms rays %.3e -> %.1f Mrays/s; iters %d; gen/ext/shade ms %.2f %.2f %.2f; nodes/ray %.2f prims/ray %.2f; mean %.6f census_miss %d" % (
                mode, q, dt * 1e3, st["device_ms"], st["rays"], st["rays"] / st["device_ms"] / 1e3, st["iterations"],
                st["generate_ms"], st["extend_ms"], st["shade_ms"], st["nodes_visited"] / max(1, st["rays"]), st["prims_tested"] / max(1, st["rays"]),
                float(img.mean()), st["census_mismatch_pixels"]), flush=True)
        sc.close()
