"""Development timing probe (not a test): per-config device time and phase split, queue sweep."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from rayrs_b200 import scenes, api, _ffi

keys = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2"]
queues = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
spp_over = int(sys.argv[3]) if len(sys.argv) > 3 else 0
modes = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]  # extra flags: 4 = split kernels
hdri = scenes.synthetic_hdri(2048, 1024)
for key in keys:
    cfg = scenes.CONFIGS[key]
    for spec in cfg.specs():
        t0 = time.time()
        sc = spec.scene(hdri, with_f64=False)
        print(key, spec.name, "scene build+upload %.2f s (bvh %.2f s) nodes %d depth %d" % (time.time() - t0, sc.build_seconds, sc.n_nodes, sc.max_depth), flush=True)
        cam = spec.camera()
        spp = spp_over or cfg.spp
        for mode in modes:
          for q in queues:
            for it in range(3):
                t0 = time.time()
                img = api.render_gpu(cam, sc, spp, cfg.max_bounces, queue_capacity=q, flags=mode | (_ffi.RRS_FLAG_TIME_PHASES if it == 2 else 0))
                dt = time.time() - t0
                st = sc.stats()
            print("   flags %d queue %9d: wall %.1f ms device %.2f ms rays %.3e -> %.1f Mrays/s; iters %d launches %d; gen/ext/shade ms %.2f %.2f %.2f; mean %.5f" % (
                mode, q, dt * 1e3, st["device_ms"], st["rays"], st["rays"] / st["device_ms"] / 1e3, st["iterations"], st["kernel_launches"],
                st["generate_ms"], st["extend_ms"], st["shade_ms"], float(img.mean())), flush=True)
        sc.close()
