"""One render of one BASELINE configuration (for ncu): python scripts/ncu_one.py c4 [spp] [flags]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from rayrs_b200 import scenes, api
key = sys.argv[1]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 0
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = scenes.CONFIGS[key]
hdri = scenes.synthetic_hdri(2048, 1024)
for spec in cfg.specs():
    sc = spec.scene(hdri, with_f64=False)
    cam = spec.camera()
    for _ in range(2):
        api.render_gpu(cam, sc, spp or cfg.spp, cfg.max_bounces, flags=flags)
    st = sc.stats()
    print(key, spec.name, "device %.2f ms rays %.3e -> %.1f Mrays/s" % (st["device_ms"], st["rays"], st["rays"] / st["device_ms"] / 1e3))
    sc.close()
