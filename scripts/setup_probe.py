import sys, time
sys.path.insert(0, '.')
from rayrs_b200 import scenes
hdri = scenes.synthetic_hdri(2048, 1024)
for key in ("c4", "c5"):
    spec = scenes.CONFIGS[key].specs()[0]
    for dev in (False, True, True):
        t0 = time.time()
        sc = spec.scene(hdri, with_f64=False, device_build=dev, topology=False)
        dt = time.time() - t0
        print(key, "device_build", dev, "scene() %.3f s; build_seconds %.3f" % (dt, sc.build_seconds), {k: round(v, 3) for k, v in sc.build_timing.items()}, flush=True)
        sc.close()
