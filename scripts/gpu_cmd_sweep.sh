python -m pytest tests -m gpu -x -q -k "not full_size" 2>&1 | tail -3
python scripts/gpu_dev.py c4 | grep -v "scene build" | tee gpurun_out/sweep_fma.log
python scripts/gpu_dev.py c5 0 16 | grep -v "scene build" | tee -a gpurun_out/sweep_fma.log
python - <<PYEOF
import sys; sys.path.insert(0, "/root/repo")
from rayrs_b200 import scenes, api
cfg = scenes.CONFIGS["c4"]; hd = scenes.synthetic_hdri(2048, 1024)
spec = cfg.specs()[0]; sc = spec.scene(hd, with_f64=False); cam = spec.camera()
api.render_gpu(cam, sc, 4, 50, flags=1)
st = sc.stats()
print("rays", st["rays"], "nodes/ray", st["nodes_visited"]/st["rays"], "prims/ray", st["prims_tested"]/st["rays"])
PYEOF
