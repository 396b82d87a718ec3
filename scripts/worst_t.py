"""Development probe: the rays with the largest relative t error of the fp32 traversal against the oracle."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1])); sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import numpy as np, oracle
from rayrs_b200 import scenes
from test_gpu_intersect import SCENES, fixed_ray_set
hdri = scenes.synthetic_hdri(256, 128)
names = sys.argv[1].split(",")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
for name in names:
    spec = SCENES[name]()
    sc = spec.scene(hdri)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
    for seed in (7, 8, 9):
        rays = fixed_ray_set(spec, osc, n, seed=seed)
        oid, ot = osc.intersect(rays)
        gid, gt = sc.intersect(rays, 32)
        stable = osc.intersect_stable(rays)
        ok = stable & (oid >= 0)
        rel = np.where(ok, np.abs(gt - ot) / np.where(ot > 0, ot, 1), 0)
        idx = np.argsort(-rel)[:6]
        print(name, "seed", seed, "stable", int(stable.sum()), "count>1e-5:", int((rel > 1e-5).sum()), "count>3e-6:", int((rel > 3e-6).sum()))
        tab = spec.tables()
        for i in idx:
            print("   rel %.3e t %.9g gpu %.9g id %d gid %d ray %s first_half %s" % (rel[i], ot[i], gt[i], oid[i], gid[i], np.array2string(rays[i], precision=6), i < n // 2))
