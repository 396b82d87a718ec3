"""Algorithmic bytes per ray of each BASELINE configuration, counted by the
oracle on the rays its own path tracer generates:

    B_ray = 32*N_nodes + S_prim*N_prims + 144 + 64*P_miss + 16*paths_per_ray

N_nodes = child boxes tested and N_prims = primitives tested per ray by a front-to-back, t-pruned
traversal of the reference tree; S_prim = 48 B triangle, 16 B sphere, 32 B plane; 144 B = ray
record 32 B written+read, hit record 8 B written+read, path state 32 B read+written; 64 B = 4 HDRI
texels for rays that escape; 16 B = the float4 radiance reduction per path.
Per-kernel split used by bench.py's roofline:
    extend   : 32 + 8 + 32*N_nodes + S_prim*N_prims
    shade    : 56 + 48 + 64*P_miss + 16*paths_per_ray
    generate : 48*paths_per_ray
Run on the CPU box (uses the oracle; test infrastructure):  python scripts/bytes_per_ray.py [c1 c2 ...]
Writes profiles/bytes_per_ray.json.
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import oracle  # noqa: E402
from rayrs_b200 import scenes  # noqa: E402

OUT = ROOT / "profiles" / "bytes_per_ray.json"


def measure(key, n_rays=200_000):
    cfg = scenes.CONFIGS[key]
    hdri = scenes.synthetic_hdri(2048, 1024)
    tot = dict(rays=0, paths=0, boxes=0, prims=[0, 0, 0], hits=0)
    for spec in cfg.specs():
        t0 = time.time()
        osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
        cam17 = oracle.camera_new(**spec.camera_args)
        W, H = cfg.width, cfg.height
        stride = max(1, (W * H) // 60_000) | 1
        rays = osc.collect_path_rays(cam17, W, H, 1, cfg.max_bounces, pixel_stride=stride, cap=n_rays)
        paths = len(range(0, W * H, stride))
        b, p, h = osc.traversal_counts(rays)
        tot["rays"] += len(rays); tot["paths"] += paths; tot["boxes"] += b; tot["hits"] += h
        for k in range(3):
            tot["prims"][k] += p[k]
        print(f"  {spec.name}: {len(rays)} rays from {paths} paths, build+count {time.time() - t0:.1f} s", flush=True)
        osc.close()
    r = tot["rays"]
    n_nodes = tot["boxes"] / r
    prim_bytes = (16 * tot["prims"][0] + 32 * tot["prims"][1] + 48 * tot["prims"][2]) / r
    n_prims = sum(tot["prims"]) / r
    p_miss = 1.0 - tot["hits"] / r
    ppr = tot["paths"] / r
    extend = 32 + 8 + 32 * n_nodes + prim_bytes
    shade = 56 + 48 + 64 * p_miss + 16 * ppr
    gen = 48 * ppr
    return {
        "description": cfg.description, "sample_rays": r, "rays_per_path": 1.0 / ppr, "N_nodes": n_nodes, "N_prims": n_prims,
        "prim_bytes_per_ray": prim_bytes, "P_miss": p_miss,
        "bytes_per_ray": 32 * n_nodes + prim_bytes + 144 + 64 * p_miss + 16 * ppr,
        "kernel_bytes_per_ray": {"extend": extend, "shade": shade, "generate": gen},
    }


if __name__ == "__main__":
    keys = sys.argv[1:] or ["c1", "c2", "c3"]
    data = json.loads(OUT.read_text()) if OUT.exists() else {}
    for k in keys:
        print(k, flush=True)
        prev = data.get(k, {})
        data[k] = measure(k)
        for keep in ("ncu_dram_bytes_per_launch",):
            if keep in prev:
                data[k][keep] = prev[keep]
        print(json.dumps(data[k], indent=1), flush=True)
        OUT.write_text(json.dumps(data, indent=1) + "\n")
