"""Render one BASELINE configuration ONCE at its full size and spp, samples split over the ranks it is launched with
(strong scaling of the sample budget: 4096 spp on 8 GPUs = 512 spp each), and print one JSON line:

    python scripts/render_config.py c5                                              # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29555 scripts/render_config.py c5

Reports the device time of the slowest rank, the rays of all ranks, the merged image's census and mean, and the
8-bit output stage's counters — the whole path a user takes: render_distributed -> reduce -> resolve -> to_raw_bytes.
"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from rayrs_b200 import api, scenes
    from rayrs_b200.multigpu import sample_range
    key = sys.argv[1] if len(sys.argv) > 1 else "c5"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = scenes.CONFIGS[key]
    hdri = scenes.synthetic_hdri(2048, 1024)
    W, H = cfg.width, cfg.height
    out = []
    for spec in cfg.specs():
        t0 = time.time()
        sc = spec.scene(hdri, device=local, with_f64=False)
        setup = time.time() - t0
        cam = spec.camera()
        first, count = sample_range(rank, world, cfg.spp)
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        api.render_accumulate(cam, sc, count, cfg.max_bounces, acc.data_ptr(), stream, sample_offset=first, spp_total=cfg.spp)
        st = sc.stats()
        if world > 1:
            dist.reduce(acc, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        stats = torch.tensor([float(st["rays"]), st["device_ms"], wall], dtype=torch.float64, device=dev)
        mx = stats.clone()
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        if rank == 0:
            img = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
            api.resolve(sc, acc.data_ptr(), W, H, cfg.spp, img.data_ptr(), True, stream)
            rgb8, census = api.to_raw_bytes(sc, acc.data_ptr(), W, H, cfg.spp, stream_ptr=stream)
            rs = sc.stats()
            rays = float(stats[0].item())
            out.append({"scene": spec.name, "primitives": sc.n_prims, "nodes": sc.n_nodes, "scene_setup_s": setup,
                        "spp_total": cfg.spp, "spp_per_gpu": count, "rays": rays, "device_ms_max": float(mx[1].item()),
                        "wall_s_max_incl_reduce": float(mx[2].item()), "Mrays_per_s": rays / float(mx[2].item()) / 1e6,
                        "census_ok": bool((acc[..., 3] == float(cfg.spp)).all().item()), "nan_pixels": rs["nan_pixels"],
                        "negative_pixels": rs["negative_pixels"], "mean_radiance": float(img.mean().item()),
                        "rgb8_mean": float(rgb8.mean()), "clamped_pixels": census["clamped"]})
        sc.close()
    if rank == 0:
        print(json.dumps({"config": cfg.description, "n_gpus": world, "width": W, "height": H, "results": out}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
