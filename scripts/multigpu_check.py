"""N-GPU parity check of the sample split, run under torchrun with the NCCL backend:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/multigpu_check.py

Every rank renders its global sample range with the CUDA backend, ONE reduce(sum) merges the fp32
radiance buffers on rank 0 — once through the library's own multi-GPU render call (rrs_render_multi over an
RrsComm whose NCCL id torch.distributed only carries: "lib"), once with torch.distributed doing the reduce
("torch") — and rank 0 compares each merged image with the same spp rendered on one GPU:
same set of paths (RNG keyed by the global sample index), so the images agree up to fp32 summation
order, and the per-pixel terminated-path census equals spp everywhere.  Prints one line and exits
non-zero on a mismatch.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from rayrs_b200 import api, scenes
    from rayrs_b200.multigpu import render_distributed

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    hdri = scenes.synthetic_hdri(512, 256)
    ok = True

    def exchange(raw):  # rank 0's NCCL id, 128 bytes, carried by torch.distributed
        box = [raw]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        return box[0]

    comm = api.Comm(world, rank, local, exchange)
    comm_ptr = comm.ptr
    for spec, spp in ((scenes.cook_torrance_spheres_plastic(320, 128), 30), (scenes.mixed_scene(60, 60, 256, 144), 13)):
        sc = spec.scene(hdri, device=local, with_f64=False)
        cam = spec.camera()
        img, acc = render_distributed(cam, sc, spp, 50, device=dev)
        # the library path: one call, sample split + ncclReduce + resolve inside rrs_render_multi
        W, H = cam.x_pixels(), cam.y_pixels()
        lib_img = torch.empty((H, W, 3), dtype=torch.float32, device=dev) if rank == 0 else None
        api.render_multi(cam, [sc.handle], comm_ptr, spp, 50, out_ptr=lib_img.data_ptr() if rank == 0 else 0, out_is_device=True,
                         streams=[torch.cuda.current_stream(dev).cuda_stream])
        lib_stats = sc.stats()
        if rank == 0:
            single = api.render_gpu(cam, sc, spp, 50).astype(np.float64)
            census = acc[..., 3].cpu().numpy()
            for name, merged, census_ok in (("torch", img.cpu().numpy().astype(np.float64), bool(np.array_equal(census, np.full_like(census, float(spp))))),
                                            ("lib", lib_img.cpu().numpy().astype(np.float64), lib_stats["census_mismatch_pixels"] == 0)):
                err = float(np.max(np.abs(merged - single) / (np.abs(single) + 1e-2)))
                good = census_ok and err < 1e-4
                ok &= good
                print(f"multigpu_check[{name}] world={world} scene={spec.name} spp={spp}: max rel diff vs 1 GPU {err:.2e}, "
                      f"census {'ok' if census_ok else 'BAD'} -> {'PASS' if good else 'FAIL'}", flush=True)
        sc.close()
    comm.close()
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
