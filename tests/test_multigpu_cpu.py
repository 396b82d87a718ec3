"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise sample_range, the
single reduce(sum) and the GPU-count independence of the counter-based RNG.  The per-rank
renderer is injected (the CPU oracle stands in for the CUDA backend, which needs a GPU)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_sample_range_partitions():
    from rayrs_b200.multigpu import sample_range
    for spp in (1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count = sample_range(r, world, spp)
                seen.extend(range(first, first + count))
            assert seen == list(range(spp))
    assert sample_range(3, 8, 4096) == (1536, 512)
    with pytest.raises(ValueError):
        sample_range(2, 2, 8)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import oracle
    from rayrs_b200 import scenes
    from rayrs_b200.multigpu import render_distributed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hdri = scenes.synthetic_hdri(64, 32)
    spec = scenes.cook_torrance_spheres_plastic(40, 24)
    cam = spec.camera()
    osc = oracle.OracleScene(spec.tables(), hdri.pixels)
    spp = 10  # not divisible by 3: ranks get 4/3/3

    def accumulate(acc, first, count):
        img, _ = osc.render(cam.derived17(), 40, 24, count, sample_offset=first, nthreads=1)
        acc[..., :3] += torch.from_numpy(img * count).float()
        acc[..., 3] += count

    img, acc = render_distributed(cam, None, spp, accumulate=accumulate, device=torch.device("cpu"))
    if rank == 0:
        full, _ = osc.render(cam.derived17(), 40, 24, spp, nthreads=1)
        np.save(os.path.join(out_dir, f"img_w{world}.npy"), img.numpy())
        np.save(os.path.join(out_dir, f"full_w{world}.npy"), full)
        np.save(os.path.join(out_dir, f"cnt_w{world}.npy"), acc[..., 3].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sample_split_equals_single_render(world, tmp_path, native_built):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    img = np.load(tmp_path / f"img_w{world}.npy")
    full = np.load(tmp_path / f"full_w{world}.npy")
    cnt = np.load(tmp_path / f"cnt_w{world}.npy")
    assert np.array_equal(cnt, np.full_like(cnt, 10.0))
    # same paths, summed in a different order and through fp32 buffers
    assert np.allclose(img, full, rtol=1e-5, atol=1e-6)
