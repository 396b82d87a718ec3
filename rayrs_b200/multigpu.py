"""Sample-split rendering across the GPUs of one box.

One process per GPU (torchrun).  Samples of a pixel are independent, so rank g renders the
contiguous global sample range sample_range(g, G, spp) of EVERY pixel into its own fp32
radiance-sum buffer (scene replicated on every GPU); ONE reduce(sum) over NCCL / NVLink merges
the buffers on rank 0, which divides by spp.  The RNG is keyed by the global sample index, so
the set of paths — and, up to fp32 summation order, the image — is independent of G.

torch is plumbing here (device buffers, streams, torch.distributed); the rendering is the CUDA
backend behind the C ABI.

The PRODUCT path for several GPUs is inside the library: rrs_render_multi over an RrsComm (api.render_multi,
api.Comm, Scene(devices=[...]); csrc/multi.cu) — one call, NCCL bound by the library itself; bench.py and the
C / Rust hosts use that.  This module keeps the same split driven from Python (render_distributed): it is what the
gloo tests on CPU exercise (tests/test_multigpu_cpu.py) and the torch-side cross-check of scripts/multigpu_check.py.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


def sample_range(rank: int, world: int, spp: int) -> tuple[int, int]:
    """(first global sample, number of samples) of `rank`: contiguous, disjoint, covering [0, spp);
    the first spp % world ranks take one extra sample."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(int(spp), world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def render_distributed(camera, scene, spp: int, max_bounces: int = 50, *, seed: Optional[int] = None,
                       queue_capacity: int = 0, flags: int = 0, accumulate: Optional[Callable] = None,
                       device=None, dst: int = 0, resolve_on_device: bool = True):
    """Render `spp` samples per pixel split over the ranks of the default process group.

    Returns (image, sum_buffer): on rank `dst` image is the H x W x 3 mean radiance (torch tensor on
    `device`); on other ranks None.  `accumulate(sum_buffer, first_sample, n_samples)` may replace the
    CUDA backend (the CPU/gloo tests inject a stand-in); by default it is rrs_render_accumulate.
    """
    import torch
    import torch.distributed as dist
    from . import api

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first, count = sample_range(rank, world, spp)
    H, W = camera.y_pixels(), camera.x_pixels()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device=device)
    if accumulate is None:
        stream = torch.cuda.current_stream(device).cuda_stream
        kw = dict(sample_offset=first, spp_total=spp, queue_capacity=queue_capacity, flags=flags)
        if seed is not None:
            kw["seed"] = seed
        if count > 0:
            api.render_accumulate(camera, scene, count, max_bounces, acc.data_ptr(), stream, **kw)
    elif count > 0:
        accumulate(acc, first, count)
    if world > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM)  # the path's single exchange step
    if rank != dst:
        return None, acc
    if accumulate is None and resolve_on_device:
        out = torch.empty((H, W, 3), dtype=torch.float32, device=device)
        api.resolve(scene, acc.data_ptr(), W, H, spp, out.data_ptr(), True, torch.cuda.current_stream(device).cuda_stream)
        return out, acc
    return acc[..., :3] / float(spp), acc
