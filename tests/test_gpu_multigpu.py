"""N > 1 on real GPUs: torchrun + NCCL sample split equals the single-GPU render (scripts/multigpu_check.py).
Skipped on a one-GPU box; the gloo tests in test_multigpu_cpu.py cover the host logic on CPU."""
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_nccl_sample_split_matches_single_gpu(native_built):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "scripts" / "multigpu_check.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stdout[-3000:]
    assert p.stdout.count("PASS") == 4 and "FAIL" not in p.stdout  # 2 scenes x (library call, torch reduce)


def test_one_process_drives_two_gpus_through_one_render_call(native_built, hdri_small):
    """A scene that lives on two GPUs (built and flattened once, uploaded twice: rrs_scene_create_multi) renders
    through the same render_gpu call as a one-GPU scene — rrs_render_multi over an in-process communicator
    (rrs_comm_init_all) — and gives the one-GPU image; census exact on both."""
    import numpy as np
    import torch
    from rayrs_b200 import api, scenes
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    for spec, spp in ((scenes.cook_torrance_spheres_plastic(320, 128), 30), (scenes.mixed_scene(60, 60, 256, 144), 13)):
        cam = spec.camera()
        one = spec.scene(hdri_small, device=0, with_f64=False)
        two = spec.scene(hdri_small, devices=[0, 1], with_f64=False)
        a = api.render_gpu(cam, one, spp, 50).astype(np.float64)
        b = api.render_gpu(cam, two, spp, 50).astype(np.float64)
        s0, s1 = two.stats(0), two.stats(1)
        assert one.stats()["census_mismatch_pixels"] == 0 and s0["census_mismatch_pixels"] == 0
        assert s0["rays"] > 0 and s1["rays"] > 0 and abs(s0["rays"] + s1["rays"] - one.stats()["rays"]) <= 1e-3 * one.stats()["rays"]
        err = float(np.max(np.abs(a - b) / (np.abs(a) + 1e-2)))
        print(f"[{spec.name}] two GPUs in one process vs one GPU: max rel diff {err:.2e}; rays {s0['rays']} + {s1['rays']}")
        assert err < 1e-4
        one.close()
        two.close()
