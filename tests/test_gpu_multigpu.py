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
    assert p.stdout.count("PASS") == 2 and "FAIL" not in p.stdout
