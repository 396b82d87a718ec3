"""In-tree build of the native libraries (nvcc for sm_100a, g++ for the host mirror).

    python -m rayrs_b200.build            # build what is out of date
    python -m rayrs_b200.build --force

Outputs (git-ignored, but they travel to the GPU box with the gpurun snapshot):
    rayrs_b200/librayrs_b200.so   CUDA kernels + C ABI (include/rayrs_b200.h)
    rayrs_b200/librayrs_host.so   C++ host mirror of the rayrs-lib scene API
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
HOST = ROOT / "host"
BUILD = ROOT / "_build"

GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = GENCODE + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
                        "--expt-relaxed-constexpr"]
# per-file extra flags: the fp64 verification TU must not contract a*b+c (Rust never does)
# wavefront.cu: approximate fp32 division / square root (<= 2 ulp) — every fp32 result on this path is
# compared against an f64 oracle at 1e-5 relative, and the decisions that must be exact (watertight
# edge functions, the f64 sphere path) use explicit round-to-nearest intrinsics
EXTRA = {"verify_f64.cu": ["-fmad=false"], "nee.cu": ["-fmad=false"], "bvh_build.cu": ["-fmad=false"], "wavefront.cu": ["-prec-div=false", "-prec-sqrt=false"]}

CUDA_LIB = ROOT / "librayrs_b200.so"
HOST_LIB = ROOT / "librayrs_host.so"


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _newer(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(d.stat().st_mtime <= t for d in deps if d.exists())


def _run(cmd: list[str], log: Path | None = None) -> None:
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log is not None:
        log.write_text(proc.stdout)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build_cuda(force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "rayrs_b200.h", Path(__file__)]
    sources = sorted(CSRC.glob("*.cu"))
    objs = []
    nvcc = _nvcc()
    for src in sources:
        obj = BUILD / (src.stem + ".o")
        objs.append(obj)
        if not force and _newer(obj, [src] + headers):
            continue
        _run([nvcc] + NVCC_FLAGS + EXTRA.get(src.name, []) + os.environ.get("RRS_NVCC_EXTRA", "").split() + ["-c", str(src), "-o", str(obj)],
             log=BUILD / (src.stem + ".ptxas.log"))
    if force or not _newer(CUDA_LIB, objs):
        _run([nvcc] + GENCODE + ["-shared", "-o", str(CUDA_LIB)] + [str(o) for o in objs])
    return CUDA_LIB


def build_host(force: bool = False) -> Path:
    BUILD.mkdir(exist_ok=True)
    sources = sorted(HOST.glob("*.cpp"))
    headers = sorted(HOST.glob("*.hpp")) + [ROOT.parent / "include" / "rayrs_b200.h"]
    if not sources:
        return HOST_LIB
    if not force and _newer(HOST_LIB, sources + headers + [CUDA_LIB]):
        return HOST_LIB
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wall",
           "-o", str(HOST_LIB)] + [str(s) for s in sources] + [
        "-L" + str(ROOT), "-lrayrs_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd)
    return HOST_LIB


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", CUDA_LIB, HOST_LIB)
