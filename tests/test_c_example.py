"""examples/render_c.c — a C99 host over the C ABI with nothing of this repo's Python or C++ host code in between (the
shape of the Rust shim in INTEGRATION.md): it must compile against include/rayrs_b200.h as plain C, link against
librayrs_b200.so, fail loudly without a device, and on a GPU render the image the Python face renders from the host
mirror's own flattening of the same scene."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def render_c(native_built, tmp_path_factory):
    exe = tmp_path_factory.mktemp("c_example") / "render_c"
    lib_dir = ROOT / "rayrs_b200"
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}",
           str(ROOT / "examples" / "render_c.c"), f"-L{lib_dir}", "-lrayrs_b200", "-lm", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def _hdri_file(tmp_path, w=256, h=128):
    from rayrs_b200 import scenes
    hdri = scenes.synthetic_hdri(w, h)
    path = tmp_path / "hdri.f32"
    np.ascontiguousarray(hdri.pixels, dtype=np.float32).tofile(path)
    return hdri, path


def test_compiles_as_c99_and_refuses_to_run_without_a_device(render_c, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _, hpath = _hdri_file(tmp_path, 8, 4)
    p = subprocess.run([str(render_c), str(hpath), "8", "4", "64", "64", "4", str(tmp_path / "o.f32")], capture_output=True, text=True)
    assert p.returncode == 1
    assert "rrs_scene_create" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_host_renders_the_same_image_as_the_python_face(render_c, tmp_path):
    from rayrs_b200 import api, scenes
    W, H, spp = 192, 128, 32
    hdri, hpath = _hdri_file(tmp_path)
    out = tmp_path / "out.f32"
    p = subprocess.run([str(render_c), str(hpath), "256", "128", str(W), str(H), str(spp), str(out)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    print(p.stdout.strip())
    assert "axis ray hits object 1" in p.stdout and "nan 0 negative 0" in p.stdout
    img_c = np.fromfile(out, dtype=np.float32).reshape(H, W, 3)
    spec = scenes.diffuse_single_sphere(W, H)
    sc = spec.scene(hdri)
    img_py = api.render_gpu(spec.camera(), sc, spp, 50)
    sc.close()
    # same flattening, same camera fields, same seed: the same paths (the accumulator's atomic order is the only freedom)
    assert np.allclose(img_c, img_py, rtol=1e-5, atol=1e-6)
