"""examples/render_c.c — a C99 host over the C ABI with nothing of this repo's Python or C++ host code in between (the
shape of the Rust shim in rust/gpu.rs): it must compile against include/rayrs_b200.h as plain C, link against
librayrs_b200.so and fail loudly without a device.  Its `row7` scene goes through a tree build and the flattening
algorithm written in C: the arrays it produces must equal the host mirror's bit for bit (CPU), and on a GPU its closest
hits and its image must equal the ones rendered from the host mirror's flattening; with two GPUs the same binary renders
through rrs_render_multi and must reproduce the one-GPU image."""
import ctypes as C
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


def _run(exe, scene, hpath, hw, hh, W, H, spp, ngpus, out, *extra):
    return subprocess.run([str(exe), scene, str(hpath), str(hw), str(hh), str(W), str(H), str(spp), str(ngpus), str(out),
                           *[str(e) for e in extra]], capture_output=True, text=True)


def _row7_spec(W, H):
    from rayrs_b200 import scenes
    from rayrs_b200.api import BvhHeuristic
    spec = scenes.cook_torrance_spheres_metallic(W, H)
    spec.heuristic = BvhHeuristic.Midpoint()  # the C example restates BvhTree::build_midpoint
    return spec


def test_compiles_as_c99_and_refuses_to_run_without_a_device(render_c, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    _, hpath = _hdri_file(tmp_path, 8, 4)
    p = _run(render_c, "single", hpath, 8, 4, 64, 64, 4, 1, tmp_path / "o.f32")
    assert p.returncode == 1
    assert "rrs_scene_create" in p.stderr and "no CPU fallback" in p.stderr


def _compare_flat(raw, sc):
    """Every byte of the C example's arrays against the host mirror's — except the material INDEX of a primitive: the
    host mirror numbers its material table in order of first use along the DFS primitive order, the example keeps the
    caller's numbering; the two must be the same pairing under a one-to-one renaming."""
    from rayrs_b200 import _ffi
    nodes, nodes64, order, topo, boxes, prims = sc.flat()
    n_prims, n_nodes, max_depth = (int(x) for x in np.frombuffer(raw[:12], dtype=np.uint32))
    assert (n_prims, n_nodes, max_depth) == (sc.n_prims, sc.n_nodes, sc.max_depth)
    off = 12
    c_prims = (_ffi.RrsPrim * n_prims).from_buffer_copy(raw[off:off + C.sizeof(_ffi.RrsPrim) * n_prims])
    off += C.sizeof(_ffi.RrsPrim) * n_prims
    rename = {}
    for a, b in zip(c_prims, prims):
        assert (a.type, a.obj_id, a.emission, list(a.v)) == (b.type, b.obj_id, b.emission, list(b.v))
        assert rename.setdefault(a.material, b.material) == b.material
    assert len(set(rename.values())) == len(rename)
    for arr, size in ((nodes, C.sizeof(_ffi.RrsNode) * sc.n_nodes), (nodes64, C.sizeof(_ffi.RrsNodeF64) * sc.n_nodes)):
        assert raw[off:off + size] == bytes(arr), type(arr)
        off += size
    assert off == len(raw)
    return nodes


def _scene_file(path, spec):
    """The object list of a SceneSpec in the example's .bin format (RrsPrim / RrsMaterial records + camera)."""
    from rayrs_b200 import _ffi
    t = spec.tables()
    n, m = t.objs.shape[0], t.mats.shape[0]
    prims = (_ffi.RrsPrim * n)()
    for i, row in enumerate(t.objs):
        p = prims[i]
        p.type, p.obj_id, p.material, p.emission = int(row[0]), i, int(row[1]), int(row[2])
        v = list(row[3:12])
        if p.type == 0:      # sphere: the table carries the radius, RrsPrim the squared radius
            v = [v[0] * v[0], v[1], v[2], v[3], 0, 0, 0, 0, 0]
        p.v[:] = v
    mats = (_ffi.RrsMaterial * m)()
    for i, row in enumerate(t.mats):
        q = mats[i]
        q.tag, q.fresnel_kind = int(row[0]), int(row[6])
        q.color[:] = list(row[1:4])
        q.alpha, q.ior = float(row[4]), float(row[5])
        q.spec_color[:] = list(row[7:10])
    c = spec.camera_args
    with open(path, "wb") as f:
        f.write(np.array([n, m], dtype=np.uint32).tobytes())
        f.write(bytes(prims))
        f.write(bytes(mats))
        f.write(np.array([*c["origin"], *c["lookat"], c["fov"]], dtype=np.float64).tobytes())


def _awkward_spec(W, H):
    """A tree with every flattening rule in it: 1-object sides (bare LeafNode children), a dead node (a group of
    coplanar planes has a zero-extent box the reference can never enter), leaf groups, three levels."""
    from rayrs_b200 import scenes
    from rayrs_b200.api import Axis, BvhHeuristic, Emission, Fresnel, Material, Object
    grey = Material.lambertian_diffuse((0.7, 0.7, 0.7))
    metal = Material.cook_torrance((1, 1, 1), 0.1, Fresnel.schlick_metallic((0.8, 0.6, 0.3)))
    objs = [Object.sphere(0.6, (-20.0, 0.6, 0.0), metal)]                                       # far left, alone: a bare LeafNode
    objs += [Object.sphere(0.5, (1.5 * i, 0.5, 0.3 * i), grey if i % 2 else metal) for i in range(5)]
    objs += [Object.plane(Axis.Y, 9.0 + i, 9.8 + i, -1.0, 1.0, 0.25, grey) for i in range(4)]    # coplanar: a dead group
    objs += [Object.triangle((15 + i, 0.1, 3), (15.8 + i, 0.1, 3.2), (15.4 + i, 1.2 + 0.1 * i, 3.1), metal) for i in range(5)]
    cam = dict(origin=(0.0, 8.0, 24.0), up=(0.0, 1.0, 0.0), lookat=(0.0, 0.5, 0.0), fov=80.0, width=W / 254.0, height=H / 254.0, ppi=100)
    return scenes.SceneSpec("awkward", cam, objs, BvhHeuristic.Midpoint())


def test_c_flattening_equals_the_host_mirrors(render_c, tmp_path):
    """The flattening rules of INTEGRATION.md written in C (tree build over index ranges, DFS primitive order,
    breadth-first node numbering behind a virtual root, bare-leaf flags, dead nodes, outward f32 rounding, max_depth)
    against the C++ host mirror: every byte of RrsPrim / RrsNode / RrsNodeF64 equal — on the seven-sphere row and on a
    scene built to contain every rule.  No GPU needed: the example writes the arrays before it asks for a device."""
    from rayrs_b200 import _ffi
    hdri, hpath = _hdri_file(tmp_path, 8, 4)
    flat = tmp_path / "flat.bin"
    p = _run(render_c, "row7", hpath, 8, 4, 64, 32, 1, 1, tmp_path / "o.f32", flat)
    assert flat.exists(), p.stderr
    sc = _row7_spec(64, 32).scene(hdri, upload=False)
    _compare_flat(flat.read_bytes(), sc)
    sc.close()
    # the awkward scene, through the example's scene-file mode
    spec = _awkward_spec(64, 32)
    scene_bin = tmp_path / "awkward.bin"
    _scene_file(scene_bin, spec)
    flat2 = tmp_path / "flat2.bin"
    p = _run(render_c, scene_bin, hpath, 8, 4, 64, 32, 1, 1, tmp_path / "o.f32", flat2)
    assert flat2.exists(), p.stderr
    sc = spec.scene(hdri, upload=False)
    nodes = _compare_flat(flat2.read_bytes(), sc)
    refs = [r for nd in nodes for r in (nd.ref0, nd.ref1)]
    assert any(nd.flags for nd in nodes), "no bare LeafNode child in the scene"
    assert sc.dead_nodes >= 1 and refs.count(_ffi.RRS_REF_EMPTY) >= 2, "no dead node in the scene"
    assert any((r & _ffi.RRS_REF_LEAF) and r != _ffi.RRS_REF_EMPTY and ((r >> 28) & 7) >= 1 for r in refs)
    assert sc.max_depth >= 4
    sc.close()


@pytest.mark.gpu
def test_c_host_renders_the_same_image_as_the_python_face(render_c, tmp_path):
    from rayrs_b200 import api, scenes
    W, H, spp = 192, 128, 32
    hdri, hpath = _hdri_file(tmp_path)
    out = tmp_path / "out.f32"
    p = _run(render_c, "single", hpath, 256, 128, W, H, spp, 1, out)
    assert p.returncode == 0, p.stderr
    print(p.stdout.strip())
    assert "axis ray hits object 1" in p.stdout and "nan 0 negative 0 census 0" in p.stdout
    img_c = np.fromfile(out, dtype=np.float32).reshape(H, W, 3)
    spec = scenes.diffuse_single_sphere(W, H)
    sc = spec.scene(hdri)
    img_py = api.render_gpu(spec.camera(), sc, spp, 50)
    sc.close()
    # same flattening, same camera fields, same seed: the same paths (the accumulator's atomic order is the only freedom)
    assert np.allclose(img_c, img_py, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
def test_c_flattened_row_of_spheres_matches_the_host_mirror_on_the_gpu(render_c, tmp_path):
    """row7 from the C flattener: closest hits of 2^16 rays and the image equal the host mirror's scene."""
    from rayrs_b200 import api
    W, H, spp = 320, 128, 16
    hdri, hpath = _hdri_file(tmp_path)
    spec = _row7_spec(W, H)
    sc = spec.scene(hdri)
    cam = spec.camera()
    rng = np.random.default_rng(3)
    n = 1 << 16
    org = np.array([0.0, 1.0, 0.0]) + rng.uniform(-1.0, 1.0, (n, 3)) * np.array([12.0, 1.5, 6.0])
    d = rng.normal(size=(n, 3))
    rays = np.concatenate([org, d], axis=1).astype(np.float32).astype(np.float64)
    rays_path, hits_path, out = tmp_path / "rays.f64", tmp_path / "hits.bin", tmp_path / "out.f32"
    rays.tofile(rays_path)
    p = _run(render_c, "row7", hpath, 256, 128, W, H, spp, 1, out, tmp_path / "flat.bin", rays_path, n, hits_path)
    assert p.returncode == 0, p.stderr
    print(p.stdout.strip())
    assert "census 0" in p.stdout
    raw = hits_path.read_bytes()
    ids_c = np.frombuffer(raw[:4 * n], dtype=np.int32)
    t_c = np.frombuffer(raw[4 * n:], dtype=np.float64)
    ids_py, t_py = sc.intersect(rays, 32)
    assert np.array_equal(ids_c, ids_py) and np.array_equal(t_c, t_py)
    assert (ids_c >= 1).mean() > 0.05 and (ids_c == 0).mean() > 0.05
    img_c = np.fromfile(out, dtype=np.float32).reshape(H, W, 3)
    img_py = api.render_gpu(cam, sc, spp, 50)
    assert np.allclose(img_c, img_py, rtol=1e-5, atol=1e-6)
    sc.close()


@pytest.mark.gpu
def test_c_flattened_awkward_scene_matches_the_host_mirror_on_the_gpu(render_c, tmp_path):
    """bare LeafNode children, a dead node, triangles and planes: the C flattener's scene gives the host mirror's hits."""
    from rayrs_b200 import _ffi
    W, H, spp = 256, 128, 8
    hdri, hpath = _hdri_file(tmp_path)
    spec = _awkward_spec(W, H)
    scene_bin = tmp_path / "awkward.bin"
    _scene_file(scene_bin, spec)
    sc = spec.scene(hdri)
    rng = np.random.default_rng(4)
    n = 1 << 16
    org = np.array([-2.0, 1.0, 1.0]) + rng.uniform(-1.0, 1.0, (n, 3)) * np.array([22.0, 1.5, 5.0])
    rays = np.concatenate([org, rng.normal(size=(n, 3))], axis=1).astype(np.float32).astype(np.float64)
    rays_path, hits_path, out = tmp_path / "rays.f64", tmp_path / "hits.bin", tmp_path / "out.f32"
    rays.tofile(rays_path)
    p = _run(render_c, scene_bin, hpath, 256, 128, W, H, spp, 1, out, tmp_path / "flat.bin", rays_path, n, hits_path)
    assert p.returncode == 0, p.stderr
    raw = hits_path.read_bytes()
    ids_c = np.frombuffer(raw[:4 * n], dtype=np.int32)
    t_c = np.frombuffer(raw[4 * n:], dtype=np.float64)
    ids_py, t_py = sc.intersect(rays, 32)
    assert np.array_equal(ids_c, ids_py) and np.array_equal(t_c, t_py)
    # objects under the dead node are unreachable, as in the reference: no ray reports them
    nodes, _, order, _, _, _ = sc.flat()
    reach, todo = set(), [0]
    while todo:
        nd = nodes[todo.pop()]
        for r in (nd.ref0, nd.ref1):
            if r == _ffi.RRS_REF_EMPTY:
                continue
            if r & _ffi.RRS_REF_LEAF:
                reach.update(int(order[(r & 0x0FFFFFFF) + k]) for k in range(((r >> 28) & 7) + 1))
            else:
                todo.append(r)
    dead = sorted(set(range(sc.n_prims)) - reach)
    assert dead and not np.isin(ids_c, dead).any(), dead
    assert len(np.unique(ids_c)) >= 10
    sc.close()


@pytest.mark.gpu
def test_c_host_renders_on_two_gpus_through_one_call(render_c, tmp_path):
    """rrs_scene_create_multi + rrs_comm_init_all + rrs_render_multi from pure C: the two-GPU image equals the
    one-GPU image (same global sample indices; fp32 summation order is the only freedom), census exact."""
    from rayrs_b200 import _ffi
    if _ffi.cuda_lib().rrs_device_count() < 2:
        pytest.skip("needs two GPUs")
    W, H, spp = 320, 128, 16
    _, hpath = _hdri_file(tmp_path)
    o1, o2 = tmp_path / "o1.f32", tmp_path / "o2.f32"
    p1 = _run(render_c, "row7", hpath, 256, 128, W, H, spp, 1, o1)
    p2 = _run(render_c, "row7", hpath, 256, 128, W, H, spp, 2, o2)
    assert p1.returncode == 0 and p2.returncode == 0, p1.stderr + p2.stderr
    print(p2.stdout.strip())
    assert "on 2 GPU(s)" in p2.stdout and "census 0" in p2.stdout
    a = np.fromfile(o1, dtype=np.float32).reshape(H, W, 3)
    b = np.fromfile(o2, dtype=np.float32).reshape(H, W, 3)
    assert np.max(np.abs(a - b) / (np.abs(a) + 1e-3)) <= 1e-5
