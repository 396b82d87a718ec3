"""Image parity of the GPU render call (rrs_render through render_gpu) against the oracle.

Two protocols:
 * sample-matched: the oracle consumes the SAME counter-based uniforms as the GPU, so for opaque
   materials the two trace the same paths and the images agree far below the Monte-Carlo noise;
 * statistical: for the dielectric scenes the f64 rounding-noise coin flip of the sphere
   re-entry cannot be matched ray by ray, so the GPU image must sit within the
   oracle-vs-oracle noise floor, and per-object mean radiance must agree within its standard error.
"""
import numpy as np
import pytest

import oracle
from rayrs_b200 import api, scenes
from conftest import relrmse

pytestmark = pytest.mark.gpu


def _pair(spec, hdri):
    sc = spec.scene(hdri)
    osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
    return sc, osc, spec.camera()


OPAQUE = {
    "diffuse_single_sphere": (lambda: scenes.diffuse_single_sphere(192, 128), 64, 8),
    "copper_single_sphere": (lambda: scenes.copper_single_sphere(192, 128), 64, 50),
    "spheres_metallic": (lambda: scenes.cook_torrance_spheres_metallic(240, 96), 64, 50),
    "spheres_plastic": (lambda: scenes.cook_torrance_spheres_plastic(240, 96), 64, 50),
    "copper_torus_5k": (lambda: scenes.copper_torus(50, 50, 192, 128), 32, 50),
    "emissive_room": (lambda: scenes.emissive_room(240, 160), 64, 50),   # Emission::Emissive + box_geom planes (11 primitives: BVH path)
}


@pytest.mark.parametrize("name", sorted(OPAQUE))
def test_sample_matched_image(name, hdri_small):
    builder, spp, mb = OPAQUE[name]
    spec = builder()
    sc, osc, cam = _pair(spec, hdri_small)
    W, H = cam.x_pixels(), cam.y_pixels()
    img = api.render_gpu(cam, sc, spp, mb).astype(np.float64)
    st = sc.stats()
    ref, ost = osc.render(cam.derived17(), W, H, spp, max_bounces=mb)
    noise, _ = osc.render(cam.derived17(), W, H, spp, max_bounces=mb, seed=999)
    e, e0 = relrmse(img, ref), relrmse(noise, ref)
    # same paths: the residual is fp32 rounding plus the rare path whose branch decision flips
    assert e < 0.02 * e0, (e, e0)
    rel = np.abs(img - ref).max(axis=2) / (np.abs(ref).max(axis=2) + 0.1)
    assert (rel > 1e-3).mean() < 0.01
    assert abs(st["rays"] - ost["rays"]) <= 2e-4 * ost["rays"]
    assert st["paths"] == W * H * spp
    assert st["nan_pixels"] == ost["nan_pixels"] == 0 and st["negative_pixels"] == 0
    assert abs(img.mean() - ref.mean()) < 2e-4 * ref.mean()
    print(f"[{name}] relRMSE matched {e:.2e} vs noise floor {e0:.2e}; rays gpu {st['rays']} oracle {ost['rays']}; "
          f"pixels off by >1e-3: {(rel > 1e-3).sum()}")
    sc.close()


DIELECTRIC = {
    "glass_single_sphere": (lambda: scenes.glass_single_sphere(160, 96), 64),
    "cook_torrance_glass_single_sphere": (lambda: scenes.cook_torrance_glass_single_sphere(160, 96), 64),
    "spheres_frosted_glass": (lambda: scenes.cook_torrance_spheres_frosted_glass(240, 80), 64),
    "material_test": (lambda: scenes.material_test(280, 56), 64),
    "spheres_ct_refract": (lambda: scenes.cook_torrance_spheres_cook_torrance_refract(240, 80), 64),
    "mixed_small": (lambda: scenes.mixed_scene(40, 40, 240, 135), 48),
    "glass_torus_3200": (lambda: scenes.glass_torus(40, 40, 192, 128), 48),   # glass_suzanne: a refractive triangle mesh
}


@pytest.mark.parametrize("name", sorted(DIELECTRIC))
def test_statistical_image(name, hdri_small):
    builder, spp = DIELECTRIC[name]
    spec = builder()
    sc, osc, cam = _pair(spec, hdri_small)
    W, H = cam.x_pixels(), cam.y_pixels()
    c17 = cam.derived17()
    A, _ = osc.render(c17, W, H, spp, seed=101, rng_mode=oracle.RNG_WIDE)
    B, _ = osc.render(c17, W, H, spp, seed=202, rng_mode=oracle.RNG_WIDE)
    G = api.render_gpu(cam, sc, spp, 50, seed=303).astype(np.float64)
    e0, eg = relrmse(A, B), relrmse(G, A)
    # tolerance = what oracle-vs-oracle measures at the same spp, not a guessed constant
    assert eg <= 1.25 * e0, (eg, e0)
    # bias probe: mean radiance per primary-hit object (region means converge much faster than pixels)
    rows, cols = np.mgrid[0:H, 0:W]
    pr = oracle.primary_rays(c17, W, H, rows.ravel(), cols.ravel(), np.zeros(W * H, dtype=np.uint32))
    ids, _ = osc.intersect(pr)
    ids = ids.reshape(H, W)
    n_obj = min(int(ids.max()) + 1, 8)  # floor + the spheres (+ first mesh triangles are lumped below)
    worst = 0.0
    for k in list(range(-1, n_obj)) + ["mesh"]:
        mask = (ids >= n_obj) if k == "mesh" else (ids == k)
        if mask.sum() < 200:
            continue
        var_pix = ((A[mask] - B[mask]) ** 2) / 2.0          # per-pixel variance estimate of one render
        se = np.sqrt(var_pix.sum(axis=0)) / mask.sum()      # standard error of the region mean
        diff = np.abs(G[mask].mean(axis=0) - A[mask].mean(axis=0))
        z = diff / (np.sqrt(2.0) * se + 1e-12)
        worst = max(worst, float(z.max()))
        assert (z < 4.5).all(), (name, k, z, G[mask].mean(axis=0), A[mask].mean(axis=0))
    # 4x the samples: the distance to the oracle must shrink like noise (~2x), not stall like bias
    A4, _ = osc.render(c17, W, H, 4 * spp, seed=404, rng_mode=oracle.RNG_WIDE)
    G4 = api.render_gpu(cam, sc, 4 * spp, 50, seed=505).astype(np.float64)
    eg4 = relrmse(G4, A4)
    assert eg4 < 0.65 * eg, (eg4, eg)
    print(f"[{name}] relRMSE gpu-vs-oracle {eg:.4f} (noise floor {e0:.4f}); at 4x spp {eg4:.4f}; worst region z {worst:.2f}")
    sc.close()


def test_config1_full_size(native_built):
    """BASELINE config 1 at full size: 512x512, 64 spp, depth 8, 2048x1024 HDRI."""
    hdri = scenes.synthetic_hdri(2048, 1024)
    cfg = scenes.CONFIGS["c1"]
    spec = cfg.specs()[0]
    sc, osc, cam = _pair(spec, hdri)
    img = api.render_gpu(cam, sc, cfg.spp, cfg.max_bounces).astype(np.float64)
    st = sc.stats()
    ref, ost = osc.render(cam.derived17(), 512, 512, cfg.spp, max_bounces=cfg.max_bounces)
    e = relrmse(img, ref)
    assert e < 2e-3, e
    assert abs(st["rays"] - ost["rays"]) <= 1e-4 * ost["rays"]
    print(f"[c1 full] relRMSE (sample matched) {e:.2e}; rays {st['rays']}; oracle {ost['seconds']:.2f} s on "
          f"{oracle.hardware_threads()} threads = {ost['rays'] / ost['seconds'] / 1e6:.2f} Mrays/s; gpu {st['device_ms']:.2f} ms")
    sc.close()


def test_shapes_queues_and_limits(hdri_small):
    """ragged image sizes (tile padding), tiny queues (many regeneration rounds), depth limits."""
    spec = scenes.cook_torrance_spheres_plastic(37, 19)
    sc, osc, cam = _pair(spec, hdri_small)
    c17 = cam.derived17()
    assert (cam.x_pixels(), cam.y_pixels()) == (37, 19)
    ref, ost = osc.render(c17, 37, 19, 40)
    base = api.render_gpu(cam, sc, 40, 50).astype(np.float64)
    assert relrmse(base, ref) < 5e-3
    rays0 = sc.stats()["rays"]
    for q in (1024, 4096, 1 << 16):
        for flags in (0, api._ffi.RRS_FLAG_FORCE_QUEUES, api._ffi.RRS_FLAG_SPLIT_KERNELS,
                      api._ffi.RRS_FLAG_SPLIT_KERNELS | api._ffi.RRS_FLAG_TIME_PHASES):
            img = api.render_gpu(cam, sc, 40, 50, queue_capacity=q, flags=flags).astype(np.float64)
            st = sc.stats()
            assert st["rays"] == rays0                      # the set of paths does not depend on queue size / kernel form
            assert np.allclose(img, base, rtol=2e-5, atol=1e-6)  # only the fp32 summation order differs
            if flags in (0, api._ffi.RRS_FLAG_FORCE_QUEUES):
                assert st["kernel_launches"] == 2  # fused wavefront: 1 render + 1 resolve
    # depth limits: max_bounces = 1 traces exactly one ray per path; 0 renders black (lib.rs:525,559)
    for mb in (1, 2, 3):
        img = api.render_gpu(cam, sc, 8, mb).astype(np.float64)
        r, o = osc.render(c17, 37, 19, 8, max_bounces=mb)
        assert sc.stats()["rays"] == o["rays"]
        assert relrmse(img, r) < 5e-3
    assert sc.stats()["rays"] > 0
    img = api.render_gpu(cam, sc, 8, 1)
    assert sc.stats()["rays"] == 37 * 19 * 8
    img = api.render_gpu(cam, sc, 4, 0)
    assert not img.any() and sc.stats()["rays"] == 0
    # spp = 1
    img = api.render_gpu(cam, sc, 1, 50).astype(np.float64)
    r, _ = osc.render(c17, 37, 19, 1)
    assert relrmse(img, r) < 5e-3
    # invalid arguments fail loudly with a status, not a crash
    with pytest.raises(api._ffi.RayrsError):
        api.render_gpu(cam, sc, 4, 300)
    # render size that disagrees with Camera::x_pixels/y_pixels (straight through the C ABI)
    import ctypes as C
    p = api.render_params(cam, 4, 50)
    p.width = 40
    buf = np.empty((19, 40, 3), dtype=np.float32)
    rc = api._ffi.cuda_lib().rrs_render(sc.handle, C.byref(cam.c), C.byref(p), buf.ctypes.data)
    assert rc == api._ffi.RRS_ERR_INVALID and b"x_pixels" in api._ffi.cuda_lib().rrs_last_error()
    sc.close()


def test_sample_split_accumulate_and_census(hdri_small):
    """The multi-GPU building block on one GPU: disjoint sample ranges accumulated into one device
    buffer equal a single render; every pixel's path counter equals spp."""
    import torch
    spec = scenes.cook_torrance_spheres_metallic(96, 40)
    sc, osc, cam = _pair(spec, hdri_small)
    W, H, spp = 96, 40, 32
    full = api.render_gpu(cam, sc, spp, 50).astype(np.float64)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda:0")
    stream = torch.cuda.current_stream().cuda_stream
    for g in range(4):  # 4 "ranks", 8 samples each, global sample indices
        api.render_accumulate(cam, sc, spp // 4, 50, acc.data_ptr(), stream, sample_offset=g * (spp // 4), spp_total=spp)
    torch.cuda.synchronize()
    a = acc.cpu().numpy().astype(np.float64)
    assert np.array_equal(a[..., 3], np.full((H, W), float(spp)))  # census: every path terminated exactly once
    assert np.allclose(a[..., :3] / spp, full, rtol=2e-5, atol=1e-6)
    out = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
    api.resolve(sc, acc.data_ptr(), W, H, spp, out.data_ptr(), True, stream)
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy(), a[..., :3] / spp, rtol=1e-6)
    # and the split result is what the oracle computes for the full sample range
    ref, _ = osc.render(cam.derived17(), W, H, spp)
    assert relrmse(a[..., :3] / spp, ref) < 5e-3
    sc.close()


def test_path_loop_and_queues_trace_the_same_paths(hdri_small):
    """Small scenes can render with the register-resident path loop (k_pathloop) instead of the queued wavefront
    kernel; both forms must trace the same set of paths and agree up to fp32 rounding."""
    F = api._ffi
    for builder, a, b in ((lambda: scenes.cook_torrance_spheres_plastic(120, 48), F.RRS_FLAG_FORCE_PATHLOOP, 0),
                          (lambda: scenes.diffuse_single_sphere(96, 64), F.RRS_FLAG_FORCE_PATHLOOP, F.RRS_FLAG_FORCE_QUEUES),
                          (lambda: scenes.cook_torrance_spheres_frosted_glass(120, 40), F.RRS_FLAG_FORCE_PATHLOOP, 0),
                          (lambda: scenes.glass_single_sphere(96, 64), F.RRS_FLAG_FORCE_PATHLOOP, F.RRS_FLAG_FORCE_QUEUES),
                          # all nine Material variants: the path loop shades with material_evaluate_cases, the
                          # queued kernel with material_evaluate_staged
                          (lambda: scenes.material_test(168, 40), F.RRS_FLAG_FORCE_PATHLOOP, 0)):
        spec = builder()
        sc = spec.scene(hdri_small, with_f64=False)
        cam = spec.camera()
        img_a = api.render_gpu(cam, sc, 24, 50, flags=a).astype(np.float64)
        st_a = sc.stats()
        img_b = api.render_gpu(cam, sc, 24, 50, flags=b).astype(np.float64)
        st_b = sc.stats()
        assert {st_a["kernel_form"], st_b["kernel_form"]} == {F.RRS_FORM_PATHLOOP, F.RRS_FORM_WAVEFRONT}
        assert st_a["kernel_launches"] == st_b["kernel_launches"] == 2
        # the two kernels shade with two forms of Material::evaluate (one arm per variant / shared stages): the
        # same operations, but the compiler contracts them into FMAs differently, so a path in a few thousand takes the
        # other side of a decision (xi < F, Russian roulette) — everything else is identical
        assert st_a["paths"] == st_b["paths"] and abs(st_a["rays"] - st_b["rays"]) <= 5e-4 * st_a["rays"]
        rel = np.abs(img_a - img_b).max(axis=2) / (np.abs(img_b).max(axis=2) + 0.1)
        assert (rel > 1e-4).mean() < 2e-2, (rel > 1e-4).mean()
        assert abs(img_a.mean() - img_b.mean()) < 1e-3 * img_b.mean()
        sc.close()


def _block_mean_with_f8_shift(img, k):
    """Average k x k blocks of a full-size image so that block (r, c) covers the film area of pixel (r, c) of the
    same view rendered k times smaller.  The camera index of pixel row r is H - r, so low-res
    row r_lo corresponds to full-size rows k*r_lo - (k-1) .. k*r_lo (and the same for columns)."""
    H, W, _ = img.shape
    pad = np.zeros((H + k - 1, W + k - 1, 3))
    pad[k - 1:, k - 1:] = img
    return pad[: H, : W].reshape(H // k, k, W // k, k, 3).mean(axis=(1, 3))


FULL = {
    # key: (block size for the low-resolution oracle comparison, oracle spp, spp override (0 = the config's))
    "c2": (16, 192, 0),
    "c3": (24, 192, 0),
    "c4": (24, 96, 0),
    "c5": (48, 64, 64),   # 4K, 4M triangles: full resolution and scene, 64 of the 4096 spp (the rest only repeats the same kernel)
}


@pytest.mark.parametrize("key", sorted(FULL))
def test_full_size_config_properties(key, native_built):
    """BASELINE configurations at their full resolution / scene size, where the oracle cannot follow ray by ray:
    size-independent properties instead — every pixel terminates exactly spp paths (census), no NaN / negative
    pixel, two disjoint sample halves add up to the one-shot render (linearity in samples, GPU-count independence),
    and the block-averaged image equals the oracle's render of the same view at 1/k resolution within its noise."""
    import torch
    k, ospp, spp_over = FULL[key]
    cfg = scenes.CONFIGS[key]
    spp = spp_over or cfg.spp
    hdri = scenes.synthetic_hdri(2048, 1024)
    W, H = cfg.width, cfg.height
    stream = torch.cuda.current_stream().cuda_stream
    for spec in cfg.specs():
        sc = spec.scene(hdri, with_f64=False)
        cam = spec.camera()
        one = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        api.render_accumulate(cam, sc, spp, cfg.max_bounces, one.data_ptr(), stream, spp_total=spp)
        st = sc.stats()
        two = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        h1 = spp // 2
        api.render_accumulate(cam, sc, h1, cfg.max_bounces, two.data_ptr(), stream, sample_offset=0, spp_total=spp)
        rays_two = sc.stats()["rays"]
        api.render_accumulate(cam, sc, spp - h1, cfg.max_bounces, two.data_ptr(), stream, sample_offset=h1, spp_total=spp)
        rays_two += sc.stats()["rays"]
        torch.cuda.synchronize()
        assert bool((one[..., 3] == float(spp)).all()) and bool((two[..., 3] == float(spp)).all())   # census
        assert st["paths"] == W * H * spp and rays_two == st["rays"]                                # same set of paths
        assert bool(torch.isfinite(one).all()) and bool((one >= 0).all())
        assert torch.allclose(one[..., :3], two[..., :3], rtol=5e-5, atol=1e-4 * spp)                # fp32 summation order only
        img = (one[..., :3] / spp).cpu().numpy().astype(np.float64)
        # the same view at 1/k resolution by the oracle (F9: the field of view does not depend on the film size)
        lw, lh = scenes.film(W // k, H // k)
        lo = type(spec)(spec.name, dict(spec.camera_args, width=lw, height=lh), spec.objects, spec.heuristic)
        lcam = lo.camera()
        assert (lcam.x_pixels(), lcam.y_pixels()) == (W // k, H // k)
        osc = oracle.OracleScene(spec.tables(), hdri.pixels, heuristic=(spec.heuristic.kind, spec.heuristic.splits), build_mode=1)
        A, _ = osc.render(lcam.derived17(), W // k, H // k, ospp, max_bounces=cfg.max_bounces, seed=11, rng_mode=oracle.RNG_WIDE)
        B, _ = osc.render(lcam.derived17(), W // k, H // k, ospp, max_bounces=cfg.max_bounces, seed=12, rng_mode=oracle.RNG_WIDE)
        G = _block_mean_with_f8_shift(img, k)
        inner = (slice(1, None), slice(1, None))   # the first block row / column is cut by the shift
        e0, eg = relrmse(A[inner], B[inner]), relrmse(G[inner], A[inner])
        # G is nearly noise-free (k*k*spp samples per block): its distance to A is A's own noise, e0 / sqrt(2)
        assert eg <= 0.85 * e0, (key, spec.name, eg, e0)
        m_g, m_o = G[inner].mean(), 0.5 * (A[inner].mean() + B[inner].mean())
        assert abs(m_g - m_o) <= 0.01 * m_o, (m_g, m_o)
        print(f"[{key} {spec.name}] {W}x{H} spp {spp}: census ok, halves == one-shot, rays {st['rays']:.3e} in {st['device_ms']:.1f} ms; "
              f"block-mean vs oracle relRMSE {eg:.4f} (oracle noise {e0:.4f}), mean {m_g:.5f} vs {m_o:.5f}")
        sc.close()
        osc.close()
