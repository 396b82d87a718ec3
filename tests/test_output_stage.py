"""Output stage, Image::to_raw_bytes (rayrs-lib/src/image.rs:193-222): the numpy restatement on hand-checked
values (CPU), and the device kernel rrs_to_raw_bytes against it, byte for byte (GPU)."""
import numpy as np
import pytest

import oracle


def test_oracle_to_raw_bytes_known_answers():
    nan = float("nan")
    img = np.array([[[0.0, 1.0, 0.5], [2.0, -1.0, 0.25], [nan, 0.2, 1e-9], [1.0 - 1e-12, 0.999, 0.001]]])
    out, census = oracle.to_raw_bytes(img, 1.0 / 2.2)
    g = 1.0 / 2.2
    expect = [[0, 255, int(255.99 * 0.5 ** g)], [255, 0, int(255.99 * 0.25 ** g)],
              [255, int(255.99 * 0.2 ** g), int(255.99 * 1e-9 ** g)],      # NaN clips to 1 (f64::min drops NaN)
              [255, int(255.99 * 0.999 ** g), int(255.99 * 0.001 ** g)]]
    assert out.tolist() == [expect]
    assert out[0, 0].tolist() == [0, 255, 186] and out[0, 1, 2] == 136      # 0.5^(1/2.2) = 0.7297, 0.25^(1/2.2) = 0.5325
    assert census == {"clamped": 1, "nan": 1, "negative": 1}
    # gamma 1: plain quantisation, 255.99 * x truncated
    out1, _ = oracle.to_raw_bytes(np.array([[[0.5, 0.00390, 0.00391]]]), 1.0)
    assert out1.tolist() == [[[127, 0, 1]]]


@pytest.mark.gpu
def test_device_to_raw_bytes_matches_the_restatement(native_built, hdri_small):
    import torch
    from rayrs_b200 import api, scenes
    spec = scenes.diffuse_single_sphere(64, 48)
    sc = spec.scene(hdri_small, with_f64=False)
    H, W, spp = 203, 317, 7
    rng = np.random.default_rng(5)
    mean = rng.random((H, W, 3)) ** 3 * 1.3                       # plenty of values on both sides of 1
    mean[rng.random((H, W)) < 0.01] = np.nan
    mean[rng.random((H, W)) < 0.01] *= -1.0
    mean[0, 0] = [0.0, 1.0, 0.5]
    acc = np.zeros((H, W, 4), dtype=np.float32)
    acc[..., :3] = (mean * spp).astype(np.float32)
    acc[..., 3] = spp
    d_acc = torch.from_numpy(acc).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    # what the host would hold: the f32 mean of rrs_resolve
    d_mean = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
    api.resolve(sc, d_acc.data_ptr(), W, H, spp, d_mean.data_ptr(), True, stream)
    held = d_mean.cpu().numpy()
    for gamma in (1.0 / 2.2, 1.0, 0.5):
        want, wc = oracle.to_raw_bytes(held, gamma)
        got, gc = api.to_raw_bytes(sc, d_acc.data_ptr(), W, H, spp, gamma, stream_ptr=stream)
        assert gc == wc
        assert np.array_equal(got, want), int((got != want).sum())
        d_out = torch.zeros((H, W, 3), dtype=torch.uint8, device="cuda")
        none, gc2 = api.to_raw_bytes(sc, d_acc.data_ptr(), W, H, spp, gamma, out_ptr=d_out.data_ptr(), stream_ptr=stream)
        assert none is None and gc2 == wc and np.array_equal(d_out.cpu().numpy(), want)
    # a real render goes through the same call
    cam = spec.camera()
    acc2 = torch.zeros((48, 64, 4), dtype=torch.float32, device="cuda")
    api.render_accumulate(cam, sc, 8, 8, acc2.data_ptr(), stream)
    d_img = torch.empty((48, 64, 3), dtype=torch.float32, device="cuda")
    api.resolve(sc, acc2.data_ptr(), 64, 48, 8, d_img.data_ptr(), True, stream)
    got, gc = api.to_raw_bytes(sc, acc2.data_ptr(), 64, 48, 8, stream_ptr=stream)
    want, wc = oracle.to_raw_bytes(d_img.cpu().numpy())
    assert np.array_equal(got, want) and gc == wc
    sc.close()
