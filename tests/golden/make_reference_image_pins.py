"""Measurements of the reference's OWN output — the six renders it ships under examples/*.png — as a small fixture:

    python tests/golden/make_reference_image_pins.py        # needs /root/reference (this container), PIL, scipy

The PNGs were rendered by rayrs itself with an HDRI that is not shipped and at an unknown spp, so their COLOURS pin
nothing.  Their GEOMETRY does: where the silhouettes of the spheres fall in the image depends only on Camera::new
(lib.rs:99-133, the F9 field-of-view quirk and the ppc rounding included), Camera::x_pixels / y_pixels,
generate_primary_ray (lib.rs:202-210), the mirrored pixel mapping of the tile loop (main.rs:71-76, F8), the scene
constants of test_scenes.rs and Sphere::intersect — none of which any reference TEST pins.  And which end of the sphere
rows is the mirror-like one pins the orientation of that mapping and the roughness assignment.

What is stored (the images themselves are not copied): per image, the pixel coordinates of its strong luminance edges
(smoothed-gradient maxima along the gradient direction, a Canny-style edge map) and, for the sphere rows, the mean
gradient magnitude of every 4x4 pixel block (where the image is sharp).  tests/test_reference_images.py holds the
oracle's and the GPU's primary-ray hit masks to them.
"""
import sys
from pathlib import Path

import numpy as np
from PIL import Image
from scipy.ndimage import gaussian_filter

REF = Path("/root/reference/examples")
OUT = Path(__file__).resolve().parent / "reference_image_pins.npz"
IMAGES = ["diffuse_single_sphere", "copper_sphere", "cook_torrance_glass_sphere", "spheres_metallic", "spheres_plastic",
          "cook_torrance_spheres_frosted_glass"]
SIGMA, THRESHOLD, BLOCK = 1.5, 1.5, 4


def edge_map(lum):
    s = gaussian_filter(lum, SIGMA)
    gy, gx = np.gradient(s)
    g = np.hypot(gx, gy)
    ang = np.arctan2(gy, gx)
    dx, dy = np.rint(np.cos(ang)).astype(int), np.rint(np.sin(ang)).astype(int)
    H, W = g.shape
    yy, xx = np.mgrid[0:H, 0:W]

    def at(y, x):
        return g[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)]
    keep = (g >= at(yy + dy, xx + dx)) & (g >= at(yy - dy, xx - dx)) & (g > THRESHOLD)
    return keep, g


def main():
    if not REF.exists():
        sys.exit("the reference tree is not present: nothing to measure")
    out = {"sigma_threshold_block": np.array([SIGMA, THRESHOLD, BLOCK])}
    for name in IMAGES:
        img = np.asarray(Image.open(REF / f"{name}.png").convert("RGB")).astype(np.float64)
        lum = img.mean(axis=2)
        keep, g = edge_map(lum)
        ys, xs = np.nonzero(keep)
        out[f"{name}/size"] = np.array([lum.shape[1], lum.shape[0]], dtype=np.int32)  # width, height
        out[f"{name}/edge_xy"] = np.stack([xs, ys], axis=1).astype(np.uint16)
        if name.startswith("spheres_"):
            H, W = g.shape
            hb, wb = H // BLOCK, W // BLOCK
            out[f"{name}/sharpness_4x4"] = g[:hb * BLOCK, :wb * BLOCK].reshape(hb, BLOCK, wb, BLOCK).mean(axis=(1, 3)).astype(np.float16)
        print(name, lum.shape[::-1], "edge pixels", xs.size)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
