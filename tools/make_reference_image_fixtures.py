#!/usr/bin/env python
"""Fixtures derived from two pictures the reference publishes (README.md:6-12 links img/): what tests/test_reference_images.py
compares a render with.  Run in the container that has /root/reference; the .npz travels, the pictures do not.

  img/earth.png           900x600  `Scene::Earth` (main.rs:246-254, camera :667-680): a sphere with the earth map
                          -> `earth_half`: RGB box-averaged 2x2 to 450x300, uint8
  img/TextureMapping.png  900x600  `Scene::TwoSpheres` (main.rs:212-227, camera :642-649): two checkered spheres
                          -> `checker_dark`: one bit per pixel, set where the pixel is a dark (0.3, 0.3, 1) square
                             ((R+1)/(B+1) < 0.65: the ratio does not depend on how bright the square is lit)
  img/CornellBox.png      600x600  `Scene::CornellBox` (main.rs:278-311, camera :700-719) at a revision whose tall box
                          was still white (HEAD: Metal, main.rs:306)
                          -> `cornell_quarter`: RGB box-averaged 4x4 to 150x150, uint8
  img/volume.png          600x600  `Scene::CornellSmoke` (main.rs:313-346) at a revision whose smoke still scattered
                          (the `old method` of main.rs:82-84; at HEAD Isotropic has no scatter_mc_method and absorbs, §Q6)
                          -> `smoke_quarter`: RGB box-averaged 4x4 to 150x150, uint8

All four were rendered by an older revision than HEAD (a white-to-blue sky gradient behind the scene instead of the constant
background of main.rs:670, unknown spp, unseeded rand): the sky and the noise are not comparable, the geometry is - camera,
sphere intersection, get_sphere_uv, the nearest-texel lookup, the JPEG decode, the checker's sin product, the gamma-2
8-bit output - and that is what the tests use."""
import os
import sys

import numpy as np
from PIL import Image

SRC = "/root/reference/img"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "reference_images.npz")


def quarter(name):
    a = np.asarray(Image.open(os.path.join(SRC, name)).convert("RGB")).astype(np.float64)
    assert a.shape == (600, 600, 3)
    return np.rint(a.reshape(150, 4, 150, 4, 3).mean(axis=(1, 3))).astype(np.uint8)


def main():
    earth = np.asarray(Image.open(os.path.join(SRC, "earth.png")).convert("RGB")).astype(np.float64)
    h, w, _ = earth.shape
    assert (w, h) == (900, 600)
    earth_half = np.rint(earth.reshape(h // 2, 2, w // 2, 2, 3).mean(axis=(1, 3))).astype(np.uint8)
    chk = np.asarray(Image.open(os.path.join(SRC, "TextureMapping.png")).convert("RGB")).astype(np.float64)
    assert chk.shape == (600, 900, 3)
    dark = (chk[..., 0] + 1.0) / (chk[..., 2] + 1.0) < 0.65
    box = np.asarray(Image.open(os.path.join(SRC, "CornellBox.png")).convert("RGB")).astype(np.float64)
    assert box.shape == (600, 600, 3)
    cornell_quarter = np.rint(box.reshape(150, 4, 150, 4, 3).mean(axis=(1, 3))).astype(np.uint8)
    np.savez_compressed(OUT, earth_half=earth_half, checker_dark=np.packbits(dark), checker_shape=np.array(dark.shape),
                        cornell_quarter=cornell_quarter, smoke_quarter=quarter("volume.png"))
    print("wrote", os.path.normpath(OUT), os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
