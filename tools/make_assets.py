#!/usr/bin/env python3
"""Regenerates assets/ from the reference checkout (run in the build container only).

The GPU box has no /root/reference, and this image has no libjpeg headers, so the two
data files the reference's scenes read at run time are converted once, here:

  earthmap.jpg (src/main.rs:248,491: image::open(..).to_rgb8())
      -> assets/earthmap_1024x512.rgb   tightly packed RGB8, row 0 at the top
  teapot.obj   (src/mesh.rs:40: tobj::load_obj)
      -> assets/teapot.obj              byte-identical copy

They are input data, not source code.  The JPEG is decoded with PIL (libjpeg-turbo);
the `image` crate's decoder may differ by an LSB or two per texel, which cannot
affect parity because the oracle and the device library read the same decoded bytes.
"""
import hashlib
import os
import shutil
import sys

from PIL import Image

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "assets")


def main():
    os.makedirs(OUT, exist_ok=True)
    im = Image.open(os.path.join(REF, "earthmap.jpg")).convert("RGB")
    assert im.size == (1024, 512), im.size
    raw = im.tobytes()
    with open(os.path.join(OUT, "earthmap_1024x512.rgb"), "wb") as f:
        f.write(raw)
    shutil.copyfile(os.path.join(REF, "teapot.obj"), os.path.join(OUT, "teapot.obj"))
    os.chmod(os.path.join(OUT, "teapot.obj"), 0o644)
    for name in ("earthmap_1024x512.rgb", "teapot.obj"):
        with open(os.path.join(OUT, name), "rb") as f:
            print(name, hashlib.sha256(f.read()).hexdigest())


if __name__ == "__main__":
    main()
