#!/usr/bin/env python3
"""Output stage at 4K (BASELINE configs[4] is 3840x2160): format_color + P3 text on the GPU
(rt_encode_ppm) against the host writer (fprintf per pixel, what the reference's println! loop does).
  python tools/encode_probe.py [width height]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402
import torch  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
spp = 1024
rng = np.random.default_rng(1)
img = (rng.uniform(0, 1.1, size=(H, W, 3)) ** 2 * spp).astype(np.float32)
hs = rt.HostScene("cornell")
dev = rt.DeviceScene(hs.scene_desc)
t = torch.from_numpy(img).cuda()
torch.cuda.synchronize()
for k in range(3):
    t0 = time.perf_counter()
    ppm = dev.encode_ppm(W, H, spp, sums_ptr=t.data_ptr())
    t_gpu = time.perf_counter() - t0
for k in range(2):
    t0 = time.perf_counter()
    rgb = dev.encode_rgb8(W, H, spp, sums_ptr=t.data_ptr())
    t_rgb = time.perf_counter() - t0
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "o.ppm")
    t0 = time.perf_counter()
    rt.write_ppm(path, img, spp)
    t_host = time.perf_counter() - t0
    same = open(path, "rb").read() == ppm
print("%dx%d: rt_encode_ppm %.1f ms (%d bytes, incl. D2H to pageable memory)  rt_encode_rgb8 %.1f ms  host writer %.1f ms  identical: %s"
      % (W, H, t_gpu * 1e3, len(ppm), t_rgb * 1e3, t_host * 1e3, same))
