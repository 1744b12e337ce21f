#!/usr/bin/env python3
"""A/B of the two pipelines on every scene: same image (bit for bit), device time of each.
  python tools/pipeline_ab.py [spp_scale]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
CASES = [("cornell", 64), ("cornell_smoke", 64), ("random", 64), ("final", 32), ("mesh", 4)]
for name, spp in CASES:
    if only and name not in only:
        continue
    spp = max(int(spp * scale), 1)
    hs = rt.HostScene(name)
    dev = rt.DeviceScene(hs.scene_desc)
    res = {}
    for label, flag in (("megakernel", rt._abi.FLAG_MEGAKERNEL), ("wavefront", rt._abi.FLAG_WAVEFRONT)):
        opts = rt.render_opts(seed=1, integrator=hs.integrator, flags=flag)
        dev.render(hs.camera, hs.width, hs.height, max(spp // 4, 1), hs.max_depth, opts)
        img, st = dev.render(hs.camera, hs.width, hs.height, spp, hs.max_depth, opts)
        res[label] = (img, st)
        print("%-14s %-10s %dx%d spp %d: %8.2f ms  %8.1f Mpaths/s  %8.1f Mrays/s  launches %d" % (
            name, label, hs.width, hs.height, spp, st.render_ms, st.paths / st.render_ms / 1e3, st.rays / st.render_ms / 1e3,
            st.kernel_launches), flush=True)
    same = np.array_equal(res["megakernel"][0], res["wavefront"][0], equal_nan=True)
    print("%-14s identical images: %s   speed-up wavefront/megakernel: %.2fx" % (
        name, same, res["megakernel"][1].render_ms / res["wavefront"][1].render_ms), flush=True)
