#!/usr/bin/env python3
"""Longer renders than gpu_probe.py (tails amortised): scene:spp pairs on argv."""
import os, sys, zlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402
for arg in sys.argv[1:]:
    name, spp = arg.split(":")
    spp = int(spp)
    hs = rt.HostScene(name)
    dev = rt.DeviceScene(hs.scene_desc)
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    dev.render(hs.camera, hs.width, hs.height, max(spp // 16, 1), hs.max_depth, opts)
    img, st = dev.render(hs.camera, hs.width, hs.height, spp, hs.max_depth, opts)
    print("%-14s spp %4d  %9.1f ms  %8.1f Mpaths/s  %8.1f Mrays/s  launches %d  image crc %08x" % (name, spp, st.render_ms, st.paths / st.render_ms / 1e3, st.rays / st.render_ms / 1e3, st.kernel_launches, zlib.crc32(img.tobytes())), flush=True)
