#!/usr/bin/env python3
"""Render one scene twice (warm-up + measured) — the command ncu wraps to capture render_kernel.
  python tools/profile_scene.py <scene> <spp> [width height]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

name, spp = sys.argv[1], int(sys.argv[2])
hs = rt.HostScene(name)
w, h = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (hs.width, hs.height)
dev = rt.DeviceScene(hs.scene_desc)
opts = rt.render_opts(seed=1, integrator=hs.integrator)
dev.render(hs.camera, w, h, max(spp // 4, 1), hs.max_depth, opts)
_, st = dev.render(hs.camera, w, h, spp, hs.max_depth, opts)
print("%s %dx%d spp %d: %.2f ms  %.1f Mpaths/s  %.1f Mrays/s" % (name, w, h, spp, st.render_ms, st.paths / st.render_ms / 1e3, st.rays / st.render_ms / 1e3))
