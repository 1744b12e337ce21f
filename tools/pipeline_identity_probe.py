#!/usr/bin/env python3
"""Megakernel against wavefront at full image sizes: are the fp32 sum images bit-identical?
  python tools/pipeline_identity_probe.py final:8 final:64 final:512 random:100"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

for arg in sys.argv[1:]:
    name, spp = arg.split(":")
    spp = int(spp)
    hs = rt.HostScene(name)
    dev = rt.DeviceScene(hs.scene_desc)
    w, h = hs.width, hs.height
    out = {}
    for flag, label in ((rt._abi.FLAG_MEGAKERNEL, "megakernel"), (rt._abi.FLAG_WAVEFRONT, "wavefront")):
        img, st = dev.render(hs.camera, w, h, spp, hs.max_depth, rt.render_opts(seed=1, integrator=hs.integrator, flags=flag))
        out[label] = (img, st)
        print("%s %s x%d: %.1f ms, paths %d rays %d nonfinite %d  %s" % (name, label, spp, st.render_ms, st.paths, st.rays, st.nonfinite_samples, dev.render_info.get("chunks")))
    a, b = out["megakernel"][0], out["wavefront"][0]
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    bad = np.argwhere(~same.all(axis=2))
    print("  differing pixels: %d of %d" % (len(bad), w * h))
    for y, x in bad[:8]:
        print("   (%d, %d): megakernel %s wavefront %s" % (x, y, a[y, x], b[y, x]))
    dev.close()
