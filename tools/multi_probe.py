#!/usr/bin/env python3
"""Single-process multi-GPU entry (rt_scene_group_create / rt_render_multi) on every GPU of the box:
device time against one GPU of the same group, and the image against the one-GPU image.
  python tools/multi_probe.py [scene:spp ...]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

cases = [a.split(":") for a in sys.argv[1:]] or [["cornell", "1000"], ["final", "256"], ["mesh", "32"]]
n = rt.device_count()
for name, spp in cases:
    spp = int(spp)
    hs = rt.HostScene(name)
    W, H, depth = hs.width, hs.height, hs.max_depth
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    t0 = time.perf_counter()
    group = rt.SceneGroup(hs.scene_desc)  # all devices; one compile, n uploads
    t_create = time.perf_counter() - t0
    group.render(hs.camera, W, H, max(spp // 8, n), depth, opts, want_sums=False)  # warm-up: contexts, modules, scratch
    t0 = time.perf_counter()
    img, st = group.render(hs.camera, W, H, spp, depth, opts)
    t_wall = time.perf_counter() - t0
    root = group.scene(0)
    root.render(hs.camera, W, H, max(spp // 8, 1), depth, opts, want_sums=False)
    one, s1 = root.render(hs.camera, W, H, spp, depth, opts)
    rel = float(np.abs(img - one).max() / max(1.0, float(np.abs(one).max())))
    last = None
    if n > 1:  # the last GPU on its own, half the samples: is it as fast as the root?
        peer = group.scene(n - 1)
        half = rt.render_opts(seed=1, integrator=hs.integrator, sample_begin=0, sample_count=max(spp // n, 1))
        peer.render(hs.camera, W, H, spp, depth, half, want_sums=False)
        _, sl = peer.render(hs.camera, W, H, spp, depth, half, want_sums=False)
        last = round(sl.render_ms, 2)
    print(json.dumps({"scene": name, "spp": spp, "n_gpus": n, "group_create_s": round(t_create, 3),
                      "multi_render_ms": round(st.render_ms, 2), "multi_wall_ms": round(t_wall * 1e3, 2),
                      "one_gpu_render_ms": round(s1.render_ms, 2), "speed_up": round(s1.render_ms / st.render_ms, 3),
                      "mpaths_per_s": round(st.paths / st.render_ms / 1e3, 1), "mrays_per_s": round(st.rays / st.render_ms / 1e3, 1),
                      "last_gpu_alone_block_ms": last, "rays_equal": int(st.rays) == int(s1.rays), "max_rel_diff_vs_one_gpu": rel,
                      "info": root.render_info}), flush=True)
    group.close()
