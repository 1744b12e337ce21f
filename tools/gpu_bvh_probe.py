#!/usr/bin/env python3
"""rt_scene_create with the host's SAH build against rt_scene_create_ex(RT_CREATE_GPU_BVH): time to a resident scene
and device time of a render on each tree, same image expected.   python tools/gpu_bvh_probe.py [scene [spp [detail]]]"""
import os
import sys
import time
import zlib

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mesh"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
detail = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hs = rt.HostScene(name, construction_seed=1, mesh_detail=detail)
opts = rt.render_opts(seed=1, integrator=hs.integrator)
rt.DeviceScene(hs.scene_desc).close()  # CUDA context, module load
for gpu in (False, True, False, True):
    t0 = time.perf_counter()
    dev = rt.DeviceScene(hs.scene_desc, gpu_bvh=gpu)
    t1 = time.perf_counter()
    dev.render(hs.camera, hs.width, hs.height, max(spp // 8, 1), hs.max_depth, opts)
    img, st = dev.render(hs.camera, hs.width, hs.height, spp, hs.max_depth, opts)
    print("%-6s %-9s create %7.1f ms   render %dx%dx%d %8.1f ms  %8.1f Mpaths/s  %8.1f Mrays/s  crc %08x" % (
        name, "gpu tree" if gpu else "host SAH", (t1 - t0) * 1e3, hs.width, hs.height, spp, st.render_ms,
        st.paths / st.render_ms / 1e3, st.rays / st.render_ms / 1e3, zlib.crc32(img.tobytes())), flush=True)
    dev.close()
