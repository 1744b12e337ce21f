"""Phase times of the whole-job path (host scene graph -> flatten -> rt_scene_group_create -> rt_render_multi ->
rt_encode_ppm -> destroy) with RTB200_MULTI_TIMING=1: where the wall time of `rtb200_render > image.ppm` goes."""
import os
import sys
import time

os.environ["RTB200_MULTI_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginrust_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
W, H, spp = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (600, 600, 1000)
hs = rt.HostScene(name)
opts = rt.render_opts(seed=1, integrator=hs.integrator)
for it in range(3):
    t0 = time.perf_counter()
    ppm, stats = hs.render_ppm(W, H, spp, 100, opts, n_gpus=1)
    print("run %d: %.1f ms wall, %.1f ms on device, %d bytes" % (it, (time.perf_counter() - t0) * 1e3, stats.render_ms, len(ppm)), file=sys.stderr)
