"""The whole-job path in a process of its own (no torch): host scene graph -> flatten -> rt_scene_group_create ->
rt_render_multi -> rt_encode_ppm (format_color + P3 text on the GPU) -> destroy, the path `rtb200_render > image.ppm`
takes.  One warm-up run (CUDA context, module load), then `runs` timed runs.  Phase times go to stderr
(RTB200_MULTI_TIMING), one JSON line to stdout: {"ms": [...], "device_ms": [...], "bytes": n, "first_ms": ...}.

    python tools/ppm_phase_probe.py [scene [width height spp [depth [runs]]]]

bench.py runs this as a child for its `e2e_ppm` leg."""
import json
import os
import sys
import time

os.environ["RTB200_MULTI_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracinginrust_b200 as rt  # noqa: E402


def main(argv):
    name = argv[1] if len(argv) > 1 else "cornell"
    hs = rt.HostScene(name)
    W, H, spp = (int(x) for x in argv[2:5]) if len(argv) > 4 else (hs.width, hs.height, hs.spp)
    depth = int(argv[5]) if len(argv) > 5 else hs.max_depth
    runs = int(argv[6]) if len(argv) > 6 else 3
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    ms, dev_ms, n_bytes, first = [], [], 0, None
    for it in range(runs + 1):
        t0 = time.perf_counter()
        ppm, stats = hs.render_ppm(W, H, spp, depth, opts, n_gpus=1)
        dt = (time.perf_counter() - t0) * 1e3
        print("run %d: %.1f ms wall, %.1f ms on device, %d bytes" % (it, dt, stats.render_ms, len(ppm)), file=sys.stderr)
        n_bytes = len(ppm)
        if it == 0:
            first = dt
            if dt > 20000.0:  # a render of many seconds: the warm-up run is the measurement
                ms.append(dt)
                dev_ms.append(stats.render_ms)
                break
        else:
            ms.append(dt)
            dev_ms.append(stats.render_ms)
            if dt > 2000.0:  # one timed run is enough when a run takes seconds
                break
    print(json.dumps({"ms": ms, "device_ms": dev_ms, "bytes": n_bytes, "first_ms": first}), flush=True)


if __name__ == "__main__":
    main(sys.argv)
