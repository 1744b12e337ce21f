#!/usr/bin/env python3
"""Quick throughput probe on one GPU: renders each config at its full resolution with a
reduced sample count and prints Mpaths/s and Mrays/s (device time, CUDA events)."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

SPP = {"random": 64, "cornell": 64, "cornell_smoke": 64, "final": 64, "mesh": 8}


def main():
    names = sys.argv[1:] or ["cornell", "cornell_smoke", "random", "final", "mesh"]
    for name in names:
        t0 = time.time()
        hs = rt.HostScene(name)
        t1 = time.time()
        dev = rt.DeviceScene(hs.scene_desc)
        t2 = time.time()
        opts = rt.render_opts(seed=1, integrator=hs.integrator)
        spp = SPP.get(name, 16)
        dev.render(hs.camera, hs.width, hs.height, max(spp // 8, 1), hs.max_depth, opts)  # warm-up
        img, st = dev.render(hs.camera, hs.width, hs.height, spp, hs.max_depth, opts)
        print("%-14s %4dx%-4d spp %3d  build %.2fs compile+upload %.2fs (%.1f MB)  render %.1f ms  %.1f Mpaths/s  %.1f Mrays/s  seg/path %.2f  nonfinite %d"
              % (name, hs.width, hs.height, spp, t1 - t0, t2 - t1, dev.device_bytes / 1e6, st.render_ms,
                 st.paths / st.render_ms / 1e3, st.rays / st.render_ms / 1e3, st.rays / st.paths, st.nonfinite_samples), flush=True)


if __name__ == "__main__":
    main()
