#!/usr/bin/env python3
"""Host-side scene compile of one scene, timed (no GPU needed): RTB200_COMPILE_TIMING=1 prints the phases to stderr.
  python tools/compile_probe.py mesh [repeats]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import raytracinginrust_b200 as rt  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mesh"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t = time.time()
hs = rt.HostScene(name)
print("host scene graph + flatten %.3f s" % (time.time() - t))
ts = []
for _ in range(reps):
    t = time.time()
    blob = rt.compile_scene(hs.scene_desc)
    ts.append(time.time() - t)
    n = blob.nbytes
    del blob
print("%s: rt_compile %.3f s min, %.3f s median of %d, blob %.1f MB, %d host threads" % (
    name, min(ts), sorted(ts)[len(ts) // 2], reps, n / 1e6, len(os.sched_getaffinity(0))))
