#!/usr/bin/env python3
"""Writes tests/golden/paths_<scene>.npz: oracle outputs on fixed (pixel, sample) paths.

There are no reference-run fixtures (no Rust toolchain, reference RNG is OS-seeded), so
these vectors pin the ORACLE against accidental change and give the GPU tests a
committed target that does not need the oracle library at run time.  Regenerate with
    python tests/golden/make_golden.py
after any deliberate change to oracle/oracle.cpp or the scene constructors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.dirname(HERE))

import oracle_py as orc  # noqa: E402
import raytracinginrust_b200 as rt  # noqa: E402
from util import SCENES, host_scene, random_path_ids  # noqa: E402

W = H = 64
SPP = 64
N = 192
SEED = 7


def main():
    for name in SCENES:
        hs = host_scene(rt, name)
        osc = orc.OracleScene(hs.scene_desc)
        opts = rt.render_opts(seed=SEED, integrator=hs.integrator)
        px, py, s = random_path_ids(N, W, H, SPP, seed=1234)
        rays = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
        hits = osc.trace_first_hit(rays)
        rgb, seg = osc.path_radiance(hs.camera, W, H, 50, opts, px, py, s)
        np.savez_compressed(os.path.join(HERE, "paths_%s.npz" % name), px=px, py=py, sample=s, rays=rays, hits=hits,
                            rgb=rgb, segments=seg, width=W, height=H, max_depth=50, seed=SEED,
                            integrator=hs.integrator)
        print(name, "hits", int((hits["node"] >= 0).sum()), "mean segments %.2f" % seg.mean(), "mean rgb", rgb.mean(axis=0))


if __name__ == "__main__":
    main()
