"""CPU tier: random scene graphs through the scene compiler and the device source, against the oracle.

The five configs and the reference's other scenes exercise the wrapper / container combinations main.rs happens
to build.  The boundary accepts any graph a `flatten()` visitor can emit (include/rtb200.h), so seeded random
graphs - lists in BVHs in rotated translated lists, flipped and moving primitives, media bounded by instanced
boxes, both integrators - are compiled by compile.cpp and traced by the g++ build of trace.cuh
(tests/native), and must return what the oracle's literal object tree returns: same object, same t, same
radiance.  No GPU; `test_gpu_parity.py::test_random_scene_graphs_on_device` runs the same graphs on the B200.
"""
import os
import sys

import numpy as np
import pytest

from graph_fuzz import GraphMaker, fuzz_camera
from util import compare_hits, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "native"))

N_GRAPHS = 40


@pytest.fixture(scope="module")
def toh():
    import trace_on_host as m
    return m


@pytest.mark.parametrize("seed", range(N_GRAPHS))
def test_random_scene_graph(rt, orc, toh, seed):
    g = GraphMaker(rt, 1000 + seed)
    sd = g.make()
    comp, osc = toh.CompiledOnHost(sd), orc.OracleScene(sd)
    counts = comp.check_tables()
    # 1. first hits: same object, same face, same numbers
    rays = g.rays(30000)
    hd, ho = comp.trace_first_hit(rays), osc.trace_first_hit(rays)
    r = compare_hits(hd, ho)
    print(seed, counts, r)
    assert r["hits"] > 100
    assert r["id_mismatch"] == 0 and r["front_face_mismatch"] == 0 and r["material_mismatch"] == 0
    assert r["t_max_rel"] <= 1e-9 and r["normal_max_abs"] <= 1e-9 and r["uv_max_abs"] <= 1e-9
    # 2. per-path radiance under the same Philox streams, both integrators (main.rs:41-120 and :84-85)
    cam = fuzz_camera(rt)
    W = H = 64
    ids = np.random.default_rng(seed)
    px, py, s = (ids.integers(0, W, 3000, dtype=np.uint32), ids.integers(0, H, 3000, dtype=np.uint32),
                 ids.integers(0, 64, 3000, dtype=np.uint32))
    for integrator in (rt.INTEGRATOR_HEAD, rt.INTEGRATOR_LEGACY):
        opts = rt.render_opts(seed=seed + 1, integrator=integrator)
        rd, segd = comp.path_radiance(cam, W, H, 50, opts, px, py, s)
        ro, sego = osc.path_radiance(cam, W, H, 50, opts, px, py, s)
        nan_d, nan_o = np.isnan(rd).any(axis=1), np.isnan(ro).any(axis=1)
        err = rel_err(np.nan_to_num(rd), np.nan_to_num(ro), floor=1e-9).max(axis=1)
        ok = ((err <= 1e-4) & ~nan_d & ~nan_o) | (nan_d & nan_o)
        print(seed, "integrator", integrator, "ok %.5f max err %.2e segments %.3f / %.3f" % (ok.mean(), err.max(), segd.mean(), sego.mean()))
        assert ok.mean() >= 0.999
    comp.close()
    osc.close()
