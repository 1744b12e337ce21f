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
    # the second half of the graphs adds Perlin / image / nested textures, the PBR material and light lists with
    # objects the reference cannot sample
    g = GraphMaker(rt, 1000 + seed, rich=seed >= N_GRAPHS // 2)
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
        # A light the reference cannot sample sends its rays along exactly (1, 0, 0) (hit.rs:29-30).  From a point
        # that lies exactly in an axis-aligned face such a ray meets the reference's unguarded arithmetic - 0/0 in
        # AARect::hit passes every rejecting comparison (rect.rs:51-58) - which the search code does not follow
        # (DESIGN.md "Known deviations"); those paths are already inf / NaN in both.
        assert ok.mean() >= (0.99 if g.has_default_light and integrator == rt.INTEGRATOR_HEAD else 0.999)
    comp.close()
    osc.close()


def test_flat_boxes_under_a_bvh_are_never_entered(rt, orc, toh):
    """bvh.rs:56-63 gives every object a Leaf node whose box is tested first, and AABB::hit (aabb.rs:31) rejects with
    `t_out <= t_in`: an axis-aligned flat triangle directly under a BVH (an OBJ cube loaded by mesh.rs and wrapped by
    main.rs:442) has a box without extent on one axis and is invisible in the reference.  The scene compiler leaves
    such objects out, so the CUDA path shows the same hole; the same triangles in a plain list are hit by both."""
    A = rt._abi

    def scene(container):
        b = rt.SceneBuilder()
        m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
        quad = [b.triangle((-2, 1, -2), (2, 1, -2), (2, 1, 2), m), b.triangle((-2, 1, -2), (2, 1, 2), (-2, 1, 2), m)]  # flat in y
        tilted = b.triangle((-2, 0, -2), (2, 0.5, -2), (0, 0.25, 2), m)
        ball = b.sphere((0, -3, 0), 1.0, m)
        light = b.flip(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 9, b.diffuse_light(b.constant_texture((4, 4, 4)))))
        kids = quad + [tilted, ball]
        world = b.list([getattr(b, container)(kids), light])
        return b.finish(world, b.list([light])), quad, tilted

    rng = np.random.default_rng(3)
    n = 4000
    rays = np.zeros(n, dtype=A.RAY_DTYPE)
    rays["origin"] = np.stack([rng.uniform(-1.5, 1.5, n), np.full(n, 6.0), rng.uniform(-1.5, 1.5, n)], axis=1)
    rays["direction"] = np.stack([rng.uniform(-0.05, 0.05, n), np.full(n, -1.0), rng.uniform(-0.05, 0.05, n)], axis=1)
    for container in ("bvh", "list"):
        sd, quad, tilted = scene(container)
        comp, osc = toh.CompiledOnHost(sd), orc.OracleScene(sd)
        comp.check_tables()
        hd, ho = comp.trace_first_hit(rays), osc.trace_first_hit(rays)
        assert np.array_equal(hd["node"], ho["node"]) and np.array_equal(hd["t"], ho["t"])
        on_quad = np.isin(ho["node"], quad)
        if container == "bvh":
            assert not on_quad.any() and (ho["node"] == tilted).sum() > n // 4  # the rays fall through the flat quad
        else:
            assert on_quad.all()


def test_mutated_descriptions_are_answered_with_a_status(rt, orc, toh):
    """rt_scene_create validates on the CPU before anything touches a GPU (include/rtb200.h): a description with
    out-of-range indices, unknown kinds, huge counts, NaN parameters or truncated tables is either rejected with a
    status or - when the damage happens to be harmless - compiled into tables that pass every structural check.
    (1 000 such descriptions were also run under -fsanitize=address,undefined: no report.)"""
    from graph_fuzz import mutated_description
    accepted = rejected = 0
    for seed in range(120):
        sd = mutated_description(rt, seed)
        try:
            comp = toh.CompiledOnHost(sd)
        except toh.TraceOnHostError as e:
            assert e.status in (rt._abi.RT_ERR_BAD_ARGUMENT, rt._abi.RT_ERR_EMPTY_SCENE, rt._abi.RT_ERR_UNSUPPORTED), (seed, str(e))
            assert str(e).split(": ", 1)[1], seed  # a message, always
            rejected += 1
            continue
        comp.check_tables()
        comp.close()
        accepted += 1
    assert rejected > 40 and accepted > 10


def test_moving_sphere_outside_its_time_range_is_not_culled(rt, orc, toh):
    """sphere.rs:144-146: center(time) is a plain linear map, so a MovingSphere with (time0, time1) = (0.3, 0.7) seen at
    shutter time 0.95 sits well beyond center1 - outside the box of sphere.rs:191-201.  In a list the reference has no
    box to cull it with; the compiler's own bounds (group bounds, SAH boxes) have to cover it.  (Found by the random
    scene graphs: 3 rays of 14 million.)"""
    A = rt._abi
    b = rt.SceneBuilder()
    m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
    rng = np.random.default_rng(5)
    kids = [b.moving_sphere((0.0, 0.0, 0.0), (4.0, 0.0, 0.0), 0.3, 0.7, 0.5, m)]
    kids += [b.sphere(c, 0.3, m) for c in rng.uniform(-6, 6, (20, 3)) + np.array([0.0, 100.0, 0.0])]  # > 8 primitives: a SAH BVH, the moving sphere alone in its leaf
    light = b.flip(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 30, b.diffuse_light(b.constant_texture((4, 4, 4)))))
    sd = b.finish(b.list(kids + [light]), b.list([light]))
    comp, osc = toh.CompiledOnHost(sd), orc.OracleScene(sd)
    comp.check_tables()
    # the bounds are built for shutter times in [0, 1]: rt_render* refuse another shutter for such a scene (api.cu:
    # check_shutter), and no scene of main.rs is one
    from util import host_scene
    assert comp.shutter_limited
    assert not toh.CompiledOnHost(host_scene(rt, "final").scene_desc).shutter_limited
    assert not toh.CompiledOnHost(host_scene(rt, "random").scene_desc).shutter_limited
    n = 3000
    rays = np.zeros(n, dtype=A.RAY_DTYPE)
    times = rng.uniform(0.0, 1.0, n)
    centre_x = 4.0 * (times - 0.3) / 0.4  # from -3 at time 0 to +7 at time 1
    rays["origin"] = np.stack([centre_x + rng.uniform(-0.3, 0.3, n), np.full(n, -0.0) + rng.uniform(-0.3, 0.3, n), np.full(n, -9.0)], axis=1)
    rays["direction"] = np.array([0.001, 0.002, 1.0])  # (not axis-parallel: 0 * inf in a slab test only ever accepts)
    rays["time"] = times
    hd, ho = comp.trace_first_hit(rays), osc.trace_first_hit(rays)
    assert (ho["node"] == kids[0]).mean() > 0.95  # the rays are aimed at where the sphere is at their time
    assert ((times < 0.3) | (times > 0.7)).mean() > 0.5
    assert np.array_equal(hd["node"], ho["node"]) and np.array_equal(hd["t"], ho["t"])


def test_inverted_cubes_show_what_six_rects_show(rt, orc, toh):
    """cube.rs:14-30 builds six AARects, and a rect with an inverted range never passes `a < a0 || a > a1` (rect.rs:55):
    a cube with min > max on one axis shows two faces, on two or three axes nothing.  The compiler turns such a cube
    into those rects (or nothing) instead of a box primitive: same object, t, normal and uv as the reference; only the
    parity hook's face number is lost (0).  Found by fuzzing degenerate parameters."""
    A = rt._abi
    seen_inverted = [0, 0, 0, 0]
    for seed in range(12):
        rng = np.random.default_rng(100 + seed)
        b = rt.SceneBuilder()
        m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
        kids = []
        for c in rng.uniform(-5, 5, (int(rng.integers(3, 14)), 3)):
            e = rng.uniform(-2, 2, 3)
            seen_inverted[int((e < 0).sum())] += 1
            kids.append(b.cube(c, c + e, m))
        light = b.flip(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 9, b.diffuse_light(b.constant_texture((4, 4, 4)))))
        sd = b.finish(b.list(kids + [light]), b.list([light]))
        comp, osc = toh.CompiledOnHost(sd), orc.OracleScene(sd)
        comp.check_tables()
        n = 20000
        rays = np.zeros(n, dtype=A.RAY_DTYPE)
        o = rng.uniform(-9, 9, (n, 3))
        rays["origin"], rays["direction"] = o, rng.uniform(-5, 5, (n, 3)) - o
        hd, ho = comp.trace_first_hit(rays), osc.trace_first_hit(rays)
        for f in ("node", "t", "normal", "front_face", "u", "v", "material"):
            assert np.array_equal(hd[f], ho[f]), (seed, f)
    assert min(seen_inverted) > 5
