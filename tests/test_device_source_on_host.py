"""CPU tier: the device SOURCE of the hot path against the oracle, without a GPU.

tests/native/trace_on_host.cpp compiles csrc/device/trace.cuh (every __device__ function the CUDA kernels call)
and csrc/device/compile.cpp (the scene compiler: groups, chains, SAH BVH, reference-order ranks) with g++ and
runs them behind per-thread loops.  These tests are the three correctness checks of test_gpu_parity.py on that
build, plus the structural invariants of the compiled tables - they pin the LOGIC of the device code and of the
host scene compiler here, where there is no GPU.  They are not parity tests of the CUDA path: those are the
`-m gpu` tests, which call the sm_100a kernels through the C ABI.  Nothing here is a CPU path of the product
(test_abi.py::test_product_does_not_reference_test_infrastructure).
"""
import os
import sys

import numpy as np
import pytest

from util import SCENES, compare_hits, host_scene, random_path_ids, rel_err, secondary_rays

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "native"))

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EXTRA_SCENES = ["light_room", "two_spheres", "two_perlin_spheres", "earth", "progress_showcase", "cornell_pbr"]
_cache = {}


@pytest.fixture(scope="module")
def toh():
    import trace_on_host as m
    return m


def scenes(rt, orc, toh, name):
    hs = host_scene(rt, name)
    if name not in _cache:
        _cache[name] = (toh.CompiledOnHost(hs.scene_desc), orc.OracleScene(hs.scene_desc))
    return (hs,) + _cache[name]


@pytest.mark.parametrize("name", SCENES + EXTRA_SCENES)
def test_compiled_tables_are_well_formed(rt, orc, toh, name):
    """What the kernels rely on without checking: every primitive in exactly one group and one leaf, leaf and
    child boxes (fp32, rounded outward) containing what is below them, the tree shallower than the traversal
    stack, ranks (the reference's traversal order) unique inside a sub-scene."""
    hs, comp, _ = scenes(rt, orc, toh, name)
    n = comp.check_tables()
    print(name, n)
    assert n["prims"] > 0 and n["world_groups"] >= 1 and n["groups"] >= n["world_groups"]
    assert n["bvh_depth"] < 64
    expect_media = {"cornell_smoke": 2, "final": 2}.get(name, 0)
    assert n["media"] == expect_media
    if name in ("cornell", "cornell_smoke", "final", "mesh", "light_room", "progress_showcase", "cornell_pbr"):
        assert n["lights"] >= 1
    if name in ("random", "final", "mesh"):
        assert n["nodes"] > 0  # the reference builds a BVH here; so does the compiler
    if name == "cornell":
        assert n["nodes"] == 0  # flat lists stay flat (18 rect tests per ray in the reference, hit.rs:58-72)


def test_mesh_at_full_detail_is_well_formed(rt, orc, toh):
    """configs[4] at the detail the bench renders (393k stand-in triangles + the teapot): parallel SAH build."""
    hs = rt.HostScene("mesh", construction_seed=1)
    comp = toh.CompiledOnHost(hs.scene_desc)
    n = comp.check_tables()
    assert n["prims"] > 300000 and n["nodes"] > 50000 and n["bvh_depth"] < 64
    comp.close()


@pytest.mark.parametrize("name", SCENES)
def test_camera_rays_on_host_match_oracle(rt, orc, toh, name):
    hs, _, _ = scenes(rt, orc, toh, name)
    W, H = 97, 61
    opts = rt.render_opts(seed=11, integrator=hs.integrator)
    px, py, s = random_path_ids(5000, W, H, 1000, seed=5)
    a = toh.camera_rays(hs.camera, W, H, opts, px, py, s)
    b = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    assert np.array_equal(a["time"], b["time"])
    assert rel_err(a["origin"], b["origin"], floor=1.0).max() < 1e-14
    assert np.abs(a["direction"] - b["direction"]).max() < 1e-12


@pytest.mark.parametrize("name", SCENES + EXTRA_SCENES)
def test_first_hit_ids_on_host(rt, orc, toh, name):
    """Check 1 on the host build: the SAH tree, fp32 boxes, composed group transforms and reference-order ranks
    of compile.cpp + the search / resolve code of trace.cuh return the object the reference's own median-split
    BVH returns, on primary and on secondary rays."""
    hs, comp, osc = scenes(rt, orc, toh, name)
    W, H = 256, 256
    opts = rt.render_opts(seed=2, integrator=hs.integrator)
    px, py, s = random_path_ids(40000, W, H, 64, seed=9)
    rays = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    hd, ho = comp.trace_first_hit(rays), osc.trace_first_hit(rays)
    r = compare_hits(hd, ho)
    print(name, "primary", r)
    assert r["id_mismatch"] == 0
    assert r["front_face_mismatch"] == 0 and r["material_mismatch"] == 0
    assert r["t_max_rel"] <= 1e-5 and r["normal_max_abs"] <= 1e-5 and r["uv_max_abs"] <= 1e-5
    rays2 = secondary_rays(ho, rays, seed=3)
    r2 = compare_hits(comp.trace_first_hit(rays2), osc.trace_first_hit(rays2))
    print(name, "secondary", r2)
    assert r2["id_mismatch"] <= max(1, r2["n"] // 20000)
    assert r2["t_max_rel"] <= 1e-5 and r2["normal_max_abs"] <= 1e-5


@pytest.mark.parametrize("name", SCENES + EXTRA_SCENES)
def test_path_radiance_on_host(rt, orc, toh, name):
    """Check 2 on the host build: path_begin / path_step (the iterative ray_color) under the oracle's Philox
    streams; a path the reference turns into NaN must be NaN here too (§Q10)."""
    hs, comp, osc = scenes(rt, orc, toh, name)
    W, H, depth = 128, 128, 100
    opts = rt.render_opts(seed=5, integrator=hs.integrator)
    px, py, s = random_path_ids(6000, W, H, 256, seed=21)
    rd, sd = comp.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    ro, so = osc.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    nan_d, nan_o = np.isnan(rd).any(axis=1), np.isnan(ro).any(axis=1)
    err = rel_err(np.nan_to_num(rd), np.nan_to_num(ro), floor=1e-9).max(axis=1)
    ok = ((err <= 1e-4) & ~nan_d & ~nan_o) | (nan_d & nan_o)
    print(name, "paths within 1e-4: %.6f, max err %.2e, mean segments host build %.3f oracle %.3f"
          % (ok.mean(), err.max(), sd.mean(), so.mean()))
    assert ok.mean() >= 0.999
    # Without DFMA contraction and with the same libm the device source follows the oracle to a few ulps (the
    # iterative form multiplies the throughput in another order than the recursion): measured max 1e-15 on 20 000
    # paths of every scene, no chaotic flips.  What the GPU adds on top is contraction and CUDA's libm only.
    assert np.array_equal(nan_d, nan_o)
    assert np.median(err) == 0.0 and np.quantile(err, 0.999) < 1e-12


@pytest.mark.parametrize("name", SCENES)
def test_golden_fixture_on_host(rt, orc, toh, name):
    """The committed oracle vectors (tests/golden/make_golden.py) against the host build of the device source."""
    hs, comp, _ = scenes(rt, orc, toh, name)
    g = np.load(os.path.join(GOLDEN, "paths_%s.npz" % name))
    W, H, depth = int(g["width"]), int(g["height"]), int(g["max_depth"])
    opts = rt.render_opts(seed=int(g["seed"]), integrator=int(g["integrator"]),
                          flags=rt._abi.FLAG_TRACE_ZERO_THROUGHPUT)
    hits = comp.trace_first_hit(g["rays"])
    assert np.array_equal(hits["node"], g["hits"]["node"]) and np.array_equal(hits["face"], g["hits"]["face"])
    rgb, seg = comp.path_radiance(hs.camera, W, H, depth, opts, g["px"], g["py"], g["sample"])
    err = rel_err(rgb, g["rgb"], floor=1e-9).max(axis=1)
    assert (err <= 1e-4).mean() >= 0.99
    assert (seg == g["segments"]).mean() >= 0.99


@pytest.mark.parametrize("name", ["cornell", "cornell_smoke", "random"])
def test_render_on_host_matches_oracle_image(rt, orc, toh, name):
    """Check 3 at a small size: the per-pixel sample loop (megakernel.inl's item loop, rows top-down)."""
    hs, comp, osc = scenes(rt, orc, toh, name)
    W, H, spp, depth = 37, 29, 8, 100
    opts = rt.render_opts(seed=8, integrator=hs.integrator)
    img, stats = comp.render(hs.camera, W, H, spp, depth, opts)
    ref, rays = osc.render(hs.camera, W, H, spp, depth, opts)
    assert stats["paths"] == W * H * spp
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    assert float((err <= 1e-4).mean()) >= 0.995
    a = orc.format_image(img, spp).astype(np.float64)
    b = orc.format_image(ref, spp).astype(np.float64)
    assert float(np.sqrt(np.mean((a - b) ** 2)) / 255.0) <= 0.01
    # a sample sub-range is the multi-GPU partition (api.cu: sample_block): the blocks add up to the render
    lo, _ = comp.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=8, integrator=hs.integrator, sample_begin=0, sample_count=3))
    hi, _ = comp.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=8, integrator=hs.integrator, sample_begin=3, sample_count=5))
    assert rel_err(lo + hi, img, floor=1e-9).max() < 1e-12


def _big_triangle_bvh(rt, n, spoil=None):
    """n random triangles under one BVH::new (a mesh, main.rs:442) + a light; spoil(b, nodes) may break one."""
    A = rt._abi
    rng = np.random.default_rng(11)
    b = rt.SceneBuilder()
    m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
    c = rng.uniform(-50, 50, (n, 3))
    e = rng.uniform(-1, 1, (n, 2, 3))
    tris = [b.triangle(c[i], c[i] + e[i, 0], c[i] + e[i, 1], m) for i in range(n)]
    if spoil:
        spoil(b, tris)
    light = b.flip(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 60, b.diffuse_light(b.constant_texture((4, 4, 4)))))
    return b.finish(b.list([b.bvh(tris), light]), b.list([light]))


def test_parallel_compile_equals_serial_compile(rt, orc, toh, monkeypatch):
    """Large primitive BVHs are emitted, boxed, permuted and encoded by several host threads (compile.cpp:
    parallel_for, emit_many): the tables are the ones the serial walk produces, byte for byte, and two compiles agree."""
    sd = _big_triangle_bvh(rt, 70000)
    a = toh.CompiledOnHost(sd)
    n = a.check_tables()
    assert n["prims"] == 70001 and n["nodes"] > 8000
    b = toh.CompiledOnHost(sd)
    monkeypatch.setenv("RTB200_COMPILE_SERIAL", "1")
    c = toh.CompiledOnHost(sd)
    monkeypatch.delenv("RTB200_COMPILE_SERIAL")
    assert a.tables_hash() == b.tables_hash() == c.tables_hash()


def test_parallel_compile_reports_the_errors_of_the_serial_walk(rt, orc, toh):
    A = rt._abi

    def bad_material(b, tris):
        b.nodes[tris[40000]].material = 99

    def nan_vertex(b, tris):
        b.nodes[tris[12345]].v[4] = float("nan")

    for spoil, word in ((bad_material, "material"), (nan_vertex, "NaN")):
        with pytest.raises(toh.TraceOnHostError) as ei:
            toh.CompiledOnHost(_big_triangle_bvh(rt, 70000, spoil))
        assert ei.value.status == A.RT_ERR_BAD_ARGUMENT and word in str(ei.value)


@pytest.mark.parametrize("name", SCENES + EXTRA_SCENES)
def test_search_resolve_split_is_the_fused_world_hit(rt, orc, toh, name):
    """wavefront.inl: the search / resolve split of the wavefront stages (extend: world_search; shade: resolve_hit /
    resolve_medium + path_shade) gives the radiance and segment count of the megakernel's fused world_hit, bit for bit."""
    hs, comp, _ = scenes(rt, orc, toh, name)
    W, H, depth = 96, 96, 100
    no_lights = name in ("random", "two_spheres", "two_perlin_spheres", "earth")  # HEAD needs a light list (§Q7)
    for integrator in ((rt.INTEGRATOR_LEGACY,) if no_lights else (rt.INTEGRATOR_HEAD, rt.INTEGRATOR_LEGACY)):
        opts = rt.render_opts(seed=17, integrator=integrator)
        px, py, s = random_path_ids(4000, W, H, 128, seed=33)
        a, sa = comp.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
        b, sb = comp.path_radiance_split(hs.camera, W, H, depth, opts, px, py, s)
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(sa, sb), (name, integrator)

