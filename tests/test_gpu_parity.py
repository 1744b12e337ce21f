"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs (BASELINE.json's three correctness checks)."""
import os

import numpy as np
import pytest

from util import SCENES, compare_hits, host_scene, random_path_ids, rel_err, secondary_rays

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_dev_cache = {}


def scenes(rt, orc, name):
    hs = host_scene(rt, name)
    if name not in _dev_cache:
        _dev_cache[name] = (rt.DeviceScene(hs.scene_desc, device=0), orc.OracleScene(hs.scene_desc))
    return (hs,) + _dev_cache[name]


@pytest.mark.parametrize("name", SCENES)
def test_camera_rays_match_oracle(rt, orc, name):
    hs, dev, _ = scenes(rt, orc, name)
    W, H = 97, 61  # ragged on purpose
    opts = rt.render_opts(seed=11, integrator=hs.integrator)
    px, py, s = random_path_ids(20000, W, H, 1000, seed=5)
    a = dev.camera_rays(hs.camera, W, H, opts, px, py, s)
    b = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    assert np.array_equal(a["time"], b["time"])
    assert rel_err(a["origin"], b["origin"], floor=1.0).max() < 1e-14
    assert np.abs(a["direction"] - b["direction"]).max() < 1e-12


@pytest.mark.parametrize("name", SCENES)
def test_first_hit_bit_exact_ids(rt, orc, name):
    """Check 1: per-pixel primary-ray first hits bit-exact on object id against the reference's
    own BVH (restated in the oracle); t and normal within 1e-5 relative."""
    hs, dev, osc = scenes(rt, orc, name)
    W, H = 256, 256
    opts = rt.render_opts(seed=2, integrator=hs.integrator)
    px, py, s = random_path_ids(200000, W, H, 64, seed=9)
    rays = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    hd, ho = dev.trace_first_hit(rays), osc.trace_first_hit(rays)
    r = compare_hits(hd, ho)
    print(name, "primary", r)
    assert r["id_mismatch"] == 0
    assert r["front_face_mismatch"] == 0 and r["material_mismatch"] == 0
    assert r["t_max_rel"] <= 1e-5 and r["normal_max_abs"] <= 1e-5 and r["uv_max_abs"] <= 1e-5
    # secondary rays leaving the surfaces in random directions (self-intersection regime, t_min = 1e-5)
    rays2 = secondary_rays(ho, rays, seed=3)
    hd2, ho2 = dev.trace_first_hit(rays2), osc.trace_first_hit(rays2)
    r2 = compare_hits(hd2, ho2)
    print(name, "secondary", r2)
    # a ray that starts on a surface re-hits it at t ~ 1e-13/|d_k|; whether that lands above
    # t_min = 1e-5 flips with the last ulp, so allow a vanishing fraction here
    assert r2["id_mismatch"] <= max(1, r2["n"] // 20000)
    assert r2["t_max_rel"] <= 1e-5 and r2["normal_max_abs"] <= 1e-5


@pytest.mark.parametrize("name", SCENES)
def test_path_radiance_matches_oracle(rt, orc, name):
    """Check 2: per-path radiance under identical Philox sequences within 1e-4 relative."""
    hs, dev, osc = scenes(rt, orc, name)
    W, H, depth = 128, 128, 100
    opts = rt.render_opts(seed=5, integrator=hs.integrator)
    px, py, s = random_path_ids(30000, W, H, 256, seed=21)
    rd, sd = dev.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    ro, so = osc.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    err = rel_err(rd, ro, floor=1e-9).max(axis=1)
    both_nan = np.isnan(rd).any(axis=1) & np.isnan(ro).any(axis=1)
    ok = (err <= 1e-4) | both_nan
    print(name, "paths within 1e-4: %.6f, median err %.2e, mean segments gpu %.3f oracle %.3f"
          % (ok.mean(), np.nanmedian(err), sd.mean(), so.mean()))
    # f64 on both sides; the only disagreements are chaotic flips at the last ulp
    assert ok.mean() >= 0.9995
    assert np.nanmedian(err) < 1e-10


@pytest.mark.parametrize("name", SCENES)
def test_golden_fixture(rt, orc, name):
    """The committed oracle vectors (tests/golden/make_golden.py): device first hits and radiance."""
    hs, dev, _ = scenes(rt, orc, name)
    g = np.load(os.path.join(GOLDEN, "paths_%s.npz" % name))
    W, H, depth = int(g["width"]), int(g["height"]), int(g["max_depth"])
    # the reference keeps tracing zero-throughput paths (§Q11): ask the device to do the same so
    # that the segment counts are comparable
    opts = rt.render_opts(seed=int(g["seed"]), integrator=int(g["integrator"]),
                          flags=rt._abi.FLAG_TRACE_ZERO_THROUGHPUT)
    hits = dev.trace_first_hit(g["rays"])
    assert np.array_equal(hits["node"], g["hits"]["node"]) and np.array_equal(hits["face"], g["hits"]["face"])
    rgb, seg = dev.path_radiance(hs.camera, W, H, depth, opts, g["px"], g["py"], g["sample"])
    err = rel_err(rgb, g["rgb"], floor=1e-9).max(axis=1)
    assert (err <= 1e-4).mean() >= 0.99
    assert (seg == g["segments"]).mean() >= 0.99


@pytest.mark.parametrize("name", SCENES)
def test_render_matches_oracle_image(rt, orc, name):
    """Check 3 at a size the oracle finishes in seconds: the accumulated image."""
    hs, dev, osc = scenes(rt, orc, name)
    W, H, spp, depth = 61, 45, 24, 100
    opts = rt.render_opts(seed=8, integrator=hs.integrator)
    img, stats = dev.render(hs.camera, W, H, spp, depth, opts)
    ref, rays = osc.render(hs.camera, W, H, spp, depth, opts)
    assert stats.paths == W * H * spp
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    frac = float((err <= 1e-4).mean())
    a = rt.format_image(img, spp).astype(np.float64)
    b = orc.format_image(ref, spp).astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)) / 255.0)
    print(name, "pixels within 1e-4: %.5f  8-bit RMSE %.5f  rays gpu %d oracle %d" % (frac, rmse, stats.rays, rays))
    assert frac >= 0.995
    assert rmse <= 0.01


def test_render_is_deterministic_and_additive(rt, orc):
    hs, dev, _ = scenes(rt, orc, "cornell")
    W, H, spp, depth = 120, 80, 32, 100
    a, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=4))
    b, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=4))
    assert np.array_equal(a, b), "two runs must be bit-identical"
    c, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=5))
    assert not np.array_equal(a, c)
    # disjoint sample ranges add up to the full render (the multi-GPU partition)
    parts = [dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=4, sample_begin=s0, sample_count=n))[0]
             for s0, n in ((0, 8), (8, 8), (16, 16))]
    tot = parts[0].astype(np.float64) + parts[1] + parts[2]
    assert rel_err(tot, a, floor=1e-6).max() < 1e-5


def test_error_codes(rt, orc):
    hs = host_scene(rt, "random")  # empty light list
    dev = rt.DeviceScene(hs.scene_desc)
    with pytest.raises(rt.RtError) as e:
        dev.render(hs.camera, 16, 16, 1, 5, rt.render_opts(integrator=rt.INTEGRATOR_HEAD))
    assert e.value.status == rt._abi.RT_ERR_NO_LIGHTS
    with pytest.raises(rt.RtError) as e:
        dev.render(hs.camera, 1, 16, 1, 5, rt.render_opts(integrator=rt.INTEGRATOR_LEGACY))
    assert e.value.status == rt._abi.RT_ERR_BAD_ARGUMENT
    # depth 0 renders black (main.rs:42-45)
    img, stats = dev.render(hs.camera, 16, 16, 2, 0, rt.render_opts(integrator=rt.INTEGRATOR_LEGACY))
    assert not img.any() and stats.rays == 0


def test_cornell_full_resolution_low_spp(rt, orc):
    """BASELINE config 2 at its full 600x600 with a sample budget the oracle finishes in seconds."""
    hs, dev, osc = scenes(rt, orc, "cornell")
    W, H, spp, depth = hs.width, hs.height, 8, hs.max_depth
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    img, stats = dev.render(hs.camera, W, H, spp, depth, opts)
    ref, rays = osc.render(hs.camera, W, H, spp, depth, opts)
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    assert (err <= 1e-4).mean() >= 0.999
    assert abs(int(stats.rays) - rays) <= rays * 0.2  # the device stops zero-throughput paths early


def test_end_to_end_render_entry(rt, orc):
    """render(world, camera, width, height, spp, max_depth) through the host layer equals the
    two-step scene_create + render path."""
    hs, dev, _ = scenes(rt, orc, "cornell_smoke")
    opts = rt.render_opts(seed=6, integrator=hs.integrator)
    a, sa = hs.render(48, 48, 16, 50, opts)
    b, _ = dev.render(hs.camera, 48, 48, 16, 50, opts)
    assert np.array_equal(a, b)
    assert sa.paths == 48 * 48 * 16


@pytest.mark.parametrize("name,spp", [("cornell", 1000), ("cornell_smoke", 1000), ("random", 800), ("final", 64)])
def test_converged_image_full_config(rt, orc, name, spp):
    """Check 3 at BASELINE's own sizes: config 1-3 at their full resolution AND spp, config 4 at its
    full 800x800 with the spp the CPU oracle can finish in under a minute.  Same Philox streams on
    both sides, so the images are comparable pixel by pixel, not just statistically."""
    hs, dev, osc = scenes(rt, orc, name)
    W, H, depth = hs.width, hs.height, hs.max_depth
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    img, stats = dev.render(hs.camera, W, H, spp, depth, opts)
    ref, rays = osc.render(hs.camera, W, H, spp, depth, opts)
    assert stats.paths == W * H * spp
    a = rt.format_image(img, spp).astype(np.float64)
    b = orc.format_image(ref, spp).astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)) / 255.0)
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    differing = int((a != b).any(axis=2).sum())
    print("%s %dx%dx%d: 8-bit RMSE %.6f, pixels with any differing byte %d of %d, pixels within 1e-4 rel %.6f, nonfinite gpu %d"
          % (name, W, H, spp, rmse, differing, W * H, (err <= 1e-4).mean(), stats.nonfinite_samples))
    assert rmse <= 0.01
    assert (err <= 1e-4).mean() >= 0.999


def test_mesh_config_reduced(rt, orc):
    """Config 5 (3840x2160x1024) is out of the oracle's reach; same scene and aspect at 480x270, 32 spp."""
    hs, dev, osc = scenes(rt, orc, "mesh")
    W, H, spp, depth = 480, 270, 32, hs.max_depth
    opts = rt.render_opts(seed=1, integrator=hs.integrator)
    img, stats = dev.render(hs.camera, W, H, spp, depth, opts)
    ref, _ = osc.render(hs.camera, W, H, spp, depth, opts)
    a = rt.format_image(img, spp).astype(np.float64)
    b = orc.format_image(ref, spp).astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)) / 255.0)
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    print("mesh %dx%dx%d: 8-bit RMSE %.6f, pixels within 1e-4 rel %.6f" % (W, H, spp, rmse, (err <= 1e-4).mean()))
    assert rmse <= 0.01 and (err <= 1e-4).mean() >= 0.999


@pytest.mark.parametrize("name", SCENES)
def test_wavefront_equals_megakernel(rt, orc, name):
    """The two pipelines run the same device arithmetic with the same summation order: the
    wavefront stages (generate / extend / shade over ray queues) must reproduce the megakernel's
    image bit for bit, and count the same paths and segments."""
    hs, dev, _ = scenes(rt, orc, name)
    W, H, spp, depth = 97, 61, 20, 100  # ragged: partial 8x4 tiles, a pool larger than the image
    abi = rt._abi
    a, sa = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=3, integrator=hs.integrator, flags=abi.FLAG_MEGAKERNEL))
    b, sb = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=3, integrator=hs.integrator, flags=abi.FLAG_WAVEFRONT))
    assert sa.paths == sb.paths == W * H * spp
    assert sa.rays == sb.rays
    assert sa.nonfinite_samples == sb.nonfinite_samples
    assert np.array_equal(a, b, equal_nan=True)
    # a sample sub-range (the multi-GPU partition) and a depth cut
    o = dict(seed=3, integrator=hs.integrator, sample_begin=5, sample_count=9)
    a, _ = dev.render(hs.camera, W, H, spp, 7, rt.render_opts(flags=abi.FLAG_MEGAKERNEL, **o))
    b, _ = dev.render(hs.camera, W, H, spp, 7, rt.render_opts(flags=abi.FLAG_WAVEFRONT, **o))
    assert np.array_equal(a, b, equal_nan=True)


def test_wavefront_matches_oracle_image(rt, orc):
    """Check 3 through the wavefront pipeline on the scene with media, BVHs, textures and motion blur."""
    hs, dev, osc = scenes(rt, orc, "final")
    W, H, spp, depth = 61, 45, 24, 100
    opts = rt.render_opts(seed=8, integrator=hs.integrator, flags=rt._abi.FLAG_WAVEFRONT)
    img, stats = dev.render(hs.camera, W, H, spp, depth, opts)
    ref, _ = osc.render(hs.camera, W, H, spp, depth, opts)
    err = rel_err(img, ref, floor=1e-6).max(axis=2)
    assert float((err <= 1e-4).mean()) >= 0.995
    assert stats.kernel_launches > 4


@pytest.mark.parametrize("name", SCENES)
def test_feature_variants_identical(rt, orc, name, monkeypatch):
    """The pipelines are compiled once per scene feature set (variants.h); the build picked for a
    scene must give the image of the build that has every feature compiled in, bit for bit."""
    hs, dev, _ = scenes(rt, orc, name)
    W, H, spp, depth = 80, 60, 12, 100
    for flag in (rt._abi.FLAG_MEGAKERNEL, rt._abi.FLAG_WAVEFRONT):
        opts = rt.render_opts(seed=9, integrator=hs.integrator, flags=flag)
        monkeypatch.delenv("RTB200_VARIANT", raising=False)
        a, sa = dev.render(hs.camera, W, H, spp, depth, opts)
        monkeypatch.setenv("RTB200_VARIANT", "vall")
        b, sb = dev.render(hs.camera, W, H, spp, depth, opts)
        monkeypatch.delenv("RTB200_VARIANT", raising=False)
        assert np.array_equal(a, b, equal_nan=True)
        assert sa.rays == sb.rays


EXTRA_SCENES = ["light_room", "two_spheres", "two_perlin_spheres", "earth", "progress_showcase", "cornell_pbr"]


@pytest.mark.parametrize("name", EXTRA_SCENES)
def test_extra_scenes_match_oracle(rt, orc, name):
    """the reference's remaining scene constructors (main.rs:212-276, :515-562; sphere
    lights, Perlin and image textures outside the final scene) and the PBR material + PDF::BRDF
    (mat.rs:86-197, pdf.rs:20-60,97-130,151-160) - the same three checks as the five configs."""
    hs, dev, osc = scenes(rt, orc, name)
    W, H, depth = 96, 96, 100
    opts = rt.render_opts(seed=4, integrator=hs.integrator)
    # 1. first hits
    px, py, s = random_path_ids(40000, W, H, 64, seed=12)
    rays = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    r = compare_hits(dev.trace_first_hit(rays), osc.trace_first_hit(rays))
    assert r["id_mismatch"] == 0 and r["t_max_rel"] <= 1e-5 and r["normal_max_abs"] <= 1e-5 and r["uv_max_abs"] <= 1e-5
    # 2. per-path radiance (a path the reference turns into NaN must be NaN here too, §Q10)
    px, py, s = random_path_ids(20000, W, H, 256, seed=13)
    rd, sd = dev.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    ro, so = osc.path_radiance(hs.camera, W, H, depth, opts, px, py, s)
    nan_d, nan_o = np.isnan(rd).any(axis=1), np.isnan(ro).any(axis=1)
    err = rel_err(np.nan_to_num(rd), np.nan_to_num(ro), floor=1e-9).max(axis=1)
    ok = ((err <= 1e-4) & ~nan_d & ~nan_o) | (nan_d & nan_o)
    print(name, "paths ok %.6f  nan gpu %.4f oracle %.4f  median err %.2e" % (ok.mean(), nan_d.mean(), nan_o.mean(), np.median(err)))
    assert ok.mean() >= 0.999
    # 3. the image, both pipelines
    for flag in (rt._abi.FLAG_MEGAKERNEL, rt._abi.FLAG_WAVEFRONT):
        o = rt.render_opts(seed=4, integrator=hs.integrator, flags=flag)
        img, stats = dev.render(hs.camera, 48, 48, 16, depth, o)
        ref, _ = osc.render(hs.camera, 48, 48, 16, depth, o)
        nd, no = np.isnan(img).any(axis=2), np.isnan(ref).any(axis=2)
        e = rel_err(np.nan_to_num(img), np.nan_to_num(ref), floor=1e-6).max(axis=2)
        good = ((e <= 1e-4) & ~nd & ~no) | (nd & no)
        assert good.mean() >= 0.99, (name, flag, good.mean())
        a = rt.format_image(img, 16).astype(np.float64)
        b = orc.format_image(ref, 16).astype(np.float64)
        assert float(np.sqrt(np.mean((a - b) ** 2)) / 255.0) <= 0.01


def test_render_info_reports_pipeline_and_variant(rt, orc):
    """rt_render_info: which pipeline build a render ran (the choice is per scene)."""
    abi = rt._abi
    expect = {"cornell": ("megakernel", "vflat"), "final": ("wavefront", "vnextweek"), "mesh": ("megakernel", "vmesh"),
              "random": ("megakernel", "vspheres"), "cornell_pbr": ("megakernel", "vall")}
    for name, (pipeline, variant) in expect.items():
        hs, dev, _ = scenes(rt, orc, name)
        dev.render(hs.camera, 32, 32, 2, 10, rt.render_opts(seed=1, integrator=hs.integrator))
        info = dev.render_info
        assert (info["pipeline"], info["variant"]) == (pipeline, variant), (name, info)
    # a tree over spheres without media takes the wavefront pipeline once the render is long (api.cu kWavefrontLongPaths)
    hs, dev, _ = scenes(rt, orc, "random")
    a, sa = dev.render(hs.camera, 1000, 1000, 80, 50, rt.render_opts(seed=1, integrator=hs.integrator))
    assert dev.render_info["pipeline"] == "wavefront" and dev.render_info["variant"] == "vspheres"
    b, sb = dev.render(hs.camera, 1000, 1000, 80, 50, rt.render_opts(seed=1, integrator=hs.integrator, flags=abi.FLAG_MEGAKERNEL))
    assert dev.render_info["pipeline"] == "megakernel" and sa.paths == sb.paths
    assert (a != b).any(axis=2).sum() <= 2  # (the same image; DESIGN.md §5.2 on the last fp32 bit of a pixel in a million)
    hs, dev, _ = scenes(rt, orc, "cornell")
    dev.render(hs.camera, 32, 32, 2, 10, rt.render_opts(seed=1, flags=abi.FLAG_WAVEFRONT))
    assert dev.render_info["pipeline"] == "wavefront" and int(dev.render_info["pool_slots"]) == 32 * 32 * 2
    with pytest.raises(rt.RtError):
        dev.render(hs.camera, 32, 32, 2, 10, rt.render_opts(seed=1, flags=abi.FLAG_WAVEFRONT | abi.FLAG_MEGAKERNEL))


@pytest.mark.parametrize("seed", range(8))
def test_random_scene_graphs_on_device(rt, orc, seed):
    """The seeded random graphs of test_scene_graph_fuzz.py (any wrapper / container nesting a flatten() visitor can
    emit, media bounded by instanced boxes, both integrators) through the CUDA path: same object as the oracle's
    literal object tree on identical rays, same radiance per path, and both pipelines rendering the same image."""
    from graph_fuzz import GraphMaker, fuzz_camera
    # seeds 4-7: Perlin / image / nested textures, PBR, light lists with objects the reference cannot sample
    g = GraphMaker(rt, 1000 + seed, rich=seed >= 4)
    sd = g.make()
    dev, osc = rt.DeviceScene(sd, device=0), orc.OracleScene(sd)
    rays = g.rays(30000)
    r = compare_hits(dev.trace_first_hit(rays), osc.trace_first_hit(rays))
    print(seed, r)
    # random geometry has no exact ties; a candidate within an ulp of t_min or of another one may still flip under
    # DFMA contraction (DESIGN.md "Precision policy")
    assert r["id_mismatch"] <= 1 and r["front_face_mismatch"] == 0 and r["material_mismatch"] == 0
    assert r["t_max_rel"] <= 1e-5 and r["normal_max_abs"] <= 1e-5 and r["uv_max_abs"] <= 1e-5
    cam = fuzz_camera(rt)
    W = H = 64
    ids = np.random.default_rng(seed)
    px, py, s = (ids.integers(0, W, 6000, dtype=np.uint32), ids.integers(0, H, 6000, dtype=np.uint32),
                 ids.integers(0, 64, 6000, dtype=np.uint32))
    for integrator in (rt.INTEGRATOR_HEAD, rt.INTEGRATOR_LEGACY):
        opts = rt.render_opts(seed=seed + 1, integrator=integrator)
        rd, _ = dev.path_radiance(cam, W, H, 50, opts, px, py, s)
        ro, _ = osc.path_radiance(cam, W, H, 50, opts, px, py, s)
        nan_d, nan_o = np.isnan(rd).any(axis=1), np.isnan(ro).any(axis=1)
        err = rel_err(np.nan_to_num(rd), np.nan_to_num(ro), floor=1e-9).max(axis=1)
        ok = ((err <= 1e-4) & ~nan_d & ~nan_o) | (nan_d & nan_o)
        print(seed, "integrator", integrator, "ok %.5f median err %.2e" % (ok.mean(), np.median(err)))
        # (1, 0, 0) rays of an unsampleable light: see test_scene_graph_fuzz.py and DESIGN.md "Known deviations"
        assert ok.mean() >= (0.99 if g.has_default_light and integrator == rt.INTEGRATOR_HEAD else 0.998)
    a, sa = dev.render(cam, 40, 30, 6, 50, rt.render_opts(seed=3, flags=rt._abi.FLAG_MEGAKERNEL))
    b, sb = dev.render(cam, 40, 30, 6, 50, rt.render_opts(seed=3, flags=rt._abi.FLAG_WAVEFRONT))
    assert np.array_equal(a, b, equal_nan=True) and sa.rays == sb.rays
    dev.close()
    osc.close()


def test_shutter_outside_unit_range_is_refused_for_extrapolating_spheres(rt, orc):
    """A MovingSphere with (time0, time1) != (0, 1) extrapolates (sphere.rs:144-146); the compiled bounds cover shutter
    times in [0, 1], so another shutter is refused (RT_ERR_UNSUPPORTED) instead of culling the sphere wrongly."""
    A = rt._abi
    b = rt.SceneBuilder()
    m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
    ball = b.moving_sphere((0.0, 0.0, 0.0), (4.0, 0.0, 0.0), 0.3, 0.7, 0.5, m)
    light = b.flip(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 30, b.diffuse_light(b.constant_texture((4, 4, 4)))))
    dev = rt.DeviceScene(b.finish(b.list([ball, light]), b.list([light])), device=0)
    ok_cam = rt.camera_new((0, 0, -9), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 9.0, 0.0, 1.0)
    bad_cam = rt.camera_new((0, 0, -9), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 9.0, 0.0, 2.0)
    img, _ = dev.render(ok_cam, 16, 16, 2, 10, rt.render_opts(seed=1))
    assert np.isfinite(img).all()
    with pytest.raises(rt.RtError) as ei:
        dev.render(bad_cam, 16, 16, 2, 10, rt.render_opts(seed=1))
    assert ei.value.status == A.RT_ERR_UNSUPPORTED and "shutter" in str(ei.value)
    dev.close()



@pytest.mark.parametrize("name,oracle_spp", [("final", 8), ("mesh", 1)])
def test_full_size_properties_of_the_two_largest_configs(rt, orc, name, oracle_spp):
    """Configs 4 and 5 at BASELINE's FULL size (800x800x10000 and 3840x2160x1024) are out of the oracle's reach
    (minutes to hours of CPU), so the full-size renders are checked through size-independent properties:
    (1) additivity - four disjoint sample blocks, the multi-GPU partition, add up to the one-call render;
    (2) the oracle's first `oracle_spp` samples ARE the first samples of the full render (same Philox keys), so a
        render restricted to that sample range must equal the oracle's image pixel by pixel at full resolution;
    (3) unbiasedness at full size - the mean radiance of the full render equals the mean radiance of that short
        oracle render within its Monte-Carlo error (a wrong weight anywhere in the path shows as a shifted mean)."""
    hs = rt.HostScene(name, construction_seed=1)  # the bench's scene: the mesh at full detail (393k + 1k triangles)
    dev, osc = rt.DeviceScene(hs.scene_desc, device=0), orc.OracleScene(hs.scene_desc)
    W, H, spp, depth = hs.width, hs.height, hs.spp, hs.max_depth
    full, st = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=1, integrator=hs.integrator))
    assert st.paths == W * H * spp
    # (1)
    acc = np.zeros((H, W, 3), dtype=np.float64)
    q = spp // 4
    for k in range(4):
        n = q if k < 3 else spp - 3 * q
        part, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=1, integrator=hs.integrator, sample_begin=k * q, sample_count=n))
        acc += part
    finite = np.isfinite(full) & np.isfinite(acc)
    assert finite.mean() > 0.9999
    assert rel_err(acc[finite], full[finite], floor=1e-3).max() < 1e-5
    # (2)
    head, _ = dev.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=1, integrator=hs.integrator, sample_begin=0, sample_count=oracle_spp))
    ref, _ = osc.render(hs.camera, W, H, spp, depth, rt.render_opts(seed=1, integrator=hs.integrator, sample_begin=0, sample_count=oracle_spp))
    err = rel_err(head, ref, floor=1e-6).max(axis=2)
    assert (err <= 1e-4).mean() >= 0.999
    # (3) per-pixel means, clipped at 20x the light's radiance so that single fireflies do not set the variance
    a = np.clip(np.nan_to_num(full.astype(np.float64) / spp), 0, 300).mean(axis=2)
    b = np.clip(np.nan_to_num(ref / oracle_spp), 0, 300).mean(axis=2)
    diff = a - b
    sigma = diff.std() / np.sqrt(diff.size)
    print("%s full size: mean radiance gpu %.6f oracle(%d spp) %.6f, difference %.2e = %.2f sigma" % (name, a.mean(), oracle_spp, b.mean(), diff.mean(), diff.mean() / sigma))
    assert abs(diff.mean()) <= 5.0 * sigma + 1e-4 * b.mean()
    dev.close()
    osc.close()


@pytest.mark.parametrize("detail", [1, 0])
def test_gpu_built_bvh_renders_the_same_image(rt, orc, detail):
    """BVH::new on the GPU (rt_scene_create_ex + RT_CREATE_GPU_BVH: Morton codes, radix sort,
    Karras hierarchy, bottom-up refit).  Boxes only cull and every primitive stays reachable, so a render on the
    linear tree must find the same winner for every ray as one on the host's SAH tree: first hits equal the oracle's
    ids, and the image is the host-tree image bit for bit.  detail 0 is the bench's mesh (393k + 1k triangles)."""
    hs = rt.HostScene("mesh", construction_seed=1, mesh_detail=detail)
    host_tree = rt.DeviceScene(hs.scene_desc, device=0)
    gpu_tree = rt.DeviceScene(hs.scene_desc, device=0, gpu_bvh=True)
    W, H, spp, depth = 160, 90, 4, 50
    opts = rt.render_opts(seed=2, integrator=hs.integrator)
    a, sa = host_tree.render(hs.camera, W, H, spp, depth, opts)
    b, sb = gpu_tree.render(hs.camera, W, H, spp, depth, opts)
    assert gpu_tree.render_info.get("gpu_built_trees") == "1" and "gpu_built_trees" not in host_tree.render_info
    assert np.array_equal(a, b, equal_nan=True)
    assert (sa.paths, sa.rays) == (sb.paths, sb.rays)
    px, py, s = random_path_ids(50000, W, H, 64, seed=4)
    rays = orc.camera_rays(hs.camera, W, H, opts, px, py, s)
    hg, hh = gpu_tree.trace_first_hit(rays), host_tree.trace_first_hit(rays)
    assert np.array_equal(hg["node"], hh["node"]) and np.array_equal(hg["t"], hh["t"])
    if detail == 1:
        osc = orc.OracleScene(hs.scene_desc)
        r = compare_hits(hg, osc.trace_first_hit(rays))
        assert r["id_mismatch"] == 0 and r["t_max_rel"] <= 1e-5
        osc.close()
    host_tree.close()
    gpu_tree.close()


