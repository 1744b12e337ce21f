"""The oracle against every known answer available for this path (CPU only).

The reference holds exactly one table of known answers for this path: the six
get_sphere_uv values in the comment at src/sphere.rs:12-17.  Everything else below is
analytic (closed-form ray/primitive geometry, invariants) or third-party (Philox4x32-10
from Random123 / cuRAND).
"""
import json
import math
import os

import numpy as np
import pytest

from util import host_scene, rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- the reference's own known answers ---------------------------------------------------
@pytest.mark.parametrize("p,uv", [((1, 0, 0), (0.50, 0.50)), ((-1, 0, 0), (0.00, 0.50)), ((0, 1, 0), (0.50, 1.00)),
                                  ((0, -1, 0), (0.50, 0.00)), ((0, 0, 1), (0.25, 0.50)), ((0, 0, -1), (0.75, 0.50))])
def test_get_sphere_uv_table(orc, p, uv):
    """src/sphere.rs:12-17"""
    u, v = orc.sphere_uv(p)
    assert abs(u - uv[0]) < 1e-15 and abs(v - uv[1]) < 1e-15


# ---- Philox4x32-10 -------------------------------------------------------------------------
def test_philox_random123_kat(orc):
    """Random123 kat_vectors for philox4x32_10."""
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                             [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_matches_curand_golden(orc):
    """tests/golden/philox_curand.json was produced by NVIDIA's curand_Philox4x32_10 (gen_philox_golden.cu)."""
    vecs = json.load(open(os.path.join(GOLDEN, "philox_curand.json")))
    assert len(vecs) == 67
    for e in vecs:
        assert orc.philox4x32_10(e["ctr"], e["key"]) == e["out"]


def test_draw_layout(orc):
    """A draw is (u53(w0,w1), u53(w2,w3), w1 & 0x7ff, w3 & 0x7ff) of philox(ctr=(bounce,slot,sub,seed), key=(pixel,sample))."""
    seed, pixel, sample, bounce, slot, sub = 9, 1234, 77, 3, 4, 0
    w = orc.philox4x32_10([bounce, slot, sub, seed], [pixel, sample])
    a, b, ba, bb = orc.draw(seed, pixel, sample, bounce, slot, sub)
    assert a == (((w[0] << 32) | w[1]) >> 11) * 2.0 ** -53
    assert b == (((w[2] << 32) | w[3]) >> 11) * 2.0 ** -53
    assert ba == w[1] & 0x7FF and bb == w[3] & 0x7FF
    assert 0.0 <= a < 1.0 and 0.0 <= b < 1.0


def test_draw_uniformity(orc):
    xs = np.array([orc.draw(1, p, 0, 0, 0, 0)[0] for p in range(4000)])
    assert abs(xs.mean() - 0.5) < 0.02 and abs(xs.var() - 1 / 12) < 0.01
    coins = np.array([orc.draw(1, p, 0, 0, 4, 0)[2] & 1 for p in range(4000)])
    assert abs(coins.mean() - 0.5) < 0.03


# ---- vec.rs / onb.rs / mat.rs helpers ------------------------------------------------------
def test_onb_is_orthonormal(orc):
    rng = np.random.default_rng(0)
    for n in list(rng.normal(size=(50, 3))) + [np.array([1.0, 0, 0]), np.array([0.95, 0.1, 0]), np.array([0, 0, -3.0])]:
        m = orc.onb(n)
        assert np.allclose(m @ m.T, np.eye(3), atol=1e-14)
        assert np.allclose(m[2], n / np.linalg.norm(n), atol=1e-15)  # w = normalize(n), onb.rs:9
        assert np.allclose(np.cross(m[2], m[1]), m[0], atol=1e-15)  # u = w x v, onb.rs:16


def test_onb_helper_axis(orc):
    """onb.rs:10-14: helper axis is Y when |w.x| > 0.9, else X."""
    m = orc.onb([1.0, 0.0, 0.0])
    assert np.allclose(m[1], np.cross([1, 0, 0], [0, 1, 0]))
    m = orc.onb([0.0, 1.0, 0.0])
    assert np.allclose(m[1], np.cross([0, 1, 0], [1, 0, 0]))


def test_reflect_refract(orc):
    v = np.array([1.0, -1.0, 0.0]) / math.sqrt(2)
    n = np.array([0.0, 1.0, 0.0])
    assert np.allclose(orc.reflect(v, n), [v[0], -v[1], 0.0], atol=1e-16)
    # Snell: sin(t) = eta sin(i)
    eta = 1.0 / 1.5
    r = orc.refract(v, n, eta)
    assert abs(np.linalg.norm(r) - 1.0) < 1e-15
    assert abs(r[0] - eta * v[0]) < 1e-15 and r[1] < 0
    # eta = 1 passes straight through
    assert np.allclose(orc.refract(v, n, 1.0), v, atol=1e-15)


def test_schlick(orc):
    r0 = ((1 - 1.5) / (1 + 1.5)) ** 2
    assert abs(orc.reflectance(1.0, 1.5) - r0) < 1e-16
    assert abs(orc.reflectance(0.0, 1.5) - 1.0) < 1e-15
    c = 0.3
    assert abs(orc.reflectance(c, 1.5) - (r0 + (1 - r0) * (1 - c) ** 5)) < 1e-15


def test_random_cosine_direction(orc):
    """pdf.rs:8-18: unit length, z = sqrt(1-r2)."""
    rng = np.random.default_rng(1)
    for r1, r2 in rng.random((100, 2)):
        d = orc.random_cosine_direction(r1, r2)
        assert abs(np.linalg.norm(d) - 1.0) < 1e-14
        assert abs(d[2] - math.sqrt(1 - r2)) < 1e-16


def test_format_color_edges(orc):
    """vec.rs:125-131 incl. §Q10: NaN -> 0, inf -> 255, negatives -> 0, >= 1 -> 255."""
    assert orc.format_color([0.0, 1.0, 4.0], 1) == [0, 255, 255]
    assert orc.format_color([float("nan"), float("inf"), -1.0], 1) == [0, 255, 0]
    assert orc.format_color([0.25 * 800, 0.5 * 800, 0.81 * 800], 800) == [128, int(256 * math.sqrt(0.5)), int(256 * 0.9)]
    img = np.array([[[float("nan"), 1e9, 0.04]]])
    assert orc.format_image(img, 1).tolist() == [[[0, 255, 51]]]


# ---- analytic ray / primitive geometry -------------------------------------------------------
def _scene(rt, orc, build):
    b = rt.SceneBuilder()
    world, lights = build(b)
    sd = b.finish(world, lights)
    return sd, orc.OracleScene(sd)


def _rays(rt, o, d, t=0.0):
    rays = np.zeros(len(o), dtype=rt.RAY_DTYPE)
    rays["origin"], rays["direction"], rays["time"] = o, d, t
    return rays


def test_sphere_hit_analytic(rt, orc):
    def build(b):
        m = b.lambertian(b.constant_texture((0.5, 0.5, 0.5)))
        s = b.sphere((0, 0, 0), 2.0, m)
        return b.list([s]), b.list([])
    sd, osc = _scene(rt, orc, build)
    h = osc.trace_first_hit(_rays(rt, [(0, 0, -10), (0, 0, 0), (0, 3, -10), (0, 0, -10)],
                                  [(0, 0, 2), (0, 0, 1), (0, 0, 1), (0, 0, -1)]))
    # direction is not normalised: t is in units of |d| (ray.rs:26-28)
    assert h["node"].tolist() == [0, 0, -1, -1]
    assert abs(h["t"][0] - 4.0) < 1e-15 and np.allclose(h["position"][0], (0, 0, -2)) and h["front_face"][0] == 1
    assert np.allclose(h["normal"][0], (0, 0, -1))
    # from inside: far root, normal flipped against the ray, front_face false (hit.rs:34-41)
    assert abs(h["t"][1] - 2.0) < 1e-15 and h["front_face"][1] == 0 and np.allclose(h["normal"][1], (0, 0, -1))
    # uv of the hit point (0,0,-1)*: sphere.rs:12-17 -> (0.75, 0.5)
    assert abs(h["u"][0] - 0.75) < 1e-15 and abs(h["v"][0] - 0.5) < 1e-15


def test_rect_and_flip(rt, orc):
    def build(b):
        m = b.diffuse_light(b.constant_texture((1, 1, 1)))
        r = b.rect(rt._abi.PLANE_XZ, 0, 2, 0, 4, 5.0, m)
        return b.list([b.flip(r)]), b.list([])
    sd, osc = _scene(rt, orc, build)
    h = osc.trace_first_hit(_rays(rt, [(1, 0, 1), (1, 10, 1), (3, 0, 1), (2, 0, 4)], [(0, 1, 0), (0, -2, 0), (0, 1, 0), (0, 1, 0)]))
    assert h["node"].tolist() == [0, 0, -1, 0]  # edges are inside (rect.rs:57)
    assert h["t"].tolist()[:2] == [5.0, 2.5]
    assert abs(h["u"][0] - 0.5) < 1e-16 and abs(h["v"][0] - 0.25) < 1e-16
    # FlipNormal flips front_face only, never the normal (§Q2)
    # ray 0 travels along +y: back face (hit.rs:35), flipped to front; ray 1 the opposite
    assert h["front_face"].tolist()[:2] == [1, 0]
    assert np.allclose(h["normal"][0], (0, -1, 0)) and np.allclose(h["normal"][1], (0, 1, 0))


def test_triangle_moller_trumbore(rt, orc):
    def build(b):
        m = b.lambertian(b.constant_texture((1, 1, 1)))
        t = b.triangle((0, 0, 0), (1, 0, 0), (0, 1, 0), m)
        return b.list([t]), b.list([])
    sd, osc = _scene(rt, orc, build)
    h = osc.trace_first_hit(_rays(rt, [(0.25, 0.5, 3), (0.25, 0.5, -3), (0.8, 0.8, 3)], [(0, 0, -1), (0, 0, 1), (0, 0, -1)]))
    assert h["node"].tolist() == [0, 0, -1]  # two-sided, outside the hypotenuse misses
    assert np.allclose(h["t"][:2], 3.0) and np.allclose(h["u"][:2], 0.25) and np.allclose(h["v"][:2], 0.5)
    assert np.allclose(h["normal"][0], (0, 0, 1)) and np.allclose(h["normal"][1], (0, 0, -1))
    assert h["front_face"].tolist()[:2] == [1, 0]


def test_cube_faces_and_instance_transforms(rt, orc):
    """Translate(Rotate(Cube)) equals the cube hit by the inversely transformed ray (translate.rs, rotate.rs)."""
    ang = 25.0

    def plain_scene(b):
        m = b.lambertian(b.constant_texture((1, 1, 1)))
        return b.list([b.cube((0, 0, 0), (1, 2, 3), m)]), b.list([])

    def inst_scene(b):
        m = b.lambertian(b.constant_texture((1, 1, 1)))
        cube = b.cube((0, 0, 0), (1, 2, 3), m)
        return b.list([b.translate(b.rotate(rt._abi.AXIS_Y, cube, ang), (10, 0, 5))]), b.list([])
    _, o_plain = _scene(rt, orc, plain_scene)
    _, o_inst = _scene(rt, orc, inst_scene)
    rng = np.random.default_rng(3)
    target = rng.uniform((-0.3, -0.3, -0.3), (1.3, 2.3, 3.3), size=(600, 3))
    o = np.array([0.5, 1.0, 1.5]) + 8.0 * (lambda v: v / np.linalg.norm(v, axis=1, keepdims=True))(rng.normal(size=(600, 3)))
    d = (target - o) * rng.uniform(0.3, 2.0, size=(600, 1))
    plain = o_plain.trace_first_hit(_rays(rt, o, d))
    # rotate.rs:82-98: object = R(world - offset); so world = R^-1 object + offset
    th = math.radians(ang)
    c, s = math.cos(th), math.sin(th)

    def to_world(v, point):
        x = c * v[:, 0] + s * v[:, 2]
        z = -s * v[:, 0] + c * v[:, 2]
        w = np.stack([x, v[:, 1], z], axis=1)
        return w + np.array([10, 0, 5]) if point else w
    inst = o_inst.trace_first_hit(_rays(rt, to_world(o, True), to_world(d, False)))
    hit = plain["node"] == 0
    assert hit.sum() > 300 and (~hit).sum() > 20
    assert np.array_equal(inst["node"] >= 0, hit)
    assert np.array_equal(inst["face"][hit], plain["face"][hit])
    assert rel_err(inst["t"][hit], plain["t"][hit]).max() < 1e-12
    assert np.abs(inst["position"][hit] - to_world(plain["position"][hit], True)).max() < 1e-12
    # §Q3 (rotate.rs:88,98-102): Rotate re-runs set_face_normal with the OBJECT-space ray against the
    # rotated-back normal, so the orientation is decided by d_obj . (R^-1 n_obj), not d_obj . n_obj.
    n_w = to_world(plain["normal"][hit], False)
    s_q3 = (d[hit] * n_w).sum(axis=1) < 0
    expect_n = np.where(s_q3[:, None], n_w, -n_w)
    assert np.abs(inst["normal"][hit] - expect_n).max() < 1e-12
    assert np.array_equal(inst["front_face"][hit], s_q3.astype(np.int32))
    quirk = s_q3 != (plain["front_face"][hit] == plain["front_face"][hit])  # object-space hits are always forwarded
    assert 0 < quirk.sum() < hit.sum() // 4  # the quirk only bites some oblique hits on the rotated faces
    # all six faces are reachable; cube.rs:17-25 order: +z -z +y -y +x -x
    faces = o_plain.trace_first_hit(_rays(rt, [(0.5, 1, 9), (0.5, 1, -9), (0.5, 9, 1), (0.5, -9, 1), (9, 1, 1), (-9, 1, 1)],
                                          [(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)]))
    assert faces["face"].tolist() == [0, 1, 2, 3, 4, 5]


def test_bvh_equals_linear_list(rt, orc):
    """The reference's BVH (bvh.rs) must find the same closest hit as its HittableList (hit.rs:59-71)."""
    rng = np.random.default_rng(5)
    cen = rng.uniform(-10, 10, size=(200, 3))
    rad = rng.uniform(0.2, 1.5, size=200)

    def build_with(kind):
        def build(b):
            m = b.lambertian(b.constant_texture((1, 1, 1)))
            kids = [b.sphere(c, r, m) for c, r in zip(cen, rad)]
            kids.append(b.cube((-3, -3, -3), (-1, 0, 2), m))
            kids.append(b.triangle((0, 0, 5), (4, 0, 5), (0, 4, 6), m))
            return (b.bvh(kids) if kind == "bvh" else b.list(kids)), b.list([])
        return build
    _, o_bvh = _scene(rt, orc, build_with("bvh"))
    _, o_list = _scene(rt, orc, build_with("list"))
    o = rng.uniform(-15, 15, size=(5000, 3))
    d = rng.normal(size=(5000, 3))
    a, b = o_bvh.trace_first_hit(_rays(rt, o, d)), o_list.trace_first_hit(_rays(rt, o, d))
    assert (a["node"] >= 0).sum() > 1000
    assert np.array_equal(a["node"], b["node"]) and np.array_equal(a["t"], b["t"])
    depth, nodes = o_bvh.bvh_stats(202)
    assert nodes == 2 * 202 - 1 and depth == 9  # one object per leaf, halving split (bvh.rs:56-70)


def test_rect_light_pdf(rt, orc):
    """rect.rs:91-101: pdf = distance^2 / (cosine * area)."""
    hs = host_scene(rt, "cornell")
    osc = orc.OracleScene(hs.scene_desc)
    o = np.array([278.0, 100.0, 279.5])
    v = np.array([0.1, 2.0, -0.05])
    t = (554.0 - o[1]) / v[1]
    expect = (t * t * v.dot(v)) / ((abs(v[1]) / np.linalg.norm(v)) * (343 - 213) * (332 - 227))
    assert abs(osc.light_pdf(o, v) - expect) / expect < 1e-14
    assert osc.light_pdf(o, [1.0, 0.01, 0.0]) == 0.0  # misses the light


def test_medium_absorbs_under_head_integrator(rt, orc):
    """§Q6: Isotropic has no scatter_mc_method, so a medium hit returns black under HEAD, and the
    legacy integrator scatters."""
    hs = host_scene(rt, "cornell_smoke")
    osc = orc.OracleScene(hs.scene_desc)
    px = np.full(400, 32, dtype=np.uint32)
    py = np.full(400, 20, dtype=np.uint32)  # looks at the boxes
    s = np.arange(400, dtype=np.uint32)
    head, seg_h = osc.path_radiance(hs.camera, 64, 64, 50, rt.render_opts(seed=1, integrator=0), px, py, s)
    legacy, seg_l = osc.path_radiance(hs.camera, 64, 64, 50, rt.render_opts(seed=1, integrator=1), px, py, s)
    assert np.isfinite(head).all()
    assert (seg_h == 1).sum() > 20  # paths absorbed at the first medium hit
    assert seg_l.mean() > seg_h.mean()


def test_depth_zero_and_one(rt, orc):
    hs = host_scene(rt, "cornell")
    osc = orc.OracleScene(hs.scene_desc)
    px = np.arange(64, dtype=np.uint32)
    rgb, seg = osc.path_radiance(hs.camera, 64, 64, 0, rt.render_opts(), px, px, px)
    assert not rgb.any() and not seg.any()  # main.rs:42-45
    rgb, seg = osc.path_radiance(hs.camera, 64, 64, 1, rt.render_opts(), px, px, px)
    assert (seg == 1).all()


def test_head_integrator_needs_lights(rt, orc):
    """§Q7: random_scene has an empty light list; the reference unwrap()s and panics."""
    hs = host_scene(rt, "random")
    osc = orc.OracleScene(hs.scene_desc)
    px = np.zeros(1, dtype=np.uint32)
    with pytest.raises(orc.OracleError):
        osc.path_radiance(hs.camera, 16, 16, 5, rt.render_opts(integrator=0), px, px, px)
    rgb, _ = osc.path_radiance(hs.camera, 16, 16, 5, rt.render_opts(integrator=1), px, px, px)
    assert np.isfinite(rgb).all()


def test_textures(rt, orc):
    hs = host_scene(rt, "final")
    osc = orc.OracleScene(hs.scene_desc)
    d = hs.scene_desc.struct
    kinds = [d.textures[i].kind for i in range(d.n_textures)]
    img = kinds.index(rt._abi.TEX_IMAGE)
    noise = kinds.index(rt._abi.TEX_NOISE)
    raw = np.fromfile(os.path.join(rt.ASSETS_DIR, "earthmap_1024x512.rgb"), dtype=np.uint8).reshape(512, 1024, 3)
    # texture.rs:99-121: i = u*W, j = (1-v)*H, clamped; nearest texel / 255
    for u, v in [(0.0, 1.0), (0.5, 0.5), (0.999999, 0.000001), (1.0, 0.0), (0.3, 0.7), (-2.0, 5.0)]:
        i = min(int(min(max(u, 0), 1) * 1024), 1023)
        j = min(int(min(max(1 - v, 0), 1) * 512), 511)
        assert np.array_equal(osc.texture(img, u, v, (0, 0, 0)), raw[j, i] / 255.0)
    # texture.rs:77: marble is grey in [0,1]
    rng = np.random.default_rng(2)
    for p in rng.uniform(100, 400, size=(50, 3)):
        c = osc.texture(noise, 0, 0, p)
        assert c[0] == c[1] == c[2] and 0.0 <= c[0] <= 1.0
    # texture.rs:45-54 checker on the random scene's ground
    hs2 = host_scene(rt, "random")
    o2 = orc.OracleScene(hs2.scene_desc)
    d2 = hs2.scene_desc.struct
    chk = [d2.textures[i].kind for i in range(d2.n_textures)].index(rt._abi.TEX_CHECKER)
    for p in rng.uniform(-5, 5, size=(50, 3)):
        sines = math.sin(10 * p[0]) * math.sin(10 * p[1]) * math.sin(10 * p[2])
        expect = (1.0, 1.0, 1.0) if sines < 0 else (0.3, 0.3, 1.0)
        assert np.allclose(o2.texture(chk, 0, 0, p), expect)


@pytest.mark.parametrize("name", ["random", "cornell", "cornell_smoke", "final", "mesh"])
def test_oracle_reproduces_golden_fixture(rt, orc, name):
    """tests/golden/paths_*.npz pin the oracle (and the scene constructors) against accidental change."""
    hs = host_scene(rt, name)
    osc = orc.OracleScene(hs.scene_desc)
    g = np.load(os.path.join(GOLDEN, "paths_%s.npz" % name))
    W, H, depth = int(g["width"]), int(g["height"]), int(g["max_depth"])
    opts = rt.render_opts(seed=int(g["seed"]), integrator=int(g["integrator"]))
    rays = orc.camera_rays(hs.camera, W, H, opts, g["px"], g["py"], g["sample"])
    for f in ("origin", "direction", "time"):
        assert np.array_equal(rays[f], g["rays"][f])
    hits = osc.trace_first_hit(rays)
    assert np.array_equal(hits["node"], g["hits"]["node"]) and np.array_equal(hits["t"], g["hits"]["t"])
    rgb, seg = osc.path_radiance(hs.camera, W, H, depth, opts, g["px"], g["py"], g["sample"])
    assert np.array_equal(seg, g["segments"])
    assert np.allclose(rgb, g["rgb"], rtol=1e-12, atol=0, equal_nan=True)


def test_oracle_render_is_sum_of_paths(rt, orc):
    """main.rs:811-830: a pixel is the sum of its samples; rows are emitted top-down (main.rs:772)."""
    hs = host_scene(rt, "cornell")
    osc = orc.OracleScene(hs.scene_desc)
    W, H, spp = 12, 10, 6
    opts = rt.render_opts(seed=2)
    img, rays = osc.render(hs.camera, W, H, spp, 20, opts)
    ii, jj, ss = np.meshgrid(np.arange(W), np.arange(H), np.arange(spp), indexing="ij")
    rgb, seg = osc.path_radiance(hs.camera, W, H, 20, opts, ii.ravel(), jj.ravel(), ss.ravel())
    acc = np.zeros((H, W, 3))
    for (i, j), c in zip(zip(ii.ravel(), jj.ravel()), rgb):
        acc[H - 1 - j, i] += c
    assert np.allclose(img, acc, rtol=1e-13)
    assert rays == seg.sum()
    half = osc.render(hs.camera, W, H, spp, 20, rt.render_opts(seed=2, sample_begin=0, sample_count=3))[0] + \
        osc.render(hs.camera, W, H, spp, 20, rt.render_opts(seed=2, sample_begin=3, sample_count=3))[0]
    assert np.allclose(img, half, rtol=1e-13)


# ---- the PBR material's scalar helpers (mat.rs:10-44) -------------------------------------------
def test_pbr_schlick_and_smith(orc):
    assert orc.pbr_scalar(0, 0.0) == 1.0 and orc.pbr_scalar(0, 1.0) == 0.0
    assert orc.pbr_scalar(0, 0.5) == pytest.approx(1.0 / 32.0, rel=1e-15)
    assert orc.pbr_scalar(0, -3.0) == 1.0 and orc.pbr_scalar(0, 7.0) == 0.0  # clamp(0, 1), mat.rs:11
    # smithG_GGX(1, a) = 1 / (1 + sqrt(a^2 + 1 - a^2)) = 1/2 for every a (mat.rs:36-40)
    for a in (0.0, 0.25, 0.9):
        assert orc.pbr_scalar(3, 1.0, a) == pytest.approx(0.5, rel=1e-15)
    # the anisotropic form with ax = ay = a, v in the x-z plane, equals 1 / (n.v + sqrt(sin^2 a^2 + cos^2))
    c, sn, a = 0.6, 0.8, 0.3
    assert orc.pbr_scalar(4, c, sn, 0.0, a, a) == pytest.approx(1.0 / (c + np.sqrt((sn * a) ** 2 + c * c)), rel=1e-15)


def _hemisphere_integral(fn, n_theta=2000, n_phi=720):
    """Integral over the hemisphere of fn * cos(theta) d_omega (midpoint rule).  fn(cos_theta) for a
    lobe that does not depend on phi, else fn(cos_theta, hx, hy)."""
    th = (np.arange(n_theta) + 0.5) * (np.pi / 2) / n_theta
    ph = (np.arange(n_phi) + 0.5) * (2 * np.pi) / n_phi
    d_theta = (np.pi / 2) / n_theta
    tot = 0.0
    for t in th:
        ct, st = np.cos(t), np.sin(t)
        if fn.__code__.co_argcount == 1:
            ring = fn(ct) * 2 * np.pi
        else:
            ring = sum(fn(ct, st * np.cos(p), st * np.sin(p)) for p in ph) * (2 * np.pi) / n_phi
        tot += ring * ct * st * d_theta
    return tot


def test_pbr_gtr2_aniso_is_a_normalised_distribution(orc):
    """GTR_2_aniso (mat.rs:32-34) integrates to 1 against cos(theta) over the hemisphere."""
    iso = _hemisphere_integral(lambda ct, hx, hy: orc.pbr_scalar(2, ct, hx, hy, 0.3, 0.3), n_theta=1500, n_phi=8)
    assert iso == pytest.approx(1.0, abs=2e-3)
    aniso = _hemisphere_integral(lambda ct, hx, hy: orc.pbr_scalar(2, ct, hx, hy, 0.25, 0.6), n_theta=600, n_phi=240)
    assert aniso == pytest.approx(1.0, abs=5e-3)


def test_pbr_gtr1_keeps_the_reference_log2(orc):
    """GTR_1 divides by log2(a^2) where Burley's normalisation has ln(a^2) (mat.rs:22): the
    reference's lobe therefore integrates to ln 2, not 1.  a >= 1 is the uniform 1/pi."""
    a = 0.1
    got = _hemisphere_integral(lambda ct: orc.pbr_scalar(1, ct, a), n_theta=20000)
    assert got == pytest.approx(np.log(2.0), rel=2e-3)
    assert orc.pbr_scalar(1, 0.3, 1.0) == pytest.approx(1.0 / np.pi, rel=1e-15)
    assert orc.pbr_scalar(1, 0.3, 2.5) == pytest.approx(1.0 / np.pi, rel=1e-15)


def test_pbr_paths_reproduce_the_reference_nan(rt, orc):
    """PDF::BRDF reflects the WORLD-space incoming direction about a TANGENT-space half vector
    (pdf.rs:35,59), so many sampled directions leave below the surface: brdf = 0 and pdf = 0, and
    main.rs:104 divides 0 by 0.  The restatement keeps that (no NaN guard, §Q10)."""
    hs = rt.HostScene("cornell_pbr")
    osc = orc.OracleScene(hs.scene_desc)
    rng = np.random.default_rng(3)
    px, py, s = (rng.integers(0, 64, 4000, dtype=np.uint32) for _ in range(3))
    rgb, seg = osc.path_radiance(hs.camera, 64, 64, 100, rt.render_opts(seed=2), px, py, s)
    nan = np.isnan(rgb).any(axis=1)
    assert 0.02 < nan.mean() < 0.5
    assert np.isfinite(rgb[~nan]).all() and (rgb[~nan] >= 0.0).all()
