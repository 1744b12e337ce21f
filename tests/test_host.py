"""Host layer (the C++ stand-in for the Rust host): Camera::new, scene constructors, OBJ
loader, format_color and the PPM writer.  CPU only."""
import math
import os

import numpy as np
import pytest

from util import host_scene


def test_camera_new_matches_formulas(rt):
    """src/camera.rs:19-49 recomputed with numpy."""
    lookfrom, lookat, vup = np.array([278.0, 278.0, -800.0]), np.array([278.0, 278.0, 0.0]), np.array([0.0, 1.0, 0.0])
    vfov, aspect, aperture, focus = 40.0, 1.0, 0.05, 10.0
    cam = rt.camera_new(lookfrom, lookat, vup, vfov, aspect, aperture, focus, 0.0, 1.0)
    theta = math.pi / 180.0 * vfov
    vh = 2.0 * math.tan(theta / 2.0)
    vw = vh * aspect
    cw = (lookfrom - lookat) / np.linalg.norm(lookfrom - lookat)
    cu = np.cross(vup, cw)
    cu /= np.linalg.norm(cu)
    cv = np.cross(cw, cu)
    h, v = focus * vw * cu, focus * vh * cv
    llc = lookfrom - h / 2 - v / 2 - focus * cw
    assert np.allclose(cam.origin[:], lookfrom) and np.allclose(cam.horizontal[:], h, atol=1e-14)
    assert np.allclose(cam.vertical[:], v, atol=1e-14) and np.allclose(cam.lower_left_corner[:], llc, atol=1e-12)
    assert cam.lens_radius == aperture / 2 and (cam.time0, cam.time1) == (0.0, 1.0)
    hs = host_scene(rt, "cornell")  # main.rs:700-705
    assert np.allclose(hs.camera.lower_left_corner[:], llc, atol=1e-12)


def _kinds(rt, hs):
    d = hs.scene_desc.struct
    return [d.nodes[i].kind for i in range(d.n_nodes)]


def test_cornell_scene_structure(rt):
    """src/main.rs:278-311: 6 rects (one flipped light), two Translate(Rotate(Cube)) instances, one light."""
    hs = host_scene(rt, "cornell")
    A = rt._abi
    k = _kinds(rt, hs)
    assert k.count(A.NODE_RECT) == 6 and k.count(A.NODE_CUBE) == 2 and k.count(A.NODE_ROTATE) == 2
    assert k.count(A.NODE_TRANSLATE) == 2 and k.count(A.NODE_FLIP) == 1 and k.count(A.NODE_LIST) == 2
    d = hs.scene_desc.struct
    world, lights = d.nodes[d.world], d.nodes[d.lights]
    assert world.count == 8 and lights.count == 1
    # the light in `lights` is the same object as in `world` (rect_light.clone(), main.rs:293,308)
    assert d.child_index[lights.child] in [d.child_index[world.child + i] for i in range(world.count)]
    rot = [d.nodes[i] for i in range(d.n_nodes) if d.nodes[i].kind == A.NODE_ROTATE]
    assert sorted(r.v[0] for r in rot) == [-18.0, 15.0] and all(r.axis == A.AXIS_Y for r in rot)
    mats = [d.materials[i] for i in range(d.n_materials)]
    metal = [m for m in mats if m.kind == A.MAT_METAL]
    assert len(metal) == 1 and metal[0].fuzz == 0.0 and list(metal[0].albedo) == [0.8, 0.85, 0.88]  # main.rs:286,306
    assert (hs.integrator, hs.width, hs.height, hs.spp, hs.max_depth) == (A.INTEGRATOR_HEAD, 600, 600, 1000, 100)


def test_random_scene_structure(rt):
    """src/main.rs:153-210: ground + 23x23 grid + 3 big spheres in one BVH, empty light list."""
    hs = host_scene(rt, "random")
    A = rt._abi
    k = _kinds(rt, hs)
    assert k.count(A.NODE_SPHERE) + k.count(A.NODE_MOVING_SPHERE) == 1 + 23 * 23 + 3
    assert k.count(A.NODE_BVH) == 1
    frac_moving = k.count(A.NODE_MOVING_SPHERE) / 529.0
    assert 0.7 < frac_moving < 0.9  # choose_mat < 0.8
    d = hs.scene_desc.struct
    assert d.nodes[d.lights].count == 0 and d.nodes[d.world].kind == A.NODE_BVH
    assert hs.integrator == A.INTEGRATOR_LEGACY and list(d.background) == [0.7, 0.8, 1.0]
    for i in range(d.n_nodes):
        n = d.nodes[i]
        if n.kind == A.NODE_MOVING_SPHERE:  # main.rs:173-174
            assert n.v[0] == n.v[3] and n.v[2] == n.v[5] and 0.0 <= n.v[4] - n.v[1] < 0.01 and n.v[8] == 0.2


def test_final_scene_structure(rt):
    """src/main.rs:453-513."""
    hs = host_scene(rt, "final")
    A = rt._abi
    k = _kinds(rt, hs)
    assert k.count(A.NODE_CUBE) == 400 and k.count(A.NODE_SPHERE) == 1000 + 6 and k.count(A.NODE_MOVING_SPHERE) == 1
    assert k.count(A.NODE_MEDIUM) == 2 and k.count(A.NODE_BVH) == 2
    d = hs.scene_desc.struct
    assert d.nodes[d.world].count == 11 and d.n_images == 1 and d.n_perlin == 1
    assert (d.images[0].width, d.images[0].height) == (1024, 512) and d.n_texel_bytes == 1024 * 512 * 3
    med = [d.nodes[i] for i in range(d.n_nodes) if d.nodes[i].kind == A.NODE_MEDIUM]
    assert sorted(m.v[0] for m in med) == [0.0001, 0.2]
    # the dielectric boundary sphere is both in the world and the boundary of the first medium (main.rs:484-486)
    kids = [d.child_index[d.nodes[d.world].child + i] for i in range(11)]
    assert any(m.child in kids for m in med)
    p = d.perlin[0]
    assert sorted(p.perm_x) == list(range(256)) and sorted(p.perm_y) == list(range(256)) and sorted(p.perm_z) == list(range(256))
    rv = np.array(p.ranvec[:]).reshape(256, 3)
    assert (np.linalg.norm(rv, axis=1) < 1.0).all()  # un-normalised in-ball vectors (perlin.rs:13-19)


def test_mesh_scene_and_obj_loader(rt):
    nv, nt = rt.obj_triangle_count(os.path.join(rt.ASSETS_DIR, "teapot.obj"))
    assert (nv, nt) == (530, 1024)  # SURVEY §2 asset facts
    hs = host_scene(rt, "mesh")
    A = rt._abi
    k = _kinds(rt, hs)
    assert k.count(A.NODE_BVH) == 2 and k.count(A.NODE_RECT) == 6
    assert k.count(A.NODE_TRIANGLE) == 1024 + 2 * 96 * 128  # teapot + stand-in at detail 1
    assert (hs.width, hs.height, hs.spp) == (3840, 2160, 1024)
    with pytest.raises(rt.RtError):
        rt.obj_triangle_count("/nonexistent.obj")


def test_obj_loader_rules(rt, tmp_path):
    """mesh.rs:40-52 / tobj: f32 positions, fan triangulation, first model only, negative indices."""
    p = tmp_path / "t.obj"
    p.write_text("# c\no first\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1 4//1\nf -4 -3 -2\n"
                 "o second\nv 5 5 5\nf 1 2 5\n")
    assert rt.obj_triangle_count(str(p)) == (4, 3)
    for bad in ("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 4294967297\n",
                "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 -4\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n", "v 0 0 0\n", ""):
        p.write_text(bad)
        with pytest.raises(rt.RtError):  # mesh.rs:40: expect("Failed to load obj file") panics; here a status
            rt.obj_triangle_count(str(p))


def test_construction_seed_is_deterministic(rt):
    a, b, c = rt.HostScene("random", 5), rt.HostScene("random", 5), rt.HostScene("random", 6)

    def centres(hs):
        d = hs.scene_desc.struct
        return [tuple(d.nodes[i].v[:4]) for i in range(d.n_nodes)]
    assert centres(a) == centres(b) and centres(a) != centres(c)


def test_unknown_scene(rt):
    with pytest.raises(rt.RtError):
        rt.HostScene("nope")


def test_format_image_matches_oracle_and_edges(rt, orc):
    """format_color, vec.rs:125-131."""
    rng = np.random.default_rng(0)
    img = rng.uniform(0, 2000, size=(17, 9, 3)).astype(np.float32)
    img[0, 0] = [np.nan, np.inf, -3.0]
    img[0, 1] = [0.0, 800.0, 799.0]
    a = rt.format_image(img, 800)
    b = orc.format_image(img.astype(np.float64), 800)
    assert np.array_equal(a, b)
    assert a[0, 0].tolist() == [0, 255, 0] and a[0, 1].tolist() == [0, 255, int(256 * math.sqrt(799 / 800))]


def test_ppm_writer(rt, tmp_path):
    """main.rs:767-769,832: P3 header, one 'r g b' line per pixel, rows as given (top row first)."""
    img = np.zeros((2, 3, 3), dtype=np.float32)
    img[0, 0] = [4.0, 1.0, 0.0]
    img[1, 2] = [0.25 * 4, 0.0, 4.0]
    path = str(tmp_path / "o.ppm")
    rt.write_ppm(path, img, 4)
    lines = open(path).read().split("\n")
    assert lines[:3] == ["P3", "3 2", "255"]
    assert lines[3] == "255 128 0" and lines[8] == "128 0 255" and len([l for l in lines if l]) == 3 + 6
