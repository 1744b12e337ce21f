"""The oracle against tests/second_hand.py, a restatement of the Cornell path written a second time, in Python, from
the reference's sources (and building the two Cornell scenes and their camera itself from src/main.rs): per-path
radiance of identical (pixel, sample) paths must agree to rounding.  Anchors what the closed-form tests do not reach:
the 50/50 mixture weighting of main.rs:92-98, Rotate's re-orientation with the rotated ray (§Q3) and the clamping
of ConstantMedium's two boundary hits (medium.rs:32-58)."""
import numpy as np
import pytest

import second_hand as sh
from util import host_scene

W = H = 64
DEPTH = 50
SEED = 7


def _oracle_paths(rt, orc, name, px, py, s):
    hs = host_scene(rt, name)
    osc = orc.OracleScene(hs.scene_desc)
    opts = rt.render_opts(seed=SEED, integrator=rt.INTEGRATOR_HEAD)
    rgb, seg = osc.path_radiance(hs.camera, W, H, DEPTH, opts, px, py, s)
    return hs, rgb, seg


def _ids(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.integers(0, W, n, dtype=np.uint32), rng.integers(0, H, n, dtype=np.uint32), rng.integers(0, 1000, n, dtype=np.uint32))


def _compare(scene, rgb, px, py, s):
    mine = np.array([sh.path_radiance(scene, W, H, DEPTH, SEED, int(i), int(j), int(k)) for i, j, k in zip(px, py, s)])
    assert np.isfinite(mine).all() and np.isfinite(rgb).all()
    err = np.abs(mine - rgb) / np.maximum(np.abs(rgb), 1e-12)
    lit = (rgb > 0).any(axis=1)
    print("paths %d, lit %d, max rel err %.3e, identical %.4f" % (len(px), lit.sum(), err.max(), (mine == rgb).all(axis=1).mean()))
    assert lit.sum() > len(px) // 10  # the comparison is not one of zeros
    assert err.max() <= 1e-12
    assert (mine == 0.0).all(axis=1).tolist() == (rgb == 0.0).all(axis=1).tolist()


def test_philox_restated_twice(orc):
    for ctr, key in (((0, 0, 0, 0), (0, 0)), ((1, 4, 0, 7), (4095, 999)), ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2)):
        assert list(sh.philox4x32_10(ctr, key)) == orc.philox4x32_10(list(ctr), list(key))
    a, b, ba, bb = sh.Draws(3, 77, 5).draw(2, sh.SLOT_SCATTER, 0)
    assert (a, b, ba, bb) == orc.draw(3, 77, 5, 2, sh.SLOT_SCATTER, 0)


def test_camera_restated_twice(rt):
    hs = host_scene(rt, "cornell")
    cam = sh.cornell_box()[3]
    c = hs.camera
    for mine, theirs in ((cam.origin, c.origin), (cam.llc, c.lower_left_corner), (cam.horizontal, c.horizontal),
                         (cam.vertical, c.vertical), (cam.cu, c.cu), (cam.cv, c.cv)):
        assert tuple(mine) == tuple(theirs)
    assert (cam.lens_radius, cam.time0, cam.time1) == (c.lens_radius, c.time0, c.time1)


def test_cornell_box_paths(rt, orc):
    px, py, s = _ids(1500, 1)
    _, rgb, _ = _oracle_paths(rt, orc, "cornell", px, py, s)
    _compare(sh.cornell_box(), rgb, px, py, s)


def test_cornell_smoke_paths(rt, orc):
    px, py, s = _ids(1500, 2)
    hs, rgb, _ = _oracle_paths(rt, orc, "cornell_smoke", px, py, s)
    d = hs.scene_desc.struct
    media = [k for k in range(d.n_nodes) if d.nodes[k].kind == rt._abi.NODE_MEDIUM]
    assert len(media) == 2
    _compare(sh.cornell_box_with_smoke(media), rgb, px, py, s)


def _general(rt, orc, hs, legacy, n_paths, seed, depth=DEPTH, nan_paths=False, min_nonzero=None):
    world, lights, background = sh.scene_from_desc(rt._abi, hs.scene_desc.struct)
    cam = sh.CameraPod(hs.camera)
    px, py, s = _ids(n_paths, seed)
    osc = orc.OracleScene(hs.scene_desc)
    integrator = rt.INTEGRATOR_LEGACY if legacy else rt.INTEGRATOR_HEAD
    rgb, _ = osc.path_radiance(hs.camera, W, H, depth, rt.render_opts(seed=SEED, integrator=integrator), px, py, s)
    mine = np.array([sh.path_radiance_general(world, lights, background, cam, W, H, depth, SEED, int(i), int(j), int(k), legacy)
                     for i, j, k in zip(px, py, s)])
    finite = np.isfinite(mine).all(axis=1) & np.isfinite(rgb).all(axis=1)
    if nan_paths:  # the PBR material: 0/0 at main.rs:104 for directions sampled below the surface - the SAME paths, channel by channel
        assert np.array_equal(np.isnan(mine), np.isnan(rgb)) and np.array_equal(np.isinf(mine), np.isinf(rgb))
        print("non-finite paths %d of %d, the same ones" % ((~finite).sum(), n_paths))
    else:
        assert finite.mean() > 0.999
    err = np.abs(mine[finite] - rgb[finite]) / np.maximum(np.abs(rgb[finite]), 1e-12)
    print("paths %d, nonzero %d, max rel err %.3e, identical %.4f" % (n_paths, (rgb > 0).any(axis=1).sum(), err.max(), (mine == rgb).all(axis=1).mean()))
    assert (rgb > 0).any(axis=1).sum() > (n_paths // 10 if min_nonzero is None else min_nonzero)
    assert err.max() <= 1e-10  # (sin / atan2 / acos of libm on both sides; recursion order is the reference's on both)
    osc.close()


def test_rtiow_random_spheres_legacy_paths(rt, orc):
    """Config 1: spheres, moving spheres, glass, fuzzy metal, the checker ground, the legacy integrator - read from the
    same RtSceneDesc the oracle gets, evaluated by the second restatement (a BVH node read as the list of its members)."""
    _general(rt, orc, host_scene(rt, "random"), True, 400, 3)


def test_cornell_scenes_through_the_scene_reader(rt, orc):
    """The reader itself: the two Cornell scenes read from the description agree as well as the hand-built ones."""
    _general(rt, orc, host_scene(rt, "cornell"), False, 500, 4)
    _general(rt, orc, host_scene(rt, "cornell_smoke"), False, 500, 5)


def test_mesh_scene_head_paths(rt, orc):
    """Config 5 at the tests' mesh size: triangles (Moeller-Trumbore as written in tri.rs), walls, the rect light."""
    _general(rt, orc, host_scene(rt, "mesh"), False, 40, 6, depth=12)


def test_next_week_final_scene_head_paths(rt, orc):
    """Config 4: the ground boxes, the moving sphere, glass, fuzzy metal, two spherical media, the earth image texture,
    the marble sphere (Perlin, restated here once more), 1000 spheres under Rotate + Translate."""
    _general(rt, orc, host_scene(rt, "final"), False, 150, 8)


def test_perlin_and_earth_scenes_under_the_sky(rt, orc):
    """Every path of these two ends in the sky colour, so every texel and every marble value on the way is in the result
    (main.rs:229-262: two_perlin_sphere, earth)."""
    _general(rt, orc, host_scene(rt, "two_perlin_spheres"), True, 300, 9)
    _general(rt, orc, host_scene(rt, "earth"), True, 300, 10)


def test_pbr_material_and_sphere_lights(rt, orc):
    """What the five configs do not use but HEAD has: the PBR material (ScatterRecord::Microfacet, mat.rs:118-200 with the
    three-lobe PDF::BRDF of pdf.rs:20-63,103-161) and Sphere::pdf_value / random (sphere.rs:27-36,103-119)."""
    _general(rt, orc, host_scene(rt, "cornell_pbr"), False, 400, 11, nan_paths=True)
    _general(rt, orc, host_scene(rt, "progress_showcase"), False, 300, 12, nan_paths=True)
    _general(rt, orc, host_scene(rt, "light_room"), False, 300, 13)


def test_scattering_smoke_under_the_legacy_integrator(rt, orc):
    """What img/volume.png shows (tests/test_reference_images.py): ConstantMedium with Isotropic::scatter (mat.rs:418-421)
    through the `old method` of main.rs:82-84.  Brute force against a small lamp: few paths carry light, all are compared."""
    _general(rt, orc, host_scene(rt, "cornell_smoke"), True, 2500, 21, min_nonzero=8)
