"""CPU tier: the oracle (and the g++ build of the device source) against closed-form radiometry.

The reference has no test vectors, so beyond its one table the oracle is pinned by restating the cited
lines faithfully.  These tests add an independent anchor: small scenes whose expected radiance has a closed form, so
that the integrator, the mixture-PDF light sampling (pdf.rs, rect.rs:91-111, sphere.rs:27-36,104-119), the material
pdfs and the medium's free-path sampling (medium.rs:42-45) are checked against physics, not against themselves.
"""
import math
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "native"))


@pytest.fixture(scope="module")
def toh():
    import trace_on_host as m
    return m


def both(rt, orc, toh, sd):
    return [("oracle", orc.OracleScene(sd)), ("device source on host", toh.CompiledOnHost(sd))]


def centre_paths(n, W, H):
    return (np.full(n, W // 2, dtype=np.uint32), np.full(n, H // 2, dtype=np.uint32), np.arange(n, dtype=np.uint32))


def rect_form_factor(a, b, h):
    """Differential element -> parallel rectangle [-a/2,a/2]x[-b/2,b/2] at height h above it (four corner terms)."""
    x, y = a / 2.0, b / 2.0
    corner = (x / math.hypot(x, h) * math.atan(y / math.hypot(x, h)) + y / math.hypot(y, h) * math.atan(x / math.hypot(y, h))) / (2.0 * math.pi)
    return 4.0 * corner


def test_furnace_legacy_integrator_returns_the_albedo(rt, orc, toh):
    """main.rs:84-85 on a convex Lambertian sphere in a uniform white environment: the scattered ray leaves the sphere
    and meets the background, so EVERY path that hits returns exactly albedo * 1 - no variance at all."""
    b = rt.SceneBuilder()
    rho = (0.25, 0.5, 0.75)
    s = b.sphere((0, 0, 0), 1.0, b.lambertian(b.constant_texture(rho)))
    sd = b.finish(b.list([s]), b.list([]), background=(1.0, 1.0, 1.0))
    cam = rt.camera_new((0, 0, -5), (0, 0, 0), (0, 1, 0), 10.0, 1.0, 0.0, 5.0)
    px, py, smp = centre_paths(2000, 33, 33)
    for name, sc in both(rt, orc, toh, sd):
        rgb, seg = sc.path_radiance(cam, 33, 33, 50, rt.render_opts(seed=7, integrator=rt.INTEGRATOR_LEGACY), px, py, smp)
        assert np.array_equal(seg, np.full(2000, 2)), name
        assert np.array_equal(rgb, np.tile(np.array(rho), (2000, 1))), name


def test_rect_light_over_a_diffuse_floor_matches_the_form_factor(rt, orc, toh):
    """main.rs:92-98: a Lambertian floor (albedo rho) under a flipped XZ rect light of radiance E, black background.
    The only light path is floor -> light, so L = rho * E * F with F the point-to-rectangle form factor.  The mixture
    of the light pdf (rect.rs:91-101) and the cosine pdf must average to it whatever the weights."""
    A = rt._abi
    rho, E, a, c, h = 0.6, 5.0, 3.0, 2.0, 2.5
    b = rt.SceneBuilder()
    floor = b.rect(A.PLANE_XZ, -500, 500, -500, 500, 0.0, b.lambertian(b.constant_texture((rho, rho, rho))))
    light = b.flip(b.rect(A.PLANE_XZ, -a / 2, a / 2, -c / 2, c / 2, h, b.diffuse_light(b.constant_texture((E, E, E)))))
    sd = b.finish(b.list([floor, light]), b.list([light]))
    # looks at the floor point under the centre of the light from the side, below the light's plane
    cam = rt.camera_new((6.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0, 1, 0), 0.05, 1.0, 0.0, 6.0)
    n = 60000
    px, py, smp = centre_paths(n, 3, 3)
    expect = rho * E * rect_form_factor(a, c, h)
    for name, sc in both(rt, orc, toh, sd):
        rgb, seg = sc.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=3, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
        assert np.isfinite(rgb).all(), name
        mean, sem = rgb[:, 0].mean(), rgb[:, 0].std() / math.sqrt(n)
        print(name, "L = %.5f +- %.5f, closed form %.5f" % (mean, sem, expect))
        assert abs(mean - expect) < 4.0 * sem + 1e-3 * expect, name
        assert sem < 0.01 * expect, name  # importance sampling keeps the variance low


def test_sphere_light_over_a_diffuse_floor(rt, orc, toh):
    """Sphere::pdf_value / random (sphere.rs:27-36,104-119; random_to_sphere): irradiance of a floor point from a
    uniform sphere of radiance E whose centre is straight above it is pi * E * (r/d)^2 - so L = rho * E * (r/d)^2."""
    A = rt._abi
    rho, E, r, d = 0.8, 4.0, 1.0, 3.0
    b = rt.SceneBuilder()
    floor = b.rect(A.PLANE_XZ, -500, 500, -500, 500, 0.0, b.lambertian(b.constant_texture((rho, rho, rho))))
    lamp = b.sphere((0.0, d, 0.0), r, b.diffuse_light(b.constant_texture((E, E, E))))
    sd = b.finish(b.list([floor, lamp]), b.list([lamp]))
    cam = rt.camera_new((6.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0, 1, 0), 0.05, 1.0, 0.0, 6.0)
    n = 60000
    px, py, smp = centre_paths(n, 3, 3)
    expect = rho * E * (r / d) ** 2
    for name, sc in both(rt, orc, toh, sd):
        rgb, _ = sc.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=5, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
        mean, sem = rgb[:, 0].mean(), rgb[:, 0].std() / math.sqrt(n)
        print(name, "L = %.5f +- %.5f, closed form %.5f" % (mean, sem, expect))
        assert abs(mean - expect) < 4.0 * sem + 1e-3 * expect, name


def test_medium_transmittance_is_beer_lambert(rt, orc, toh):
    """medium.rs:42-45: hit_distance = -(1/density) ln(xi) against the chord through the boundary.  With a black
    phase-function albedo under the legacy integrator a path returns the white background iff it crosses the slab
    unscattered, so the mean is exp(-density * thickness)."""
    density, thickness = 0.35, 2.0
    b = rt.SceneBuilder()
    slab = b.cube((-50, -50, 0.0), (50, 50, thickness), b.lambertian(b.constant_texture((1, 1, 1))))
    fog = b.medium(slab, density, b.constant_texture((0.0, 0.0, 0.0)))
    sd = b.finish(b.list([fog]), b.list([]), background=(1.0, 1.0, 1.0))
    cam = rt.camera_new((0, 0, -10), (0, 0, 0), (0, 1, 0), 0.05, 1.0, 0.0, 10.0)
    n = 40000
    px, py, smp = centre_paths(n, 3, 3)
    expect = math.exp(-density * thickness)
    for name, sc in both(rt, orc, toh, sd):
        rgb, seg = sc.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=9, integrator=rt.INTEGRATOR_LEGACY), px, py, smp)
        assert set(np.unique(rgb[:, 0])) <= {0.0, 1.0}, name
        mean = rgb[:, 0].mean()
        sem = math.sqrt(expect * (1 - expect) / n)
        print(name, "T = %.5f +- %.5f, Beer-Lambert %.5f" % (mean, sem, expect))
        assert abs(mean - expect) < 4.0 * sem, name


def test_mirror_and_glass_conserve_a_uniform_environment(rt, orc, toh):
    """Specular arms (main.rs:89-91; mat.rs:280-293, :343-374): in a uniform white environment a fuzz-free mirror
    returns exactly its albedo and a glass sphere exactly 1 (attenuation 1 at every interface), whatever the path."""
    b = rt.SceneBuilder()
    mirror = b.sphere((-1.5, 0, 0), 1.0, b.metal((0.8, 0.6, 0.4), 0.0))
    glass = b.sphere((1.5, 0, 0), 1.0, b.dielectric(1.5))
    lamp = b.flip(b.rect(rt._abi.PLANE_XZ, -0.1, 0.1, -0.1, 0.1, 50.0, b.diffuse_light(b.constant_texture((1, 1, 1)))))
    sd = b.finish(b.list([mirror, glass, lamp]), b.list([lamp]), background=(1.0, 1.0, 1.0))
    n = 3000
    smp = np.arange(n, dtype=np.uint32)
    for name, sc in both(rt, orc, toh, sd):
        for target, expect in (((-1.5, 0, 0), (0.8, 0.6, 0.4)), ((1.5, 0, 0), (1.0, 1.0, 1.0))):
            cam = rt.camera_new((target[0], 0.0, -6.0), target, (0, 1, 0), 6.0, 1.0, 0.0, 6.0)
            px, py, _ = centre_paths(n, 5, 5)
            rgb, seg = sc.path_radiance(cam, 5, 5, 100, rt.render_opts(seed=2, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
            assert (seg >= 2).all(), name
            done = seg < 100  # total internal reflection can trap a path until the depth limit (returns black)
            assert done.mean() > 0.99, name
            assert np.allclose(rgb[done], np.array(expect), rtol=0, atol=1e-12), (name, target)


def test_thin_lens_camera_geometry(rt, orc, toh):
    """camera.rs:19-59: every ray starts on the lens disk (radius aperture/2, perpendicular to the view axis), passes
    at parameter 1 through the focal plane at focus_dist, inside its pixel's footprint there, and carries a shutter
    time uniform in [time0, time1)."""
    lookfrom, lookat = np.array([3.0, 2.0, -7.0]), np.array([0.5, 0.0, 1.0])
    vfov, aspect, aperture, focus, t0, t1 = 35.0, 1.5, 0.8, 6.5, 0.25, 0.75
    cam = rt.camera_new(lookfrom, lookat, (0, 1, 0), vfov, aspect, aperture, focus, t0, t1)
    W, H, n = 90, 60, 40000
    rng = np.random.default_rng(1)
    px, py, s = rng.integers(0, W, n, dtype=np.uint32), rng.integers(0, H, n, dtype=np.uint32), np.arange(n, dtype=np.uint32)
    w = (lookfrom - lookat) / np.linalg.norm(lookfrom - lookat)
    u = np.cross((0, 1, 0), w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    half_h = focus * math.tan(math.radians(vfov) / 2.0)
    half_w = aspect * half_h
    opts = rt.render_opts(seed=4)
    for name, rays in (("oracle", orc.camera_rays(cam, W, H, opts, px, py, s)), ("device source on host", toh.camera_rays(cam, W, H, opts, px, py, s))):
        off = rays["origin"] - lookfrom
        assert np.abs(off @ w).max() < 1e-12, name                       # the lens lies in the plane through lookfrom
        r = np.hypot(off @ u, off @ v)
        assert r.max() <= aperture / 2 + 1e-12 and r.max() > 0.49 * aperture, name
        assert abs((r ** 2).mean() - (aperture / 2) ** 2 / 2) < 0.01 * (aperture / 2) ** 2, name  # uniform on the disk: E r^2 = R^2/2
        p = rays["origin"] + rays["direction"] - lookfrom                 # parameter 1: the focal plane
        assert np.abs(p @ w + focus).max() < 1e-10, name
        sx, sy = (p @ u + half_w) / (2 * half_w), (p @ v + half_h) / (2 * half_h)   # (s, t) of main.rs:817-818
        assert ((sx * (W - 1) >= px - 1e-9) & (sx * (W - 1) < px + 1 + 1e-9)).all(), name
        assert ((sy * (H - 1) >= py - 1e-9) & (sy * (H - 1) < py + 1 + 1e-9)).all(), name
        t = rays["time"]
        assert t.min() >= t0 and t.max() < t1 and abs(t.mean() - 0.5 * (t0 + t1)) < 0.005, name


def _perlin_rs(p, scale, ranvec, px_, py_, pz_):
    """perlin.rs:77-109 + :39-56 read anew in Python: Hermite smoothing applied in `perlin` AND again in
    `perlin_interp`, weights from the already smoothed (u, v, w), `floor(x) as usize` saturating negatives to 0."""
    def usize(x):
        return 0 if (x != x or x <= 0.0) else int(x)
    x, y, z = scale * p[0], scale * p[1], scale * p[2]
    u, v, w = x - math.floor(x), y - math.floor(y), z - math.floor(z)
    u, v, w = u * u * (3.0 - 2.0 * u), v * v * (3.0 - 2.0 * v), w * w * (3.0 - 2.0 * w)
    i, j, k = usize(math.floor(x)), usize(math.floor(y)), usize(math.floor(z))
    uu, vv, ww = u * u * (3.0 - 2.0 * u), v * v * (3.0 - 2.0 * v), w * w * (3.0 - 2.0 * w)
    accum = 0.0
    for di in range(2):
        for dj in range(2):
            for dk in range(2):
                c = ranvec[px_[(i + di) & 255] ^ py_[(j + dj) & 255] ^ pz_[(k + dk) & 255]]
                weight = (u - di, v - dj, w - dk)
                dot = c[0] * weight[0] + c[1] * weight[1] + c[2] * weight[2]
                accum += (di * uu + (1 - di) * (1.0 - uu)) * (dj * vv + (1 - dj) * (1.0 - vv)) * (dk * ww + (1 - dk) * (1.0 - ww)) * dot
    return accum


def _noise_texture_rs(p, scale, tables):
    """texture.rs:77 over perlin.rs:111-121 (turb: seven octaves, the point doubled per octave, abs of the sum)."""
    accum, weight, q = 0.0, 1.0, list(p)
    for _ in range(7):
        accum += weight * _perlin_rs(q, scale, *tables)
        weight *= 0.5
        q = [2.0 * c for c in q]
    return 0.5 * (1.0 + math.sin(scale * p[2] + 10.0 * abs(accum)))


def test_marble_texture_against_perlin_rs(rt, orc, toh):
    """An independent restatement of perlin.rs / texture.rs:77 (above) against the oracle and the device source: a
    marble sphere in a white furnace under the legacy integrator returns exactly the texture at the hit point."""
    rng = np.random.default_rng(12)
    ranvec = rng.uniform(-1, 1, (256, 3)) * rng.uniform(0.1, 1.0, (256, 1))
    perms = [rng.permutation(256) for _ in range(3)]
    scale = 1.7
    b = rt.SceneBuilder()
    marble = b.lambertian(b.noise_texture(scale, ranvec, *perms))
    # the centre is chosen so that hit points have negative, small and large coordinates (the usize saturation at 0)
    ball = b.sphere((1.0, -0.5, 2.0), 3.0, marble)
    sd = b.finish(b.list([ball]), b.list([]), background=(1.0, 1.0, 1.0))
    cam = rt.camera_new((1.0, -0.5, -9.0), (1.0, -0.5, 2.0), (0, 1, 0), 30.0, 1.0, 0.0, 9.0)
    W = H = 48
    ids = np.random.default_rng(3)
    n = 600
    px, py, s = ids.integers(8, 40, n, dtype=np.uint32), ids.integers(8, 40, n, dtype=np.uint32), np.arange(n, dtype=np.uint32)
    opts = rt.render_opts(seed=21, integrator=rt.INTEGRATOR_LEGACY)
    rays = orc.camera_rays(cam, W, H, opts, px, py, s)
    tables = ([tuple(v) for v in ranvec], [int(x) for x in perms[0]], [int(x) for x in perms[1]], [int(x) for x in perms[2]])
    for name, sc in both(rt, orc, toh, sd):
        hits = sc.trace_first_hit(rays)
        assert (hits["node"] >= 0).all(), name
        assert (hits["position"].min(axis=0) < -0.5).all() and (hits["position"].max(axis=0) > 1).any()
        rgb, seg = sc.path_radiance(cam, W, H, 50, opts, px, py, s)
        want = np.array([_noise_texture_rs(p, scale, tables) for p in hits["position"]])
        assert (seg == 2).all(), name
        assert np.abs(rgb - want[:, None]).max() < 1e-13, (name, np.abs(rgb - want[:, None]).max())
        assert want.min() >= 0.0 and want.max() <= 1.0 and want.std() > 0.05  # it is marble, not a constant


def test_legacy_lambertian_scatter_is_cosine_distributed(rt, orc, toh):
    """mat.rs:213-223: normal + unit vector is a cosine-weighted direction.  Under the legacy integrator a floor of
    albedo rho below a square lamp of radiance 1 (black sky) returns rho when the scattered ray reaches the lamp and 0
    otherwise, so the mean is rho times the cosine-weighted solid angle of the lamp over pi - the form factor again,
    this time reached by sampling the material instead of the light."""
    A = rt._abi
    rho, h, r = 0.7, 1.0, 2.0  # a disc-like square light straight above the shaded point
    b = rt.SceneBuilder()
    floor = b.rect(A.PLANE_XZ, -500, 500, -500, 500, 0.0, b.lambertian(b.constant_texture((rho, rho, rho))))
    lamp = b.flip(b.rect(A.PLANE_XZ, -r, r, -r, r, h, b.diffuse_light(b.constant_texture((1.0, 1.0, 1.0)))))
    sd = b.finish(b.list([floor, lamp]), b.list([lamp]))
    cam = rt.camera_new((6.0, 0.5, 0.0), (0.0, 0.0, 0.0), (0, 1, 0), 0.05, 1.0, 0.0, 6.0)
    n = 80000
    px, py, smp = centre_paths(n, 3, 3)
    expect = rho * rect_form_factor(2 * r, 2 * r, h)
    for name, sc in both(rt, orc, toh, sd):
        rgb, seg = sc.path_radiance(cam, 3, 3, 50, rt.render_opts(seed=13, integrator=rt.INTEGRATOR_LEGACY), px, py, smp)
        assert set(np.unique(np.round(rgb[:, 0], 12))) <= {0.0, round(rho, 12)}, name  # hit the lamp or the black sky
        mean, sem = rgb[:, 0].mean(), rgb[:, 0].std() / math.sqrt(n)
        print(name, "L = %.5f +- %.5f, closed form %.5f" % (mean, sem, expect))
        assert abs(mean - expect) < 4.0 * sem, name


def test_glass_reflectance_at_normal_incidence_is_schlick_r0(rt, orc, toh):
    """mat.rs:343-374: a ray along the radius of a glass sphere meets the first interface at normal incidence and is
    reflected with probability R0 = ((1 - n) / (1 + n))^2 (Schlick at cos = 1, mat.rs:303-307).  A reflected path is
    the one with two segments (hit, then the sky); every path returns 1 in a white environment (attenuation 1)."""
    ior = 1.5
    b = rt.SceneBuilder()
    ball = b.sphere((0, 0, 0), 1.0, b.dielectric(ior))
    lamp = b.flip(b.rect(rt._abi.PLANE_XZ, -0.1, 0.1, -0.1, 0.1, 50.0, b.diffuse_light(b.constant_texture((1, 1, 1)))))
    sd = b.finish(b.list([ball, lamp]), b.list([lamp]), background=(1.0, 1.0, 1.0))
    cam = rt.camera_new((0.0, 0.0, -6.0), (0.0, 0.0, 0.0), (0, 1, 0), 0.01, 1.0, 0.0, 6.0)
    n = 60000
    px, py, smp = centre_paths(n, 3, 3)
    r0 = ((1.0 - ior) / (1.0 + ior)) ** 2
    for name, sc in both(rt, orc, toh, sd):
        rgb, seg = sc.path_radiance(cam, 3, 3, 100, rt.render_opts(seed=19, integrator=rt.INTEGRATOR_HEAD), px, py, smp)
        reflected = (seg == 2).mean()
        sem = math.sqrt(r0 * (1 - r0) / n)
        print(name, "first-interface reflectance %.5f +- %.5f, Schlick R0 %.5f" % (reflected, sem, r0))
        assert abs(reflected - r0) < 4.0 * sem, name
        assert np.allclose(rgb, 1.0, atol=1e-12), name  # attenuation 1 everywhere, white environment
