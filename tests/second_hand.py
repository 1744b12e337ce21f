"""A second, independent restatement of the Cornell-box path in plain Python floats (IEEE f64, nothing fused) - TEST
INFRASTRUCTURE, written from the reference's sources alone, not from oracle/oracle.cpp:

  ray_color                         src/main.rs:41-120   (HEAD integrator, all arms a Cornell scene reaches)
  HittableList / FlipNormal         src/hit.rs:58-133
  AARect (hit, pdf_value, random)   src/rect.rs:26-111
  Cube                              src/cube.rs:14-37
  Translate / Rotate                src/translate.rs:21-30, src/rotate.rs:31-106 (incl. the face re-orientation with
                                    the ROTATED ray)
  ConstantMedium                    src/medium.rs:26-61 (both boundary queries and the clamping of hit1 / hit2)
  Lambertian / Metal / DiffuseLight / Isotropic   src/mat.rs:212-422
  PDF::{Cosine, Hittable, Mixture}, random_cosine_direction   src/pdf.rs:8-18, 62-176
  ONB                               src/onb.rs:8-37
  Camera::new / get_ray             src/camera.rs:19-59
  the sample closure                src/main.rs:811-820
  cornell_box / cornell_box_with_smoke and their cameras      src/main.rs:278-346, 700-705, 714-719

r2-q adds the rest of what configs 1 and 5 reach, and a reader of the RtSceneDesc so that those scenes need not be
built twice: Sphere / MovingSphere (src/sphere.rs:11-25,38-94,122-188), Triangle (src/tri.rs:24-57), Dielectric
(src/mat.rs:303-374), fuzzy Metal, CheckTexture / NoiseTexture over Perlin / ImageTexture (src/texture.rs:45-121,
src/perlin.rs:39-121), the legacy integrator (the "old method" of
src/main.rs:84-85 with Material::scatter, mat.rs:213-223,269-278,317-341,418-421).  A BVH node is read as the list of
its members (bvh.rs only culls; the oracle's own tests assert BVH == list on these scenes).

The one thing it shares with the oracle is the convention that replaces thread_rng: Philox4x32-10,
restated here as well, addressed by (pixel, sample | bounce, slot, sub, seed).  The oracle's per-path radiance must
equal this file's to rounding (tests/test_oracle_second_hand.py) - an anchor for the mixture weighting, the Rotate
quirk and the medium clamping that does not come from the first restatement's author reading his own code.
"""
import math

INF = float("inf")
F64_MAX = 1.7976931348623157e308

SLOT_PIXEL, SLOT_LENS, SLOT_TIME, SLOT_MEDIUM, SLOT_SCATTER, SLOT_BALL = range(6)


# ---------------------------------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al. 2011; the constants are those of Random123 / cuRAND) and the draw convention
# ---------------------------------------------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def u53(hi, lo):
    return float(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0)


class Draws:
    """The random numbers of one path: key (pixel, sample), counter (bounce, slot, sub, seed)."""

    def __init__(self, seed, pixel, sample):
        self.seed, self.pixel, self.sample = seed, pixel, sample

    def draw(self, bounce, slot, sub):
        w = philox4x32_10((bounce, slot, sub, self.seed), (self.pixel, self.sample))
        return u53(w[0], w[1]), u53(w[2], w[3]), w[1] & 0x7FF, w[3] & 0x7FF


# ---------------------------------------------------------------------------------------------------------------
# Vec3 (src/vec.rs) on tuples
# ---------------------------------------------------------------------------------------------------------------
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def mul(a, s): return (a[0] * s, a[1] * s, a[2] * s)          # Vec3 * f64
def smul(s, a): return (s * a[0], s * a[1], s * a[2])         # f64 * Vec3
def vmul(a, b): return (a[0] * b[0], a[1] * b[1], a[2] * b[2])
def fdiv(x, y):
    """IEEE division (Python raises where f64 gives inf or NaN; the PBR material reaches 0/0, main.rs:104)."""
    try:
        return x / y
    except ZeroDivisionError:
        if x != x or x == 0.0:
            return float("nan")
        return math.copysign(float("inf"), x) * math.copysign(1.0, y)


def fsqrt(x):  # f64::sqrt of a negative number is NaN
    return math.sqrt(x) if x >= 0.0 else float("nan")


def div(a, s): return (fdiv(a[0], s), fdiv(a[1], s), fdiv(a[2], s))
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def length(a): return math.sqrt(dot(a, a))
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def normalized(a): return div(a, length(a))
def reflect(v, n): return add(v, smul(-dot(v, n) * 2.0, n))   # vec.rs:112-114: self + (-self.dot(n) * 2.0 * n)


def powi(x, n):
    """f64::powi: square-and-multiply (compiler-rt __powidf2 / LLVM's expansion), not libm pow - Python's `x ** 2` goes
    through pow(), which is within an ulp of x * x but not always equal to it."""
    r = 1.0
    while True:
        if n & 1:
            r *= x
        n >>= 1
        if n == 0:
            return r
        x *= x


class Ray:
    def __init__(self, o, d, time):
        self.o, self.d, self.time = o, d, time

    def at(self, t):
        return add(self.o, smul(t, self.d))


class Hit:
    __slots__ = ("p", "normal", "t", "u", "v", "front_face", "material")

    def set_face_normal(self, r, outward):  # hit.rs:34-41
        self.front_face = dot(r.d, outward) < 0.0
        self.normal = outward if self.front_face else smul(-1.0, outward)


# ---------------------------------------------------------------------------------------------------------------
# Hittables
# ---------------------------------------------------------------------------------------------------------------
PLANE_AXES = {"YZ": (0, 1, 2), "XZ": (1, 0, 2), "XY": (2, 0, 1)}  # rect.rs:26-32: (k, a, b)


class AARect:
    def __init__(self, plane, a0, a1, b0, b1, k, material):
        self.axes, self.a0, self.a1, self.b0, self.b1, self.k, self.material = PLANE_AXES[plane], a0, a1, b0, b1, k, material

    def hit(self, r, t_min, t_max, ctx):  # rect.rs:49-78
        ki, ai, bi = self.axes
        t = (self.k - r.o[ki]) / r.d[ki]
        if t < t_min or t > t_max:
            return None
        a = r.o[ai] + t * r.d[ai]
        b = r.o[bi] + t * r.d[bi]
        if a < self.a0 or a > self.a1 or b < self.b0 or b > self.b1:
            return None
        h = Hit()
        h.u = (a - self.a0) / (self.a1 - self.a0)
        h.v = (b - self.b0) / (self.b1 - self.b0)
        h.p = r.at(t)
        h.t = t
        n = [0.0, 0.0, 0.0]
        n[ki] = 1.0
        h.material = self.material
        h.set_face_normal(r, tuple(n))
        return h

    def pdf_value(self, o, v):  # rect.rs:91-101
        rec = self.hit(Ray(o, v, 0.0), 0.001, INF, None)
        if rec is None:
            return 0.0
        area = (self.a1 - self.a0) * (self.b1 - self.b0)
        distance_squared = powi(rec.t, 2) * powi(length(v), 2)
        cosine = abs(dot(v, rec.normal)) / length(v)
        return distance_squared / (cosine * area) if cosine != 0.0 else 0.0

    def random(self, o, r1, r2):  # rect.rs:103-111; gen_range(lo..hi) = lo + (hi - lo) * u
        ki, ai, bi = self.axes
        p = [0.0, 0.0, 0.0]
        p[ai] = self.a0 + (self.a1 - self.a0) * r1
        p[bi] = self.b0 + (self.b1 - self.b0) * r2
        p[ki] = self.k
        return sub(tuple(p), o)


class HittableList:
    def __init__(self, items=None):
        self.list = list(items or [])

    def push(self, h):
        self.list.append(h)

    def hit(self, r, t_min, t_max, ctx):  # hit.rs:58-72
        best, closest = None, t_max
        for obj in self.list:
            rec = obj.hit(r, t_min, closest, ctx)
            if rec is not None:
                closest, best = rec.t, rec
        return best

    def pdf_value(self, o, v):  # hit.rs:90-92
        return sum(h.pdf_value(o, v) for h in self.list) / float(len(self.list))


def cube(pmin, pmax, material):  # cube.rs:14-30: six rects in this order
    return HittableList([
        AARect("XY", pmin[0], pmax[0], pmin[1], pmax[1], pmax[2], material),
        AARect("XY", pmin[0], pmax[0], pmin[1], pmax[1], pmin[2], material),
        AARect("XZ", pmin[0], pmax[0], pmin[2], pmax[2], pmax[1], material),
        AARect("XZ", pmin[0], pmax[0], pmin[2], pmax[2], pmin[1], material),
        AARect("YZ", pmin[1], pmax[1], pmin[2], pmax[2], pmax[0], material),
        AARect("YZ", pmin[1], pmax[1], pmin[2], pmax[2], pmin[0], material)])


class FlipNormal:  # hit.rs:112-133
    def __init__(self, inner):
        self.inner = inner

    def hit(self, r, t_min, t_max, ctx):
        rec = self.inner.hit(r, t_min, t_max, ctx)
        if rec is not None:
            rec.front_face = not rec.front_face
        return rec

    def pdf_value(self, o, v):
        return self.inner.pdf_value(o, v)

    def random(self, o, r1, r2):
        return self.inner.random(o, r1, r2)


class Translate:  # translate.rs:21-30
    def __init__(self, inner, offset):
        self.inner, self.offset = inner, offset

    def hit(self, r, t_min, t_max, ctx):
        rec = self.inner.hit(Ray(sub(r.o, self.offset), r.d, r.time), t_min, t_max, ctx)
        if rec is not None:
            rec.p = add(rec.p, self.offset)
        return rec


ROT_AXES = {"X": (0, 1, 2), "Y": (1, 0, 2), "Z": (2, 0, 1)}  # rotate.rs:15-21: (r, a, b)


class Rotate:
    def __init__(self, axis, inner, angle):  # rotate.rs:33-37
        self.axes, self.inner = ROT_AXES[axis], inner
        radians = (math.pi / 180.0) * angle
        self.sin_theta, self.cos_theta = math.sin(radians), math.cos(radians)

    def hit(self, r, t_min, t_max, ctx):  # rotate.rs:77-106
        _, a, b = self.axes
        c, s = self.cos_theta, self.sin_theta
        o, d = list(r.o), list(r.d)
        o[a] = c * r.o[a] - s * r.o[b]
        o[b] = s * r.o[a] + c * r.o[b]
        d[a] = c * r.d[a] - s * r.d[b]
        d[b] = s * r.d[a] + c * r.d[b]
        rotated = Ray(tuple(o), tuple(d), r.time)
        rec = self.inner.hit(rotated, t_min, t_max, ctx)
        if rec is None:
            return None
        p, n = list(rec.p), list(rec.normal)
        p[a] = c * rec.p[a] + s * rec.p[b]
        p[b] = -s * rec.p[a] + c * rec.p[b]
        n[a] = c * rec.normal[a] + s * rec.normal[b]
        n[b] = -s * rec.normal[a] + c * rec.normal[b]
        rec.p = tuple(p)
        rec.set_face_normal(rotated, tuple(n))  # with the ROTATED ray, as written (§Q3)
        return rec


class ConstantMedium:  # medium.rs:26-61
    def __init__(self, boundary, density, phase_material, draw_sub):
        self.boundary, self.density, self.phase, self.draw_sub = boundary, density, phase_material, draw_sub

    def hit(self, r, t_min, t_max, ctx):
        hit1 = self.boundary.hit(r, -F64_MAX, F64_MAX, ctx)
        if hit1 is None:
            return None
        hit2 = self.boundary.hit(r, hit1.t + 0.0001, F64_MAX, ctx)
        if hit2 is None:
            return None
        t1, t2 = hit1.t, hit2.t
        if t1 < t_min:
            t1 = t_min
        if t2 > t_max:
            t2 = t_max
        if not t1 < t2:
            return None
        dist_inside = (t2 - t1) * length(r.d)
        draws, bounce = ctx
        xi = draws.draw(bounce, SLOT_MEDIUM, self.draw_sub)[0]
        hit_distance = -(1.0 / self.density) * math.log(xi)
        if not hit_distance < dist_inside:
            return None
        h = Hit()
        h.t = t1 + hit_distance / length(r.d)
        h.p = r.at(h.t)
        h.u = h.v = 0.0
        h.front_face = False
        h.normal = (1.0, 0.0, 0.0)
        h.material = self.phase
        return h


# ---------------------------------------------------------------------------------------------------------------
# Materials: ("lambertian", albedo) | ("metal", albedo, fuzz) | ("light", emit) | ("isotropic", albedo)
# ---------------------------------------------------------------------------------------------------------------
def onb_from_w(n):  # onb.rs:8-21
    w = normalized(n)
    a = (0.0, 1.0, 0.0) if abs(w[0]) > 0.9 else (1.0, 0.0, 0.0)
    v = normalized(cross(w, a))
    u = cross(w, v)
    return u, v, w


def onb_local(uvw, a):  # onb.rs:35-37
    u, v, w = uvw
    return add(add(smul(a[0], u), smul(a[1], v)), smul(a[2], w))


def random_cosine_direction(r1, r2):  # pdf.rs:8-18
    z = math.sqrt(1.0 - r2)
    phi = 2.0 * math.pi * r1
    return (math.cos(phi) * math.sqrt(r2), math.sin(phi) * math.sqrt(r2), z)


def ray_color(ray, background, world, lights, depth, draws, max_depth):  # main.rs:41-120
    if depth <= 0:
        return (0.0, 0.0, 0.0)
    bounce = max_depth - depth
    rec = world.hit(ray, 0.00001, INF, (draws, bounce))
    if rec is None:
        return background
    m = rec.material
    kind = m[0]
    emitted = (m[1] if rec.front_face else (0.0, 0.0, 0.0)) if kind == "light" else (0.0, 0.0, 0.0)  # mat.rs:395-401
    if kind == "metal":  # mat.rs:280-293: ScatterRecord::Specular, or None when the ray ends up below the surface
        reflected = normalized(reflect(ray.d, rec.normal))
        assert m[2] == 0.0, "fuzz > 0 draws random_in_unit_sphere: not needed for the Cornell scenes"
        scattered = Ray(rec.p, reflected, ray.time)
        if dot(scattered.d, rec.normal) > 0.0:
            return vmul(m[1], ray_color(scattered, background, world, lights, depth - 1, draws, max_depth))
        return emitted
    if kind == "lambertian":  # mat.rs:232-249 + main.rs:92-98
        uvw = onb_from_w(rec.normal)                       # PDF::cosine_pdf(rec.normal)
        a, b, bits_a, bits_b = draws.draw(bounce, SLOT_SCATTER, 0)
        if bits_a & 1:                                     # pdf.rs:169: gen::<bool>() -> p0 = the hittable pdf
            light = lights.list[(bits_b * len(lights.list)) >> 11]   # hit.rs:94-96: choose
            direction = light.random(rec.p, a, b)
        else:
            direction = onb_local(uvw, random_cosine_direction(a, b))
        scattered = Ray(rec.p, direction, ray.time)
        cosine = dot(normalized(direction), uvw[2])        # pdf.rs:131-139
        cosine_pdf = cosine / math.pi if cosine > 0.0 else 0.0
        pdf_value = 0.5 * lights.pdf_value(rec.p, direction) + 0.5 * cosine_pdf   # pdf.rs:143-145
        scattering_pdf = max(dot(rec.normal, normalized(scattered.d)), 0.0) / math.pi   # mat.rs:246-249
        nxt = ray_color(scattered, background, world, lights, depth - 1, draws, max_depth)
        # emitted + attenuation * scattering_pdf * ray_color(..) / pdf_value, left to right (main.rs:97)
        return add(emitted, div(vmul(mul(m[1], scattering_pdf), nxt), pdf_value))
    # DiffuseLight and Isotropic have no scatter_mc_method (trait default None, mat.rs:60-62): emitted
    return emitted


# ---------------------------------------------------------------------------------------------------------------
# Camera (camera.rs:19-59) and the sample closure (main.rs:811-820)
# ---------------------------------------------------------------------------------------------------------------
class Camera:
    def __init__(self, lookfrom, lookat, vup, vfov, aspect_ratio, aperture, focus_dist, time0, time1):
        theta = math.pi / 180.0 * vfov
        viewport_height = 2.0 * math.tan(theta / 2.0)
        viewport_width = viewport_height * aspect_ratio
        cw = normalized(sub(lookfrom, lookat))
        cu = normalized(cross(vup, cw))
        cv = cross(cw, cu)
        h = smul(focus_dist * viewport_width, cu)
        v = smul(focus_dist * viewport_height, cv)
        self.origin, self.horizontal, self.vertical = lookfrom, h, v
        self.llc = sub(sub(sub(lookfrom, div(h, 2.0)), div(v, 2.0)), smul(focus_dist, cw))
        self.cu, self.cv, self.lens_radius, self.time0, self.time1 = cu, cv, aperture / 2.0, time0, time1

    def get_ray(self, s, t, draws):
        it = 0
        while True:  # vec.rs:96-105
            a, b, _, _ = draws.draw(0, SLOT_LENS, it)
            p = (-1.0 + (1.0 - -1.0) * a, -1.0 + (1.0 - -1.0) * b, 0.0)
            if length(p) < 1.0:
                break
            it += 1
        rd = smul(self.lens_radius, p)
        offset = add(mul(self.cu, rd[0]), mul(self.cv, rd[1]))
        time = self.time0 + draws.draw(0, SLOT_TIME, 0)[0] * (self.time1 - self.time0)
        o = add(self.origin, offset)
        return Ray(o, sub(add(add(self.llc, smul(s, self.horizontal)), smul(t, self.vertical)), o), time)


def path_radiance(scene, width, height, max_depth, seed, i, j, sample):
    """ray_color of sample `sample` of pixel (i, j), j counted bottom-up as in main.rs:772-777."""
    world, lights, background, camera = scene
    draws = Draws(seed, j * width + i, sample)
    ru, rv, _, _ = draws.draw(0, SLOT_PIXEL, 0)
    u = (float(i) + ru) / float(width - 1)
    v = (float(j) + rv) / float(height - 1)
    return ray_color(camera.get_ray(u, v, draws), background, world, lights, max_depth, draws, max_depth)


# ---------------------------------------------------------------------------------------------------------------
# Scenes: main.rs:278-311 and :313-346, cameras :700-705 and :714-719
# ---------------------------------------------------------------------------------------------------------------
def _room():
    red, white, green = ("lambertian", (0.65, 0.05, 0.05)), ("lambertian", (0.73, 0.73, 0.73)), ("lambertian", (0.12, 0.45, 0.15))
    light = ("light", (15.0, 15.0, 15.0))
    rect_light = FlipNormal(AARect("XZ", 213.0, 343.0, 227.0, 332.0, 554.0, light))
    world = HittableList()
    world.push(AARect("YZ", 0.0, 555.0, 0.0, 555.0, 555.0, green))
    world.push(AARect("YZ", 0.0, 555.0, 0.0, 555.0, 0.0, red))
    world.push(rect_light)
    world.push(AARect("XZ", 0.0, 555.0, 0.0, 555.0, 0.0, white))
    world.push(AARect("XZ", 0.0, 555.0, 0.0, 555.0, 555.0, white))
    world.push(AARect("XY", 0.0, 555.0, 0.0, 555.0, 555.0, white))
    return world, HittableList([rect_light]), white


def _camera():
    return Camera((278.0, 278.0, -800.0), (278.0, 278.0, 0.0), (0.0, 1.0, 0.0), 40.0, 1.0, 0.05, 10.0, 0.0, 1.0)


def cornell_box():
    world, lights, white = _room()
    metal = ("metal", (0.8, 0.85, 0.88), 0.0)
    world.push(Translate(Rotate("Y", cube((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white), -18.0), (130.0, 0.0, 65.0)))
    world.push(Translate(Rotate("Y", cube((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), metal), 15.0), (265.0, 0.0, 295.0)))
    return world, lights, (0.0, 0.0, 0.0), _camera()


def cornell_box_with_smoke(medium_draw_subs):
    """medium_draw_subs: the two sub-slot numbers of the media's free-path draws (the medium's node
    index in the scene description; a property of the convention, not of the reference)."""
    world, lights, white = _room()
    box1 = Translate(Rotate("Y", cube((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white), -18.0), (130.0, 0.0, 65.0))
    box2 = Translate(Rotate("Y", cube((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), white), 15.0), (265.0, 0.0, 295.0))
    world.push(ConstantMedium(box1, 0.01, ("isotropic", (1.0, 1.0, 1.0)), medium_draw_subs[0]))
    world.push(ConstantMedium(box2, 0.01, ("isotropic", (0.0, 0.0, 0.0)), medium_draw_subs[1]))
    return world, lights, (0.0, 0.0, 0.0), _camera()


# ---------------------------------------------------------------------------------------------------------------
# r2-q: spheres, triangles, glass, fuzzy metal, checker textures, the legacy integrator, scenes read from an RtSceneDesc
# ---------------------------------------------------------------------------------------------------------------
def get_sphere_uv(p):  # sphere.rs:11-25
    phi = math.atan2(-p[2], p[0]) + math.pi
    theta = math.acos(-p[1])
    return phi / (2.0 * math.pi), theta / math.pi


class Sphere:
    def __init__(self, center, radius, material):
        self.c, self.radius, self.material = center, radius, material

    def center(self, time):
        return self.c

    def hit(self, r, t_min, t_max, ctx):  # sphere.rs:56-94 (== :150-188 with center(time))
        center = self.center(r.time)
        oc = sub(r.o, center)
        a = powi(length(r.d), 2)
        half_b = dot(oc, r.d)
        c = powi(length(oc), 2) - powi(self.radius, 2)
        discriminant = powi(half_b, 2) - a * c
        if discriminant < 0.0:
            return None
        sqrt_d = math.sqrt(discriminant)
        root = (-half_b - sqrt_d) / a
        if root < t_min or root > t_max:
            root = (-half_b + sqrt_d) / a
            if root < t_min or root > t_max:
                return None
        h = Hit()
        h.t = root
        h.p = r.at(root)
        h.material = self.material
        outward = div(sub(h.p, center), self.radius)
        h.set_face_normal(r, outward)
        h.u, h.v = get_sphere_uv(outward)
        return h


def _sphere_pdf_value(self, o, v):  # sphere.rs:103-111
    if self.hit(Ray(o, v, 0.0), 0.001, 1.7976931348623157e308, None) is None:
        return 0.0
    cos_theta_max = fsqrt(1.0 - fdiv(powi(self.radius, 2), powi(length(sub(self.c, o)), 2)))
    solid_angle = 2.0 * math.pi * (1.0 - cos_theta_max)
    return fdiv(1.0, solid_angle)


def _sphere_random(self, o, r1, r2):  # sphere.rs:113-118 over random_to_sphere, :27-36
    direction = sub(self.c, o)
    distance_squared = powi(length(direction), 2)
    uvw = onb_from_w(direction)
    z = 1.0 + r2 * (fsqrt(1.0 - fdiv(powi(self.radius, 2), distance_squared)) - 1.0)
    phi = 2.0 * math.pi * r1
    x = math.cos(phi) * fsqrt(1.0 - powi(z, 2))
    y = math.sin(phi) * fsqrt(1.0 - powi(z, 2))
    return onb_local(uvw, (x, y, z))


Sphere.pdf_value = _sphere_pdf_value
Sphere.random = _sphere_random


class MovingSphere(Sphere):
    def __init__(self, c0, c1, t0, t1, radius, material):
        self.c0, self.c1, self.t0, self.t1, self.radius, self.material = c0, c1, t0, t1, radius, material

    def center(self, time):  # sphere.rs:144-146
        return add(self.c0, smul((time - self.t0) / (self.t1 - self.t0), sub(self.c1, self.c0)))


class Triangle:
    def __init__(self, v0, v1, v2, material):
        self.v, self.material = (v0, v1, v2), material

    def hit(self, r, t_min, t_max, ctx):  # tri.rs:24-57
        s = sub(r.o, self.v[0])
        e1 = sub(self.v[1], self.v[0])
        e2 = sub(self.v[2], self.v[0])
        s1 = cross(r.d, e2)
        s2 = cross(s, e1)
        s1_e1 = dot(s1, e1)
        if s1_e1 == 0.0:
            return None  # x / 0: every comparison of the reference then fails or yields inf / NaN; not reached by these scenes' rays
        t = dot(s2, e2) / s1_e1
        b1 = dot(s1, s) / s1_e1
        b2 = dot(s2, r.d) / s1_e1
        if t < t_min or t > t_max:
            return None
        if b1 < 0.0 or b2 < 0.0 or (1.0 - b1 - b2) < 0.0:
            return None
        h = Hit()
        h.t, h.u, h.v = t, b1, b2
        h.p = r.at(t)
        h.material = self.material
        h.set_face_normal(r, normalized(cross(e1, e2)))
        return h


def as_usize(x):  # Rust `as usize`: saturating, NaN -> 0
    if x != x or x <= 0.0:
        return 0
    return int(x) if x < 18446744073709551615.0 else 0xFFFFFFFFFFFFFFFF


def perlin_noise(tables, p, scale):  # perlin.rs:77-109 and perlin_interp :39-56 (the Hermite curve is applied twice, as written)
    ranvec, perm_x, perm_y, perm_z = tables
    u = scale * p[0] - math.floor(scale * p[0])
    v = scale * p[1] - math.floor(scale * p[1])
    w = scale * p[2] - math.floor(scale * p[2])
    u = u * u * (3.0 - 2.0 * u)
    v = v * v * (3.0 - 2.0 * v)
    w = w * w * (3.0 - 2.0 * w)
    i, j, k = as_usize(math.floor(scale * p[0])), as_usize(math.floor(scale * p[1])), as_usize(math.floor(scale * p[2]))
    uu = u * u * (3.0 - 2.0 * u)
    vv = v * v * (3.0 - 2.0 * v)
    ww = w * w * (3.0 - 2.0 * w)
    accum = 0.0
    for di in range(2):
        for dj in range(2):
            for dk in range(2):
                c = ranvec[perm_x[(i + di) & 255] ^ perm_y[(j + dj) & 255] ^ perm_z[(k + dk) & 255]]
                weight = (u - float(di), v - float(dj), w - float(dk))
                accum += (float(di) * uu + float(1 - di) * (1.0 - uu)) * (float(dj) * vv + float(1 - dj) * (1.0 - vv)) * \
                         (float(dk) * ww + float(1 - dk) * (1.0 - ww)) * dot(c, weight)
    return accum


def perlin_turb(tables, p, scale, depth):  # perlin.rs:111-121
    accum, temp_p, weight = 0.0, p, 1.0
    for _ in range(depth):
        accum += weight * perlin_noise(tables, temp_p, scale)
        weight *= 0.5
        temp_p = mul(temp_p, 2.0)
    return abs(accum)


def texture_value(tex, u, v, p):
    """tex: ("constant", color) | ("checker", odd, even) | ("noise", scale, tables) | ("image", width, height, bytes)
    (texture.rs:23-27, 45-54, 71-79, 99-121)"""
    while tex[0] == "checker":
        sines = math.sin(10.0 * p[0]) * math.sin(10.0 * p[1]) * math.sin(10.0 * p[2])
        tex = tex[1] if sines < 0.0 else tex[2]
    if tex[0] == "noise":
        return mul(mul((1.0, 1.0, 1.0), 0.5), 1.0 + math.sin(tex[1] * p[2] + 10.0 * perlin_turb(tex[2], p, tex[1], 7)))
    if tex[0] == "image":
        width, height, data = tex[1], tex[2], tex[3]
        i = as_usize(min(max(u, 0.0), 1.0) * float(width))
        j = as_usize(min(max(1.0 - v, 0.0), 1.0) * float(height))
        i = min(i, width - 1)
        j = min(j, height - 1)
        idx = 3 * i + 3 * width * j
        return (data[idx] / 255.0, data[idx + 1] / 255.0, data[idx + 2] / 255.0)
    return tex[1]


def random_in_unit_sphere(draws, bounce):  # vec.rs:70-85 with the BALL slots of DESIGN.md §3
    it = 0
    while True:
        a, b, _, _ = draws.draw(bounce, SLOT_BALL, 2 * it)
        c = draws.draw(bounce, SLOT_BALL, 2 * it + 1)[0]
        v = (-1.0 + (1.0 - -1.0) * a, -1.0 + (1.0 - -1.0) * b, -1.0 + (1.0 - -1.0) * c)
        if length(v) < 1.0:
            return v
        it += 1


def near_zero(a):  # vec.rs:107-110
    return abs(a[0]) < 1.0e-8 and abs(a[1]) < 1.0e-8 and abs(a[2]) < 1.0e-8


def refract(v, n, etai_over_etat):  # vec.rs:116-121
    cos_theta = min(dot(smul(-1.0, v), n), 1.0)
    r_out_perp = smul(etai_over_etat, add(v, smul(cos_theta, n)))
    r_out_para = smul(-1.0 * math.sqrt(abs(1.0 - powi(length(r_out_perp), 2))), n)
    return add(r_out_perp, r_out_para)


def dielectric_direction(ir, r_in, rec, draws, bounce):  # mat.rs:317-341 == :343-366
    refraction_ratio = 1.0 / ir if rec.front_face else ir
    unit_direction = normalized(r_in.d)
    cos_theta = min(dot(smul(-1.0, unit_direction), rec.normal), 1.0)
    sin_theta = math.sqrt(1.0 - powi(cos_theta, 2))
    cannot_refract = refraction_ratio * sin_theta > 1.0
    r0 = powi((1.0 - refraction_ratio) / (1.0 + refraction_ratio), 2)  # mat.rs:303-307
    reflectance = r0 + (1.0 - r0) * powi(1.0 - cos_theta, 5)
    will_reflect = draws.draw(bounce, SLOT_SCATTER, 0)[0] < reflectance
    if cannot_refract or will_reflect:
        return reflect(unit_direction, rec.normal)
    return refract(unit_direction, rec.normal, refraction_ratio)


def material_texture(m, rec):
    return texture_value(m[1], rec.u, rec.v, rec.p)


def scatter_legacy(m, ray, rec, draws, bounce):
    """Material::scatter (the old method): (attenuation, scattered ray) or None."""
    kind = m[0]
    if kind == "lambertian":  # mat.rs:213-223
        d = add(rec.normal, normalized(random_in_unit_sphere(draws, bounce)))
        if near_zero(d):
            d = rec.normal
        return material_texture(m, rec), Ray(rec.p, d, ray.time)
    if kind == "metal":  # mat.rs:269-278
        reflected = normalized(reflect(ray.d, rec.normal))
        d = add(reflected, smul(m[2], random_in_unit_sphere(draws, bounce))) if m[2] != 0.0 else reflected
        return (m[1], Ray(rec.p, d, ray.time)) if dot(d, rec.normal) > 0.0 else None
    if kind == "dielectric":
        return (1.0, 1.0, 1.0), Ray(rec.p, dielectric_direction(m[1], ray, rec, draws, bounce), ray.time)
    if kind == "isotropic":  # mat.rs:418-421
        return material_texture(m, rec), Ray(rec.p, random_in_unit_sphere(draws, bounce), ray.time)
    return None  # DiffuseLight (mat.rs:391-393)


def ray_color_legacy(ray, background, world, depth, draws, max_depth):  # main.rs:41-47, 84-85, 111-119
    if depth <= 0:
        return (0.0, 0.0, 0.0)
    bounce = max_depth - depth
    rec = world.hit(ray, 0.00001, INF, (draws, bounce))
    if rec is None:
        return background
    m = rec.material
    emitted = (material_texture(m, rec) if rec.front_face else (0.0, 0.0, 0.0)) if m[0] == "light" else (0.0, 0.0, 0.0)
    sc = scatter_legacy(m, ray, rec, draws, bounce)
    if sc is None:
        return emitted
    attenuation, scattered = sc
    return add(emitted, vmul(attenuation, ray_color_legacy(scattered, background, world, depth - 1, draws, max_depth)))


def _clamp01(x):
    return min(max(x, 0.0), 1.0)


def schlick_fresnel(u):  # mat.rs:10-14
    m = _clamp01(1.0 - u)
    m2 = powi(m, 2)
    return m2 * m2 * m


def gtr_1(n_dot_h, a):  # mat.rs:16-24
    if a >= 1.0:
        return 1.0 / math.pi
    a2 = a * a
    t = 1.0 + (a2 - 1.0) * n_dot_h * n_dot_h
    return fdiv(a2 - 1.0, math.pi * math.log2(a2) * t)


def gtr_2_aniso(n_dot_h, h_dot_x, h_dot_y, ax, ay):  # mat.rs:32-34
    return fdiv(1.0, math.pi * ax * ay * powi(powi(h_dot_x / ax, 2) + powi(h_dot_y / ay, 2) + n_dot_h * n_dot_h, 2))


def smith_g_ggx(n_dot_v, alpha_g):  # mat.rs:36-40
    a = alpha_g * alpha_g
    b = n_dot_v * n_dot_v
    return fdiv(1.0, n_dot_v + fsqrt(a + b - a * b))


def smith_g_ggx_aniso(n_dot_v, v_dot_x, v_dot_y, ax, ay):  # mat.rs:42-44
    return fdiv(1.0, n_dot_v + math.sqrt(powi(v_dot_x * ax, 2) + powi(v_dot_y * ay, 2) + powi(n_dot_v, 2)))


def mix(a, b, t):  # mat.rs:50-52
    return a * (1.0 - t) + b * t


def vmix(a, b, t):  # vec.rs:60-68
    return (a[0] * (1.0 - t) + b[0] * t, a[1] * (1.0 - t) + b[1] * t, a[2] * (1.0 - t) + b[2] * t)


def _powf(x, y):  # f64::powf is libm's pow; a negative base with a fractional exponent is NaN there, an exception here
    try:
        return math.pow(x, y)
    except ValueError:
        return float("nan")


def _aniso_alphas(roughness, anisotropic):  # mat.rs:170-172 == pdf.rs:43-45,121-123
    aspect = math.sqrt(1.0 - anisotropic * 0.9)
    return max(powi(roughness, 2) / aspect, 0.001), max(powi(roughness, 2) * aspect, 0.001)


def pbr_brdf(m, r_in, r_out, rec):  # mat.rs:134-198
    _, base, metallic, subsurface, specular, roughness, specular_tint, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss = m
    l = mul(normalized(r_in.d), -1.0)
    v = normalized(r_out.d)
    x, y, n = onb_from_w(rec.normal)
    n_dot_v = dot(n, v)
    n_dot_l = dot(n, l)
    if n_dot_l < 0.0 or n_dot_v < 0.0:
        return (0.0, 0.0, 0.0)
    h = normalized(add(l, v))
    n_dot_h = dot(n, h)
    l_dot_h = dot(l, h)
    c = texture_value(base, rec.u, rec.v, rec.p)
    cd_lin = (_powf(c[0], 2.2), _powf(c[1], 2.2), _powf(c[2], 2.2))  # mon_to_lin, mat.rs:46-48
    cd_lum = 0.3 * cd_lin[0] + 0.6 * cd_lin[1] + 0.1 * cd_lin[2]
    white = (1.0, 1.0, 1.0)
    c_tint = div(cd_lin, cd_lum) if cd_lum > 0.0 else white
    c_spec0 = vmix(mul(mul(vmix(white, c_tint, specular_tint), 0.08), specular), cd_lin, metallic)
    c_sheen = vmix(white, c_tint, sheen_tint)
    fresnel_l = schlick_fresnel(n_dot_l)
    fresnel_v = schlick_fresnel(n_dot_v)
    fresnel_diffuse_90 = 0.5 + 2.0 * l_dot_h * l_dot_h * roughness
    fresnel_diffuse = mix(1.0, fresnel_diffuse_90, fresnel_l) * mix(1.0, fresnel_diffuse_90, fresnel_v)
    fss90 = l_dot_h * l_dot_h * roughness
    fss = mix(1.0, fss90, fresnel_l) * mix(1.0, fss90, fresnel_v)
    subface_scatter = 1.25 * (fss * (fdiv(1.0, n_dot_l + n_dot_v) - 0.5) + 0.5)
    ax, ay = _aniso_alphas(roughness, anisotropic)
    d_specular = gtr_2_aniso(n_dot_h, dot(h, x), dot(h, y), ax, ay)
    fresnel_h = schlick_fresnel(l_dot_h)
    f_specular = vmix(c_spec0, white, fresnel_h)
    g_specular = smith_g_ggx_aniso(n_dot_l, dot(l, x), dot(l, y), ax, ay) * smith_g_ggx_aniso(n_dot_v, dot(v, x), dot(v, y), ax, ay)
    fresnel_sheen = smul(fresnel_h * sheen, c_sheen)  # f64 * f64 * Vec3, left to right
    d_reflect = gtr_1(n_dot_h, mix(0.1, 0.001, clearcoat_gloss))
    f_reflect = mix(0.04, 1.0, fresnel_h)
    g_reflect = smith_g_ggx(n_dot_l, 0.25) * smith_g_ggx(n_dot_v, 0.25)
    # ((1/pi) * mix(..) * cd_lin + sheen) * (1 - metallic) + g * f * d + 0.25 * clearcoat * g_r * f_r * d_r, as written
    diffuse = add(smul((1.0 / math.pi) * mix(fresnel_diffuse, subface_scatter, subsurface), cd_lin), fresnel_sheen)
    spec = mul(smul(g_specular, f_specular), d_specular)
    coat = mul(mul(mul(mul((0.25, 0.25, 0.25), clearcoat), g_reflect), f_reflect), d_reflect)
    return add(add(mul(diffuse, 1.0 - metallic), spec), coat)


def brdf_pdf_value(uvw, r_in_d, m, r_out):  # pdf.rs:103-129
    roughness, anisotropic, clearcoat_gloss = m[5], m[7], m[11]
    cosine = dot(normalized(r_out), uvw[2])
    if cosine <= 0.0:
        return 0.0
    diffuse_pdf = cosine / math.pi
    l = mul(normalized(r_in_d), -1.0)
    v = normalized(r_out)
    x, y, n = uvw
    n_dot_l = dot(n, l)
    h = normalized(add(l, v))
    n_dot_h = dot(n, h)
    if n_dot_h <= 0.0:
        return 0.0
    ax, ay = _aniso_alphas(roughness, anisotropic)
    specular_pdf = fdiv(gtr_2_aniso(n_dot_h, dot(h, x), dot(h, y), ax, ay) * abs(n_dot_h) * 0.25, n_dot_l)
    clearcoat_pdf = fdiv(gtr_1(n_dot_h, mix(0.1, 0.001, clearcoat_gloss)) * abs(n_dot_h) * 0.25, n_dot_l)
    return (diffuse_pdf + specular_pdf + clearcoat_pdf) / 3.0


def brdf_pdf_generate(uvw, r_in_d, m, selector, r1, r2):  # pdf.rs:152-161 over :20-63
    roughness, anisotropic, clearcoat_gloss = m[5], m[7], m[11]
    if selector < 0.333:
        return onb_local(uvw, random_cosine_direction(r1, r2))
    if selector < 0.666:  # GTR_1_direction
        a = mix(0.1, 0.001, clearcoat_gloss)
        a2 = a * a
        cos_theta = math.sqrt(max(0.001, (1.0 - _powf(a2, 1.0 - r1)) / (1.0 - a2)))
        sin_theta = math.sqrt(max(0.001, 1.0 - cos_theta * cos_theta))
        phi = math.pi * 2.0 * r2
        wh = (sin_theta * math.cos(phi), sin_theta * math.sin(phi), cos_theta)  # spherical_direction(.., sin_phi, cos_phi)
        return onb_local(uvw, reflect(r_in_d, wh))
    ax, ay = _aniso_alphas(roughness, anisotropic)  # GTR_2_aniso_direction
    phi = math.atan(ay / ax * math.tan(2.0 * math.pi * r2 + 0.5 * math.pi))
    if r2 > 0.5:
        phi += math.pi
    sin_phi, cos_phi = math.sin(phi), math.cos(phi)
    ax_2, ay_2 = ax * ax, ay * ay
    a2 = fdiv(1.0, cos_phi * cos_phi / ax_2 + sin_phi * sin_phi / ay_2)
    tan_theta_2 = fdiv(a2 * r1, 1.0 - r1)
    cos_theta = fdiv(1.0, fsqrt(1.0 + tan_theta_2))
    sin_theta = math.sqrt(max(0.001, 1.0 - cos_theta * cos_theta))
    wh = (sin_theta * math.cos(phi), sin_theta * math.sin(phi), cos_theta)
    return onb_local(uvw, reflect(r_in_d, wh))


def ray_color_general(ray, background, world, lights, depth, draws, max_depth):
    """main.rs:41-120 for every material of configs 1-3 and 5 (the Cornell-only ray_color above plus Dielectric,
    fuzzy Metal and textured Lambertian / DiffuseLight)."""
    if depth <= 0:
        return (0.0, 0.0, 0.0)
    bounce = max_depth - depth
    rec = world.hit(ray, 0.00001, INF, (draws, bounce))
    if rec is None:
        return background
    m = rec.material
    kind = m[0]
    emitted = (material_texture(m, rec) if rec.front_face else (0.0, 0.0, 0.0)) if kind == "light" else (0.0, 0.0, 0.0)
    if kind in ("metal", "dielectric"):  # ScatterRecord::Specular (mat.rs:280-293, :343-374)
        sc = scatter_legacy(m, ray, rec, draws, bounce)
        if sc is None:
            return emitted
        return vmul(sc[0], ray_color_general(sc[1], background, world, lights, depth - 1, draws, max_depth))
    if kind == "lambertian":
        uvw = onb_from_w(rec.normal)
        a, b, bits_a, bits_b = draws.draw(bounce, SLOT_SCATTER, 0)
        if bits_a & 1:
            direction = lights.list[(bits_b * len(lights.list)) >> 11].random(rec.p, a, b)
        else:
            direction = onb_local(uvw, random_cosine_direction(a, b))
        scattered = Ray(rec.p, direction, ray.time)
        cosine = dot(normalized(direction), uvw[2])
        cosine_pdf = cosine / math.pi if cosine > 0.0 else 0.0
        pdf_value = 0.5 * lights.pdf_value(rec.p, direction) + 0.5 * cosine_pdf
        scattering_pdf = max(dot(rec.normal, normalized(scattered.d)), 0.0) / math.pi
        nxt = ray_color_general(scattered, background, world, lights, depth - 1, draws, max_depth)
        return add(emitted, div(vmul(mul(material_texture(m, rec), scattering_pdf), nxt), pdf_value))
    if kind == "pbr":  # ScatterRecord::Microfacet, main.rs:99-105 over mat.rs:119-132
        uvw = onb_from_w(rec.normal)
        a, b, bits_a, bits_b = draws.draw(bounce, SLOT_SCATTER, 0)
        if bits_a & 1:
            direction = lights.list[(bits_b * len(lights.list)) >> 11].random(rec.p, a, b)
        else:
            direction = brdf_pdf_generate(uvw, ray.d, m, draws.draw(bounce, SLOT_SCATTER, 1)[0], a, b)
        scattered = Ray(rec.p, direction, ray.time)
        pdf_value = 0.5 * lights.pdf_value(rec.p, direction) + 0.5 * brdf_pdf_value(uvw, ray.d, m, direction)
        nxt = ray_color_general(scattered, background, world, lights, depth - 1, draws, max_depth)
        return add(emitted, div(vmul(pbr_brdf(m, ray, scattered, rec), nxt), pdf_value))
    return emitted


def scene_from_desc(abi, desc):
    """The object graph of an RtSceneDesc (ctypes struct), in the classes of this file."""
    def tex(i):
        t = desc.textures[i]
        if t.kind == abi.TEX_CONSTANT:
            return ("constant", tuple(t.color))
        if t.kind == abi.TEX_CHECKER:
            return ("checker", tex(t.a), tex(t.b))
        if t.kind == abi.TEX_NOISE:
            pt = desc.perlin[t.a]
            ranvec = [(pt.ranvec[3 * k], pt.ranvec[3 * k + 1], pt.ranvec[3 * k + 2]) for k in range(256)]
            return ("noise", t.scale, (ranvec, list(pt.perm_x), list(pt.perm_y), list(pt.perm_z)))
        if t.kind == abi.TEX_IMAGE:
            im = desc.images[t.a]
            n = 3 * im.width * im.height
            return ("image", im.width, im.height, bytes(bytearray(desc.texels[im.offset + k] for k in range(n))))
        raise NotImplementedError("texture kind %d" % t.kind)

    def mat(i):
        m = desc.materials[i]
        if m.kind == abi.MAT_LAMBERTIAN:
            return ("lambertian", tex(m.texture))
        if m.kind == abi.MAT_METAL:
            return ("metal", tuple(m.albedo), m.fuzz)
        if m.kind == abi.MAT_DIELECTRIC:
            return ("dielectric", m.ir)
        if m.kind == abi.MAT_DIFFUSE_LIGHT:
            return ("light", tex(m.texture))
        if m.kind == abi.MAT_ISOTROPIC:
            return ("isotropic", tex(m.texture))
        if m.kind == abi.MAT_PBR:  # rtb200.h RT_PBR_*: metallic, subsurface, specular, roughness, specular_tint, anisotropic,
            q = m.pbr              # sheen, sheen_tint, clearcoat, clearcoat_gloss
            return ("pbr", tex(m.texture)) + tuple(q[k] for k in range(10))
        raise NotImplementedError("material kind %d" % m.kind)

    plane = {abi.PLANE_YZ: "YZ", abi.PLANE_XZ: "XZ", abi.PLANE_XY: "XY"}
    axis = {abi.AXIS_X: "X", abi.AXIS_Y: "Y", abi.AXIS_Z: "Z"}

    def node(i):
        n = desc.nodes[i]
        v = n.v
        if n.kind == abi.NODE_SPHERE:
            return Sphere((v[0], v[1], v[2]), v[3], mat(n.material))
        if n.kind == abi.NODE_MOVING_SPHERE:
            return MovingSphere((v[0], v[1], v[2]), (v[3], v[4], v[5]), v[6], v[7], v[8], mat(n.material))
        if n.kind == abi.NODE_RECT:
            return AARect(plane[n.axis], v[0], v[1], v[2], v[3], v[4], mat(n.material))
        if n.kind == abi.NODE_TRIANGLE:
            return Triangle((v[0], v[1], v[2]), (v[3], v[4], v[5]), (v[6], v[7], v[8]), mat(n.material))
        if n.kind == abi.NODE_CUBE:
            return cube((v[0], v[1], v[2]), (v[3], v[4], v[5]), mat(n.material))
        if n.kind in (abi.NODE_LIST, abi.NODE_BVH):
            return HittableList([node(desc.child_index[n.child + k]) for k in range(n.count)])
        if n.kind == abi.NODE_TRANSLATE:
            return Translate(node(n.child), (v[0], v[1], v[2]))
        if n.kind == abi.NODE_ROTATE:
            return Rotate(axis[n.axis], node(n.child), v[0])
        if n.kind == abi.NODE_FLIP:
            return FlipNormal(node(n.child))
        if n.kind == abi.NODE_MEDIUM:
            return ConstantMedium(node(n.child), v[0], mat(n.material), i)
        raise NotImplementedError("node kind %d" % n.kind)

    return node(desc.world), node(desc.lights), tuple(desc.background)


class CameraPod:
    """Camera::get_ray (camera.rs:51-59) over the nine fields of an RtCamera."""

    def __init__(self, c):
        self.origin, self.llc = tuple(c.origin), tuple(c.lower_left_corner)
        self.horizontal, self.vertical, self.cu, self.cv = tuple(c.horizontal), tuple(c.vertical), tuple(c.cu), tuple(c.cv)
        self.lens_radius, self.time0, self.time1 = c.lens_radius, c.time0, c.time1

    get_ray = Camera.get_ray


def path_radiance_general(world, lights, background, camera, width, height, max_depth, seed, i, j, sample, legacy):
    draws = Draws(seed, j * width + i, sample)
    ru, rv, _, _ = draws.draw(0, SLOT_PIXEL, 0)
    u = (float(i) + ru) / float(width - 1)
    v = (float(j) + rv) / float(height - 1)
    ray = camera.get_ray(u, v, draws)
    if legacy:
        return ray_color_legacy(ray, background, world, max_depth, draws, max_depth)
    return ray_color_general(ray, background, world, lights, max_depth, draws, max_depth)
