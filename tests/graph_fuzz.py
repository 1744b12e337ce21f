"""Seeded random scene graphs for the parity tests (CPU tier: test_scene_graph_fuzz.py; GPU: test_gpu_parity.py)."""
import numpy as np


class GraphMaker:
    """A seeded random scene graph inside the cube [-10, 10]^3 with one rect light above it."""

    def __init__(self, rt, seed, rich=False):
        self.rt, self.A, self.rich = rt, rt._abi, rich
        self.has_default_light = False
        self.rng = np.random.default_rng(seed)
        self.b = rt.SceneBuilder()
        b = self.b
        grey = b.constant_texture((0.6, 0.6, 0.6))
        check = b.check_texture(b.constant_texture((0.2, 0.3, 0.1)), b.constant_texture((0.9, 0.9, 0.9)))
        self.materials = [b.lambertian(grey), b.lambertian(check), b.lambertian(b.constant_texture((0.7, 0.2, 0.2))),
                          b.metal((0.8, 0.85, 0.88), 0.0), b.metal((0.7, 0.6, 0.5), 0.3), b.dielectric(1.5)]
        if rich:
            # Perlin::new (perlin.rs:5-24): in-ball gradients, three shuffled permutations; an 8x4 image; textures
            # nested in a checker; the Disney-style material (mat.rs:86-197) with PDF::BRDF
            r = self.rng
            ranvec = r.uniform(-1, 1, (256, 3)) * r.uniform(0.2, 1.0, (256, 1))
            marble = b.noise_texture(float(r.uniform(0.1, 4.0)), ranvec, r.permutation(256), r.permutation(256), r.permutation(256))
            image = b.image_texture(r.integers(0, 256, 8 * 4 * 3, dtype=np.uint8).tobytes(), 8, 4)
            self.materials += [b.lambertian(marble), b.lambertian(image), b.lambertian(b.check_texture(marble, image)),
                               b.pbr(grey, metallic=0.3, specular=0.5, roughness=0.4, clearcoat=0.2, clearcoat_gloss=0.7),
                               b.pbr(image, metallic=0.0, subsurface=0.2, roughness=0.8, anisotropic=0.5, sheen=0.3)]
        self.light_material = b.diffuse_light(b.constant_texture((7.0, 7.0, 7.0)))

    def u(self, lo, hi, n=None):
        return self.rng.uniform(lo, hi, n)

    def material(self):
        return self.materials[int(self.rng.integers(len(self.materials)))]

    def primitive(self, under_bvh, scale=1.0):
        b, A = self.b, self.A
        c = self.u(-8, 8, 3) * scale
        kind = int(self.rng.integers(5))
        if kind == 0:
            # a negative radius is the RTiOW hollow-glass trick: the normal points inwards (sphere.rs:75-76), and the
            # inverted box c -+ r keeps the sphere out of any BVH (aabb.rs:31)
            # (below a BVH its inverted box also shrinks the union box of a list around it: the §Q5 class again)
            sign = -1.0 if self.rich and not under_bvh and self.rng.random() < 0.15 else 1.0
            return b.sphere(c, sign * self.u(0.3, 2.0) * scale, self.material())
        if kind == 1:
            # center(time) extrapolates outside [time0, time1] (sphere.rs:144-146, §Q18); camera times are in [0, 1)
            # (MovingSphere::bounding_box only covers center0 / center1, sphere.rs:191-201: below a BVH the
            # extrapolated sphere sticks out of the reference's box, so the time range stays [0, 1] there)
            t0, t1 = (0.0, 1.0) if (under_bvh or not self.rich) else (float(self.u(-0.5, 0.4)), float(self.u(0.6, 1.5)))
            return b.moving_sphere(c, c + self.u(-0.5, 0.5, 3), t0, t1, self.u(0.3, 1.5) * scale, self.material())
        if kind == 2:
            a0, b0 = self.u(-8, 6, 2) * scale
            # §Q5: AARect::bounding_box ignores the plane (rect.rs:83-89), so the reference's own BVH culls XZ / YZ
            # rects it should hit; the compiler builds correct bounds on purpose.  Below a BVH only XY
            # rects - where the reference's box is right - are comparable.
            plane = A.PLANE_XY if under_bvh else int(self.rng.integers(3))
            return b.rect(plane, a0, a0 + self.u(0.5, 4) * scale, b0, b0 + self.u(0.5, 4) * scale,
                          self.u(-8, 8) * scale, self.material())
        if kind == 3:
            return b.triangle(c, c + self.u(-3, 3, 3) * scale, c + self.u(-3, 3, 3) * scale, self.material())
        return b.cube(c, c + self.u(0.3, 3, 3) * scale, self.material())

    def wrap(self, node):
        """Zero to three wrappers in random order (translate.rs, rotate.rs, hit.rs:99-133)."""
        b = self.b
        for _ in range(int(self.rng.integers(4))):
            w = int(self.rng.integers(3))
            if w == 0:
                node = b.translate(node, self.u(-3, 3, 3))
            elif w == 1:
                node = b.rotate(int(self.rng.integers(3)), node, self.u(-180, 180))
            else:
                node = b.flip(node)
        return node

    def subtree(self, depth, under_bvh):
        b = self.b
        r = self.rng.random()
        if depth >= 3 or r < 0.35:
            return self.wrap(self.primitive(under_bvh))
        n = int(self.rng.integers(1, 9))
        is_bvh = r < 0.7
        kids = [self.subtree(depth + 1, under_bvh or is_bvh) for _ in range(n)]
        node = b.bvh(kids, 0.0, 1.0) if is_bvh else b.list(kids)
        return self.wrap(node)

    def medium(self):
        b = self.b
        c = self.u(-5, 5, 3)
        kind = int(self.rng.integers(3))
        if kind == 0:
            boundary = b.sphere(c, self.u(1.0, 3.0), self.materials[0])
        elif kind == 1:
            boundary = b.cube(c, c + self.u(1.0, 4.0, 3), self.materials[0])
        else:  # the Cornell smoke construction: Translate(Rotate(Cube)) (main.rs:330-343)
            boundary = b.translate(b.rotate(self.A.AXIS_Y, b.cube((0, 0, 0), self.u(1.0, 4.0, 3), self.materials[0]),
                                            self.u(-40, 40)), c)
        return b.medium(boundary, self.u(0.05, 0.6), b.constant_texture(self.u(0.1, 1.0, 3)))

    def make(self):
        b, A = self.b, self.A
        world_is_bvh = self.rng.random() < 0.4
        top = [self.subtree(0, world_is_bvh) for _ in range(int(self.rng.integers(2, 7)))]
        if self.rng.random() < 0.5:
            top += [self.medium() for _ in range(int(self.rng.integers(1, 3)))]
        light = b.flip(b.rect(A.PLANE_XZ, -3, 3, -3, 3, 11.0, self.light_material))
        lights, emitters = [light], [light]
        if self.rng.random() < 0.3:  # a sphere light next to it (sphere.rs:27-36,104-119)
            sl = b.sphere((6.0, 9.0, -4.0), 1.2, self.light_material)
            emitters.append(sl)
            lights.append(sl)
        if self.rich and self.rng.random() < 0.5:
            # Lights the reference's pdf cannot sample: only AARect and Sphere implement pdf_value / random and only
            # FlipNormal forwards them (hit.rs:126-132); a Cube or a translated rect in the light list falls back to
            # the trait defaults (hit.rs:29-30: pdf 0, direction (1,0,0)).
            odd = b.cube((-7.0, 8.0, 5.0), (-5.0, 9.0, 7.0), self.light_material) if self.rng.random() < 0.5 else \
                b.translate(b.rect(A.PLANE_XZ, -1, 1, -1, 1, 10.0, self.light_material), (4.0, 0.0, 4.0))
            emitters.append(odd)
            lights.append(odd)
            self.has_default_light = True
        self.rng.shuffle(top)
        # the XZ light rect stays outside any BVH (§Q5), like in every scene of main.rs
        world = b.list([b.bvh(top, 0.0, 1.0)] + emitters) if world_is_bvh else b.list(top + emitters)
        if self.rng.random() < 0.3:
            world = b.list([world])
        bg = (0.0, 0.0, 0.0) if self.rng.random() < 0.5 else (0.7, 0.8, 1.0)
        return b.finish(world, b.list(lights), background=bg)

    def rays(self, n):
        """Origins around and inside the scene, un-normalised directions, times in [0, 1)."""
        rays = np.zeros(n, dtype=self.A.RAY_DTYPE)
        o = self.u(-14, 14, (n, 3))
        target = self.u(-8, 8, (n, 3))
        rays["origin"] = o
        rays["direction"] = (target - o) * self.u(0.05, 2.0, (n, 1))
        rays["time"] = self.rng.random(n)
        return rays


def fuzz_camera(rt):
    """Looks at the cube the graphs live in from outside it."""
    return rt.camera_new((0.0, 3.0, -26.0), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 40.0, 1.0, 0.1, 26.0, 0.0, 1.0)


WEIRD_U32 = [0, 1, 2, 3, 5, 7, 0xFFFFFFFF, 0x7FFFFFFF, 0x80000000, 1000, 65535]
WEIRD_F64 = [float("nan"), float("inf"), -float("inf"), 0.0, -0.0, 1e308, -1e308, 1e-320]


def mutated_description(rt, seed):
    """A random graph with one to five fields of its description overwritten by out-of-range indices, unknown kinds,
    huge counts, NaN / inf parameters, and now and then a table declared shorter than it is: what a buggy flatten()
    could hand to rt_scene_create.  The compiler has to answer with a status, whatever it gets."""
    rng = np.random.default_rng(seed)
    g = GraphMaker(rt, seed, rich=True)
    g.make()
    b = g.b
    nodes, mats, texs = b.nodes, b.materials, b.textures
    pick = lambda seq: seq[int(rng.integers(len(seq)))]  # noqa: E731
    for _ in range(int(rng.integers(1, 6))):
        what = int(rng.integers(9))
        if what == 0:
            pick(nodes).kind = int(pick(WEIRD_U32 + list(range(12))))
        elif what == 1:
            pick(nodes).material = int(pick(WEIRD_U32))
        elif what == 2:
            pick(nodes).child = int(pick(WEIRD_U32 + [int(rng.integers(len(nodes)))]))
        elif what == 3:
            pick(nodes).count = int(pick(WEIRD_U32))
        elif what == 4:
            pick(nodes).axis = int(pick(WEIRD_U32))
        elif what == 5:
            pick(nodes).v[int(rng.integers(10))] = float(pick(WEIRD_F64))
        elif what == 6:
            pick(mats).kind = int(pick(WEIRD_U32))
        elif what == 7:
            pick(mats).texture = int(pick(WEIRD_U32))
        else:
            t, f = pick(texs), int(rng.integers(3))
            if f == 0:
                t.kind = int(pick(WEIRD_U32))
            elif f == 1:
                t.a = int(pick(WEIRD_U32 + [int(rng.integers(len(texs)))]))
            else:
                t.b = int(pick(WEIRD_U32 + [int(rng.integers(len(texs)))]))
    if rng.random() < 0.2 and b.child_index:
        b.child_index[int(rng.integers(len(b.child_index)))] = int(pick(WEIRD_U32 + [int(rng.integers(len(nodes)))]))
    # make() creates the world list second to last and the light list last
    world = len(nodes) - 2 if rng.random() < 0.8 else int(pick([int(rng.integers(len(nodes))), 0xFFFFFFFF]))
    sd = b.finish(world, len(nodes) - 1)
    if rng.random() < 0.15:
        d, f = sd.desc, int(rng.integers(5))
        if f == 0:
            d.n_nodes = int(rng.integers(0, d.n_nodes + 1))
        elif f == 1:
            d.n_child_index = int(rng.integers(0, d.n_child_index + 1))
        elif f == 2:
            d.n_materials = int(rng.integers(0, d.n_materials + 1))
        elif f == 3:
            d.n_textures = int(rng.integers(0, d.n_textures + 1))
        else:
            d.n_perlin = 0
    return sd
