// scene_api.cpp — flatten() implementations, construction RNG, Perlin tables,
// OBJ reader, Camera::new, format_color and the PPM writer.  See scene_api.hpp.
#include "scene_api.hpp"

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace rtb200 {

// ---------------------------------------------------------------------------
// SceneRng: Philox4x32-10, key = (seed, stream), counter = draw number
// ---------------------------------------------------------------------------
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = 0xD2511F53ull * c[0];
        uint64_t p1 = 0xCD9E8D57ull * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1;
        c[3] = (uint32_t)p0;
        c[0] = n0;
        c[2] = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

double SceneRng::gen_f64() {
    uint32_t c[4] = {(uint32_t)counter_, (uint32_t)(counter_ >> 32), 0x5CE11E5u, 0u};
    ++counter_;
    philox4x32_10(c, seed_, stream_);
    uint64_t x = ((uint64_t)c[0] << 32) | c[1];
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
double SceneRng::gen_range(double lo, double hi) { return lo + (hi - lo) * gen_f64(); }
uint32_t SceneRng::gen_index_inclusive(uint32_t hi) {
    uint32_t v = (uint32_t)(gen_f64() * (double)(hi + 1u));
    return v > hi ? hi : v;
}
Vec3 SceneRng::random_vec(double lo, double hi) {
    double a = gen_range(lo, hi);
    double b = gen_range(lo, hi);
    double c = gen_range(lo, hi);
    return Vec3(a, b, c);
}
Vec3 SceneRng::random_in_unit_sphere() {
    for (;;) {
        Vec3 v = random_vec(-1.0, 1.0);
        if (v.length() < 1.0) return v;
    }
}

// perlin.rs:13-37,68-75: gradient vectors first, then the three permutations.
Perlin::Perlin(SceneRng &rng) {
    for (int i = 0; i < 256; ++i) {
        Vec3 v = rng.random_in_unit_sphere();
        table.ranvec[3 * i] = v[0];
        table.ranvec[3 * i + 1] = v[1];
        table.ranvec[3 * i + 2] = v[2];
    }
    uint32_t *perms[3] = {table.perm_x, table.perm_y, table.perm_z};
    for (uint32_t *p : perms) {
        for (uint32_t i = 0; i < 256; ++i) p[i] = i;
        for (int i = 255; i >= 0; --i) {  // perlin.rs:21-28
            uint32_t target = rng.gen_index_inclusive((uint32_t)i);
            std::swap(p[i], p[target]);
        }
    }
}

// ---------------------------------------------------------------------------
// flatten()
// ---------------------------------------------------------------------------
template <class F>
static uint32_t memoised(SceneBuilder &b, const void *self, F make) {
    auto it = b.memo.find(self);
    if (it != b.memo.end()) return it->second;
    uint32_t id = make();
    b.memo[self] = id;
    return id;
}

uint32_t ConstantTexture::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtTexture t{};
        t.kind = RT_TEX_CONSTANT;
        t.a = t.b = RT_NONE;
        t.color[0] = value[0];
        t.color[1] = value[1];
        t.color[2] = value[2];
        b.textures.push_back(t);
        return (uint32_t)(b.textures.size() - 1);
    });
}
uint32_t CheckTexture::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtTexture t{};
        t.kind = RT_TEX_CHECKER;
        t.a = odd->flatten(b);
        t.b = even->flatten(b);
        b.textures.push_back(t);
        return (uint32_t)(b.textures.size() - 1);
    });
}
uint32_t NoiseTexture::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        b.perlin.push_back(noise.table);
        RtTexture t{};
        t.kind = RT_TEX_NOISE;
        t.a = (uint32_t)(b.perlin.size() - 1);
        t.b = RT_NONE;
        t.scale = scale;
        b.textures.push_back(t);
        return (uint32_t)(b.textures.size() - 1);
    });
}
uint32_t ImageTexture::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtImage im{};
        im.width = width;
        im.height = height;
        im.offset = b.texels.size();
        b.texels.insert(b.texels.end(), data.begin(), data.end());
        b.images.push_back(im);
        RtTexture t{};
        t.kind = RT_TEX_IMAGE;
        t.a = (uint32_t)(b.images.size() - 1);
        t.b = RT_NONE;
        b.textures.push_back(t);
        return (uint32_t)(b.textures.size() - 1);
    });
}

static uint32_t push_material(SceneBuilder &b, uint32_t kind, uint32_t texture, Color albedo, double fuzz, double ir) {
    RtMaterial m{};
    m.kind = kind;
    m.texture = texture;
    m.albedo[0] = albedo[0];
    m.albedo[1] = albedo[1];
    m.albedo[2] = albedo[2];
    m.fuzz = fuzz;
    m.ir = ir;
    b.materials.push_back(m);
    return (uint32_t)(b.materials.size() - 1);
}
uint32_t Lambertian::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return push_material(b, RT_MAT_LAMBERTIAN, albedo->flatten(b), Color(), 0, 0); });
}
uint32_t Metal::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return push_material(b, RT_MAT_METAL, RT_NONE, albedo, fuzz, 0); });
}
uint32_t Dielectric::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return push_material(b, RT_MAT_DIELECTRIC, RT_NONE, Color(), 0, ir); });
}
uint32_t PBR::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        uint32_t k = push_material(b, RT_MAT_PBR, base_color->flatten(b), Color(), 0, 0);
        for (int i = 0; i < 10; ++i) b.materials[k].pbr[i] = p[i];
        return k;
    });
}
uint32_t DiffuseLight::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return push_material(b, RT_MAT_DIFFUSE_LIGHT, emit->flatten(b), Color(), 0, 0); });
}
uint32_t Isotropic::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return push_material(b, RT_MAT_ISOTROPIC, albedo->flatten(b), Color(), 0, 0); });
}

uint32_t Sphere::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_SPHERE);
        n.material = material->flatten(b);
        n.v[0] = center[0]; n.v[1] = center[1]; n.v[2] = center[2]; n.v[3] = radius;
        return b.add_node(n);
    });
}
uint32_t MovingSphere::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_MOVING_SPHERE);
        n.material = material->flatten(b);
        for (int a = 0; a < 3; ++a) {
            n.v[a] = center0[a];
            n.v[3 + a] = center1[a];
        }
        n.v[6] = time0; n.v[7] = time1; n.v[8] = radius;
        return b.add_node(n);
    });
}
uint32_t AARect::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_RECT);
        n.material = material->flatten(b);
        n.axis = plane == Plane::YZ ? RT_PLANE_YZ : plane == Plane::XZ ? RT_PLANE_XZ : RT_PLANE_XY;
        n.v[0] = a0; n.v[1] = a1; n.v[2] = b0; n.v[3] = b1; n.v[4] = k;
        return b.add_node(n);
    });
}
uint32_t Triangle::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_TRIANGLE);
        n.material = material->flatten(b);
        for (int i = 0; i < 3; ++i)
            for (int a = 0; a < 3; ++a) n.v[3 * i + a] = vertices[i][a];
        return b.add_node(n);
    });
}
uint32_t Cube::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_CUBE);
        n.material = material->flatten(b);
        for (int a = 0; a < 3; ++a) {
            n.v[a] = min[a];
            n.v[3 + a] = max[a];
        }
        return b.add_node(n);
    });
}
static uint32_t flatten_children(SceneBuilder &b, uint32_t kind, const std::vector<HittablePtr> &kids, double t0, double t1) {
    std::vector<uint32_t> ids;
    ids.reserve(kids.size());
    for (const HittablePtr &h : kids) ids.push_back(h->flatten(b));
    RtNode n = SceneBuilder::blank(kind);
    n.child = (uint32_t)b.child_index.size();
    n.count = (uint32_t)ids.size();
    n.v[0] = t0;
    n.v[1] = t1;
    b.child_index.insert(b.child_index.end(), ids.begin(), ids.end());
    return b.add_node(n);
}
uint32_t HittableList::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return flatten_children(b, RT_NODE_LIST, list, 0.0, 0.0); });
}
uint32_t BVH::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] { return flatten_children(b, RT_NODE_BVH, hit, time0, time1); });
}
uint32_t Translate::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_TRANSLATE);
        n.child = hittable->flatten(b);
        n.v[0] = offset[0]; n.v[1] = offset[1]; n.v[2] = offset[2];
        return b.add_node(n);
    });
}
uint32_t Rotate::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_ROTATE);
        n.child = hittable->flatten(b);
        n.axis = axis == Axis::X ? RT_AXIS_X : axis == Axis::Y ? RT_AXIS_Y : RT_AXIS_Z;
        n.v[0] = angle;
        return b.add_node(n);
    });
}
uint32_t FlipNormal::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_FLIP);
        n.child = hittable->flatten(b);
        return b.add_node(n);
    });
}
uint32_t ConstantMedium::flatten(SceneBuilder &b) const {
    return memoised(b, this, [&] {
        RtNode n = SceneBuilder::blank(RT_NODE_MEDIUM);
        n.child = boundary->flatten(b);
        n.material = phase_function->flatten(b);
        n.v[0] = density;
        return b.add_node(n);
    });
}

// ---------------------------------------------------------------------------
// Mesh / OBJ (src/mesh.rs)
// ---------------------------------------------------------------------------
Mesh::Mesh(const std::vector<Vec3> &positions, const std::vector<uint32_t> &indices, MaterialPtr material) {
    for (size_t i = 0; i < indices.size() / 3; ++i) {  // mesh.rs:19-26
        tris.push(Triangle::make(positions[indices[i * 3]], positions[indices[i * 3 + 1]],
                                 positions[indices[i * 3 + 2]], material));
    }
}

static long resolve_index(long idx, size_t n_vertices) {
    // OBJ indices are 1-based; negative indices count from the end.  (long all the way: an index beyond 2^31 must
    // be out of range, not wrap onto a vertex.)
    if (idx > 0) return idx - 1;
    if (idx < 0) return (long)n_vertices + idx;
    return -1;
}

void read_obj_first_model(const std::string &path, std::vector<float> &positions, std::vector<uint32_t> &indices) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Failed to load obj file: cannot open " + path);
    positions.clear();
    indices.clear();
    bool model_has_faces = false;
    std::string line;
    while (std::getline(in, line)) {
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos || line[p] == '#') continue;
        std::istringstream ss(line.substr(p));
        std::string tag;
        ss >> tag;
        if (tag == "v") {
            // tobj stores positions as f32 (mesh.rs:48-52 widens them to f64 afterwards)
            std::string a, b, c;
            ss >> a >> b >> c;
            positions.push_back(std::strtof(a.c_str(), nullptr));
            positions.push_back(std::strtof(b.c_str(), nullptr));
            positions.push_back(std::strtof(c.c_str(), nullptr));
        } else if (tag == "f") {
            std::vector<uint32_t> poly;
            std::string tok;
            while (ss >> tok) {
                long vi = std::strtol(tok.c_str(), nullptr, 10);  // "v", "v/vt", "v//vn", "v/vt/vn"
                long r = resolve_index(vi, positions.size() / 3);
                if (r < 0 || (size_t)r >= positions.size() / 3)
                    throw std::runtime_error("Failed to load obj file: face index out of range in " + path);
                poly.push_back((uint32_t)r);
            }
            if (poly.size() < 3) throw std::runtime_error("Failed to load obj file: degenerate face in " + path);
            for (size_t k = 1; k + 1 < poly.size(); ++k) {  // fan triangulation (triangulate: true)
                indices.push_back(poly[0]);
                indices.push_back(poly[k]);
                indices.push_back(poly[k + 1]);
            }
            model_has_faces = true;
        } else if (tag == "o" || tag == "g") {
            // a new object/group after faces starts models[1]; the reference uses models[0] only (§Q20)
            if (model_has_faces) break;
        }
    }
    if (indices.empty()) throw std::runtime_error("Failed to load obj file: no faces in " + path);
}

Mesh Mesh::load_obj(const std::string &path, Vec3 offset, double scale, MaterialPtr material) {
    std::vector<float> pos;
    std::vector<uint32_t> idx;
    read_obj_first_model(path, pos, idx);
    std::vector<Vec3> tri_positions;
    tri_positions.reserve(pos.size() / 3);
    for (size_t i = 0; i + 2 < pos.size(); i += 3)  // mesh.rs:48-52
        tri_positions.push_back(Point3((double)pos[i], (double)pos[i + 1], (double)pos[i + 2]) * scale + offset);
    return Mesh(tri_positions, idx, material);
}

// ---------------------------------------------------------------------------
// Camera::new (src/camera.rs:19-49)
// ---------------------------------------------------------------------------
Camera::Camera(Point3 lookfrom, Point3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
               double focus_dist, double time0, double time1) {
    const double PI = 3.14159265358979323846264338327950288;
    double theta = PI / 180.0 * vfov;
    double viewport_height = 2.0 * std::tan(theta / 2.0);
    double viewport_width = viewport_height * aspect_ratio;
    Vec3 cw = (lookfrom - lookat).normalized();
    Vec3 cu = vup.cross(cw).normalized();
    Vec3 cv = cw.cross(cu);
    Vec3 h = (focus_dist * viewport_width) * cu;
    Vec3 v = (focus_dist * viewport_height) * cv;
    Vec3 llc = lookfrom - h / 2.0 - v / 2.0 - focus_dist * cw;
    for (int a = 0; a < 3; ++a) {
        pod.origin[a] = lookfrom[a];
        pod.lower_left_corner[a] = llc[a];
        pod.horizontal[a] = h[a];
        pod.vertical[a] = v[a];
        pod.cu[a] = cu[a];
        pod.cv[a] = cv[a];
    }
    pod.lens_radius = aperture / 2.0;
    pod.time0 = time0;
    pod.time1 = time1;
}

// ---------------------------------------------------------------------------
// FlatScene
// ---------------------------------------------------------------------------
FlatScene::FlatScene(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights, Color background) {
    std::memset(&desc, 0, sizeof(desc));
    desc.abi_version = RTB200_ABI_VERSION;
    desc.world = world->flatten(b);
    desc.lights = lights->flatten(b);
    desc.background[0] = background[0];
    desc.background[1] = background[1];
    desc.background[2] = background[2];
    desc.nodes = b.nodes.data();
    desc.n_nodes = b.nodes.size();
    desc.child_index = b.child_index.data();
    desc.n_child_index = b.child_index.size();
    desc.materials = b.materials.data();
    desc.n_materials = b.materials.size();
    desc.textures = b.textures.data();
    desc.n_textures = b.textures.size();
    desc.perlin = b.perlin.data();
    desc.n_perlin = b.perlin.size();
    desc.images = b.images.data();
    desc.n_images = b.images.size();
    desc.texels = b.texels.data();
    desc.n_texel_bytes = b.texels.size();
}

// ---------------------------------------------------------------------------
// format_color (src/vec.rs:125-131) and the PPM writer (src/main.rs:767-769,832)
// ---------------------------------------------------------------------------
void format_color(const float sum[3], uint64_t samples_per_pixel, uint64_t out[3]) {
    for (int a = 0; a < 3; ++a) {
        double x = std::sqrt((double)sum[a] / (double)samples_per_pixel);
        // f64::clamp(0.0, 0.999) keeps NaN; `NaN as u64` is 0 (§Q10)
        if (x < 0.0) x = 0.0;
        if (x > 0.999) x = 0.999;
        double y = 256.0 * x;
        out[a] = (y == y && y > 0.0) ? (uint64_t)y : 0;
    }
}

void write_ppm(FILE *f, const float *rgb_sum, uint32_t width, uint32_t height, uint64_t samples_per_pixel) {
    std::fprintf(f, "P3\n%u %u\n255\n", width, height);
    for (uint64_t p = 0; p < (uint64_t)width * height; ++p) {
        uint64_t c[3];
        format_color(rgb_sum + 3 * p, samples_per_pixel, c);
        std::fprintf(f, "%llu %llu %llu\n", (unsigned long long)c[0], (unsigned long long)c[1],
                     (unsigned long long)c[2]);
    }
}

}  // namespace rtb200
