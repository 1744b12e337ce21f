// scene_api.hpp — host-side mirror of the reference's scene-construction API.
//
// The reference's host is Rust; there is no Rust toolchain in this image, so the
// host side above the C ABI (include/rtb200.h) is restated in C++ with the same
// type names, constructor arguments and argument meaning as the Rust types, so
// that a scene constructor here reads like src/main.rs:153-513.  These classes
// only *describe* geometry: none of them can intersect a ray.  Each one has a
// flatten() that serialises it into the RtSceneDesc the device library consumes
// — the method a Rust maintainer would add to `trait Hittable` (src/hit.rs:26-31),
// `trait Material` (src/mat.rs:54-77) and `trait Texture` (src/texture.rs:5-7);
// see INTEGRATION.md.
#pragma once

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/rtb200.h"

namespace rtb200 {

// ---- Vec3 (src/vec.rs:9-132): only what scene construction needs -----------
struct Vec3 {
    double e[3];
    Vec3() : e{0, 0, 0} {}
    Vec3(double a, double b, double c) : e{a, b, c} {}
    double x() const { return e[0]; }
    double y() const { return e[1]; }
    double z() const { return e[2]; }
    double operator[](int i) const { return e[i]; }
    double &operator[](int i) { return e[i]; }
    double dot(Vec3 o) const { return e[0] * o[0] + e[1] * o[1] + e[2] * o[2]; }
    double length() const { return std::sqrt(dot(*this)); }
    Vec3 cross(Vec3 o) const {
        return Vec3(e[1] * o[2] - e[2] * o[1], e[2] * o[0] - e[0] * o[2], e[0] * o[1] - e[1] * o[0]);
    }
    Vec3 normalized() const {
        double l = length();
        return Vec3(e[0] / l, e[1] / l, e[2] / l);
    }
};
typedef Vec3 Point3;
typedef Vec3 Color;
inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline Vec3 operator*(Vec3 a, double s) { return Vec3(a[0] * s, a[1] * s, a[2] * s); }
inline Vec3 operator*(double s, Vec3 a) { return Vec3(s * a[0], s * a[1], s * a[2]); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
inline Vec3 operator/(Vec3 a, double s) { return Vec3(a[0] / s, a[1] / s, a[2] / s); }

// ---- construction-time RNG --------------------------------------------------
// The reference seeds scene construction from the OS (rand::thread_rng,
// src/main.rs:154,457, src/perlin.rs:5-24), so its scenes differ run to run.
// Here construction draws come from a seeded Philox4x32-10 counter stream so a
// (scene, seed) pair always yields the same RtSceneDesc.
class SceneRng {
public:
    explicit SceneRng(uint32_t seed, uint32_t stream = 0) : seed_(seed), stream_(stream), counter_(0) {}
    double gen_f64();                              // rng.gen::<f64>()
    double gen_range(double lo, double hi);        // rng.gen_range(lo..hi)
    uint32_t gen_index_inclusive(uint32_t hi);     // rng.gen_range(0..=hi)
    Vec3 random_vec(double lo, double hi);         // Vec3::random (src/vec.rs:70-76)
    Vec3 random_in_unit_sphere();                  // src/vec.rs:78-85
private:
    uint32_t seed_, stream_;
    uint64_t counter_;
};

// ---- SceneBuilder: the flatten() target --------------------------------------
class SceneBuilder {
public:
    std::vector<RtNode> nodes;
    std::vector<uint32_t> child_index;
    std::vector<RtMaterial> materials;
    std::vector<RtTexture> textures;
    std::vector<RtPerlin> perlin;
    std::vector<RtImage> images;
    std::vector<uint8_t> texels;
    std::unordered_map<const void *, uint32_t> memo;  // object identity -> index

    uint32_t add_node(const RtNode &n) {
        nodes.push_back(n);
        return (uint32_t)(nodes.size() - 1);
    }
    static RtNode blank(uint32_t kind) {
        RtNode n{};
        n.kind = kind;
        n.material = RT_NONE;
        n.child = RT_NONE;
        return n;
    }
};

// ---- Texture (src/texture.rs) -------------------------------------------------
struct Texture {
    virtual ~Texture() {}
    virtual uint32_t flatten(SceneBuilder &b) const = 0;
};
typedef std::shared_ptr<const Texture> TexturePtr;

struct ConstantTexture : Texture {  // texture.rs:10-27
    Color value;
    explicit ConstantTexture(Color c) : value(c) {}
    static TexturePtr make(Color c) { return std::make_shared<ConstantTexture>(c); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct CheckTexture : Texture {  // texture.rs:31-54
    TexturePtr odd, even;
    CheckTexture(TexturePtr o, TexturePtr e) : odd(o), even(e) {}
    static TexturePtr make(TexturePtr o, TexturePtr e) { return std::make_shared<CheckTexture>(o, e); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Perlin {  // perlin.rs:60-75
    RtPerlin table;
    explicit Perlin(SceneRng &rng);
};
struct NoiseTexture : Texture {  // texture.rs:57-79
    Perlin noise;
    double scale;
    NoiseTexture(double s, SceneRng &rng) : noise(rng), scale(s) {}
    static TexturePtr make(double s, SceneRng &rng) { return std::make_shared<NoiseTexture>(s, rng); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct ImageTexture : Texture {  // texture.rs:83-121
    std::vector<uint8_t> data;
    uint32_t width, height;
    ImageTexture(std::vector<uint8_t> d, uint32_t w, uint32_t h) : data(std::move(d)), width(w), height(h) {}
    static TexturePtr make(std::vector<uint8_t> d, uint32_t w, uint32_t h) {
        return std::make_shared<ImageTexture>(std::move(d), w, h);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};

// ---- Material (src/mat.rs:199-422) ---------------------------------------------
struct Material {
    virtual ~Material() {}
    virtual uint32_t flatten(SceneBuilder &b) const = 0;
};
typedef std::shared_ptr<const Material> MaterialPtr;

struct Lambertian : Material {  // mat.rs:199-250
    TexturePtr albedo;
    explicit Lambertian(TexturePtr a) : albedo(a) {}
    static MaterialPtr make(TexturePtr a) { return std::make_shared<Lambertian>(a); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Metal : Material {  // mat.rs:253-294
    Color albedo;
    double fuzz;
    Metal(Color a, double f) : albedo(a), fuzz(f) {}
    static MaterialPtr make(Color a, double f) { return std::make_shared<Metal>(a, f); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Dielectric : Material {  // mat.rs:297-375
    double ir;
    explicit Dielectric(double i) : ir(i) {}
    static MaterialPtr make(double i) { return std::make_shared<Dielectric>(i); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct DiffuseLight : Material {  // mat.rs:377-402
    TexturePtr emit;
    explicit DiffuseLight(TexturePtr e) : emit(e) {}
    static MaterialPtr make(TexturePtr e) { return std::make_shared<DiffuseLight>(e); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct PBR : Material {  // mat.rs:86-197 (PBR::new, :101-115)
    TexturePtr base_color;
    double p[10];  // metallic, subsurface, specular, roughness, specular_tint, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss
    PBR(TexturePtr c, double metallic, double subsurface, double specular, double roughness, double specular_tint,
        double anisotropic, double sheen, double sheen_tint, double clearcoat, double clearcoat_gloss)
        : base_color(c), p{metallic, subsurface, specular, roughness, specular_tint, anisotropic, sheen, sheen_tint, clearcoat, clearcoat_gloss} {}
    static MaterialPtr make(TexturePtr c, double metallic, double subsurface, double specular, double roughness,
                            double specular_tint, double anisotropic, double sheen, double sheen_tint, double clearcoat,
                            double clearcoat_gloss) {
        return std::make_shared<PBR>(c, metallic, subsurface, specular, roughness, specular_tint, anisotropic, sheen, sheen_tint,
                                     clearcoat, clearcoat_gloss);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Isotropic : Material {  // mat.rs:404-422
    TexturePtr albedo;
    explicit Isotropic(TexturePtr a) : albedo(a) {}
    static MaterialPtr make(TexturePtr a) { return std::make_shared<Isotropic>(a); }
    uint32_t flatten(SceneBuilder &b) const override;
};

// ---- Hittable (src/hit.rs:26-31 and implementors) -------------------------------
struct Hittable {
    virtual ~Hittable() {}
    virtual uint32_t flatten(SceneBuilder &b) const = 0;
};
typedef std::shared_ptr<const Hittable> HittablePtr;

enum class Plane { XY, XZ, YZ };  // rect.rs:8-13
enum class Axis { X, Y, Z };      // rotate.rs:8-13

struct Sphere : Hittable {  // sphere.rs:38-54
    Point3 center;
    double radius;
    MaterialPtr material;
    Sphere(Point3 c, double r, MaterialPtr m) : center(c), radius(r), material(m) {}
    static HittablePtr make(Point3 c, double r, MaterialPtr m) { return std::make_shared<Sphere>(c, r, m); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct MovingSphere : Hittable {  // sphere.rs:122-147
    Point3 center0, center1;
    double time0, time1, radius;
    MaterialPtr material;
    MovingSphere(Point3 c0, Point3 c1, double t0, double t1, double r, MaterialPtr m)
        : center0(c0), center1(c1), time0(t0), time1(t1), radius(r), material(m) {}
    static HittablePtr make(Point3 c0, Point3 c1, double t0, double t1, double r, MaterialPtr m) {
        return std::make_shared<MovingSphere>(c0, c1, t0, t1, r, m);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct AARect : Hittable {  // rect.rs:15-46
    Plane plane;
    double a0, a1, b0, b1, k;
    MaterialPtr material;
    AARect(Plane p, double a0_, double a1_, double b0_, double b1_, double k_, MaterialPtr m)
        : plane(p), a0(a0_), a1(a1_), b0(b0_), b1(b1_), k(k_), material(m) {}
    static HittablePtr make(Plane p, double a0, double a1, double b0, double b1, double k, MaterialPtr m) {
        return std::make_shared<AARect>(p, a0, a1, b0, b1, k, m);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Triangle : Hittable {  // tri.rs:9-21
    Point3 vertices[3];
    MaterialPtr material;
    Triangle(Point3 v0, Point3 v1, Point3 v2, MaterialPtr m) : vertices{v0, v1, v2}, material(m) {}
    static HittablePtr make(Point3 v0, Point3 v1, Point3 v2, MaterialPtr m) {
        return std::make_shared<Triangle>(v0, v1, v2, m);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Cube : Hittable {  // cube.rs:7-31
    Point3 min, max;
    MaterialPtr material;
    Cube(Point3 mn, Point3 mx, MaterialPtr m) : min(mn), max(mx), material(m) {}
    static HittablePtr make(Point3 mn, Point3 mx, MaterialPtr m) { return std::make_shared<Cube>(mn, mx, m); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct HittableList : Hittable {  // hit.rs:47-57
    std::vector<HittablePtr> list;
    void push(HittablePtr h) { list.push_back(h); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct BVH : Hittable {  // bvh.rs:12-73.  The tree is built by whoever consumes the description.
    std::vector<HittablePtr> hit;
    double time0, time1;
    BVH(std::vector<HittablePtr> h, double t0, double t1) : hit(std::move(h)), time0(t0), time1(t1) {}
    static HittablePtr make(std::vector<HittablePtr> h, double t0, double t1) {
        return std::make_shared<BVH>(std::move(h), t0, t1);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Translate : Hittable {  // translate.rs:6-19
    HittablePtr hittable;
    Vec3 offset;
    Translate(HittablePtr h, Vec3 o) : hittable(h), offset(o) {}
    static HittablePtr make(HittablePtr h, Vec3 o) { return std::make_shared<Translate>(h, o); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct Rotate : Hittable {  // rotate.rs:23-31 (angle in degrees)
    Axis axis;
    HittablePtr hittable;
    double angle;
    Rotate(Axis a, HittablePtr h, double deg) : axis(a), hittable(h), angle(deg) {}
    static HittablePtr make(Axis a, HittablePtr h, double deg) { return std::make_shared<Rotate>(a, h, deg); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct FlipNormal : Hittable {  // hit.rs:99-110
    HittablePtr hittable;
    explicit FlipNormal(HittablePtr h) : hittable(h) {}
    static HittablePtr make(HittablePtr h) { return std::make_shared<FlipNormal>(h); }
    uint32_t flatten(SceneBuilder &b) const override;
};
struct ConstantMedium : Hittable {  // medium.rs:10-24
    HittablePtr boundary;
    double density;
    MaterialPtr phase_function;  // Isotropic::new(texture)
    ConstantMedium(HittablePtr bnd, double d, TexturePtr t)
        : boundary(bnd), density(d), phase_function(Isotropic::make(t)) {}
    static HittablePtr make(HittablePtr bnd, double d, TexturePtr t) {
        return std::make_shared<ConstantMedium>(bnd, d, t);
    }
    uint32_t flatten(SceneBuilder &b) const override;
};

// Mesh (src/mesh.rs:10-61): triangles of the first model of an OBJ file.
struct Mesh {
    HittableList tris;
    Mesh(const std::vector<Vec3> &positions, const std::vector<uint32_t> &indices, MaterialPtr material);
    // mesh.rs:33-61.  Throws std::runtime_error("Failed to load obj file: ...").
    static Mesh load_obj(const std::string &path, Vec3 offset, double scale, MaterialPtr material);
};
// The part of tobj 3.2.3's load_obj the reference relies on (mesh.rs:40-52):
// first model only, `v` parsed as f32, faces fan-triangulated, position indices.
void read_obj_first_model(const std::string &path, std::vector<float> &positions, std::vector<uint32_t> &indices);

// ---- Camera (src/camera.rs:5-49) ---------------------------------------------------
struct Camera {
    RtCamera pod;
    Camera(Point3 lookfrom, Point3 lookat, Vec3 vup, double vfov, double aspect_ratio, double aperture,
           double focus_dist, double time0, double time1);
};

// ---- A flattened scene that owns its storage ----------------------------------------
struct FlatScene {
    SceneBuilder b;
    RtSceneDesc desc;
    // world and lights as returned by the reference's scene constructors
    // (`(Box<dyn Hittable>, Box<dyn Hittable>)`, src/main.rs:153), plus the background.
    FlatScene(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights, Color background);
    FlatScene(const FlatScene &) = delete;
    FlatScene &operator=(const FlatScene &) = delete;
};

// ---- render(): the function BASELINE's north star introduces ------------------------
struct RenderResult {
    std::vector<float> rgb_sum;  // W*H*3, rows top-down (empty when only the PPM was asked for)
    std::string ppm;             // the P3 file, formatted on the GPU (render_ppm only)
    RtStats stats;
};
// Replaces the loop at src/main.rs:772-834.  `lights` and `background` are inputs of
// ray_color (src/main.rs:41) and so of render().  Throws std::runtime_error with
// rt_last_error() on failure (the reference panics).
RenderResult render(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights,
                    Color background, const Camera &camera, uint32_t width, uint32_t height,
                    uint32_t spp, uint32_t max_depth, const RtRenderOpts &opts, int device = 0);
// The same over `n_gpus` GPUs of the box (0 = all): one host thread, contiguous sample blocks per GPU,
// one combine kernel over NVLink peer memory (rt_scene_group_create / rt_render_multi).
RenderResult render_gpus(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights,
                         Color background, const Camera &camera, uint32_t width, uint32_t height,
                         uint32_t spp, uint32_t max_depth, const RtRenderOpts &opts, uint32_t n_gpus);
// render_gpus + format_color + the P3 text of src/main.rs:767-769,832 on the GPU: what `cargo run
// --release > image.ppm` prints, ready for one fwrite.  The fp32 sums do not come back to the host.
RenderResult render_ppm(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights,
                        Color background, const Camera &camera, uint32_t width, uint32_t height,
                        uint32_t spp, uint32_t max_depth, const RtRenderOpts &opts, uint32_t n_gpus);

// Vec3::format_color (src/vec.rs:125-131) on a sum of `samples_per_pixel` samples.
void format_color(const float sum[3], uint64_t samples_per_pixel, uint64_t out[3]);
// The P3 writer of src/main.rs:767-769,832.
void write_ppm(FILE *f, const float *rgb_sum, uint32_t width, uint32_t height, uint64_t samples_per_pixel);

// ---- scene catalogue (src/main.rs:153-513, 623-765) -----------------------------------
struct SceneSpec {
    HittablePtr world;
    std::shared_ptr<const HittableList> lights;
    Color background;
    Camera camera;
    uint32_t integrator;  // RtIntegrator the config is rendered with (§Q7)
    uint32_t width, height, spp, max_depth;  // the config's full size
};
// name: "random" (C1), "cornell" (C2), "cornell_smoke" (C3), "final" (C4), "mesh" (C5),
//       "light_room", "two_spheres" (small extras for tests).
// assets_dir holds earthmap_1024x512.rgb and teapot.obj (and Venus.obj if it exists).
SceneSpec make_scene(const std::string &name, uint32_t construction_seed, const std::string &assets_dir,
                     uint32_t mesh_detail = 0);

}  // namespace rtb200
