// tables.h — the HBM-resident layout of a compiled scene (shared by the host-side
// scene compiler and the CUDA kernels).  See DESIGN.md "Data layout in HBM".
//
// The reference walks a Box<dyn Hittable> tree (src/hit.rs:26-31) with dynamic
// dispatch.  Here the tree is compiled once into flat tables:
//   prims   tagged-union primitive records (sphere.rs, rect.rs, tri.rs, cube.rs)
//   ops     translate / rotate / flip steps      (translate.rs, rotate.rs, hit.rs:99-133)
//   chains  op sequences, outermost first: the wrappers above a primitive
//   groups  primitives that share a ray transform; linear or with a BVH
//   nodes   flattened SAH BVH2, two child boxes per node, breadth-first order
//   media   ConstantMedium records (medium.rs) with their boundary sub-scene
//   lights  the light list (pdf.rs PDF::Hittable -> hit.rs:90-96)
#pragma once
#include <stdint.h>

namespace rtb200dev {

enum PrimKind : uint32_t { PRIM_SPHERE = 0, PRIM_MSPHERE = 1, PRIM_RECT = 2, PRIM_TRI = 3, PRIM_BOX = 4 };
enum OpKind : uint32_t { OP_TRANSLATE = 0, OP_ROTATE = 1, OP_FLIP = 2 };
enum LightKind : uint32_t { LIGHT_RECT = 0, LIGHT_SPHERE = 1, LIGHT_DEFAULT = 2 };

// 128 bytes, 16-byte aligned: a whole record is eight 128-bit loads.
struct alignas(16) DPrim {
    uint32_t kind;      // PrimKind
    uint32_t material;  // index into materials
    uint32_t chain;     // index into chains (full wrapper sequence incl. flips)
    uint32_t axis;      // RECT: RtPlane
    int32_t node;       // RtSceneDesc node index (reported by the parity hooks)
    uint32_t rank;      // position in the reference's traversal order (tie-break)
    uint32_t pad0, pad1;
    // SPHERE  cx cy cz r
    // MSPHERE c0x c0y c0z c1x c1y c1z t0 t1 r
    // RECT    a0 a1 b0 b1 k
    // TRI     v0xyz e1xyz e2xyz nxyz   (e1 = v1-v0, e2 = v2-v0, n = normalize(e1 x e2): the
    //                                   same f64 expressions tri.rs:27-28,41 evaluates per hit)
    // BOX     minxyz maxxyz
    double d[12];
};

struct alignas(16) DOp {
    uint32_t kind;  // OpKind
    uint32_t axis;  // ROTATE: RtAxis
    double sin_theta, cos_theta;
    double offset[3];
};

struct DChain {
    uint32_t first_op, n_ops;
};

enum GroupFlags : uint32_t {
    GROUP_CULL = 1,     // test the group's outer-space bounds before transforming the ray
    GROUP_ROTATED = 2,  // the chain rotates: the reciprocal direction must be recomputed
    GROUP_XFORM = 4,    // the chain moves the ray at all (else the group lives in the outer space)
};

struct alignas(16) DGroup {
    uint32_t chain;       // ray transform of this group (flips are skipped when transforming)
    uint32_t first_prim;  // into prims
    uint32_t n_prims;
    int32_t bvh_root;     // node index, or -1: scan the primitives linearly
    uint32_t flags;       // GroupFlags
    uint32_t pad0, pad1, pad2;
    double bmin[3], bmax[3];  // conservative bounds in the space the ray is given in (culling only)
    // SEARCH only: the chain composed into one affine map, object = m * outer + t (row-major m).
    // The winner is re-derived op by op in reference arithmetic (chain_ray).
    double m[9], t[3];
};

// Two child boxes per node, 64 bytes = four 128-bit loads.  child >= 0: inner node; child < 0:
// leaf, ~child = (first_prim << 3) | (count-1).  The boxes only cull, so they are fp32: each
// bound is the f64 bound rounded OUTWARD, which together with the traversal's rounded-up /
// rounded-down ray origin and its relative slack on the slab distances (trace.cuh slab2f) can
// only make a box larger than the f64 box, never smaller.
constexpr int kStackSize = 64;  // traversal stack entries per ray; compile_scene refuses a tree that could need more
struct alignas(16) DBvhNode {
    float lo0[3], hi0[3];
    float lo1[3], hi1[3];
    int32_t child0, child1;
    int32_t pad0, pad1;
};

struct alignas(16) DMedium {
    uint32_t first_group, n_groups;  // boundary sub-scene
    uint32_t chain;                  // wrappers above the medium itself
    uint32_t material;               // the Isotropic phase function
    int32_t node;                    // RtSceneDesc node index; also the RNG sub-slot
    uint32_t rank;
    double density;
    // The boundary is ONE box or sphere primitive in one group under a Translate / Rotate (the smoke boxes of
    // cornell_box_with_smoke): its index, else 0xFFFFFFFF.  world_search then answers both boundary queries of medium.rs:29-30 from one slab /
    // root computation on one transformed ray instead of searching the boundary sub-scene twice.
    uint32_t convex_prim;
    uint32_t pad0, pad1, pad2;
};

struct alignas(16) DLight {
    uint32_t kind;  // LightKind
    uint32_t axis;  // RECT: RtPlane
    double d[5];    // RECT a0 a1 b0 b1 k ; SPHERE cx cy cz r
    double pad;
};

struct alignas(16) DMaterial {
    uint32_t kind;     // RtMaterialKind
    uint32_t texture;
    uint32_t needs_uv;  // the texture tree reads (u,v): only image textures do
    uint32_t costly;    // the texture tree has Perlin noise or an image: shaded in the compacted second pass (wavefront)
    double albedo[3];
    double fuzz, ir;
    double pad2;
    double pbr[10];  // RT_MAT_PBR: PBR::new's ten scalars (RT_PBR_* order)
};

struct alignas(16) DTexture {
    uint32_t kind, a, b, pad;
    double color[3];
    double scale;
};

struct DImage {
    uint32_t width, height;
    uint64_t offset;
};

struct DPerlin {
    double ranvec[256 * 3];
    uint32_t perm_x[256], perm_y[256], perm_z[256];
};

// What a scene uses.  The pipelines are compiled once per feature set (variants.h): a scene runs
// on the smallest variant that covers its features, so the Cornell box does not carry sphere,
// triangle, BVH, Perlin or legacy-integrator code through its instruction cache and registers.
enum Feat : uint32_t {
    F_SPHERE = 1, F_MSPHERE = 2, F_RECT = 4, F_TRI = 8, F_BOX = 16,
    F_BVH = 32,            // some group has a BVH
    F_TEX = 64,            // a non-constant texture (checker, noise, image)
    F_SPHERE_LIGHT = 128,  // a sphere in the light list
    F_METAL = 256, F_DIELECTRIC = 512,
    F_LEGACY = 1024, F_HEAD = 2048,  // which ray_color (a render option, added at render time)
    F_PBR = 4096,                    // the Disney-style material (mat.rs:86-197) and PDF::BRDF
    F_ALL = 0xFFFFFFFFu
};

// Pointers are device pointers once uploaded.
struct DScene {
    const DPrim *prims;
    const DOp *ops;
    const DChain *chains;
    const DGroup *groups;
    const DBvhNode *nodes;
    const DMedium *media;
    const DLight *lights;
    const DMaterial *materials;
    const DTexture *textures;
    const DImage *images;
    const DPerlin *perlin;
    const uint8_t *texels;
    uint32_t n_world_groups;  // the world sub-scene is groups [0, n_world_groups)
    uint32_t n_media;
    uint32_t n_lights;
    uint32_t n_prims;
    double background[3];
};

}  // namespace rtb200dev
