// trace_on_host.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the *device source* of the hot path (raytracinginrust_b200/csrc/device/trace.cuh: every
// __device__ function the CUDA kernels call) with g++, with the handful of CUDA intrinsics it uses
// restated below, behind loops that do what the kernel shells of kernels.cu / megakernel.inl do per
// thread.  The CPU test tier (`pytest -m "not gpu"`) then checks that source - and the tables the
// scene compiler (compile.cpp) produces - against the oracle without a GPU: first-hit ids,
// per-path radiance, images, and the structural invariants of the compiled tables.
//
// It is NOT a CPU path of the product: librtb200.so does not contain, link or load it, nothing under
// raytracinginrust_b200/ refers to it (tests/test_abi.py enforces both), and it is neither measured
// nor shipped.  The GPU parity tests (`pytest -m gpu`) stay the parity tests proper; this harness only
// lets the logic of a kernel change be checked here before GPU time is spent on it.
//
// Differences from the sm_100a build, all confined to the last ulp: g++ is run with
// -ffp-contract=off (nvcc contracts a*b+c into DFMA), libm's sin/cos/log/atan2/acos stand in for
// CUDA's, and rcp_fast's MUFU seed is the exact reciprocal.
#include <cuda_runtime.h>  // vector types and the __device__/__forceinline__ macros (empty / GCC attributes under g++)

#include <math.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

// ---- the CUDA intrinsics trace.cuh uses, for the host ----
static inline double __longlong_as_double(long long v) {
    double d;
    std::memcpy(&d, &v, sizeof d);
    return d;
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline float __double2float_rn(double x) { return (float)x; }
static inline float __double2float_ru(double x) {
    float f = (float)x;
    return ((double)f < x) ? nextafterf(f, INFINITY) : f;
}
static inline float __double2float_rd(double x) {
    float f = (float)x;
    return ((double)f > x) ? nextafterf(f, -INFINITY) : f;
}
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __int_as_float(int v) {
    float f;
    std::memcpy(&f, &v, sizeof f);
    return f;
}
template <class T>
static inline T __ldg(const T *p) { return *p; }
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
using std::max;
using std::min;

#include "../../raytracinginrust_b200/csrc/device/compile.h"
#include "../../raytracinginrust_b200/csrc/device/kernels.h"
#include "../../raytracinginrust_b200/csrc/device/trace.cuh"

// the work counter of the persistent kernels: lanes are simulated one after the other here
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long old = *p;
    *p += v;
    return old;
}
namespace rtb200dev {
inline namespace RT_VARIANT_NS {
}
}  // namespace rtb200dev

using namespace rtb200dev;

namespace {
thread_local std::string g_err;

struct HostTables {
    CompiledScene cs;
    DScene ds{};
    uint32_t features = 0;
};

bool has_pbr(const CompiledScene &cs) {
    for (const DMaterial &m : cs.materials)
        if (m.kind == RT_MAT_PBR) return true;
    return false;
}

// api.cu: hook_params / make_params (the part the device functions read)
RenderParams params(const HostTables &t, uint32_t width, uint32_t height, uint32_t max_depth, const RtRenderOpts *opts) {
    RtRenderOpts o{};
    if (opts) o = *opts;
    RenderParams P;
    std::memset(&P, 0, sizeof(P));
    P.width = width;
    P.height = height;
    P.max_depth = max_depth;
    P.seed = o.seed;
    P.integrator = o.integrator;
    P.flags = o.flags | (has_pbr(t.cs) ? RT_FLAG_TRACE_ZERO_THROUGHPUT : 0u);
    return P;
}

int fail(const std::string &m) {
    g_err = m;
    return 1;
}

// ---------------------------------------------------------------------------
// Structural invariants of the compiled tables (what the kernels rely on without checking)
// ---------------------------------------------------------------------------
struct Bounds {
    double lo[3], hi[3];
};

// f64 bounds of a primitive in its group's space (moving spheres: over their whole time range)
Bounds prim_bounds(const DPrim &p) {
    Bounds b;
    const double *d = p.d;
    switch (p.kind) {
        case PRIM_SPHERE:
            for (int a = 0; a < 3; ++a) { b.lo[a] = d[a] - fabs(d[3]); b.hi[a] = d[a] + fabs(d[3]); }
            break;
        case PRIM_MSPHERE:  // wherever center(time) puts it for a shutter time in [0, 1] (and at center0 / center1)
            for (int a = 0; a < 3; ++a) {
                const double c_at_0 = d[a] + ((0.0 - d[6]) / (d[7] - d[6])) * (d[3 + a] - d[a]);
                const double c_at_1 = d[a] + ((1.0 - d[6]) / (d[7] - d[6])) * (d[3 + a] - d[a]);
                b.lo[a] = fmin(fmin(d[a], d[3 + a]), fmin(c_at_0, c_at_1)) - fabs(d[8]);
                b.hi[a] = fmax(fmax(d[a], d[3 + a]), fmax(c_at_0, c_at_1)) + fabs(d[8]);
            }
            break;
        case PRIM_RECT: {
            int k, a0, a1;
            if (p.axis == RT_PLANE_XY) { a0 = 0; a1 = 1; k = 2; }
            else if (p.axis == RT_PLANE_XZ) { a0 = 0; a1 = 2; k = 1; }
            else { a0 = 1; a1 = 2; k = 0; }
            b.lo[a0] = d[0]; b.hi[a0] = d[1];
            b.lo[a1] = d[2]; b.hi[a1] = d[3];
            b.lo[k] = b.hi[k] = d[4];
            break;
        }
        case PRIM_TRI:
            for (int a = 0; a < 3; ++a) {
                double v0 = d[a], v1 = d[a] + d[3 + a], v2 = d[a] + d[6 + a];
                b.lo[a] = fmin(v0, fmin(v1, v2));
                b.hi[a] = fmax(v0, fmax(v1, v2));
            }
            break;
        default:  // PRIM_BOX
            for (int a = 0; a < 3; ++a) { b.lo[a] = d[a]; b.hi[a] = d[3 + a]; }
            break;
    }
    return b;
}

struct TableCheck {
    const CompiledScene &cs;
    std::vector<uint32_t> prim_owner;  // group that lists the primitive
    std::vector<uint8_t> prim_in_leaf;
    std::vector<uint8_t> node_seen;
    uint32_t max_depth = 0;
    std::string err;

    explicit TableCheck(const CompiledScene &c) : cs(c) {}

    bool bad(const std::string &m) {
        if (err.empty()) err = m;
        return false;
    }

    bool leaf(uint32_t gi, int code_node, const Bounds *within) {
        const DGroup &g = cs.groups[gi];
        uint32_t code = ~(uint32_t)code_node;
        uint32_t first = code >> 3, count = (code & 7u) + 1u;
        if (first < g.first_prim || (uint64_t)first + count > (uint64_t)g.first_prim + g.n_prims)
            return bad("group " + std::to_string(gi) + ": a leaf leaves the group's primitive range");
        for (uint32_t i = first; i < first + count; ++i) {
            if (prim_in_leaf[i]) return bad("primitive " + std::to_string(i) + " is in two leaves");
            prim_in_leaf[i] = 1;
            if (within) {
                Bounds pb = prim_bounds(cs.prims[i]);
                for (int a = 0; a < 3; ++a)
                    if (!(within->lo[a] <= pb.lo[a] && within->hi[a] >= pb.hi[a]))
                        return bad("primitive " + std::to_string(i) + " sticks out of its leaf box (axis " + std::to_string(a) + ")");
            }
        }
        return true;
    }

    // every box of the subtree lies inside `within` (the parent's box for this child)
    bool subtree(uint32_t gi, int node, const Bounds *within, uint32_t depth) {
        if (depth > max_depth) max_depth = depth;
        if (node < 0) return leaf(gi, node, within);
        if ((size_t)node >= cs.nodes.size()) return bad("node index out of range");
        if (node_seen[node]) return bad("node " + std::to_string(node) + " has two parents");
        node_seen[node] = 1;
        const DBvhNode &n = cs.nodes[node];
        Bounds b0, b1;
        for (int a = 0; a < 3; ++a) {
            b0.lo[a] = n.lo0[a]; b0.hi[a] = n.hi0[a];
            b1.lo[a] = n.lo1[a]; b1.hi[a] = n.hi1[a];
            if (!(n.lo0[a] <= n.hi0[a]) || !(n.lo1[a] <= n.hi1[a])) return bad("node " + std::to_string(node) + ": empty or NaN child box");
            if (within && !(within->lo[a] <= b0.lo[a] && within->hi[a] >= b0.hi[a] && within->lo[a] <= b1.lo[a] && within->hi[a] >= b1.hi[a]))
                return bad("node " + std::to_string(node) + ": a child box sticks out of its parent's");
        }
        return subtree(gi, n.child0, &b0, depth + 1) && subtree(gi, n.child1, &b1, depth + 1);
    }

    bool run() {
        const size_t np = cs.prims.size();
        prim_owner.assign(np, 0xFFFFFFFFu);
        prim_in_leaf.assign(np, 0);
        node_seen.assign(cs.nodes.size(), 0);
        // groups: disjoint primitive ranges that cover the table; world groups first, then the media's
        std::vector<uint8_t> group_used(cs.groups.size(), 0);
        if (cs.n_world_groups > cs.groups.size()) return bad("n_world_groups exceeds the group table");
        for (uint32_t gi = 0; gi < cs.n_world_groups; ++gi) group_used[gi] = 1;
        for (size_t mi = 0; mi < cs.media.size(); ++mi) {
            const DMedium &m = cs.media[mi];
            if ((uint64_t)m.first_group + m.n_groups > cs.groups.size()) return bad("medium: group range out of the table");
            if (m.chain >= cs.chains.size() || m.material >= cs.materials.size()) return bad("medium: chain / material out of range");
            if (!(m.density == m.density)) return bad("medium: NaN density");
            for (uint32_t gi = m.first_group; gi < m.first_group + m.n_groups; ++gi) {
                if (group_used[gi]) return bad("group " + std::to_string(gi) + " belongs to two sub-scenes");
                group_used[gi] = 1;
            }
        }
        for (size_t gi = 0; gi < cs.groups.size(); ++gi)
            if (!group_used[gi]) return bad("group " + std::to_string(gi) + " belongs to no sub-scene");
        for (size_t ci = 0; ci < cs.chains.size(); ++ci)
            if ((uint64_t)cs.chains[ci].first_op + cs.chains[ci].n_ops > cs.ops.size()) return bad("chain: op range out of the table");
        for (uint32_t gi = 0; gi < cs.groups.size(); ++gi) {
            const DGroup &g = cs.groups[gi];
            if (g.chain >= cs.chains.size()) return bad("group: chain out of range");
            if ((uint64_t)g.first_prim + g.n_prims > np) return bad("group: primitive range out of the table");
            for (uint32_t i = g.first_prim; i < g.first_prim + g.n_prims; ++i) {
                if (prim_owner[i] != 0xFFFFFFFFu) return bad("primitive " + std::to_string(i) + " is in two groups");
                prim_owner[i] = gi;
            }
            for (int a = 0; a < 3; ++a)
                if (!(g.bmin[a] <= g.bmax[a]) || !std::isfinite(g.bmin[a]) || !std::isfinite(g.bmax[a])) return bad("group: bad bounds");
            if (g.bvh_root >= 0) {
                if (cs.nodes.empty()) return bad("group with a BVH root but no nodes");
                if (!subtree(gi, g.bvh_root, nullptr, 1)) return false;
            } else if (g.bvh_root != -1) {  // a small group is a single leaf
                if (!leaf(gi, g.bvh_root, nullptr)) return false;
            } else {
                for (uint32_t i = g.first_prim; i < g.first_prim + g.n_prims; ++i) prim_in_leaf[i] = 1;  // linear scan
            }
            // every primitive of the group is reachable
            for (uint32_t i = g.first_prim; i < g.first_prim + g.n_prims; ++i)
                if (!prim_in_leaf[i]) return bad("primitive " + std::to_string(i) + " of group " + std::to_string(gi) + " is in no leaf");
        }
        for (size_t i = 0; i < np; ++i) {
            const DPrim &p = cs.prims[i];
            if (prim_owner[i] == 0xFFFFFFFFu) return bad("primitive " + std::to_string(i) + " is in no group");
            if (p.kind > PRIM_BOX) return bad("primitive: unknown kind");
            if (p.material >= cs.materials.size() || p.chain >= cs.chains.size()) return bad("primitive: material / chain out of range");
        }
        for (size_t i = 0; i < cs.nodes.size(); ++i)
            if (!node_seen[i]) return bad("node " + std::to_string(i) + " is unreachable");
        if (max_depth >= (uint32_t)kStackSize) return bad("BVH deeper than the traversal stack");
        // ranks: the reference's traversal order is a total order inside each sub-scene
        auto ranks_unique = [&](uint32_t g0, uint32_t g1, const char *what) {
            std::vector<uint32_t> r;
            for (uint32_t gi = g0; gi < g1; ++gi)
                for (uint32_t i = cs.groups[gi].first_prim; i < cs.groups[gi].first_prim + cs.groups[gi].n_prims; ++i) r.push_back(cs.prims[i].rank);
            std::sort(r.begin(), r.end());
            for (size_t i = 1; i < r.size(); ++i)
                if (r[i] == r[i - 1]) return bad(std::string(what) + ": two primitives share rank " + std::to_string(r[i]));
            return true;
        };
        if (!ranks_unique(0, cs.n_world_groups, "world")) return false;
        for (const DMedium &m : cs.media)
            if (!ranks_unique(m.first_group, m.first_group + m.n_groups, "medium boundary")) return false;
        for (const DMaterial &m : cs.materials) {
            const bool textured = m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT || m.kind == RT_MAT_ISOTROPIC || m.kind == RT_MAT_PBR;
            if (textured && m.texture >= cs.textures.size()) return bad("material: texture out of range");
        }
        for (const DImage &im : cs.images)
            if (im.offset + (uint64_t)im.width * im.height * 3 > cs.texels.size()) return bad("image: texels out of the table");
        return true;
    }
};

// FNV-1a over every compiled table: two compiles of one description must agree byte for byte (the build is parallel)
template <class T>
static void fnv(uint64_t &h, const std::vector<T> &v) {
    const unsigned char *p = reinterpret_cast<const unsigned char *>(v.data());
    for (size_t i = 0, n = v.size() * sizeof(T); i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char *toh_last_error() { return g_err.c_str(); }

int toh_scene_create(const RtSceneDesc *desc, void **out) {
    if (!desc || !out) return fail("null argument");
    *out = nullptr;
    std::unique_ptr<HostTables> t(new HostTables());
    std::string err;
    RtStatus st = compile_scene(*desc, t->cs, err);
    if (st != RT_OK) {
        g_err = err;
        return (int)st;
    }
    const CompiledScene &cs = t->cs;
    DScene &d = t->ds;
    d.prims = cs.prims.data();
    d.ops = cs.ops.data();
    d.chains = cs.chains.data();
    d.groups = cs.groups.data();
    d.nodes = cs.nodes.data();
    d.media = cs.media.data();
    d.lights = cs.lights.data();
    d.materials = cs.materials.data();
    d.textures = cs.textures.data();
    d.images = cs.images.data();
    d.perlin = cs.perlin.data();
    d.texels = cs.texels.data();
    d.n_world_groups = cs.n_world_groups;
    d.n_media = (uint32_t)cs.media.size();
    d.n_lights = (uint32_t)cs.lights.size();
    d.n_prims = (uint32_t)cs.prims.size();
    for (int a = 0; a < 3; ++a) d.background[a] = cs.background[a];
    *out = t.release();
    return 0;
}

void toh_scene_destroy(void *h) { delete (HostTables *)h; }

uint64_t toh_tables_hash(void *h) {
    const CompiledScene &cs = ((HostTables *)h)->cs;
    uint64_t x = 1469598103934665603ull;
    fnv(x, cs.prims); fnv(x, cs.ops); fnv(x, cs.chains); fnv(x, cs.groups); fnv(x, cs.nodes); fnv(x, cs.media);
    fnv(x, cs.lights); fnv(x, cs.materials); fnv(x, cs.textures); fnv(x, cs.images); fnv(x, cs.perlin); fnv(x, cs.texels);
    return x;
}

// CompiledScene::shutter_limited (compile.h): rt_render* refuse a camera whose shutter leaves [0, 1] for such a scene
int toh_shutter_limited(void *h) { return ((HostTables *)h)->cs.shutter_limited ? 1 : 0; }

// counts[0..7]: prims, groups, world groups, nodes, media, lights, chains, deepest BVH level
int toh_check_tables(void *h, uint64_t *counts) {
    const HostTables &t = *(HostTables *)h;
    TableCheck c(t.cs);
    bool ok = c.run();
    if (counts) {
        counts[0] = t.cs.prims.size();
        counts[1] = t.cs.groups.size();
        counts[2] = t.cs.n_world_groups;
        counts[3] = t.cs.nodes.size();
        counts[4] = t.cs.media.size();
        counts[5] = t.cs.lights.size();
        counts[6] = t.cs.chains.size();
        counts[7] = c.max_depth;
    }
    return ok ? 0 : fail(c.err);
}

// kernels.cu: first_hit_kernel, one iteration per thread
int toh_trace_first_hit(void *h, const RtRay *rays, uint64_t n, RtHit *hits) {
    const HostTables &t = *(HostTables *)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        Ray r;
        r.o = ld3(rays[k].origin);
        r.d = ld3(rays[k].direction);
        r.time = rays[k].time;
        Rng rng{0, 0, 0, 0};
        HitRec rec;
        RtHit o;
        std::memset(&o, 0, sizeof o);
        if (!world_hit<false, true>(t.ds, r, rng, rec)) {
            o.node = -1;
            o.material = -1;
        } else {
            o.node = rec.node;
            o.face = rec.face;
            o.material = (int32_t)rec.material;
            o.front_face = rec.front_face ? 1 : 0;
            o.t = rec.t;
            o.u = rec.u;
            o.v = rec.v;
            o.position[0] = rec.p.x; o.position[1] = rec.p.y; o.position[2] = rec.p.z;
            o.normal[0] = rec.normal.x; o.normal[1] = rec.normal.y; o.normal[2] = rec.normal.z;
        }
        hits[k] = o;
    }
    return 0;
}

// kernels.cu: camera_rays_kernel
int toh_camera_rays(const RtCamera *cam, uint32_t width, uint32_t height, const RtRenderOpts *opts, const uint32_t *px,
                    const uint32_t *py, const uint32_t *sample, uint64_t n, RtRay *rays) {
    uint32_t seed = opts ? opts->seed : 0u;
    for (uint64_t k = 0; k < n; ++k) {
        Rng rng{seed, py[k] * width + px[k], sample[k], 0};
        Ray r = camera_ray(*cam, width, height, px[k], py[k], rng);
        RtRay o;
        std::memset(&o, 0, sizeof o);
        o.origin[0] = r.o.x; o.origin[1] = r.o.y; o.origin[2] = r.o.z;
        o.direction[0] = r.d.x; o.direction[1] = r.d.y; o.direction[2] = r.d.z;
        o.time = r.time;
        rays[k] = o;
    }
    return 0;
}

// kernels.cu: path_radiance_kernel
int toh_path_radiance(void *h, const RtCamera *cam, uint32_t width, uint32_t height, uint32_t max_depth, const RtRenderOpts *opts,
                      const uint32_t *px, const uint32_t *py, const uint32_t *sample, uint64_t n, double *rgb, uint32_t *segments) {
    const HostTables &t = *(HostTables *)h;
    const RenderParams P = params(t, width, height, max_depth, opts);
    if (P.integrator == RT_INTEGRATOR_HEAD && t.cs.lights.empty()) return fail("HEAD integrator needs a non-empty light list");
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        PathState ps;
        path_begin(ps, *cam, P.width, P.height, px[k], py[k], sample[k], P.seed, P.max_depth);
        while (path_step<true>(t.ds, ps, P.integrator, P.flags)) {
        }
        rgb[3 * k] = ps.radiance.x;
        rgb[3 * k + 1] = ps.radiance.y;
        rgb[3 * k + 2] = ps.radiance.z;
        if (segments) segments[k] = ps.segments;
    }
    return 0;
}

// wavefront.inl: the search / resolve split of the wavefront stages (extend: world_search; shade: resolve_hit or
// resolve_medium + path_shade) - must be the megakernel's fused world_hit path, bit for bit.
int toh_path_radiance_split(void *h, const RtCamera *cam, uint32_t width, uint32_t height, uint32_t max_depth,
                            const RtRenderOpts *opts, const uint32_t *px, const uint32_t *py, const uint32_t *sample,
                            uint64_t n, double *rgb, uint32_t *segments) {
    const HostTables &t = *(HostTables *)h;
    const RenderParams P = params(t, width, height, max_depth, opts);
    if (P.integrator == RT_INTEGRATOR_HEAD && t.cs.lights.empty()) return fail("HEAD integrator needs a non-empty light list");
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t k = 0; k < (int64_t)n; ++k) {
        PathState ps;
        path_begin(ps, *cam, P.width, P.height, px[k], py[k], sample[k], P.seed, P.max_depth);
        for (bool alive = ps.depth_left != 0; alive;) {
            ps.segments += 1;
            Best win;
            double closest;
            world_search<true>(t.ds, ps.ray, ps.rng, win, closest);
            HitRec rec;
            const uint32_t cls = hit_class(t.ds, win.prim);
            const bool hit = cls != WF_CLS_MISS;
            if (hit) {
                win.t = closest;
                if (cls == WF_CLS_MEDIUM) resolve_medium(t.ds, ps.ray, win, closest, rec);
                else resolve_hit<false>(t.ds, ps.ray, win, closest, rec);
            }
            alive = path_shade(t.ds, ps, hit, rec, P.integrator, P.flags);
        }
        rgb[3 * k] = ps.radiance.x;
        rgb[3 * k + 1] = ps.radiance.y;
        rgb[3 * k + 2] = ps.radiance.z;
        if (segments) segments[k] = ps.segments;
    }
    return 0;
}

// megakernel.inl: what one lane does for its work items, for every pixel; samples [begin, begin+count) are
// added in sample order in f64 (the device adds per-chunk sums in chunk order: same order, other grouping).
// out: H x W x 3 f64 sums, rows top-down.  stats: paths, rays, non-finite paths.
int toh_render(void *h, const RtCamera *cam, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
               const RtRenderOpts *opts, double *out, uint64_t *stats) {
    const HostTables &t = *(HostTables *)h;
    const RenderParams P = params(t, width, height, max_depth, opts);
    if (P.integrator == RT_INTEGRATOR_HEAD && t.cs.lights.empty()) return fail("HEAD integrator needs a non-empty light list");
    uint32_t begin = opts ? opts->sample_begin : 0u;
    uint32_t count = (opts && opts->sample_count) ? opts->sample_count : (spp > begin ? spp - begin : 0u);
    unsigned long long n_paths = 0, n_rays = 0, n_bad = 0;
    const bool media = !t.cs.media.empty();
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : n_paths, n_rays, n_bad)
    for (int64_t row = 0; row < (int64_t)height; ++row) {
        for (uint32_t i = 0; i < width; ++i) {
            V3 sum = mk(0.0, 0.0, 0.0);
            for (uint32_t s = begin; s < begin + count; ++s) {
                PathState ps;
                path_begin(ps, *cam, P.width, P.height, i, P.height - 1u - (uint32_t)row, s, P.seed, P.max_depth);
                if (media) {
                    while (path_step<true>(t.ds, ps, P.integrator, P.flags)) {
                    }
                } else {
                    while (path_step<false>(t.ds, ps, P.integrator, P.flags)) {
                    }
                }
                ++n_paths;
                n_rays += ps.segments;
                if (!(std::isfinite(ps.radiance.x) && std::isfinite(ps.radiance.y) && std::isfinite(ps.radiance.z))) ++n_bad;
                sum = sum + ps.radiance;
            }
            double *dst = out + 3 * ((uint64_t)row * width + i);
            dst[0] = sum.x;
            dst[1] = sum.y;
            dst[2] = sum.z;
        }
    }
    if (stats) {
        stats[0] = n_paths;
        stats[1] = n_rays;
        stats[2] = n_bad;
    }
    return 0;
}

}  // extern "C"
#pragma GCC visibility pop
