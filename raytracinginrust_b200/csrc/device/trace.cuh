// trace.cuh — device code of the path-tracing hot path (sm_100a, f64).
//
// Everything the reference's ray_color (src/main.rs:41-120) calls, as inlined device
// functions over the flat tables of tables.h.
//
// Two kinds of arithmetic live here, and they are kept apart on purpose:
//
//  * SEARCH (s_* functions, slab tests, BVH traversal): finds WHICH primitive a ray hits
//    first.  Plain f64 with reciprocal multiplies, positive-logic accepts and a slab test
//    for cubes.  It only has to order candidates; like the BVH boxes it never produces a
//    number that reaches the image.
//  * RESOLVE / SHADE (exact_t, resolve_hit, materials, pdfs, camera): every value that
//    reaches the image — t, hit point, normal, uv, directions, pdfs, throughput — is
//    computed with the reference's own expressions in the reference's operation order
//    (true IEEE divisions included), each citing the line it restates.
//
// A different winner can only come out of SEARCH when two candidates lie within a couple
// of ulps of each other or of an interval end (DESIGN.md "Precision policy").
#pragma once
#include <cfloat>
#include <cstdint>

#include "../../../include/rtb200.h"
#include "tables.h"

// This header is compiled once per pipeline variant (variants.h): RT_VARIANT_NS names the variant,
// RT_FEAT_MASK is the feature set compiled in; feat(F_X) folds the code of unused features away.
#ifndef RT_VARIANT_NS
#define RT_VARIANT_NS vall
#endif
#ifndef RT_FEAT_MASK
#define RT_FEAT_MASK F_ALL
#endif

namespace rtb200dev {
inline namespace RT_VARIANT_NS {

#define RT_DEV __device__ __forceinline__
#ifndef RT_MICRO_OPT
// A/B switch of three instruction trims that cannot change a value (r2-k: Cornell +6.6 %, smoke +7.3 %): accept()
// as selects, no mean over a single light, the item arithmetic by reciprocal multiplication.  0 restores the old
// forms.  (Measured next and NOT kept, profiles/r2_m_instruction_trims.md: Vec3 / f64 with one shared reciprocal,
// one cosine / PI where the ONB's w is the normal, no square root for vectors of squared length exactly 1.)
#define RT_MICRO_OPT 1
#endif
__host__ __device__ __forceinline__ constexpr bool feat(uint32_t f) { return ((uint32_t)(RT_FEAT_MASK) & f) != 0u; }
#define RT_DEV_COLD static __device__ __noinline__

constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double kTMin = 0.00001;  // src/main.rs:48 (§Q1)
constexpr uint32_t kNoPrim = 0xFFFFFFFFu;
constexpr uint32_t kMediumFlag = 0x80000000u;
#define RT_INF (__longlong_as_double(0x7FF0000000000000ll))

// ---------------------------------------------------------------------------
// Vec3 (src/vec.rs)
// ---------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
RT_DEV V3 mk(double x, double y, double z) { return V3{x, y, z}; }
RT_DEV V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_DEV V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_DEV V3 operator*(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
RT_DEV V3 operator*(double s, V3 a) { return mk(s * a.x, s * a.y, s * a.z); }
RT_DEV V3 operator*(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }
// IEEE a/b.  nvcc's inline division sequence leaves its fast path for a zero numerator and
// calls a ~70-instruction subroutine; 0/b is +-0 = 0*b for finite non-zero b, so take it directly.
RT_DEV double ddiv(double a, double b) {
    if (a == 0.0) {
        double z = a * b;
        if (z == 0.0 && b != 0.0) return z;
    }
    return a / b;
}
RT_DEV V3 operator/(V3 a, double s) { return mk(ddiv(a.x, s), ddiv(a.y, s), ddiv(a.z, s)); }
RT_DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }  // vec.rs:38-40
RT_DEV double length(V3 a) { return sqrt(dot(a, a)); }                        // vec.rs:42-44
RT_DEV V3 cross(V3 a, V3 b) {                                                 // vec.rs:46-54
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
RT_DEV V3 normalized(V3 a) {  // vec.rs:56-58: self / self.length(); x/1.0 == x exactly
    double l = length(a);
    if (l == 1.0) return a;
    return a / l;
}
RT_DEV double powi2(double x) { return x * x; }
RT_DEV double powi5(double x) {
    double x2 = x * x;
    double x4 = x2 * x2;
    return x4 * x;
}
RT_DEV V3 ld3(const double *p) { return mk(p[0], p[1], p[2]); }
RT_DEV bool near_zero(V3 a) {  // vec.rs:107-110
    const double EPS = 1.0e-8;
    return fabs(a.x) < EPS && fabs(a.y) < EPS && fabs(a.z) < EPS;
}
RT_DEV V3 reflect(V3 v, V3 n) { return v + ((-dot(v, n)) * 2.0) * n; }  // vec.rs:112-114
RT_DEV V3 refract(V3 v, V3 n, double etai_over_etat) {                  // vec.rs:116-121
    double cos_theta = fmin(dot((-1.0) * v, n), 1.0);
    V3 r_out_perp = etai_over_etat * (v + cos_theta * n);
    V3 r_out_para = ((-1.0) * sqrt(fabs(1.0 - powi2(length(r_out_perp))))) * n;
    return r_out_perp + r_out_para;
}
// SEARCH-grade reciprocal: MUFU.RCP64H + two Newton steps (full precision; 1/+-0 = +-inf so that
// axis-parallel rays order correctly in the slab tests; infinities and NaN only ever reject).
RT_DEV double rcp_fast(double x) {
    double r0;
#ifdef __CUDACC__
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
#else  // g++ build of this header for the CPU test tier (tests/native): the exact reciprocal as the seed
    r0 = 1.0 / x;
#endif
    double e = fma(-x, r0, 1.0);
    double r = fma(r0, e, r0);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return x == 0.0 ? r0 : r;  // 1/+-0 = +-inf (the Newton step would turn it into NaN)
}

// ---------------------------------------------------------------------------
// Philox4x32-10, slot-addressed (DESIGN.md "Random numbers")
// ---------------------------------------------------------------------------
enum Slot : uint32_t { SLOT_PIXEL = 0, SLOT_LENS = 1, SLOT_TIME = 2, SLOT_MEDIUM = 3, SLOT_SCATTER = 4, SLOT_BALL = 5 };

struct Draw {
    double a, b;
    uint32_t bits_a, bits_b;
};
struct Rng {
    uint32_t seed, pixel, sample, bounce;
};

RT_DEV void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
RT_DEV double u53(uint32_t hi, uint32_t lo) {
    unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}
RT_DEV Draw draw(const Rng &r, uint32_t slot, uint32_t sub) {
    uint32_t c0 = r.bounce, c1 = slot, c2 = sub, c3 = r.seed;
    philox4x32_10(c0, c1, c2, c3, r.pixel, r.sample);
    Draw d;
    d.a = u53(c0, c1);
    d.b = u53(c2, c3);
    d.bits_a = c1 & 0x7FFu;
    d.bits_b = c3 & 0x7FFu;
    return d;
}
RT_DEV double gen_range(double lo, double hi, double u) { return lo + (hi - lo) * u; }

RT_DEV V3 random_in_unit_sphere(const Rng &rng) {  // vec.rs:78-85
    for (uint32_t it = 0;; ++it) {
        Draw d0 = draw(rng, SLOT_BALL, 2 * it);
        Draw d1 = draw(rng, SLOT_BALL, 2 * it + 1);
        V3 v = mk(gen_range(-1.0, 1.0, d0.a), gen_range(-1.0, 1.0, d0.b), gen_range(-1.0, 1.0, d1.a));
        if (length(v) < 1.0) return v;
    }
}
RT_DEV V3 random_in_unit_disk(const Rng &rng) {  // vec.rs:96-105
    for (uint32_t it = 0;; ++it) {
        Draw d = draw(rng, SLOT_LENS, it);
        V3 p = mk(gen_range(-1.0, 1.0, d.a), gen_range(-1.0, 1.0, d.b), 0.0);
        if (length(p) < 1.0) return p;
    }
}

// ---------------------------------------------------------------------------
// Ray (src/ray.rs) and ONB (src/onb.rs)
// ---------------------------------------------------------------------------
struct Ray {
    V3 o, d;
    double time;
};

struct ONB {
    V3 u, v, w;
};
RT_DEV ONB onb_from_w(V3 n) {  // onb.rs:8-21
    ONB o;
    o.w = normalized(n);
    V3 a = fabs(o.w.x) > 0.9 ? mk(0.0, 1.0, 0.0) : mk(1.0, 0.0, 0.0);
    o.v = normalized(cross(o.w, a));
    o.u = cross(o.w, o.v);
    return o;
}
RT_DEV V3 onb_local(const ONB &o, V3 a) { return a.x * o.u + a.y * o.v + a.z * o.w; }  // onb.rs:35-37

// ---------------------------------------------------------------------------
// Wrapper chains: Translate (translate.rs:22-30), Rotate (rotate.rs:77-106), FlipNormal (hit.rs:113-120)
// ---------------------------------------------------------------------------
// rotate.rs:82-86: into the rotated (object) frame; (a,b) axes per rotate.rs:15-21
RT_DEV V3 rot_fwd(const DOp &op, V3 v) {
    double c = op.cos_theta, s = op.sin_theta;
    if (op.axis == RT_AXIS_Y) return mk(c * v.x - s * v.z, v.y, s * v.x + c * v.z);
    if (op.axis == RT_AXIS_X) return mk(v.x, c * v.y - s * v.z, s * v.y + c * v.z);
    return mk(c * v.x - s * v.y, s * v.x + c * v.y, v.z);
}
// rotate.rs:94-98: back out of the rotated frame
RT_DEV V3 rot_back(const DOp &op, V3 v) {
    double c = op.cos_theta, s = op.sin_theta;
    if (op.axis == RT_AXIS_Y) return mk(c * v.x + s * v.z, v.y, -s * v.x + c * v.z);
    if (op.axis == RT_AXIS_X) return mk(v.x, c * v.y + s * v.z, -s * v.y + c * v.z);
    return mk(c * v.x + s * v.y, -s * v.x + c * v.y, v.z);
}
// Ray as seen below ops [first, first+n) of a chain (flips do not touch the ray).
RT_DEV void chain_ray(const DScene &sc, uint32_t first, uint32_t n, V3 &o, V3 &d) {
    for (uint32_t i = 0; i < n; ++i) {
        const DOp &op = sc.ops[first + i];
        if (op.kind == OP_TRANSLATE) {
            o = o - ld3(op.offset);  // translate.rs:23
        } else if (op.kind == OP_ROTATE) {
            o = rot_fwd(op, o);
            d = rot_fwd(op, d);
        }
    }
}

// ===========================================================================
// SEARCH: which primitive is hit first
// ===========================================================================
struct SRay {  // a ray in some group's space with its reciprocal direction
    V3 o, d, inv;
    double time;
};
struct Best {
    double t;       // search-grade t of the current winner (upper end of the search interval)
    uint32_t prim;  // index into prims, kMediumFlag|index into media, or kNoPrim
    uint32_t rank;
    int face;       // BOX: which of the six sides (cube.rs:17-25 order)
};

// Every reference test accepts t == t_max, and lists / BVH nodes keep the object visited
// last (hit.rs:64-66, bvh.rs:81-84; §Q17): an equal t only wins with a higher rank.
RT_DEV void accept(bool h, double t, uint32_t pi, uint32_t rank, int face, Best &best) {
    // selects, not a branch: a third of a warp's lanes take a candidate at a time, and the branch around four
    // moves cost more issue slots than the moves
    const bool take = h && (t < best.t || rank > best.rank);
#if !RT_MICRO_OPT
    if (take) best = Best{t, pi, rank, face};
    return;
#endif
    best.t = take ? t : best.t;
    best.prim = take ? pi : best.prim;
    best.rank = take ? rank : best.rank;
    best.face = take ? face : best.face;
}

RT_DEV V3 msphere_center(const double *pd, double time) {  // sphere.rs:144-146
    V3 c0 = ld3(pd), c1 = ld3(pd + 3);
    return c0 + ddiv(time - pd[6], pd[7] - pd[6]) * (c1 - c0);
}

// Both roots of a sphere (search-grade), then the choice Sphere::hit makes for an interval (sphere.rs:56-73): the two
// halves are separate so that a medium's two boundary queries can share the first (world_search).
struct SphereRoots {
    double t_near, t_far;
    bool valid;
};
RT_DEV SphereRoots sphere_roots(const SRay &r, V3 center, double radius) {
    V3 oc = r.o - center;
    double a = dot(r.d, r.d);
    double half_b = dot(oc, r.d);
    double c = dot(oc, oc) - radius * radius;
    double disc = half_b * half_b - a * c;
    SphereRoots s;
    s.valid = disc >= 0.0;
    double sq = sqrt(disc);
    double ia = rcp_fast(a);
    s.t_near = (-half_b - sq) * ia;
    s.t_far = (-half_b + sq) * ia;
    return s;
}
RT_DEV void s_sphere_from(const SphereRoots &s, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    if (!s.valid) return;
    double t = s.t_near;
    if (!(t >= t_min && t <= best.t)) t = s.t_far;
    accept(t >= t_min && t <= best.t, t, pi, rank, 0, best);
}
// (the leaf test proper leaves before the square root when the ray misses, and makes the far root only if the
// near one is outside the interval: RTiOW's leaves are sphere tests, most of them misses)
RT_DEV void s_sphere(const SRay &r, V3 center, double radius, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    V3 oc = r.o - center;
    double a = dot(r.d, r.d);
    double half_b = dot(oc, r.d);
    double c = dot(oc, oc) - radius * radius;
    double disc = half_b * half_b - a * c;
    if (!(disc >= 0.0)) return;
    double sq = sqrt(disc);
    double ia = rcp_fast(a);
    double t = (-half_b - sq) * ia;
    if (!(t >= t_min && t <= best.t)) t = (-half_b + sq) * ia;
    accept(t >= t_min && t <= best.t, t, pi, rank, 0, best);
}
RT_DEV void s_rect(const SRay &r, uint32_t plane, const double *pd, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    double a0 = pd[0], a1 = pd[1], b0 = pd[2], b1 = pd[3], k = pd[4];
    double t, a, b;
    if (plane == RT_PLANE_XZ) {
        t = (k - r.o.y) * r.inv.y;
        a = fma(t, r.d.x, r.o.x);
        b = fma(t, r.d.z, r.o.z);
    } else if (plane == RT_PLANE_YZ) {
        t = (k - r.o.x) * r.inv.x;
        a = fma(t, r.d.y, r.o.y);
        b = fma(t, r.d.z, r.o.z);
    } else {
        t = (k - r.o.z) * r.inv.z;
        a = fma(t, r.d.x, r.o.x);
        b = fma(t, r.d.y, r.o.y);
    }
    accept(t >= t_min && t <= best.t && a >= a0 && a <= a1 && b >= b0 && b <= b1, t, pi, rank, 0, best);
}
// Cube = six AARects in a list (cube.rs:17-25,35-37).  For a convex box the closest accepted
// side is the entry point if it lies in the interval, else the exit point: a slab test.
struct BoxSlab {
    double t_in, t_out;
    int f_in, f_out;
    bool valid;
};
RT_DEV BoxSlab box_slab(const SRay &r, const double *pd) {
    double x0 = (pd[0] - r.o.x) * r.inv.x, x1 = (pd[3] - r.o.x) * r.inv.x;
    double y0 = (pd[1] - r.o.y) * r.inv.y, y1 = (pd[4] - r.o.y) * r.inv.y;
    double z0 = (pd[2] - r.o.z) * r.inv.z, z1 = (pd[5] - r.o.z) * r.inv.z;
    // near/far per axis and the side index it belongs to (cube.rs order: +z 0, -z 1, +y 2, -y 3, +x 4, -x 5)
    bool sx = x0 <= x1, sy = y0 <= y1, sz = z0 <= z1;
    double nx = sx ? x0 : x1, fx = sx ? x1 : x0;
    double ny = sy ? y0 : y1, fy = sy ? y1 : y0;
    double nz = sz ? z0 : z1, fz = sz ? z1 : z0;
    BoxSlab s;
    s.t_in = nx;
    s.f_in = sx ? 5 : 4;
    if (ny > s.t_in) { s.t_in = ny; s.f_in = sy ? 3 : 2; }
    if (nz > s.t_in) { s.t_in = nz; s.f_in = sz ? 1 : 0; }
    s.t_out = fx;
    s.f_out = sx ? 4 : 5;
    if (fy < s.t_out) { s.t_out = fy; s.f_out = sy ? 2 : 3; }
    if (fz < s.t_out) { s.t_out = fz; s.f_out = sz ? 0 : 1; }
    s.valid = s.t_in <= s.t_out;  // false: misses the box (NaN: a ray parallel to a slab on its plane -> no hit)
    return s;
}
RT_DEV void s_box_from(const BoxSlab &s, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    if (!s.valid) return;
    bool in_ok = s.t_in >= t_min && s.t_in <= best.t;
    double t = in_ok ? s.t_in : s.t_out;
    int f = in_ok ? s.f_in : s.f_out;
    accept(t >= t_min && t <= best.t, t, pi, rank, f, best);
}
RT_DEV void s_box(const SRay &r, const double *pd, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    s_box_from(box_slab(r, pd), t_min, pi, rank, best);
}
RT_DEV void s_tri(const SRay &r, const double *pd, double t_min, uint32_t pi, uint32_t rank, Best &best) {
    V3 s = r.o - ld3(pd);
    V3 e1 = ld3(pd + 3), e2 = ld3(pd + 6);
    V3 s1 = cross(r.d, e2);
    V3 s2 = cross(s, e1);
    double inv = rcp_fast(dot(s1, e1));
    double t = dot(s2, e2) * inv;
    double b1 = dot(s1, s) * inv;
    double b2 = dot(s2, r.d) * inv;
    accept(t >= t_min && t <= best.t && b1 >= 0.0 && b2 >= 0.0 && (1.0 - b1 - b2) >= 0.0, t, pi, rank, 0, best);
}

RT_DEV void s_prim(const DScene &sc, uint32_t pi, const SRay &r, double t_min, Best &best) {
    const DPrim &p = sc.prims[pi];
    uint32_t kind = p.kind, rank = p.rank;
    if (feat(F_RECT) && kind == PRIM_RECT) s_rect(r, p.axis, p.d, t_min, pi, rank, best);
    else if (feat(F_BOX) && kind == PRIM_BOX) s_box(r, p.d, t_min, pi, rank, best);
    else if (feat(F_SPHERE) && kind == PRIM_SPHERE) s_sphere(r, ld3(p.d), p.d[3], t_min, pi, rank, best);
    else if (feat(F_TRI) && kind == PRIM_TRI) s_tri(r, p.d, t_min, pi, rank, best);
    else if (feat(F_MSPHERE)) s_sphere(r, msphere_center(p.d, r.time), p.d[8], t_min, pi, rank, best);
}

// Conservative slab test against [t_min, t_max] (culling only; NaN operands are ignored by fmin/fmax).
RT_DEV bool slab(V3 o, V3 inv, const double *lo, const double *hi, double t_min, double t_max, double &t_entry) {
    double tx0 = (lo[0] - o.x) * inv.x, tx1 = (hi[0] - o.x) * inv.x;
    double ty0 = (lo[1] - o.y) * inv.y, ty1 = (hi[1] - o.y) * inv.y;
    double tz0 = (lo[2] - o.z) * inv.z, tz1 = (hi[2] - o.z) * inv.z;
    double tin = fmax(fmax(fmin(tx0, tx1), fmin(ty0, ty1)), fmax(fmin(tz0, tz1), t_min));
    double tout = fmin(fmin(fmax(tx0, tx1), fmax(ty0, ty1)), fmin(fmax(tz0, tz1), t_max));
    t_entry = tin;
    return tin <= tout;
}

// ---- fp32 node boxes: conservative by construction --------------------------------------
// A plane distance (p - o) / d is evaluated as one FFMA, p * i - c, with i ~ 1/d in fp32 and c ~ o * i taken
// from the f64 origin and rounded in the direction that makes the result err to the safe side: the plane a ray
// ENTERS through gets c rounded up (distance under-estimated), the plane it LEAVES through c rounded down.  Which
// plane is which follows from the sign of d, so the two constants are sorted per ray into c_lo (paired with the
// box's lower bound) and c_hi.  What is left is the rounding of i (d to fp32, the reciprocal: 2^-23 relative in
// all) and of the FFMA itself (2^-24 of the result) - relative errors of the DISTANCE, because p * i - o * i =
// (p - o) * i holds exactly before rounding - and the 2^-19 relative slack of slab2f dominates both.  The bounds
// themselves are the f64 bounds rounded outward (compile.cpp).  So a box can only grow.
// An axis the ray runs (almost) parallel to, |d| < 2^-60, never culls: i = 0, c = -+inf gives (-inf, +inf).
struct FRay {
    float ix, iy, iz;        // ~ 1 / d
    float clx, cly, clz;     // subtracted from lower-bound * i
    float chx, chy, chz;     // subtracted from upper-bound * i
};
RT_DEV void fray_axis(double o, double d, float &i, float &c_lo, float &c_hi) {
    const float df = __double2float_rn(d);
    // 2^-60 <= |d| <= 2^60 (NaN fails the test too): outside, 1/d or o/d leave fp32's normal range
    if (!(fabsf(df) >= 8.673617379884035e-19f && fabsf(df) <= 1.152921504606847e18f)) {
        i = 0.0f;
        c_lo = __int_as_float(0x7f800000);        // lo * 0 - (+inf) = -inf
        c_hi = __int_as_float((int)0xff800000u);  // hi * 0 - (-inf) = +inf
        return;
    }
    i = __frcp_rn(df);
    // o * i in f64, pushed outward by more than its own rounding (2^-53 relative) before the directed conversions,
    // so that up >= o * i >= dn holds for the exact product
    const double oi = o * (double)i, m = fabs(oi) * 2.220446049250313e-16;  // 2^-52
    const float up = __double2float_ru(oi + m), dn = __double2float_rd(oi - m);
    // d > 0: the lower bound is the entry plane (c up), the upper bound the exit plane (c down); d < 0: the reverse
    c_lo = df > 0.0f ? up : dn;
    c_hi = df > 0.0f ? dn : up;
}
RT_DEV FRay make_fray(const SRay &r) {
    FRay f;
    fray_axis(r.o.x, r.d.x, f.ix, f.clx, f.chx);
    fray_axis(r.o.y, r.d.y, f.iy, f.cly, f.chy);
    fray_axis(r.o.z, r.d.z, f.iz, f.clz, f.chz);
    return f;
}
RT_DEV bool slab2f(const FRay &f, float lx, float ly, float lz, float hx, float hy, float hz, float t_min, float t_max,
                   float &t_entry) {
    const float kSlack = 1.9073486328125e-06f;  // 2^-19
    float ax = fmaf(lx, f.ix, -f.clx), bx = fmaf(hx, f.ix, -f.chx);
    float ay = fmaf(ly, f.iy, -f.cly), by = fmaf(hy, f.iy, -f.chy);
    float az = fmaf(lz, f.iz, -f.clz), bz = fmaf(hz, f.iz, -f.chz);
    float tin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    float tout = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tin = fmaxf(fmaf(-fabsf(tin), kSlack, tin), t_min);
    tout = fminf(fmaf(fabsf(tout), kSlack, tout), t_max);
    t_entry = tin;
    return tin <= tout;
}

// g.bvh_root >= 0: a BVH node; < 0: a leaf code (a small group is a single leaf).  "while-while"
// traversal: every lane first descends inner nodes until it holds a leaf (or is done), then the
// leaves are tested together, which keeps the warp together in both phases.
RT_DEV void trace_group(const DScene &sc, const DGroup &g, const SRay &r, double t_min, Best &best) {
    const int kDone = (int)0x80000000;
    int stack[kStackSize];
    int sp = 0;
    int node = g.bvh_root;
    FRay f;
    float t_min_f = 0.f, t_max_f = 0.f;
    if (!feat(F_BVH)) node = node >= 0 ? kDone : node;
    if (node >= 0) {
        f = make_fray(r);
        t_min_f = __double2float_rd(t_min);
        t_max_f = __double2float_ru(best.t);
    }
    while (node != kDone) {
        while (feat(F_BVH) && node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(sc.nodes + node);
            float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
            int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 3));
            float e0, e1;
            bool h0 = slab2f(f, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_min_f, t_max_f, e0);
            bool h1 = slab2f(f, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_min_f, t_max_f, e1);
            if (h0 && h1) {
                bool swap = e1 < e0;
                int near_c = swap ? ch.y : ch.x, far_c = swap ? ch.x : ch.y;
                if (sp < kStackSize) stack[sp++] = far_c;
                node = near_c;
            } else if (h0) {
                node = ch.x;
            } else if (h1) {
                node = ch.y;
            } else {
                node = sp ? stack[--sp] : kDone;
            }
        }
        if (node < 0 && node != kDone) {
            uint32_t code = ~(uint32_t)node;
            uint32_t first = code >> 3, count = (code & 7u) + 1u;
            double before = best.t;
            for (uint32_t i = 0; i < count; ++i) s_prim(sc, first + i, r, t_min, best);
            if (best.t != before) t_max_f = __double2float_ru(best.t);
            node = sp ? stack[--sp] : kDone;
        }
    }
}

// The ray in a group's space (search-grade: one composed affine map instead of the op chain).
RT_DEV SRay group_ray(const DGroup &g, const Ray &ray, V3 inv) {
    SRay r;
    r.o = ray.o;
    r.d = ray.d;
    r.time = ray.time;
    r.inv = inv;
    if (g.flags & GROUP_XFORM) {
        const double *m = g.m;
        V3 o = ray.o, d = ray.d;
        r.o = mk(fma(m[0], o.x, fma(m[1], o.y, fma(m[2], o.z, g.t[0]))), fma(m[3], o.x, fma(m[4], o.y, fma(m[5], o.z, g.t[1]))),
                 fma(m[6], o.x, fma(m[7], o.y, fma(m[8], o.z, g.t[2]))));
        if (g.flags & GROUP_ROTATED) {
            r.d = mk(fma(m[0], d.x, fma(m[1], d.y, m[2] * d.z)), fma(m[3], d.x, fma(m[4], d.y, m[5] * d.z)),
                     fma(m[6], d.x, fma(m[7], d.y, m[8] * d.z)));
            r.inv = mk(rcp_fast(r.d.x), rcp_fast(r.d.y), rcp_fast(r.d.z));
        }
    }
    return r;
}
// One group of a sub-scene for the ray given in the outermost space: transform, trace.
RT_DEV void trace_one_group(const DScene &sc, const DGroup &g, const Ray &ray, V3 inv, double t_min, Best &best) {
    trace_group(sc, g, group_ray(g, ray, inv), t_min, best);
}

// Closest hit over a sub-scene (a range of groups) for the ray given in the outermost space.
RT_DEV void trace_groups(const DScene &sc, uint32_t first_group, uint32_t n_groups, const Ray &ray, V3 inv, double t_min,
                         Best &best) {
#pragma unroll 1
    for (uint32_t gi = 0; gi < n_groups; ++gi) {
        const DGroup &g = sc.groups[first_group + gi];
        double e;
        // a one- or two-primitive group is cheaper to test than to cull
        if ((g.flags & GROUP_CULL) && !slab(ray.o, inv, g.bmin, g.bmax, t_min, best.t, e)) continue;
        trace_one_group(sc, g, ray, inv, t_min, best);
    }
}

// ===========================================================================
// RESOLVE: the winner's HitRecord (hit.rs:9-24) in reference arithmetic
// ===========================================================================
struct HitRec {
    V3 p, normal;
    double t, u, v;
    bool front_face;
    uint32_t material;
    int32_t node, face;
};

RT_DEV void set_face_normal(HitRec &rec, V3 ray_dir, V3 outward_normal) {  // hit.rs:34-41
    rec.front_face = dot(ray_dir, outward_normal) < 0.0;
    rec.normal = rec.front_face ? outward_normal : (-1.0) * outward_normal;
}
RT_DEV void get_sphere_uv(V3 p, double &u, double &v) {  // sphere.rs:11-25
    double phi = atan2(-p.z, p.x) + kPi;
    double theta = acos(-p.y);
    u = phi / (2.0 * kPi);
    v = theta / kPi;
}
// The six sides of a Cube in the order of cube.rs:17-25; pd = minx miny minz maxx maxy maxz
RT_DEV void box_face(const double *pd, int face, uint32_t &plane, double &a0, double &a1, double &b0, double &b1, double &k) {
    if (face < 2) {
        plane = RT_PLANE_XY; a0 = pd[0]; a1 = pd[3]; b0 = pd[1]; b1 = pd[4]; k = face == 0 ? pd[5] : pd[2];
    } else if (face < 4) {
        plane = RT_PLANE_XZ; a0 = pd[0]; a1 = pd[3]; b0 = pd[2]; b1 = pd[5]; k = face == 2 ? pd[4] : pd[1];
    } else {
        plane = RT_PLANE_YZ; a0 = pd[1]; a1 = pd[4]; b0 = pd[2]; b1 = pd[5]; k = face == 4 ? pd[3] : pd[0];
    }
}
// The nearest root in [t_min, inf) exactly as Sphere::hit computes it (sphere.rs:56-73 == :150-167).
// For the winner of a search the reference's t_max test cannot have rejected the chosen root.
RT_DEV double sphere_root_exact(V3 o, V3 d, V3 center, double radius, double t_min) {
    V3 oc = o - center;
    double a = powi2(length(d));
    double half_b = dot(oc, d);
    double c = powi2(length(oc)) - powi2(radius);
    double discriminant = powi2(half_b) - a * c;
    double sqrt_d = sqrt(fmax(discriminant, 0.0));  // search said "hit": a last-ulp negative stays a graze
    double root = (-half_b - sqrt_d) / a;
    if (root < t_min) root = (-half_b + sqrt_d) / a;
    return root;
}
// Wrapper post-processing, innermost first (translate.rs:26, rotate.rs:88-104, hit.rs:116).
// d_obj: the ray direction below the WHOLE chain (what chain_ray gives for all n_ops ops).
RT_DEV void chain_post(const DScene &sc, uint32_t chain, const Ray &world, V3 d_obj, HitRec &rec) {
    DChain c = sc.chains[chain];
    bool inner_only_flips = true;  // no Translate/Rotate below op i: its object-space ray is the chain's
    for (uint32_t i = c.n_ops; i-- > 0;) {
        const DOp &op = sc.ops[c.first_op + i];
        if (op.kind == OP_FLIP) {
            rec.front_face = !rec.front_face;
        } else if (op.kind == OP_TRANSLATE) {
            rec.p = rec.p + ld3(op.offset);
            inner_only_flips = false;
        } else {
            rec.p = rot_back(op, rec.p);
            V3 n = rot_back(op, rec.normal);
            // §Q3: face orientation is recomputed with the ray of THIS Rotate's object space
            V3 d = d_obj;
            if (!inner_only_flips) {
                V3 o = world.o;
                d = world.d;
                chain_ray(sc, c.first_op, i + 1, o, d);
            }
            set_face_normal(rec, d, n);
            inner_only_flips = false;
        }
    }
}

// The ray as the primitive's own hit() sees it: below every wrapper of its chain.
RT_DEV void object_ray(const DScene &sc, uint32_t chain, const Ray &world, V3 &o, V3 &d) {
    o = world.o;
    d = world.d;
    DChain c = sc.chains[chain];
    chain_ray(sc, c.first_op, c.n_ops, o, d);
}

// t: the winner's reference-arithmetic t (exact_t); (o, d): object_ray of the winner's chain.
template <bool WANT_UV>
RT_DEV void resolve_hit_obj(const DScene &sc, const Ray &world, const Best &best, double t, V3 o, V3 d, HitRec &rec) {
    rec.t = t;
    rec.u = 0.0;
    rec.v = 0.0;
    const DPrim &p = sc.prims[best.prim];
    rec.material = p.material;
    rec.node = p.node;
    rec.face = best.face;
    bool want_uv = WANT_UV || sc.materials[p.material].needs_uv;
    uint32_t kind = p.kind;
    if (kind == PRIM_RECT || kind == PRIM_BOX) {  // rect.rs:49-78
        uint32_t plane = p.axis;
        double a0 = p.d[0], a1 = p.d[1], b0 = p.d[2], b1 = p.d[3], k = p.d[4];
        if (kind == PRIM_BOX) box_face(p.d, best.face, plane, a0, a1, b0, b1, k);
        double oa, da, ob, db;
        V3 normal;
        if (plane == RT_PLANE_XZ) { oa = o.x; da = d.x; ob = o.z; db = d.z; normal = mk(0.0, 1.0, 0.0); }
        else if (plane == RT_PLANE_YZ) { oa = o.y; da = d.y; ob = o.z; db = d.z; normal = mk(1.0, 0.0, 0.0); }
        else { oa = o.x; da = d.x; ob = o.y; db = d.y; normal = mk(0.0, 0.0, 1.0); }
        if (want_uv) {
            double a = oa + t * da;
            double b = ob + t * db;
            rec.u = (a - a0) / (a1 - a0);
            rec.v = (b - b0) / (b1 - b0);
        }
        rec.p = o + t * d;  // r.at(t), ray.rs:26-28
        set_face_normal(rec, d, normal);
    } else if (feat(F_TRI) && kind == PRIM_TRI) {  // tri.rs:24-54; p.d = v0 e1 e2 n
        if (want_uv) {
            V3 s = o - ld3(p.d);
            V3 e1 = ld3(p.d + 3), e2 = ld3(p.d + 6);
            V3 s1 = cross(d, e2);
            V3 s2 = cross(s, e1);
            double s1_e1 = dot(s1, e1);
            rec.u = dot(s1, s) / s1_e1;
            rec.v = dot(s2, d) / s1_e1;
        }
        rec.p = o + t * d;
        set_face_normal(rec, d, ld3(p.d + 9));
    } else if (feat(F_SPHERE | F_MSPHERE)) {  // sphere.rs:56-94, :150-188
        V3 center = kind == PRIM_SPHERE ? ld3(p.d) : msphere_center(p.d, world.time);
        double radius = kind == PRIM_SPHERE ? p.d[3] : p.d[8];
        rec.p = o + t * d;
        V3 outward_normal = (rec.p - center) / radius;
        set_face_normal(rec, d, outward_normal);
        if (want_uv) get_sphere_uv(outward_normal, rec.u, rec.v);
    }
    chain_post(sc, p.chain, world, d, rec);
}
template <bool WANT_UV>
RT_DEV void resolve_hit(const DScene &sc, const Ray &world, const Best &best, double t, HitRec &rec) {
    V3 o, d;
    object_ray(sc, sc.prims[best.prim].chain, world, o, d);
    resolve_hit_obj<WANT_UV>(sc, world, best, t, o, d, rec);
}

// The reference-arithmetic t of a search winner (the part of the primitive's hit() that produces
// rec.t), for the object-space ray (o, d).  t_min: the lower end of the interval the search ran with.
RT_DEV double exact_t_obj(const DScene &sc, const Best &best, V3 o, V3 d, double time, double t_min) {
    const DPrim &p = sc.prims[best.prim];
    uint32_t kind = p.kind;
    if (kind == PRIM_RECT || kind == PRIM_BOX) {  // rect.rs:51
        uint32_t plane = p.axis;
        double k = p.d[4];
        if (kind == PRIM_BOX) {
            double a0, a1, b0, b1;
            box_face(p.d, best.face, plane, a0, a1, b0, b1, k);
        }
        double ok = plane == RT_PLANE_XZ ? o.y : (plane == RT_PLANE_YZ ? o.x : o.z);
        double dk = plane == RT_PLANE_XZ ? d.y : (plane == RT_PLANE_YZ ? d.x : d.z);
        return (k - ok) / dk;
    }
    if (!feat(F_SPHERE | F_MSPHERE | F_TRI)) return 0.0;
    if (feat(F_TRI) && kind == PRIM_TRI) {  // tri.rs:26-33
        V3 s = o - ld3(p.d);
        V3 e1 = ld3(p.d + 3), e2 = ld3(p.d + 6);
        V3 s1 = cross(d, e2);
        V3 s2 = cross(s, e1);
        return dot(s2, e2) / dot(s1, e1);
    }
    if (!feat(F_SPHERE | F_MSPHERE)) return 0.0;
    V3 center = kind == PRIM_SPHERE ? ld3(p.d) : msphere_center(p.d, time);
    return sphere_root_exact(o, d, center, kind == PRIM_SPHERE ? p.d[3] : p.d[8], t_min);
}
RT_DEV double exact_t(const DScene &sc, const Ray &world, const Best &best, double t_min) {
    V3 o, d;
    object_ray(sc, sc.prims[best.prim].chain, world, o, d);
    return exact_t_obj(sc, best, o, d, world.time, t_min);
}

RT_DEV void resolve_medium(const DScene &sc, const Ray &world, const Best &best, double t, HitRec &rec) {  // medium.rs:46-56
    const DMedium &m = sc.media[best.prim & ~kMediumFlag];
    V3 o = world.o, d = world.d;
    DChain c = sc.chains[m.chain];
    chain_ray(sc, c.first_op, c.n_ops, o, d);
    rec.t = t;
    rec.u = 0.0;
    rec.v = 0.0;
    rec.p = o + t * d;
    rec.front_face = false;
    rec.normal = mk(1.0, 0.0, 0.0);
    rec.material = m.material;
    rec.node = m.node;
    rec.face = 0;
    chain_post(sc, m.chain, world, d, rec);
}

// world.hit(ray, 0.00001, inf) (main.rs:48).  One query loop with a single search call site:
// query 0 is the world's surfaces; then, per ConstantMedium, the two boundary queries of
// medium.rs:29-30.  Media are visited after the surfaces: with slot-addressed draws the outcome
// does not depend on list order.  `closest` is closest_so_far (hit.rs:61-66) in reference arithmetic.
// The SEARCH half of world_hit: which primitive (or medium) wins, and closest_so_far in reference
// arithmetic.  Returns false on a miss.
template <bool WITH_MEDIA>
RT_DEV bool world_search(const DScene &sc, const Ray &r, const Rng &rng, Best &win, double &closest) {
    V3 inv = mk(rcp_fast(r.d.x), rcp_fast(r.d.y), rcp_fast(r.d.z));
    win = Best{RT_INF, kNoPrim, 0, 0};
    closest = RT_INF;
    double t1 = 0.0;
    bool have_t1 = false;
    const uint32_t nq = 1u + (WITH_MEDIA ? 2u * sc.n_media : 0u);
#pragma unroll 1
    for (uint32_t q = 0; q < nq; ++q) {
        uint32_t fg = 0, ng = sc.n_world_groups, mi = 0;
        double t_min = kTMin;
        Best b{RT_INF, kNoPrim, 0, 0};
        bool second = false;
        double te = 0.0;
        bool found = false, searched = false;
        if (q > 0) {
            mi = (q - 1u) >> 1;
            second = ((q - 1u) & 1u) != 0u;
            if (second && !have_t1) continue;
            const DMedium &m = sc.media[mi];
            fg = m.first_group;
            ng = m.n_groups;
            t_min = second ? t1 + 0.0001 : -DBL_MAX;  // boundary.hit(r, -MAX, MAX) ; boundary.hit(r, hit1.t + 0.0001, MAX)
            b.t = DBL_MAX;
            if (!second && m.convex_prim != kNoPrim) {
                // The boundary is one box or sphere (tables.h: DMedium::convex_prim): both queries come from ONE slab /
                // root computation on ONE transformed ray, with the selection of s_box / s_sphere applied once per
                // interval - the same comparisons on the same numbers as two searches of the boundary sub-scene,
                // without the second ray transform, primitive test and object-space ray.  This iteration answers the
                // first query and, if it hits, the second one too; the next iteration (q + 1) is skipped.
                const uint32_t pi = m.convex_prim;
                const DPrim &p = sc.prims[pi];
                const SRay gr = group_ray(sc.groups[fg], r, inv);
                V3 go, gd;
                object_ray(sc, p.chain, r, go, gd);
                Best b2{DBL_MAX, kNoPrim, 0, 0};
                if (p.kind == PRIM_BOX) {
                    const BoxSlab slab = box_slab(gr, p.d);
                    s_box_from(slab, t_min, pi, p.rank, b);
                    if (b.prim != kNoPrim) {
                        t1 = exact_t_obj(sc, b, go, gd, r.time, t_min);
                        s_box_from(slab, t1 + 0.0001, pi, p.rank, b2);
                    }
                } else {
                    const SphereRoots roots = sphere_roots(gr, ld3(p.d), p.d[3]);
                    s_sphere_from(roots, t_min, pi, p.rank, b);
                    if (b.prim != kNoPrim) {
                        t1 = exact_t_obj(sc, b, go, gd, r.time, t_min);
                        s_sphere_from(roots, t1 + 0.0001, pi, p.rank, b2);
                    }
                }
                have_t1 = false;  // the second query's iteration has nothing left to do
                ++q;
                if (b.prim == kNoPrim || b2.prim == kNoPrim) continue;
                te = exact_t_obj(sc, b2, go, gd, r.time, t1 + 0.0001);
                second = true;
                found = true;
                searched = true;
            }
        }
        if (!searched) {
            trace_groups(sc, fg, ng, r, inv, t_min, b);
            found = b.prim != kNoPrim;
            if (q > 0 && !second) have_t1 = found;
            if (!found) continue;
            te = exact_t(sc, r, b, t_min);
        }
        if (q == 0) {
            win = b;
            closest = te;
        } else if (!second) {
            t1 = te;
        } else {  // medium.rs:32-58
            const DMedium &m = sc.media[mi];
            double h1 = t1, h2 = te;
            if (h1 < kTMin) h1 = kTMin;
            if (h2 > closest) h2 = closest;
            if (h1 < h2) {
                // r.direction().length() of the ray the medium sees (below its own wrappers)
                V3 o = r.o, d = r.d;
                DChain c = sc.chains[m.chain];
                chain_ray(sc, c.first_op, c.n_ops, o, d);
                double len = length(d);
                double distance_inside_boundary = (h2 - h1) * len;
                Draw dr = draw(rng, SLOT_MEDIUM, (uint32_t)m.node);
                double hit_distance = -(1.0 / m.density) * log(dr.a);
                if (hit_distance < distance_inside_boundary) {
                    closest = h1 + hit_distance / len;
                    win.prim = kMediumFlag | mi;
                    win.rank = m.rank;
                    win.face = 0;
                }
            }
        }
    }
    return win.prim != kNoPrim;
}

template <bool WITH_MEDIA, bool WANT_UV>
RT_DEV bool world_hit(const DScene &sc, const Ray &r, const Rng &rng, HitRec &rec) {
    Best win{RT_INF, kNoPrim, 0, 0};
    if (!WITH_MEDIA) {  // one query; t and the record come from one object-space ray
        V3 inv = mk(rcp_fast(r.d.x), rcp_fast(r.d.y), rcp_fast(r.d.z));
        trace_groups(sc, 0, sc.n_world_groups, r, inv, kTMin, win);
        if (win.prim == kNoPrim) return false;
        V3 o, d;
        object_ray(sc, sc.prims[win.prim].chain, r, o, d);
        resolve_hit_obj<WANT_UV>(sc, r, win, exact_t_obj(sc, win, o, d, r.time, kTMin), o, d, rec);
        return true;
    }
    double closest;
    if (!world_search<WITH_MEDIA>(sc, r, rng, win, closest)) return false;
    if (win.prim & kMediumFlag) resolve_medium(sc, r, win, closest, rec);
    else resolve_hit<WANT_UV>(sc, r, win, closest, rec);
    return true;
}

// ---------------------------------------------------------------------------
// Textures (texture.rs, perlin.rs)
// ---------------------------------------------------------------------------
RT_DEV unsigned long long as_usize(double x) {  // Rust `as usize`: saturating, NaN -> 0
    if (!(x == x) || x <= 0.0) return 0ull;
    if (x >= 18446744073709551615.0) return 0xFFFFFFFFFFFFFFFFull;
    return (unsigned long long)x;
}
RT_DEV double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

RT_DEV double perlin_noise(const DPerlin &t, V3 p, double scale) {  // perlin.rs:77-109 + :39-56 (§Q15)
    double fx = floor(scale * p.x), fy = floor(scale * p.y), fz = floor(scale * p.z);
    double u = scale * p.x - fx, v = scale * p.y - fy, w = scale * p.z - fz;
    u = u * u * (3.0 - 2.0 * u);
    v = v * v * (3.0 - 2.0 * v);
    w = w * w * (3.0 - 2.0 * w);
    unsigned long long i = as_usize(fx), j = as_usize(fy), k = as_usize(fz);
    double uu = u * u * (3.0 - 2.0 * u);
    double vv = v * v * (3.0 - 2.0 * v);
    double ww = w * w * (3.0 - 2.0 * w);
    double accum = 0.0;
#pragma unroll
    for (int di = 0; di < 2; ++di)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int dk = 0; dk < 2; ++dk) {
                uint32_t idx = (t.perm_x[(i + di) & 255] ^ t.perm_y[(j + dj) & 255] ^ t.perm_z[(k + dk) & 255]) & 255u;
                V3 c = ld3(t.ranvec + 3 * idx);
                V3 weight = mk(u - (double)di, v - (double)dj, w - (double)dk);
                accum += ((double)di * uu + (double)(1 - di) * (1.0 - uu)) * ((double)dj * vv + (double)(1 - dj) * (1.0 - vv)) *
                         ((double)dk * ww + (double)(1 - dk) * (1.0 - ww)) * dot(c, weight);
            }
    return accum;
}
// Cold: only textured scenes reach these; keeping them out of line keeps the hot loop small.
RT_DEV_COLD double perlin_turb(const DPerlin *t, double px, double py, double pz, double scale, int depth) {  // perlin.rs:111-121
    double accum = 0.0;
    V3 temp_p = mk(px, py, pz);
    double weight = 1.0;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlin_noise(*t, temp_p, scale);
        weight *= 0.5;
        temp_p = temp_p * 2.0;
    }
    return fabs(accum);
}
RT_DEV_COLD void image_texel(const DImage *images, const uint8_t *texels, uint32_t id, double u, double v, double *rgb) {  // texture.rs:99-121
    DImage im = images[id];
    unsigned long long width = im.width, height = im.height;
    unsigned long long i = as_usize(clampd(u, 0.0, 1.0) * (double)width);
    unsigned long long j = as_usize(clampd(1.0 - v, 0.0, 1.0) * (double)height);
    if (i > width - 1) i = width - 1;
    if (j > height - 1) j = height - 1;
    const uint8_t *px = texels + im.offset + 3 * i + 3 * width * j;
    rgb[0] = (double)px[0] / 255.0;
    rgb[1] = (double)px[1] / 255.0;
    rgb[2] = (double)px[2] / 255.0;
}
RT_DEV V3 texture_value(const DScene &sc, uint32_t id, double u, double v, V3 p) {
    for (int guard = 0; guard < 16; ++guard) {
        const DTexture &t = sc.textures[id];
        uint32_t kind = t.kind;
        if (!feat(F_TEX) || kind == RT_TEX_CONSTANT) return ld3(t.color);  // texture.rs:23-27
        if (kind == RT_TEX_CHECKER) {                       // texture.rs:45-54
            double sines = sin(10.0 * p.x) * sin(10.0 * p.y) * sin(10.0 * p.z);
            id = sines < 0.0 ? t.a : t.b;
            continue;
        }
        if (kind == RT_TEX_NOISE) {  // texture.rs:71-79
            double s = 1.0 + sin(t.scale * p.z + 10.0 * perlin_turb(&sc.perlin[t.a], p.x, p.y, p.z, t.scale, 7));
            return (mk(1.0, 1.0, 1.0) * 0.5) * s;
        }
        double rgb[3];
        image_texel(sc.images, sc.texels, t.a, u, v, rgb);
        return mk(rgb[0], rgb[1], rgb[2]);
    }
    return mk(0.0, 0.0, 0.0);
}

// ---------------------------------------------------------------------------
// Lights: PDF::Hittable over the light list (pdf.rs:140-142,164-166; hit.rs:90-96)
// ---------------------------------------------------------------------------
RT_DEV double light_pdf_one(const DLight &l, V3 o, V3 v) {
    if (l.kind == LIGHT_RECT) {  // rect.rs:91-101 -> AARect::hit(Ray(o,v), 0.001, inf), rect.rs:49-58
        double a0 = l.d[0], a1 = l.d[1], b0 = l.d[2], b1 = l.d[3], k = l.d[4];
        double ok, vk, oa, va, ob, vb;
        if (l.axis == RT_PLANE_XZ) { ok = o.y; vk = v.y; oa = o.x; va = v.x; ob = o.z; vb = v.z; }
        else if (l.axis == RT_PLANE_YZ) { ok = o.x; vk = v.x; oa = o.y; va = v.y; ob = o.z; vb = v.z; }
        else { ok = o.z; vk = v.z; oa = o.x; va = v.x; ob = o.y; vb = v.y; }
        double t = (k - ok) / vk;
        if (t < 0.001 || t > RT_INF) return 0.0;
        double a = oa + t * va;
        double b = ob + t * vb;
        if (a < a0 || a > a1 || b < b0 || b > b1) return 0.0;
        double area = (a1 - a0) * (b1 - b0);
        double len = length(v);
        double distance_squared = powi2(t) * powi2(len);
        // rec.normal is +-axis_k after set_face_normal, so |v . normal| = |v_k| exactly
        double cosine = fabs(vk) / len;
        return cosine != 0.0 ? distance_squared / (cosine * area) : 0.0;
    }
    if (feat(F_SPHERE_LIGHT) && l.kind == LIGHT_SPHERE) {  // sphere.rs:104-112
        V3 center = ld3(l.d);
        V3 oc = o - center;
        double a = powi2(length(v));
        double half_b = dot(oc, v);
        double c = powi2(length(oc)) - powi2(l.d[3]);
        double discriminant = powi2(half_b) - a * c;
        if (discriminant < 0.0) return 0.0;
        double sqrt_d = sqrt(discriminant);
        double root = (-half_b - sqrt_d) / a;
        if (root < 0.001 || root > DBL_MAX) {
            root = (-half_b + sqrt_d) / a;
            if (root < 0.001 || root > DBL_MAX) return 0.0;
        }
        double cos_theta_max = sqrt(1.0 - powi2(l.d[3]) / powi2(length(center - o)));
        double solid_angle = 2.0 * kPi * (1.0 - cos_theta_max);
        return 1.0 / solid_angle;
    }
    return 0.0;  // hit.rs:29
}
RT_DEV double lights_pdf_value(const DScene &sc, V3 o, V3 v) {  // hit.rs:90-92
    if (RT_MICRO_OPT && sc.n_lights == 1u) return light_pdf_one(sc.lights[0], o, v);  // 0.0 + x == x and x / 1.0 == x exactly
    double sum = 0.0;
    for (uint32_t i = 0; i < sc.n_lights; ++i) sum += light_pdf_one(sc.lights[i], o, v);
    return sum / (double)sc.n_lights;
}
RT_DEV V3 lights_random(const DScene &sc, V3 o, const Draw &dr) {  // hit.rs:94-96
    uint32_t idx = (dr.bits_b * sc.n_lights) >> 11;
    const DLight &l = sc.lights[idx];
    if (l.kind == LIGHT_RECT) {  // rect.rs:103-111
        double a = gen_range(l.d[0], l.d[1], dr.a);
        double b = gen_range(l.d[2], l.d[3], dr.b);
        double k = l.d[4];
        V3 random_point;
        if (l.axis == RT_PLANE_XZ) random_point = mk(a, k, b);
        else if (l.axis == RT_PLANE_YZ) random_point = mk(k, a, b);
        else random_point = mk(a, b, k);
        return random_point - o;
    }
    if (feat(F_SPHERE_LIGHT) && l.kind == LIGHT_SPHERE) {  // sphere.rs:114-119 + :27-36
        V3 direction = ld3(l.d) - o;
        double distance_squared = powi2(length(direction));
        ONB uvw = onb_from_w(direction);
        double r1 = dr.a, r2 = dr.b, radius = l.d[3];
        double z = 1.0 + r2 * (sqrt(1.0 - powi2(radius) / distance_squared) - 1.0);
        double phi = 2.0 * kPi * r1;
        double x = cos(phi) * sqrt(1.0 - powi2(z));
        double y = sin(phi) * sqrt(1.0 - powi2(z));
        return onb_local(uvw, mk(x, y, z));
    }
    return mk(1.0, 0.0, 0.0);  // hit.rs:30
}

// ---------------------------------------------------------------------------
// Materials (mat.rs:199-422) and the integrator (main.rs:41-120), iteratively
// ---------------------------------------------------------------------------
RT_DEV double reflectance(double cosine, double index_of_refraction) {  // mat.rs:303-307
    double r0 = powi2((1.0 - index_of_refraction) / (1.0 + index_of_refraction));
    return r0 + (1.0 - r0) * powi5(1.0 - cosine);
}
RT_DEV V3 dielectric_direction(const DMaterial &m, V3 r_in_dir, const HitRec &rec, const Rng &rng) {  // mat.rs:343-366
    double refraction_ratio = rec.front_face ? 1.0 / m.ir : m.ir;
    V3 unit_direction = normalized(r_in_dir);
    double cos_theta = fmin(dot((-1.0) * unit_direction, rec.normal), 1.0);
    double sin_theta = sqrt(1.0 - powi2(cos_theta));
    bool cannot_refract = refraction_ratio * sin_theta > 1.0;
    Draw d = draw(rng, SLOT_SCATTER, 0);
    bool will_reflect = d.a < reflectance(cos_theta, refraction_ratio);
    if (cannot_refract || will_reflect) return reflect(unit_direction, rec.normal);
    return refract(unit_direction, rec.normal, refraction_ratio);
}
RT_DEV V3 random_cosine_direction(double r1, double r2) {  // pdf.rs:8-18
    double z = sqrt(1.0 - r2);
    double phi = 2.0 * kPi * r1;
    double s, c;
    sincos(phi, &s, &c);
    double x = c * sqrt(r2);
    double y = s * sqrt(r2);
    return mk(x, y, z);
}

// ---------------------------------------------------------------------------
// The Disney-style PBR material (mat.rs:10-52, :86-197) and PDF::BRDF (pdf.rs:20-60, :97-130, :151-160)
// ---------------------------------------------------------------------------
RT_DEV double clamp01(double x) { return x < 0.0 ? 0.0 : (x > 1.0 ? 1.0 : x); }  // f64::clamp: NaN stays NaN
RT_DEV double schlick_fresnel(double u) {  // mat.rs:10-14
    double m = clamp01(1.0 - u);
    double m2 = powi2(m);
    return m2 * m2 * m;
}
RT_DEV double GTR_1(double n_dot_h, double a) {  // mat.rs:16-24 (log2, as written)
    if (a >= 1.0) return 1.0 / kPi;
    double a2 = a * a;
    double t = 1.0 + (a2 - 1.0) * n_dot_h * n_dot_h;
    return (a2 - 1.0) / (kPi * log2(a2) * t);
}
RT_DEV double GTR_2_aniso(double n_dot_h, double h_dot_x, double h_dot_y, double ax, double ay) {  // mat.rs:32-34
    return 1.0 / (kPi * ax * ay * powi2(powi2(h_dot_x / ax) + powi2(h_dot_y / ay) + n_dot_h * n_dot_h));
}
RT_DEV double smithG_GGX(double n_dot_v, double alphaG) {  // mat.rs:36-40
    double a = alphaG * alphaG;
    double b = n_dot_v * n_dot_v;
    return 1.0 / (n_dot_v + sqrt(a + b - a * b));
}
RT_DEV double smithG_GGX_aniso(double n_dot_v, double v_dot_x, double v_dot_y, double ax, double ay) {  // mat.rs:42-44
    return 1.0 / (n_dot_v + sqrt(powi2(v_dot_x * ax) + powi2(v_dot_y * ay) + powi2(n_dot_v)));
}
RT_DEV double mixd(double a, double b, double t) { return a * (1.0 - t) + b * t; }  // mat.rs:50-52
RT_DEV V3 mixv(V3 a, V3 b, double t) {                                              // vec.rs:60-68
    return mk(a.x * (1.0 - t) + b.x * t, a.y * (1.0 - t) + b.y * t, a.z * (1.0 - t) + b.z * t);
}
RT_DEV void pbr_alpha(const DMaterial &m, double &ax, double &ay) {  // mat.rs:168-170 == pdf.rs:42-44 == :119-121
    double aspect = sqrt(1.0 - m.pbr[RT_PBR_ANISOTROPIC] * 0.9);
    ax = fmax(powi2(m.pbr[RT_PBR_ROUGHNESS]) / aspect, 0.001);
    ay = fmax(powi2(m.pbr[RT_PBR_ROUGHNESS]) * aspect, 0.001);
}
// PBR::brdf (mat.rs:133-195); cd = base_color.mapping(u, v, p)
RT_DEV_COLD V3 pbr_brdf(const DMaterial *mp, V3 cd, V3 r_in_dir, V3 r_out_dir, V3 normal) {
    const DMaterial &m = *mp;
    const double metallic = m.pbr[RT_PBR_METALLIC], subsurface = m.pbr[RT_PBR_SUBSURFACE], specular = m.pbr[RT_PBR_SPECULAR],
                 roughness = m.pbr[RT_PBR_ROUGHNESS], specular_tint = m.pbr[RT_PBR_SPECULAR_TINT], sheen = m.pbr[RT_PBR_SHEEN],
                 sheen_tint = m.pbr[RT_PBR_SHEEN_TINT], clearcoat = m.pbr[RT_PBR_CLEARCOAT], clearcoat_gloss = m.pbr[RT_PBR_CLEARCOAT_GLOSS];
    V3 l = normalized(r_in_dir) * (-1.0);
    V3 v = normalized(r_out_dir);
    ONB onb = onb_from_w(normal);
    V3 n = onb.w, x = onb.u, y = onb.v;
    double n_dot_v = dot(n, v);
    double n_dot_l = dot(n, l);
    if (n_dot_l < 0.0 || n_dot_v < 0.0) return mk(0.0, 0.0, 0.0);
    V3 h = normalized(l + v);
    double n_dot_h = dot(n, h);
    double l_dot_h = dot(l, h);
    V3 cd_lin = mk(pow(cd.x, 2.2), pow(cd.y, 2.2), pow(cd.z, 2.2));  // mon_to_lin, mat.rs:46-48
    double cd_lum = 0.3 * cd_lin.x + 0.6 * cd_lin.y + 0.1 * cd_lin.z;
    V3 c_tint = cd_lum > 0.0 ? cd_lin / cd_lum : mk(1.0, 1.0, 1.0);
    V3 c_spec0 = mixv((mixv(mk(1.0, 1.0, 1.0), c_tint, specular_tint) * 0.08) * specular, cd_lin, metallic);
    V3 c_sheen = mixv(mk(1.0, 1.0, 1.0), c_tint, sheen_tint);
    double fresnel_l = schlick_fresnel(n_dot_l);
    double fresnel_v = schlick_fresnel(n_dot_v);
    double fresnel_diffuse_90 = 0.5 + 2.0 * l_dot_h * l_dot_h * roughness;
    double fresnel_diffuse = mixd(1.0, fresnel_diffuse_90, fresnel_l) * mixd(1.0, fresnel_diffuse_90, fresnel_v);
    double fss90 = l_dot_h * l_dot_h * roughness;
    double fss = mixd(1.0, fss90, fresnel_l) * mixd(1.0, fss90, fresnel_v);
    double subface_scatter = 1.25 * (fss * (1.0 / (n_dot_l + n_dot_v) - 0.5) + 0.5);
    double ax, ay;
    pbr_alpha(m, ax, ay);
    double d_specular = GTR_2_aniso(n_dot_h, dot(h, x), dot(h, y), ax, ay);
    double fresnel_h = schlick_fresnel(l_dot_h);
    V3 f_specular = mixv(c_spec0, mk(1.0, 1.0, 1.0), fresnel_h);
    double g_specular = smithG_GGX_aniso(n_dot_l, dot(l, x), dot(l, y), ax, ay) * smithG_GGX_aniso(n_dot_v, dot(v, x), dot(v, y), ax, ay);
    V3 fresnel_sheen = (fresnel_h * sheen) * c_sheen;
    double d_reflect = GTR_1(n_dot_h, mixd(0.1, 0.001, clearcoat_gloss));
    double f_reflect = mixd(0.04, 1.0, fresnel_h);
    double g_reflect = smithG_GGX(n_dot_l, 0.25) * smithG_GGX(n_dot_v, 0.25);
    return ((((1.0 / kPi) * mixd(fresnel_diffuse, subface_scatter, subsurface)) * cd_lin + fresnel_sheen) * (1.0 - metallic) +
            (g_specular * f_specular) * d_specular) +
           (((mk(0.25, 0.25, 0.25) * clearcoat) * g_reflect) * f_reflect) * d_reflect;
}
// PDF::BRDF value (pdf.rs:97-130)
RT_DEV_COLD double brdf_pdf_value(const DMaterial *mp, V3 uu, V3 uv, V3 uw, V3 r_in, V3 r_out) {
    const DMaterial &m = *mp;
    double cosine = dot(normalized(r_out), uw);
    if (cosine <= 0.0) return 0.0;
    double diffuse_pdf = cosine / kPi;
    V3 l = normalized(r_in) * (-1.0);
    V3 v = normalized(r_out);
    double n_dot_l = dot(uw, l);
    V3 h = normalized(l + v);
    double n_dot_h = dot(uw, h);
    if (n_dot_h <= 0.0) return 0.0;
    double ax, ay;
    pbr_alpha(m, ax, ay);
    double specular_pdf = GTR_2_aniso(n_dot_h, dot(h, uu), dot(h, uv), ax, ay) * fabs(n_dot_h) * 0.25 / n_dot_l;
    double clearcoat_pdf = GTR_1(n_dot_h, mixd(0.1, 0.001, m.pbr[RT_PBR_CLEARCOAT_GLOSS])) * fabs(n_dot_h) * 0.25 / n_dot_l;
    return (diffuse_pdf + specular_pdf + clearcoat_pdf) / 3.0;
}
RT_DEV V3 spherical_direction(double sin_theta, double cos_theta, double sin_phi, double cos_phi) {  // pdf.rs:20-22
    return mk(sin_theta * cos_phi, sin_theta * sin_phi, cos_theta);
}
// pdf.rs:24-36 and :38-60.  r_in is the world-space incoming direction, wh a tangent-space half
// vector: the reference reflects one about the other as written, then maps the result through uvw.local.
RT_DEV_COLD V3 brdf_lobe_direction(const DMaterial *mp, bool aniso, V3 r_in, double r1, double r2) {
    const DMaterial &m = *mp;
    V3 wh;
    if (!aniso) {  // GTR_1_direction
        double a = mixd(0.1, 0.001, m.pbr[RT_PBR_CLEARCOAT_GLOSS]);
        double a2 = a * a;
        double cos_theta = sqrt(fmax(0.001, (1.0 - pow(a2, 1.0 - r1)) / (1.0 - a2)));
        double sin_theta = sqrt(fmax(0.001, 1.0 - cos_theta * cos_theta));
        double phi = kPi * 2.0 * r2;
        wh = spherical_direction(sin_theta, cos_theta, sin(phi), cos(phi));
    } else {  // GTR_2_aniso_direction
        double ax, ay;
        pbr_alpha(m, ax, ay);
        double phi = atan(ay / ax * tan(2.0 * kPi * r2 + 0.5 * kPi));
        if (r2 > 0.5) phi += kPi;
        double sin_phi = sin(phi);
        double cos_phi = cos(phi);
        double ax_2 = ax * ax;
        double ay_2 = ay * ay;
        double a2 = 1.0 / (cos_phi * cos_phi / ax_2 + sin_phi * sin_phi / ay_2);
        double tan_theta_2 = a2 * r1 / (1.0 - r1);
        double cos_theta = 1.0 / sqrt(1.0 + tan_theta_2);
        double sin_theta = sqrt(fmax(0.001, 1.0 - cos_theta * cos_theta));
        wh = spherical_direction(sin_theta, cos_theta, sin(phi), cos(phi));
    }
    return reflect(r_in, wh);
}

struct PathState {
    Ray ray;
    V3 beta;      // product of the factors the recursion multiplies on the way back up
    V3 radiance;  // what ray_color returns for the camera ray: written by the segment that ends the path
    Rng rng;
    uint32_t depth_left;
    uint32_t segments;
};

// One call of ray_color (one segment).  Returns false when the path ended.
//   recursion:  L = emitted + f * L_next      iteration:  radiance += beta*emitted ; beta *= f
RT_DEV bool path_shade(const DScene &sc, PathState &ps, bool hit, const HitRec &rec, uint32_t integrator, uint32_t flags);

// The path ends with nothing emitted: the recursion returns f * 0 / pdf all the way up, which is
// 0 - or NaN when some factor on the way was not finite (0/0 or x/0 at main.rs:97,104; §Q10).
RT_DEV bool path_end_black(PathState &ps) {
    ps.radiance = ps.beta * 0.0;
    return false;
}

// MEDIA: the scene has ConstantMedium objects (a compile-time property of the kernel variant, so
// that scenes without volumes do not carry the boundary-query loop).
template <bool MEDIA>
RT_DEV bool path_step(const DScene &sc, PathState &ps, uint32_t integrator, uint32_t flags) {
    ps.radiance = mk(0.0, 0.0, 0.0);
    if (ps.depth_left == 0) return false;  // main.rs:42-45: contributes black
    HitRec rec;
    ps.segments += 1;
    bool hit = world_hit<MEDIA, false>(sc, ps.ray, ps.rng, rec);  // main.rs:48
    return path_shade(sc, ps, hit, rec, integrator, flags);
}

// Everything ray_color does after world.hit returned (main.rs:62-119).
RT_DEV bool path_shade(const DScene &sc, PathState &ps, bool hit, const HitRec &rec, uint32_t integrator, uint32_t flags) {
    // Every factor of the recursion L = emitted + f * L_next is zero-emission until the path ends
    // (only DiffuseLight emits, and it never scatters), so the camera ray's radiance is written
    // once, by the last segment: 0 + beta * emitted.  It is not carried from segment to segment.
    ps.radiance = mk(0.0, 0.0, 0.0);
    if (!hit) {  // main.rs:118
        ps.radiance = ps.radiance + ps.beta * ld3(sc.background);
        return false;
    }
    const DMaterial &m = sc.materials[rec.material];
    uint32_t mkind = m.kind;
    // Material::emitted (mat.rs:70-72, :395-401)
    if (mkind == RT_MAT_DIFFUSE_LIGHT) {
        // DiffuseLight never scatters (mat.rs:391-393 / default scatter_mc_method): return emitted
        if (rec.front_face) ps.radiance = ps.radiance + ps.beta * texture_value(sc, m.texture, rec.u, rec.v, rec.p);
        else return path_end_black(ps);
        return false;
    }
    V3 new_dir = mk(0.0, 0.0, 0.0);
    V3 factor = mk(0.0, 0.0, 0.0);
    if (feat(F_METAL) && mkind == RT_MAT_METAL) {  // mat.rs:280-293 == :269-278
        V3 reflected = normalized(reflect(ps.ray.d, rec.normal));
        // random_in_unit_sphere is drawn even for fuzz == 0 (§Q12); with slot addressing the
        // draw can be skipped when its product with fuzz is exactly zero.
        new_dir = m.fuzz != 0.0 ? reflected + m.fuzz * random_in_unit_sphere(ps.rng) : reflected;
        if (!(dot(new_dir, rec.normal) > 0.0)) return path_end_black(ps);  // None -> emitted (black)
        factor = ld3(m.albedo);
    } else if (feat(F_DIELECTRIC) && mkind == RT_MAT_DIELECTRIC) {  // mat.rs:343-374 == :317-341
        new_dir = dielectric_direction(m, ps.ray.d, rec, ps.rng);
        factor = mk(1.0, 1.0, 1.0);
    } else if (feat(F_LEGACY) && (!feat(F_HEAD) || integrator == RT_INTEGRATOR_LEGACY)) {
        if (mkind == RT_MAT_LAMBERTIAN) {  // mat.rs:213-223
            new_dir = rec.normal + normalized(random_in_unit_sphere(ps.rng));
            if (near_zero(new_dir)) new_dir = rec.normal;
        } else if (mkind == RT_MAT_ISOTROPIC) {  // mat.rs:418-421
            new_dir = random_in_unit_sphere(ps.rng);
        } else {
            return path_end_black(ps);  // PBR has no legacy scatter (trait default None, mat.rs:56-58)
        }
        factor = texture_value(sc, m.texture, rec.u, rec.v, rec.p);
    } else if (feat(F_HEAD) && feat(F_PBR) && mkind == RT_MAT_PBR) {  // main.rs:99-105 with mat.rs:118-131
        ONB uvw = onb_from_w(rec.normal);  // PDF::brdf_pdf (pdf.rs:70-79)
        Draw d = draw(ps.rng, SLOT_SCATTER, 0);
        if (d.bits_a & 1u) {  // pdf.rs:169
            new_dir = lights_random(sc, rec.p, d);
        } else {  // pdf.rs:151-160: the lobe by gen_range(0.0..1.0), then the lobe's own (r1, r2)
            double lobe = draw(ps.rng, SLOT_SCATTER, 1).a;
            V3 local = lobe < 0.333 ? random_cosine_direction(d.a, d.b) : brdf_lobe_direction(&m, !(lobe < 0.666), ps.ray.d, d.a, d.b);
            new_dir = onb_local(uvw, local);
        }
        double pdf_value = 0.5 * lights_pdf_value(sc, rec.p, new_dir) + 0.5 * brdf_pdf_value(&m, uvw.u, uvw.v, uvw.w, ps.ray.d, new_dir);
        V3 f = pbr_brdf(&m, texture_value(sc, m.texture, rec.u, rec.v, rec.p), ps.ray.d, new_dir, rec.normal);
        factor = f / pdf_value;  // main.rs:104
    } else if (feat(F_HEAD)) {
        if (mkind != RT_MAT_LAMBERTIAN) return path_end_black(ps);  // Isotropic under HEAD: scatter_mc_method is None (§Q6)
        // main.rs:92-98
        V3 attenuation = texture_value(sc, m.texture, rec.u, rec.v, rec.p);
        ONB uvw = onb_from_w(rec.normal);  // PDF::cosine_pdf (pdf.rs:83-87)
        Draw d = draw(ps.rng, SLOT_SCATTER, 0);
        if (d.bits_a & 1u)  // pdf.rs:169
            new_dir = lights_random(sc, rec.p, d);
        else
            new_dir = onb_local(uvw, random_cosine_direction(d.a, d.b));
        double light_pdf = lights_pdf_value(sc, rec.p, new_dir);
        V3 unit = normalized(new_dir);
        double cosine = dot(unit, uvw.w);  // pdf.rs:131-139
        double cosine_pdf = cosine > 0.0 ? cosine / kPi : 0.0;
        double pdf_value = 0.5 * light_pdf + 0.5 * cosine_pdf;  // pdf.rs:143-145
        double spdf = fmax(dot(rec.normal, unit), 0.0) / kPi;    // mat.rs:246-249
        factor = (attenuation * spdf) / pdf_value;               // main.rs:97
    }
    ps.beta = ps.beta * factor;
    ps.ray.o = rec.p;
    ps.ray.d = new_dir;  // time is inherited (main.rs:95, mat.rs:219,270,368)
    ps.depth_left -= 1;
    ps.rng.bounce += 1;
    if (!(flags & RT_FLAG_TRACE_ZERO_THROUGHPUT) && ps.beta.x == 0.0 && ps.beta.y == 0.0 && ps.beta.z == 0.0)
        return false;  // §Q11: the reference keeps tracing; the contribution is zero either way
    if (ps.depth_left == 0) return path_end_black(ps);  // the next call is ray_color(depth = 0): black (main.rs:42-45)
    return true;
}

// The sample closure of main.rs:811-829: pixel jitter + Camera::get_ray (camera.rs:51-59)
RT_DEV Ray camera_ray(const RtCamera &c, uint32_t width, uint32_t height, uint32_t i, uint32_t j, Rng rng) {
    rng.bounce = 0;
    Draw dj = draw(rng, SLOT_PIXEL, 0);
    double u = ((double)i + dj.a) / (double)(width - 1);
    double v = ((double)j + dj.b) / (double)(height - 1);
    V3 origin = ld3(c.origin), llc = ld3(c.lower_left_corner);
    V3 horizontal = ld3(c.horizontal), vertical = ld3(c.vertical);
    V3 rd = c.lens_radius * random_in_unit_disk(rng);
    V3 offset = ld3(c.cu) * rd.x + ld3(c.cv) * rd.y;
    Draw dt = draw(rng, SLOT_TIME, 0);
    Ray r;
    r.time = c.time0 + dt.a * (c.time1 - c.time0);
    r.o = origin + offset;
    r.d = llc + u * horizontal + v * vertical - (origin + offset);
    return r;
}

RT_DEV void path_begin(PathState &ps, const RtCamera &cam, uint32_t width, uint32_t height, uint32_t i, uint32_t j,
                       uint32_t sample, uint32_t seed, uint32_t max_depth) {
    ps.rng = Rng{seed, j * width + i, sample, 0};
    ps.ray = camera_ray(cam, width, height, i, j, ps.rng);
    ps.beta = mk(1.0, 1.0, 1.0);
    ps.radiance = mk(0.0, 0.0, 0.0);
    ps.depth_left = max_depth;
    ps.segments = 0;
}

// q = n / d and n - q * d for 0 <= n < 2^53, 1 <= d < 2^32, with inv = 1.0 / d: the product is within one of the
// quotient (its relative error is below 2^-51 and the quotient below 2^33 here), one correction step makes it exact.
// Replaces a 64-bit integer division (a ~70-instruction subroutine that the item fetch ran at two active lanes).
RT_DEV uint64_t divmod_by(uint64_t n, uint64_t d, double inv, uint64_t &rem) {
#if !RT_MICRO_OPT
    rem = n % d;
    return n / d;
#endif
    uint64_t q = (uint64_t)((double)n * inv);
    int64_t r = (int64_t)(n - q * d);
    if (r < 0) {
        --q;
        r += (int64_t)d;
    } else if (r >= (int64_t)d) {
        ++q;
        r -= (int64_t)d;
    }
    rem = (uint64_t)r;
    return q;
}
// A work item is (sample chunk, pixel): item = chunk * items_per_chunk + lin.
RT_DEV uint32_t item_split(const RenderParams &P, uint64_t item, uint64_t &lin) {
    return (uint32_t)divmod_by(item, P.items_per_chunk, P.inv_items_per_chunk, lin);
}
// pixel order inside the item space: 8x4 tiles so that the 32 lanes of a warp start
// on neighbouring pixels (coherent primary rays and BVH paths)
RT_DEV bool item_pixel(const RenderParams &P, uint64_t lin, uint32_t &i, uint32_t &row) {
    const uint32_t within = (uint32_t)(lin & 31u);
    uint64_t tx;
    const uint32_t ty = (uint32_t)divmod_by(lin >> 5, P.tiles_x, P.inv_tiles_x, tx);
    i = (uint32_t)tx * 8u + (within & 7u);
    row = ty * 4u + (within >> 3);
    return i < P.width && row < P.height;
}

// What a ray found, as the class the sorting stages group by (wavefront.inl: shade tiles; sorted.inl: the lanes of a
// block): miss, medium, one class per material kind, and the costly materials (Perlin noise, image textures).
enum : uint32_t { WF_CLS_MISS = 0, WF_CLS_MEDIUM = 1, WF_CLS_MATERIAL = 2, WF_CLS_COSTLY = 7, WF_N_CLASSES = 8, WF_CLS_NONE = 0xFFu };
RT_DEV uint32_t hit_class(const DScene &sc, uint32_t prim) {
    if (prim == kNoPrim) return WF_CLS_MISS;
    if (prim & kMediumFlag) return WF_CLS_MEDIUM;
    const DMaterial &m = sc.materials[sc.prims[prim].material];
    if (feat(F_TEX) && m.costly != 0u) return WF_CLS_COSTLY;
    const uint32_t k = m.kind;  // RtMaterialKind: lambertian, metal, dielectric, light, isotropic | PBR
    return WF_CLS_MATERIAL + (k < 4u ? k : 4u);
}


}  // namespace RT_VARIANT_NS
}  // namespace rtb200dev
