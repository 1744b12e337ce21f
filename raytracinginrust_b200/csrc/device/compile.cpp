// compile.cpp — scene-graph compiler and SAH BVH builder (host side of the device
// library).  Replaces the Box<dyn Hittable> tree (src/hit.rs, src/bvh.rs:18-73) with
// the flat tables of tables.h.  Only *culling* structures are new here: the
// primitives, their parameters and the wrapper order are taken over unchanged, so
// every ray/primitive test sees the same f64 numbers the reference would.
#include "compile.h"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <future>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <functional>
#include <memory>
#include <queue>
#include <thread>
#include <unordered_map>

namespace rtb200dev {
namespace {

const double PI = 3.14159265358979323846264338327950288;
const uint32_t LINEAR_MAX = 6;  // groups with at most this many primitives are scanned linearly
const uint32_t LEAF_MAX = 2;    // primitives per BVH leaf
const int N_BINS = 16;
const int FORCE_MEDIAN_DEPTH = 32;  // bounds the tree depth (traversal stack is 64 entries)

// Splits [0, n) into contiguous blocks over the host threads (large meshes only; the result never depends on
// the split: every index is written by exactly one thread).
template <class F>
void parallel_for(size_t n, F f) {
    size_t threads = std::thread::hardware_concurrency();
    if (threads > 16) threads = 16;
    if (n < 65536 || threads < 2) {
        f(0, n);
        return;
    }
    const size_t per = (n + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < threads; ++t) {
        const size_t a = t * per, b = std::min(n, a + per);
        if (a < b) pool.emplace_back([=, &f] { f(a, b); });
    }
    f(0, std::min(n, per));
    for (std::thread &th : pool) th.join();
}

struct Box {
    double lo[3], hi[3];
    void reset() {
        for (int a = 0; a < 3; ++a) {
            lo[a] = DBL_MAX;
            hi[a] = -DBL_MAX;
        }
    }
    // plain comparisons (minsd / maxsd), not fmin / fmax: at -O2 those are libm calls, and the builder
    // makes ~10^8 of them.  A NaN operand is ignored like fmin would (the comparison is false), and
    // finalize_groups rejects non-finite bounds before any tree is built.
    void grow(const Box &b) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = b.lo[a] < lo[a] ? b.lo[a] : lo[a];
            hi[a] = b.hi[a] > hi[a] ? b.hi[a] : hi[a];
        }
    }
    void grow(const double p[3]) {
        for (int a = 0; a < 3; ++a) {
            lo[a] = p[a] < lo[a] ? p[a] : lo[a];
            hi[a] = p[a] > hi[a] ? p[a] : hi[a];
        }
    }
    double area() const {
        double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        return 2.0 * (dx * dy + dy * dz + dz * dx);
    }
    bool finite() const {
        for (int a = 0; a < 3; ++a)
            if (!std::isfinite(lo[a]) || !std::isfinite(hi[a])) return false;
        return true;
    }
    // Outward padding that dominates the rounding of an f64 slab test and gives
    // zero-thickness boxes (rects) a volume.
    void pad() {
        double m = 1e-3;
        for (int a = 0; a < 3; ++a) m = std::fmax(m, std::fmax(std::fabs(lo[a]), std::fabs(hi[a])));
        double e = 1e-9 * m;
        for (int a = 0; a < 3; ++a) {
            lo[a] -= e;
            hi[a] += e;
        }
    }
};

void rect_axes(uint32_t plane, int &k, int &a, int &b) {  // rect.rs:26-32
    switch (plane) {
        case RT_PLANE_YZ: k = 0; a = 1; b = 2; break;
        case RT_PLANE_XZ: k = 1; a = 0; b = 2; break;
        default: k = 2; a = 0; b = 1; break;
    }
}
void rotate_axes(uint32_t axis, int &r, int &a, int &b) {  // rotate.rs:15-21
    switch (axis) {
        case RT_AXIS_X: r = 0; a = 1; b = 2; break;
        case RT_AXIS_Y: r = 1; a = 0; b = 2; break;
        default: r = 2; a = 0; b = 1; break;
    }
}

Box prim_box(const DPrim &p) {
    Box b;
    b.reset();
    switch (p.kind) {
        case PRIM_SPHERE: {
            double r = std::fabs(p.d[3]);
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = p.d[a] - r;
                b.hi[a] = p.d[a] + r;
            }
            break;
        }
        case PRIM_MSPHERE: {  // union of the boxes at center0 and center1 (sphere.rs:191-201, §Q18) ...
            double r = std::fabs(p.d[8]);
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = std::fmin(p.d[a], p.d[3 + a]) - r;
                b.hi[a] = std::fmax(p.d[a], p.d[3 + a]) + r;
            }
            // ... and, unlike the reference's box, of where center(time) (sphere.rs:144-146) puts the sphere over the
            // shutter of every camera of main.rs, [0, 1): with (time0, time1) != (0, 1) the centre extrapolates beyond
            // center0 / center1, and a list in the reference has no box to cull it with.  (Found by the random scene
            // graphs.  A shutter outside [0, 1] together with such a sphere is not covered: DESIGN.md.)
            const double t0 = p.d[6], t1 = p.d[7];
            if (t0 != 0.0 || t1 != 1.0) {
                for (double time : {0.0, 1.0}) {
                    const double s = (time - t0) / (t1 - t0);
                    for (int a = 0; a < 3; ++a) {
                        const double c = p.d[a] + s * (p.d[3 + a] - p.d[a]);
                        b.lo[a] = std::fmin(b.lo[a], c - r);  // a NaN centre (time0 == time1) leaves the box as it is
                        b.hi[a] = std::fmax(b.hi[a], c + r);
                    }
                }
            }
            break;
        }
        case PRIM_RECT: {  // correct bounds on the rect's own plane (the reference's are wrong, §Q5)
            int k, a, bb;
            rect_axes(p.axis, k, a, bb);
            b.lo[a] = std::fmin(p.d[0], p.d[1]);
            b.hi[a] = std::fmax(p.d[0], p.d[1]);
            b.lo[bb] = std::fmin(p.d[2], p.d[3]);
            b.hi[bb] = std::fmax(p.d[2], p.d[3]);
            b.lo[k] = b.hi[k] = p.d[4];
            break;
        }
        case PRIM_TRI: {
            double v0[3] = {p.d[0], p.d[1], p.d[2]};
            double v1[3] = {p.d[0] + p.d[3], p.d[1] + p.d[4], p.d[2] + p.d[5]};
            double v2[3] = {p.d[0] + p.d[6], p.d[1] + p.d[7], p.d[2] + p.d[8]};
            b.grow(v0);
            b.grow(v1);
            b.grow(v2);
            // v1, v2 are re-derived from v0 + e: cover the rounding of that sum
            for (int a = 0; a < 3; ++a) {
                double e = 4e-16 * std::fmax(std::fabs(b.lo[a]), std::fabs(b.hi[a]));
                b.lo[a] -= e;
                b.hi[a] += e;
            }
            break;
        }
        case PRIM_BOX:
            for (int a = 0; a < 3; ++a) {
                b.lo[a] = std::fmin(p.d[a], p.d[3 + a]);
                b.hi[a] = std::fmax(p.d[a], p.d[3 + a]);
            }
            break;
    }
    return b;
}

// object space -> the space above the chain: inverse ops, innermost first
// (translate.rs:26, rotate.rs:94-95)
void point_to_outer(const std::vector<DOp> &ops, double p[3]) {
    for (size_t i = ops.size(); i-- > 0;) {
        const DOp &op = ops[i];
        if (op.kind == OP_TRANSLATE) {
            for (int a = 0; a < 3; ++a) p[a] += op.offset[a];
        } else if (op.kind == OP_ROTATE) {
            int r, a, b;
            rotate_axes(op.axis, r, a, b);
            double pa = op.cos_theta * p[a] + op.sin_theta * p[b];
            double pb = -op.sin_theta * p[a] + op.cos_theta * p[b];
            p[a] = pa;
            p[b] = pb;
        }
    }
}

// ---------------------------------------------------------------------------
// Binned SAH BVH2
// ---------------------------------------------------------------------------
struct TNode {  // (no default initialisers: the builder sizes its array for 2 n nodes and writes the ones it hands out)
    Box box;
    int left, right;        // inner; -1, -1 for a leaf
    uint32_t first, count;  // leaf (into the order[] permutation)
};

struct Builder {
    const std::vector<Box> &boxes;
    std::vector<uint32_t> &order;
    std::unique_ptr<TNode[]> nodes;  // 2 * count of them, not initialised; ids are handed out by `next`
    std::atomic<int> next{0};
    std::atomic<uint32_t> max_depth{0};
    std::vector<double> cx[3];
    size_t n_threads;  // (asked once: hardware_concurrency is a system call)

    Builder(const std::vector<Box> &b, std::vector<uint32_t> &o) : boxes(b), order(o) {
        n_threads = std::getenv("RTB200_COMPILE_SERIAL") ? 1 : std::min<size_t>(std::thread::hardware_concurrency(), kMaxPieces);
        for (int a = 0; a < 3; ++a) cx[a].resize(b.size());
        parallel_for(b.size(), [&](size_t i0, size_t i1) {
            for (int a = 0; a < 3; ++a)
                for (size_t i = i0; i < i1; ++i) cx[a][i] = 0.5 * (b[i].lo[a] + b[i].hi[a]);
        });
    }

    struct Bins {
        Box bb[3][N_BINS];
        uint32_t bn[3][N_BINS];
        void reset() {
            for (int ax = 0; ax < 3; ++ax)
                for (int k = 0; k < N_BINS; ++k) {
                    bb[ax][k].reset();
                    bn[ax][k] = 0;
                }
        }
    };
    static constexpr uint32_t kBigNode = 65536;  // nodes of this many primitives bin with all host threads
    static constexpr int kMaxPieces = 16;
    static constexpr uint32_t kTaskMin = 4096;

    // f(piece, i0, i1) over contiguous pieces of [first, first + count): threads for a big node, one piece otherwise.
    // What the pieces compute is merged with min / max / integer sums only, so the result never depends on the split.
    template <class F>
    int pieces(uint32_t first, uint32_t count, F f) const {
        size_t threads = n_threads;
        if (count < kBigNode || threads < 2) {
            f(0, (size_t)first, (size_t)first + count);
            return 1;
        }
        const size_t per = (count + threads - 1) / threads;
        std::vector<std::thread> pool;
        for (size_t t = 1; t < threads; ++t) {
            const size_t a = first + t * per, b = std::min((size_t)first + count, a + per);
            pool.emplace_back([=, &f] { f((int)t, a, std::max(a, b)); });
        }
        f(0, (size_t)first, std::min((size_t)first + count, (size_t)first + per));
        for (std::thread &th : pool) th.join();
        return (int)threads;
    }

    // bounds of the primitives and of their centroids over a range of order[]
    void bounds(uint32_t first, uint32_t count, Box &box, Box &cbox) const {
        Box pb[kMaxPieces], pc[kMaxPieces];
        const int n = pieces(first, count, [&](int k, size_t i0, size_t i1) {
            Box b, c;
            b.reset();
            c.reset();
            for (size_t i = i0; i < i1; ++i) {
                b.grow(boxes[order[i]]);
                double cc[3] = {cx[0][order[i]], cx[1][order[i]], cx[2][order[i]]};
                c.grow(cc);
            }
            pb[k] = b;
            pc[k] = c;
        });
        box.reset();
        cbox.reset();
        for (int k = 0; k < n; ++k) {
            box.grow(pb[k]);
            cbox.grow(pc[k]);
        }
    }

    int build(uint32_t first, uint32_t count, uint32_t depth) {
        Box box, cbox;
        bounds(first, count, box, cbox);
        return build_node(first, count, depth, box, cbox);
    }

    // The two halves of a node own disjoint ranges of order[] and disjoint node ids, so large
    // subtrees are built in parallel (the tree does not depend on the schedule).  A node gets its bounds from the
    // pass that partitioned its parent (the predicate sees every primitive exactly once and knows its side).
    int build_node(uint32_t first, uint32_t count, uint32_t depth, const Box &box, const Box &cbox) {
        int id = next.fetch_add(1);
        nodes[id].box = box;
        nodes[id].left = nodes[id].right = -1;
        nodes[id].first = nodes[id].count = 0;
        for (uint32_t seen = max_depth.load(); seen < depth && !max_depth.compare_exchange_weak(seen, depth);) {
        }
        if (count <= LEAF_MAX) {
            nodes[id].first = first;
            nodes[id].count = count;
            return id;
        }
        int axis = 0;
        double ext[3] = {cbox.hi[0] - cbox.lo[0], cbox.hi[1] - cbox.lo[1], cbox.hi[2] - cbox.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        uint32_t mid = first + count / 2;
        bool done = false;
        Box side_box[2], side_cbox[2];
        if ((int)depth < FORCE_MEDIAN_DEPTH && ext[axis] > 0.0) {
            // binned SAH over all three axes; one pass over the primitives fills the bins of every axis
            double best_cost = DBL_MAX;
            int best_axis = -1, best_bin = -1;
            double scale3[3];
            for (int ax = 0; ax < 3; ++ax) scale3[ax] = ext[ax] > 0.0 ? (double)N_BINS / ext[ax] : 0.0;
            Bins one;
            std::unique_ptr<Bins[]> many(count >= kBigNode && n_threads > 1 ? new Bins[kMaxPieces] : nullptr);
            Bins *part = many ? many.get() : &one;
            const int n_parts = pieces(first, count, [&](int k, size_t i0, size_t i1) {
                Bins &B = part[k];
                B.reset();
                for (size_t i = i0; i < i1; ++i) {
                    const uint32_t p = order[i];
                    const Box &pb = boxes[p];
                    for (int ax = 0; ax < 3; ++ax) {
                        if (!(ext[ax] > 0.0)) continue;
                        int b = (int)((cx[ax][p] - cbox.lo[ax]) * scale3[ax]);
                        b = std::min(std::max(b, 0), N_BINS - 1);
                        B.bb[ax][b].grow(pb);
                        B.bn[ax][b]++;
                    }
                }
            });
            Bins &B = part[0];
            for (int k = 1; k < n_parts; ++k)
                for (int ax = 0; ax < 3; ++ax)
                    for (int b = 0; b < N_BINS; ++b) {
                        if (part[k].bn[ax][b]) B.bb[ax][b].grow(part[k].bb[ax][b]);
                        B.bn[ax][b] += part[k].bn[ax][b];
                    }
            for (int ax = 0; ax < 3; ++ax) {
                if (!(ext[ax] > 0.0)) continue;
                double right_area[N_BINS];
                uint32_t right_n[N_BINS];
                Box acc;
                acc.reset();
                uint32_t n = 0;
                for (int k = N_BINS - 1; k > 0; --k) {
                    if (B.bn[ax][k]) acc.grow(B.bb[ax][k]);
                    n += B.bn[ax][k];
                    right_area[k] = n ? acc.area() : 0.0;
                    right_n[k] = n;
                }
                acc.reset();
                n = 0;
                for (int k = 0; k < N_BINS - 1; ++k) {
                    if (B.bn[ax][k]) acc.grow(B.bb[ax][k]);
                    n += B.bn[ax][k];
                    if (n == 0 || right_n[k + 1] == 0) continue;
                    double cost = acc.area() * (double)n + right_area[k + 1] * (double)right_n[k + 1];
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = ax;
                        best_bin = k;
                    }
                }
            }
            if (best_axis >= 0) {
                double scale = (double)N_BINS / ext[best_axis];
                double lo = cbox.lo[best_axis];
                for (int sd = 0; sd < 2; ++sd) {
                    side_box[sd].reset();
                    side_cbox[sd].reset();
                }
                auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t p) {
                    int k = (int)((cx[best_axis][p] - lo) * scale);
                    k = std::min(std::max(k, 0), N_BINS - 1);
                    const int sd = k <= best_bin ? 0 : 1;  // std::partition applies the predicate exactly once per element
                    side_box[sd].grow(boxes[p]);
                    double cc[3] = {cx[0][p], cx[1][p], cx[2][p]};
                    side_cbox[sd].grow(cc);
                    return sd == 0;
                });
                mid = (uint32_t)(it - order.begin());
                done = mid > first && mid < first + count;
            }
        }
        if (!done) {  // object median on the widest axis
            mid = first + count / 2;
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                             [&](uint32_t a, uint32_t b) { return cx[axis][a] < cx[axis][b]; });
            bounds(first, mid - first, side_box[0], side_cbox[0]);
            bounds(mid, first + count - mid, side_box[1], side_cbox[1]);
        }
        int l, r;
        // a thread per left subtree down to kTaskMin primitives (a few hundred tasks for a mesh of 400 k triangles: with
        // 32768 and at most 64 tasks the 24 threads of the GPU box waited for the largest of ~20 subtrees)
        if (mid - first >= kTaskMin && first + count - mid >= kTaskMin && n_threads > 1) {
            auto left = std::async(std::launch::async, [&] { return build_node(first, mid - first, depth + 1, side_box[0], side_cbox[0]); });
            r = build_node(mid, first + count - mid, depth + 1, side_box[1], side_cbox[1]);
            l = left.get();
        } else {
            l = build_node(first, mid - first, depth + 1, side_box[0], side_cbox[0]);
            r = build_node(mid, first + count - mid, depth + 1, side_box[1], side_cbox[1]);
        }
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
    }
};

// f64 -> f32 rounded toward -inf / +inf (the node boxes may only grow)
float f32_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}
float f32_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

// Builds the BVH of prims[first, first+count) (reordering them) and appends the
// nodes, breadth-first, to out.nodes.  Returns the root node index.
int32_t build_group_bvh(CompiledScene &out, uint32_t first, uint32_t count) {
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (getenv("RTB200_COMPILE_TIMING") && count > 10000) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[bvh %u] %s %.3f s\n", count, what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    std::vector<Box> boxes(count);
    std::vector<uint32_t> order(count);
    parallel_for(count, [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; ++i) {
            boxes[i] = prim_box(out.prims[first + i]);
            boxes[i].pad();
            order[i] = (uint32_t)i;
        }
    });
    lap("boxes");
    Builder b(boxes, order);
    b.nodes.reset(new TNode[2 * (size_t)count + 1]);
    lap("builder init");
    int root = b.build(0, count, 1);
    lap("build");
    const size_t n_tnodes = (size_t)b.next.load();
    out.max_bvh_depth = std::max(out.max_bvh_depth, b.max_depth.load());
    // permute the primitives into leaf order
    {
        std::unique_ptr<DPrim[]> tmp(new DPrim[count]);  // not value-initialised: every record is overwritten below
        parallel_for(count, [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) tmp[i] = out.prims[first + order[i]];
        });
        parallel_for(count, [&](size_t i0, size_t i1) { std::copy(tmp.get() + i0, tmp.get() + i1, out.prims.begin() + first + i0); });
    }
    lap("permute");
    // breadth-first emission of the inner nodes
    uint32_t base = (uint32_t)out.nodes.size();
    std::vector<int> bfs;  // temp-node ids of inner nodes in BFS order
    std::vector<int32_t> slot(n_tnodes, -1);
    bfs.push_back(root);
    slot[root] = 0;
    for (size_t h = 0; h < bfs.size(); ++h) {
        const TNode &n = b.nodes[bfs[h]];
        for (int c : {n.left, n.right}) {
            if (b.nodes[c].left >= 0) {
                slot[c] = (int32_t)bfs.size();
                bfs.push_back(c);
            }
        }
    }
    out.nodes.resize(base + bfs.size());
    auto encode = [&](int c) -> int32_t {
        const TNode &n = b.nodes[c];
        if (n.left >= 0) return (int32_t)(base + slot[c]);
        uint32_t code = ((first + n.first) << 3) | (n.count - 1);
        return (int32_t)~code;
    };
    parallel_for(bfs.size(), [&](size_t h0, size_t h1) {
        for (size_t h = h0; h < h1; ++h) {
            const TNode &n = b.nodes[bfs[h]];
            DBvhNode &dn = out.nodes[base + h];
            const Box &b0 = b.nodes[n.left].box, &b1 = b.nodes[n.right].box;
            for (int a = 0; a < 3; ++a) {
                dn.lo0[a] = f32_down(b0.lo[a]);
                dn.hi0[a] = f32_up(b0.hi[a]);
                dn.lo1[a] = f32_down(b1.lo[a]);
                dn.hi1[a] = f32_up(b1.hi[a]);
            }
            dn.child0 = encode(n.left);
            dn.child1 = encode(n.right);
            dn.pad0 = dn.pad1 = 0;
        }
    });
    lap("emit");
    return (int32_t)base;
}

// ---------------------------------------------------------------------------
// Graph walk
// ---------------------------------------------------------------------------
struct GroupBuild {
    std::vector<DOp> xform;  // chain without flips
    bool tree = false;       // its primitives sit below a BVH node of the scene graph (see Walker::in_bvh)
    uint32_t chain = 0;
    std::vector<DPrim> prims;
};
struct PendingMedium {
    uint32_t node;
    std::vector<DOp> stack;
    uint32_t rank;
};

bool same_ops(const std::vector<DOp> &a, const std::vector<DOp> &b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); ++i) {
        if (a[i].kind != b[i].kind || a[i].axis != b[i].axis) return false;
        if (std::memcmp(&a[i].sin_theta, &b[i].sin_theta, sizeof(double) * 5) != 0) return false;
    }
    return true;
}


// ---------------------------------------------------------------------------
// The reference's own traversal order.  On an exact-t tie the reference keeps the
// object it visits LAST (hit.rs:64-66, bvh.rs:81-84), and coincident
// faces (the touching ground boxes of final_scene) make such ties common for rays
// inside a box.  Ranks must therefore follow the reference's BVH leaf order, which
// depends on its build: bounding_box() of every child (with the reference's quirks),
// widest axis, sort by min+max, split in halves (bvh.rs:18-73).
// ---------------------------------------------------------------------------
struct RefBox {
    double lo[3], hi[3];
};
bool ref_bbox(const RtSceneDesc &d, uint32_t id, double t0, double t1, RefBox &out, uint32_t depth) {
    if (id >= d.n_nodes || depth > d.n_nodes + 1) return false;
    const RtNode &n = d.nodes[id];
    switch (n.kind) {
        case RT_NODE_SPHERE:  // sphere.rs:97-102
            for (int a = 0; a < 3; ++a) {
                out.lo[a] = n.v[a] - n.v[3];
                out.hi[a] = n.v[a] + n.v[3];
            }
            return true;
        case RT_NODE_MOVING_SPHERE:  // sphere.rs:191-201
            for (int a = 0; a < 3; ++a) {
                out.lo[a] = std::fmin(n.v[a] - n.v[8], n.v[3 + a] - n.v[8]);
                out.hi[a] = std::fmax(n.v[a] + n.v[8], n.v[3 + a] + n.v[8]);
            }
            return true;
        case RT_NODE_RECT:  // rect.rs:83-89 (§Q5: always (a0,b0,k-1e-4)..(a1,b1,k+1e-4))
            out.lo[0] = n.v[0]; out.lo[1] = n.v[2]; out.lo[2] = n.v[4] - 0.0001;
            out.hi[0] = n.v[1]; out.hi[1] = n.v[3]; out.hi[2] = n.v[4] + 0.0001;
            return true;
        case RT_NODE_TRIANGLE:  // tri.rs:59-70
            for (int a = 0; a < 3; ++a) {
                out.lo[a] = std::fmin(n.v[a], std::fmin(n.v[3 + a], n.v[6 + a]));
                out.hi[a] = std::fmax(n.v[a], std::fmax(n.v[3 + a], n.v[6 + a]));
            }
            return true;
        case RT_NODE_CUBE:  // cube.rs:39-46
            for (int a = 0; a < 3; ++a) {
                out.lo[a] = n.v[a];
                out.hi[a] = n.v[3 + a];
            }
            return true;
        case RT_NODE_LIST:   // hit.rs:73-88
        case RT_NODE_BVH: {  // bvh.rs:93-95: surrounding boxes all the way up = the union
            if (n.count == 0 || (uint64_t)n.child + n.count > d.n_child_index) return false;
            for (uint32_t i = 0; i < n.count; ++i) {
                RefBox b;
                if (!ref_bbox(d, d.child_index[n.child + i], t0, t1, b, depth + 1)) return false;
                if (i == 0) out = b;
                else
                    for (int a = 0; a < 3; ++a) {
                        out.lo[a] = std::fmin(out.lo[a], b.lo[a]);
                        out.hi[a] = std::fmax(out.hi[a], b.hi[a]);
                    }
            }
            return true;
        }
        case RT_NODE_TRANSLATE:  // translate.rs:32-40
            if (!ref_bbox(d, n.child, t0, t1, out, depth + 1)) return false;
            for (int a = 0; a < 3; ++a) {
                out.lo[a] += n.v[a];
                out.hi[a] += n.v[a];
            }
            return true;
        case RT_NODE_ROTATE: {  // rotate.rs:40-57 (§Q4): the box comes out as [f64::MIN, f64::MAX]^3
            RefBox b;
            if (!ref_bbox(d, n.child, 0.0, 1.0, b, depth + 1)) return false;
            for (int a = 0; a < 3; ++a) {
                out.lo[a] = -DBL_MAX;
                out.hi[a] = DBL_MAX;
            }
            return true;
        }
        case RT_NODE_FLIP:    // hit.rs:122-124
        case RT_NODE_MEDIUM:  // medium.rs:63-65
            return ref_bbox(d, n.child, t0, t1, out, depth + 1);
        default:
            return false;
    }
}

// Leaf order of BVH::new(hit, t0, t1): left subtree first, then right (bvh.rs:64-70,79-84).
// The reference recomputes every bounding box at every level and sorts a fresh Vec per node; the
// order it arrives at only depends on the boxes, so they are computed once, the recursion sorts
// index ranges in place, and the two halves of a large node are ordered in parallel.
struct RefOrder {
    struct Item {
        double key;    // lo + hi on the split axis of the node being sorted (bvh.rs:23-25)
        uint32_t idx;  // index into hit / boxes
    };
    const std::vector<RefBox> &boxes;
    std::vector<Item> items;  // permutation of [0, n), sorted range by range
    std::vector<Item> tmp;    // the radix sort's other buffer (a node only touches its own range)
    std::atomic<int> bad{0};  // 1: NaN extent, 2: NaN centroid

    // Stable LSD radix sort of items[first, first + n) by key (no NaN among them): the same order as
    // std::stable_sort with `x.key < y.key`, in a third of the time on the large nodes of a mesh.  -0.0 and +0.0
    // compare equal there, so both map to the same integer here.
    static uint64_t sortable(double key) {
        uint64_t u;
        key += 0.0;  // -0.0 -> +0.0
        std::memcpy(&u, &key, 8);
        return u ^ ((u >> 63) ? ~0ull : 0x8000000000000000ull);
    }
    void radix_sort(size_t first, size_t n) {
        uint32_t hist[8][256];
        std::memset(hist, 0, sizeof(hist));
        Item *a = items.data() + first, *b = tmp.data() + first;
        for (size_t i = 0; i < n; ++i) {
            const uint64_t u = sortable(a[i].key);
            for (int d = 0; d < 8; ++d) hist[d][(u >> (8 * d)) & 255u]++;
        }
        for (int d = 0; d < 8; ++d) {
            uint32_t *h = hist[d];
            bool trivial = false;
            for (int k = 0; k < 256; ++k)
                if (h[k] == n) trivial = true;  // every key has the same digit here
            if (trivial) continue;
            uint32_t acc = 0;
            for (int k = 0; k < 256; ++k) {
                const uint32_t c = h[k];
                h[k] = acc;
                acc += c;
            }
            for (size_t i = 0; i < n; ++i) b[h[(sortable(a[i].key) >> (8 * d)) & 255u]++] = a[i];
            std::swap(a, b);
        }
        if (a != items.data() + first) std::copy(a, a + n, items.data() + first);
    }

    void rec(size_t first, size_t n, int depth) {
        if (n <= 1 || bad.load(std::memory_order_relaxed)) return;
        // bvh.rs:33-48; plain comparisons instead of fmin / fmax (libm calls at -O2): a NaN bound is
        // skipped either way, and a NaN key is caught below
        double bmin[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, bmax[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (size_t i = first; i < first + n; ++i) {
            const RefBox &b = boxes[items[i].idx];
            for (int a = 0; a < 3; ++a) {
                bmin[a] = b.lo[a] < bmin[a] ? b.lo[a] : bmin[a];
                bmax[a] = b.hi[a] > bmax[a] ? b.hi[a] : bmax[a];
            }
        }
        int axis = 0;
        double range[3];
        for (int a = 0; a < 3; ++a) {
            range[a] = bmax[a] - bmin[a];
            if (range[a] != range[a]) {  // partial_cmp().unwrap(), bvh.rs:47
                bad.store(1);
                return;
            }
        }
        if (range[1] > range[axis]) axis = 1;
        if (range[2] > range[axis]) axis = 2;
        for (size_t i = first; i < first + n; ++i) {
            const RefBox &b = boxes[items[i].idx];
            const double key = b.lo[axis] + b.hi[axis];
            if (key != key) {  // partial_cmp().unwrap(), bvh.rs:26
                bad.store(2);
                return;
            }
            items[i].key = key;
        }
        // bvh.rs:51 sort_unstable_by: the order of equal keys is unspecified in the reference;
        // a stable sort fixes it (same choice as the oracle)
        if (n >= 1024 && !std::getenv("RTB200_COMPILE_SERIAL")) radix_sort(first, n);  // (the switch: the tests compare the two)
        else std::stable_sort(items.begin() + first, items.begin() + first + n, [](const Item &x, const Item &y) { return x.key < y.key; });
        const size_t half = n / 2;
        if (n >= 16384 && depth < 5) {
            auto left = std::async(std::launch::async, [&] { rec(first, half, depth + 1); });
            rec(first + half, n - half, depth + 1);
            left.get();
        } else {
            rec(first, half, depth + 1);
            rec(first + half, n - half, depth + 1);
        }
    }
};

bool ref_bvh_order(const RtSceneDesc &d, const std::vector<uint32_t> &hit, double t0, double t1, std::vector<uint32_t> &out,
                   std::string &err) {
    if (hit.empty()) {
        err = "no object in the scene";  // bvh.rs:55
        return false;
    }
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (getenv("RTB200_COMPILE_TIMING") && hit.size() > 10000) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[ref order %zu] %s %.3f s\n", hit.size(), what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    std::vector<RefBox> boxes(hit.size());
    std::atomic<int> no_box{0};
    parallel_for(hit.size(), [&](size_t i0, size_t i1) {
        for (size_t i = i0; i < i1; ++i)
            if (!ref_bbox(d, hit[i], t0, t1, boxes[i], 0)) no_box.store(1);
    });
    if (no_box.load()) {
        err = "no bounding box in bvh node";  // bvh.rs:28,61
        return false;
    }
    lap("boxes");
    RefOrder ro{boxes, std::vector<RefOrder::Item>(hit.size()), std::vector<RefOrder::Item>(hit.size())};
    for (size_t i = 0; i < hit.size(); ++i) ro.items[i] = RefOrder::Item{0.0, (uint32_t)i};
    ro.rec(0, hit.size(), 0);
    lap("sort");
    if (ro.bad.load()) {
        err = ro.bad.load() == 1 ? "NaN extent in BVH build" : "NaN centroid in BVH build";
        return false;
    }
    // Every object of a reference BVH sits in a Leaf node of its own whose box is tested first
    // (bvh.rs:56-63,77-78), and AABB::hit rejects with `t_out <= t_in` (aabb.rs:31): a box without
    // extent along some axis has t0 == t1 there and is never entered.  So an axis-aligned flat
    // triangle (or a list of them) directly under a BVH - an OBJ cube through mesh.rs and
    // main.rs:442, say - is invisible in the reference, and it is left out here.  (The one ray that
    // gets through has a zero direction component and its origin exactly in that plane: 0 * inf = NaN,
    // which f64::max / min skip.)  The object still took part in the ordering above.
    out.reserve(out.size() + hit.size());
    for (size_t i = 0; i < hit.size(); ++i) {
        const RefBox &b = boxes[ro.items[i].idx];
        if (b.lo[0] < b.hi[0] && b.lo[1] < b.hi[1] && b.lo[2] < b.hi[2]) out.push_back(hit[ro.items[i].idx]);
    }
    return true;
}

struct Walker {
    const RtSceneDesc &d;
    CompiledScene &out;
    std::string &err;
    Walker(const RtSceneDesc &desc, CompiledScene &compiled, std::string &message) : d(desc), out(compiled), err(message) {}
    RtStatus status = RT_OK;
    std::vector<DOp> stack;
    std::vector<std::vector<DOp>> chain_ops;  // interned chains
    uint32_t rank = 0;
    // > 0 below a BVH node.  The reference scans the members of a HittableList linearly and only walks a tree where
    // the scene says BVH::new, and the groups keep that distinction: what hangs directly in a list (the walls of a
    // room) is a group of its own, scanned linearly when it is small, instead of being mixed into the tree of the
    // mesh next to it - every ray would then walk that tree, where now only the rays that reach the mesh's bounds do
    // (megakernel.inl: render_deferred_kernel lets them wait for each other).
    int in_bvh = 0;
    std::vector<GroupBuild> *groups = nullptr;
    std::vector<PendingMedium> *media = nullptr;  // null while walking a medium boundary

    bool fail(RtStatus s, const std::string &m) {
        if (status == RT_OK) {
            status = s;
            err = m;
        }
        return false;
    }
    // chains and groups are looked up by a hash of their ops (a scene of n individually transformed instances has n
    // of each; the linear scans that were here made the walk quadratic in n)
    static uint64_t ops_hash(const std::vector<DOp> &ops) {
        uint64_t h = 1469598103934665603ull;
        for (const DOp &op : ops) {
            uint64_t w[7] = {op.kind, op.axis, 0, 0, 0, 0, 0};
            std::memcpy(&w[2], &op.sin_theta, sizeof(double) * 5);
            for (uint64_t x : w) h = (h ^ x) * 1099511628211ull;
        }
        return h;
    }
    std::unordered_multimap<uint64_t, uint32_t> chain_index;
    std::unordered_multimap<uint64_t, size_t> group_index;  // of the vector `groups` points at (use_groups resets it)
    void use_groups(std::vector<GroupBuild> *g) {
        groups = g;
        group_index.clear();
        for (size_t i = 0; i < g->size(); ++i) group_index.emplace(ops_hash((*g)[i].xform) * 2u + ((*g)[i].tree ? 1u : 0u), i);
        cache_valid = false;
    }
    uint32_t intern_chain(const std::vector<DOp> &ops) {
        const uint64_t h = ops_hash(ops);
        auto range = chain_index.equal_range(h);
        for (auto it = range.first; it != range.second; ++it)
            if (same_ops(chain_ops[it->second], ops)) return it->second;
        chain_ops.push_back(ops);
        chain_index.emplace(h, (uint32_t)(chain_ops.size() - 1));
        return (uint32_t)(chain_ops.size() - 1);
    }
    GroupBuild &group_for_stack() {
        std::vector<DOp> xf;
        for (const DOp &op : stack)
            if (op.kind != OP_FLIP) xf.push_back(op);
        const uint64_t key = ops_hash(xf) * 2u + (in_bvh > 0 ? 1u : 0u);
        auto &index = group_index;
        auto range = index.equal_range(key);
        for (auto it = range.first; it != range.second; ++it) {
            GroupBuild &g = (*groups)[it->second];
            if (g.tree == (in_bvh > 0) && same_ops(g.xform, xf)) return g;
        }
        index.emplace(key, groups->size());
        groups->emplace_back();
        groups->back().xform = xf;
        groups->back().tree = in_bvh > 0;
        groups->back().chain = intern_chain(xf);
        return groups->back();
    }
    bool check_material(uint32_t m) {
        if (m >= d.n_materials) return fail(RT_ERR_BAD_ARGUMENT, "material index out of range");
        return true;
    }
    // The wrapper stack only changes between siblings of a wrapper node; a mesh emits hundreds of
    // thousands of primitives under one stack, so the (chain, group) lookup is cached per stack state.
    size_t cached_depth = (size_t)-1;
    const void *cached_groups = nullptr;
    uint32_t cached_chain = 0;
    size_t cached_group = 0;
    void group_of_stack() {
        if (!(cache_valid && cached_depth == stack.size() && cached_groups == (const void *)groups)) {
            cached_chain = intern_chain(stack);
            GroupBuild &g = group_for_stack();
            cached_group = (size_t)(&g - groups->data());
            cached_depth = stack.size();
            cached_groups = (const void *)groups;
            cache_valid = true;
        }
    }
    void emit(DPrim p, uint32_t node_id) {
        group_of_stack();
        p.chain = cached_chain;
        p.node = (int32_t)node_id;
        p.rank = rank++;
        p.pad0 = p.pad1 = 0;
        for (int i = 0; i < 12; ++i)
            if (p.d[i] != p.d[i]) {
                fail(RT_ERR_BAD_ARGUMENT, "NaN primitive parameter");
                return;
            }
        (*groups)[cached_group].prims.push_back(p);
    }
    bool cache_valid = false;

    // The record of a primitive node (everything but chain / node / rank).  false: not a primitive, or a
    // parameter walk() reports an error for.
    bool prim_record(const RtNode &n, DPrim &p) const {
        std::memset(&p, 0, sizeof(p));
        p.material = n.material;
        if (n.material >= d.n_materials) return false;
        switch (n.kind) {
            case RT_NODE_SPHERE:
                p.kind = PRIM_SPHERE;
                for (int i = 0; i < 4; ++i) p.d[i] = n.v[i];
                return true;
            case RT_NODE_MOVING_SPHERE:
                p.kind = PRIM_MSPHERE;
                for (int i = 0; i < 9; ++i) p.d[i] = n.v[i];
                return true;
            case RT_NODE_RECT:
                if (n.axis > 2) return false;
                p.kind = PRIM_RECT;
                p.axis = n.axis;
                for (int i = 0; i < 5; ++i) p.d[i] = n.v[i];
                return true;
            case RT_NODE_TRIANGLE: {
                p.kind = PRIM_TRI;
                // tri.rs:27-28: e1 = v1 - v0, e2 = v2 - v0 ; tri.rs:41: normal = normalize(e1 x e2)
                double e1[3], e2[3];
                for (int a = 0; a < 3; ++a) {
                    p.d[a] = n.v[a];
                    e1[a] = n.v[3 + a] - n.v[a];
                    e2[a] = n.v[6 + a] - n.v[a];
                    p.d[3 + a] = e1[a];
                    p.d[6 + a] = e2[a];
                }
                double c[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                double len = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
                // A degenerate triangle gives a NaN normal in the reference too; keep the
                // primitive (its hit test uses only v0,e1,e2) but store a zero normal marker.
                for (int a = 0; a < 3; ++a) p.d[9 + a] = len > 0.0 ? c[a] / len : 0.0;
                return true;
            }
            case RT_NODE_CUBE:
                p.kind = PRIM_BOX;
                for (int i = 0; i < 6; ++i) p.d[i] = n.v[i];
                return true;
            default:
                return false;
        }
    }

    // A large BVH over bare primitives (a mesh: main.rs:442) is emitted in parallel: the i-th child in reference
    // order gets rank0 + i, exactly what the serial walk hands out.  Anything else - a wrapper or container among
    // the children, a parameter that is an error - returns false with nothing changed, and the serial walk runs
    // (and reports).
    bool emit_many(const std::vector<uint32_t> &order) {
        if (order.size() < 65536 || std::getenv("RTB200_COMPILE_SERIAL")) return false;  // the switch is for the tests
        group_of_stack();
        std::vector<DPrim> &dst = (*groups)[cached_group].prims;
        const size_t base = dst.size();
        dst.resize(base + order.size());
        const uint32_t rank0 = rank, chain = cached_chain;
        std::atomic<int> bad{0};
        parallel_for(order.size(), [&](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1 && !bad.load(std::memory_order_relaxed); ++i) {
                const uint32_t id = order[i];
                DPrim p;
                bool ok = id < d.n_nodes && prim_record(d.nodes[id], p);
                for (int k = 0; ok && k < 12; ++k) ok = p.d[k] == p.d[k];
                if (!ok) {
                    bad.store(1);
                    return;
                }
                p.chain = chain;
                p.node = (int32_t)id;
                p.rank = rank0 + (uint32_t)i;
                dst[base + i] = p;
            }
        });
        if (bad.load()) {
            dst.resize(base);
            return false;
        }
        rank = rank0 + (uint32_t)order.size();
        return true;
    }
    bool walk(uint32_t id, uint32_t depth) {
        if (status != RT_OK) return false;
        if (id >= d.n_nodes) return fail(RT_ERR_BAD_ARGUMENT, "node index out of range");
        if (depth > d.n_nodes + 1) return fail(RT_ERR_BAD_ARGUMENT, "cycle in scene graph");
        const RtNode &n = d.nodes[id];
        DPrim p;
        switch (n.kind) {
            case RT_NODE_CUBE: {
                // Cube::new (cube.rs:14-30) is six AARects, and a rect whose range is inverted never passes
                // `a < a0 || a > a1` (rect.rs:55).  A cube with min > max on one axis therefore shows only the two
                // faces perpendicular to that axis, and none with two inverted axes; the slab test of the device's
                // box primitive would show a whole box.  (Found by fuzzing degenerate parameters; such a cube under a
                // BVH is never entered at all - ref_bvh_order.)
                if (!check_material(n.material)) return false;
                int inverted = 0, axis = -1;
                for (int a = 0; a < 3; ++a)
                    if (n.v[a] > n.v[3 + a]) {
                        ++inverted;
                        axis = a;
                    }
                if (inverted == 0) {
                    if (!prim_record(n, p)) return fail(RT_ERR_BAD_ARGUMENT, "bad cube");
                    emit(p, id);
                } else if (inverted == 1) {
                    // the two faces at max[axis] and min[axis], in the order of cube.rs (the parity hook reports
                    // face 0 for them: DPrim has no face field for a rect)
                    const int a0 = axis == 0 ? 1 : 0, a1 = axis == 2 ? 1 : 2;  // x: (y, z)  y: (x, z)  z: (x, y)
                    for (int side = 0; side < 2; ++side) {
                        std::memset(&p, 0, sizeof(p));
                        p.material = n.material;
                        p.kind = PRIM_RECT;
                        p.axis = axis == 0 ? RT_PLANE_YZ : (axis == 1 ? RT_PLANE_XZ : RT_PLANE_XY);
                        p.d[0] = n.v[a0];
                        p.d[1] = n.v[3 + a0];
                        p.d[2] = n.v[a1];
                        p.d[3] = n.v[3 + a1];
                        p.d[4] = side == 0 ? n.v[3 + axis] : n.v[axis];
                        emit(p, id);
                    }
                }
                break;
            }
            case RT_NODE_SPHERE:
            case RT_NODE_MOVING_SPHERE:
            case RT_NODE_RECT:
            case RT_NODE_TRIANGLE:
                if (!check_material(n.material)) return false;
                if (!prim_record(n, p)) return fail(RT_ERR_BAD_ARGUMENT, "bad rect plane");
                emit(p, id);
                break;
            case RT_NODE_LIST:
            case RT_NODE_BVH: {
                if ((uint64_t)n.child + n.count > d.n_child_index) return fail(RT_ERR_BAD_ARGUMENT, "child range out of bounds");
                if (n.kind == RT_NODE_BVH) {
                    // visit the children in the reference BVH's leaf order so that ranks follow
                    // the reference's traversal order (tie-break, §Q17)
                    std::vector<uint32_t> kids(d.child_index + n.child, d.child_index + n.child + n.count), order;
                    std::string e;
                    if (!ref_bvh_order(d, kids, n.v[0], n.v[1], order, e))
                        return fail(n.count == 0 ? RT_ERR_EMPTY_SCENE : RT_ERR_BAD_ARGUMENT, e);
                    ++in_bvh;
                    cache_valid = false;
                    auto TE = std::chrono::steady_clock::now();
                    bool many = emit_many(order);
                    if (getenv("RTB200_COMPILE_TIMING") && order.size() > 10000) fprintf(stderr, "[walk] emit_many(%zu) = %d: %.3f s\n", order.size(), (int)many, std::chrono::duration<double>(std::chrono::steady_clock::now() - TE).count());
                    bool ok = true;
                    if (!many)
                        for (uint32_t id2 : order)
                            if (!walk(id2, depth + 1)) {
                                ok = false;
                                break;
                            }
                    --in_bvh;
                    cache_valid = false;
                    if (!ok) return false;
                } else {
                    for (uint32_t i = 0; i < n.count; ++i)
                        if (!walk(d.child_index[n.child + i], depth + 1)) return false;
                }
                break;
            }
            case RT_NODE_TRANSLATE: {
                DOp op;
                std::memset(&op, 0, sizeof(op));
                op.kind = OP_TRANSLATE;
                for (int a = 0; a < 3; ++a) op.offset[a] = n.v[a];
                stack.push_back(op);
                cache_valid = false;
                bool ok = walk(n.child, depth + 1);
                stack.pop_back();
                cache_valid = false;
                if (!ok) return false;
                break;
            }
            case RT_NODE_ROTATE: {
                if (n.axis > 2) return fail(RT_ERR_BAD_ARGUMENT, "bad rotate axis");
                DOp op;
                std::memset(&op, 0, sizeof(op));
                op.kind = OP_ROTATE;
                op.axis = n.axis;
                double radiants = (PI / 180.0) * n.v[0];  // rotate.rs:34-36
                op.sin_theta = std::sin(radiants);
                op.cos_theta = std::cos(radiants);
                stack.push_back(op);
                cache_valid = false;
                bool ok = walk(n.child, depth + 1);
                stack.pop_back();
                cache_valid = false;
                if (!ok) return false;
                break;
            }
            case RT_NODE_FLIP: {
                DOp op;
                std::memset(&op, 0, sizeof(op));
                op.kind = OP_FLIP;
                stack.push_back(op);
                cache_valid = false;
                bool ok = walk(n.child, depth + 1);
                stack.pop_back();
                cache_valid = false;
                if (!ok) return false;
                break;
            }
            case RT_NODE_MEDIUM: {
                if (!check_material(n.material)) return false;
                if (!media) return fail(RT_ERR_UNSUPPORTED, "a ConstantMedium inside a medium boundary is not supported");
                media->push_back(PendingMedium{id, stack, rank++});
                break;
            }
            default:
                return fail(RT_ERR_BAD_ARGUMENT, "unknown node kind");
        }
        return status == RT_OK;
    }
};

bool finalize_groups(CompiledScene &out, std::vector<GroupBuild> &gb, std::string &err, const CompileOptions &opts) {
    // List members stay apart from the tree of the same transform only while they are few enough to be scanned
    // linearly; a larger flat part would need a second tree that every ray walks next to the first, so it joins
    // the tree (measured on the Next Week final scene, 7 list members next to the 400 ground boxes: one tree 36.1 ms,
    // two trees 36.8 ms; on the mesh scene, 6 walls next to 394k triangles: 190.6 ms in one tree, 126.6 ms apart).
    for (GroupBuild &flat : gb) {
        if (flat.tree || flat.prims.size() <= LINEAR_MAX) continue;
        for (GroupBuild &tree : gb)
            if (tree.tree && !tree.prims.empty() && same_ops(tree.xform, flat.xform)) {
                tree.prims.insert(tree.prims.end(), flat.prims.begin(), flat.prims.end());
                flat.prims.clear();
                break;
            }
    }
    size_t n_live = 0;
    for (const GroupBuild &g : gb) n_live += g.prims.empty() ? 0 : 1;
    for (GroupBuild &g : gb) {
        if (g.prims.empty()) continue;
        DGroup dg;
        std::memset(&dg, 0, sizeof(dg));
        dg.chain = g.chain;
        dg.first_prim = (uint32_t)out.prims.size();
        dg.n_prims = (uint32_t)g.prims.size();
        if (out.prims.size() + g.prims.size() >= (1u << 28)) {
            err = "too many primitives";
            return false;
        }
        out.prims.insert(out.prims.end(), g.prims.begin(), g.prims.end());
        Box ob;
        ob.reset();
        for (const DPrim &p : g.prims) ob.grow(prim_box(p));
        if (!ob.finite()) {
            err = "non-finite primitive bounds";
            return false;
        }
        ob.pad();
        Box wb;
        wb.reset();
        for (int c = 0; c < 8; ++c) {
            double p[3] = {(c & 1) ? ob.hi[0] : ob.lo[0], (c & 2) ? ob.hi[1] : ob.lo[1], (c & 4) ? ob.hi[2] : ob.lo[2]};
            point_to_outer(g.xform, p);
            wb.grow(p);
        }
        wb.pad();
        for (int a = 0; a < 3; ++a) {
            dg.bmin[a] = wb.lo[a];
            dg.bmax[a] = wb.hi[a];
        }
        bool has_bvh = dg.n_prims > LINEAR_MAX;
        if (has_bvh && opts.gpu_bvh_min_prims > 0 && dg.n_prims >= opts.gpu_bvh_min_prims) {
            // the tree is left to gpu_bvh.cu: the primitives stay in walk order, their boxes go along as fp32
            CompiledScene::PendingBvh pb;
            pb.group = (uint32_t)out.groups.size();
            pb.first_prim = dg.first_prim;
            pb.n_prims = dg.n_prims;
            pb.node_base = (uint32_t)out.nodes.size();
            pb.first_box = out.pending_boxes.size() / 6;
            for (int a = 0; a < 3; ++a) {
                pb.lo[a] = ob.lo[a];
                pb.hi[a] = ob.hi[a];
            }
            out.pending_boxes.resize(out.pending_boxes.size() + 6 * (size_t)dg.n_prims);
            float *dst = out.pending_boxes.data() + 6 * pb.first_box;
            const uint32_t first = dg.first_prim;
            parallel_for(dg.n_prims, [&](size_t i0, size_t i1) {
                for (size_t i = i0; i < i1; ++i) {
                    Box b = prim_box(out.prims[first + i]);
                    b.pad();
                    for (int a = 0; a < 3; ++a) {
                        dst[6 * i + a] = f32_down(b.lo[a]);
                        dst[6 * i + 3 + a] = f32_up(b.hi[a]);
                    }
                }
            });
            DBvhNode zero;
            std::memset(&zero, 0, sizeof(zero));
            out.nodes.resize(out.nodes.size() + (size_t)dg.n_prims - 1, zero);
            dg.bvh_root = (int32_t)pb.node_base;
            out.max_bvh_depth = std::max(out.max_bvh_depth, 62u);  // a radix tree over 62-bit keys (gpu_bvh.cu)
            out.pending_bvh.push_back(pb);
        } else if (has_bvh) {
            dg.bvh_root = build_group_bvh(out, dg.first_prim, dg.n_prims);
        } else {
            // a small group is one leaf: same encoding as a BVH leaf (LINEAR_MAX <= 8)
            dg.bvh_root = (int32_t) ~((dg.first_prim << 3) | (dg.n_prims - 1));
        }
        // one or two primitives are cheaper to test than to cull; a BVH culls with its own root boxes
        if (dg.n_prims > 2 && !has_bvh) dg.flags |= GROUP_CULL;
        // (except next to other groups or under a transform: there the group's bounds save the transform and the root visit)
        if (has_bvh && (!g.xform.empty() || n_live > 1)) dg.flags |= GROUP_CULL;
        double M[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
        for (const DOp &op : g.xform) {
            dg.flags |= GROUP_XFORM;
            if (op.kind == OP_TRANSLATE) {  // p' = p - offset (translate.rs:23)
                for (int a = 0; a < 3; ++a) T[a] -= op.offset[a];
            } else if (op.kind == OP_ROTATE) {  // p' = R p (rotate.rs:82-86)
                dg.flags |= GROUP_ROTATED;
                int r, a, b;
                rotate_axes(op.axis, r, a, b);
                double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
                R[a * 3 + a] = op.cos_theta; R[a * 3 + b] = -op.sin_theta;
                R[b * 3 + a] = op.sin_theta; R[b * 3 + b] = op.cos_theta;
                double M2[9], T2[3];
                for (int i = 0; i < 3; ++i) {
                    T2[i] = R[i * 3] * T[0] + R[i * 3 + 1] * T[1] + R[i * 3 + 2] * T[2];
                    for (int j = 0; j < 3; ++j) M2[i * 3 + j] = R[i * 3] * M[j] + R[i * 3 + 1] * M[3 + j] + R[i * 3 + 2] * M[6 + j];
                }
                std::memcpy(M, M2, sizeof(M));
                std::memcpy(T, T2, sizeof(T));
            }
        }
        std::memcpy(dg.m, M, sizeof(M));
        std::memcpy(dg.t, T, sizeof(T));
        out.groups.push_back(dg);
    }
    // A linearly scanned group whose bounds ARE the sub-scene's bounds (the walls of a room around everything else)
    // gains nothing from its cull test: a ray that starts inside the scene is inside those bounds, and one from
    // outside that misses them misses everything.  The test was 35 f64 instructions per segment at full warp width
    // on the Cornell box and never rejected a ray.  (Culling only skips work: the result cannot change.)
    const size_t g0 = out.groups.size() - n_live;
    if (std::getenv("RTB200_KEEP_ENCLOSING_CULL")) return true;  // A/B switch (r2-m)
    if (n_live > 1) {
        Box all;
        all.reset();
        for (size_t k = g0; k < out.groups.size(); ++k) {
            all.grow(out.groups[k].bmin);
            all.grow(out.groups[k].bmax);
        }
        for (size_t k = g0; k < out.groups.size(); ++k) {
            DGroup &dg = out.groups[k];
            bool encloses = dg.bvh_root < 0 && (dg.flags & GROUP_CULL);
            for (int a = 0; a < 3 && encloses; ++a) encloses = dg.bmin[a] <= all.lo[a] && dg.bmax[a] >= all.hi[a];
            if (encloses) dg.flags &= ~(uint32_t)GROUP_CULL;
        }
    } else if (n_live == 1 && out.groups[g0].bvh_root < 0) {
        out.groups[g0].flags &= ~(uint32_t)GROUP_CULL;  // the only group: nothing to skip to
    }
    return true;
}

bool texture_is_costly(const RtSceneDesc &d, uint32_t id, int depth) {
    if (id >= d.n_textures || depth > 16) return false;
    const RtTexture &t = d.textures[id];
    if (t.kind == RT_TEX_NOISE || t.kind == RT_TEX_IMAGE) return true;
    if (t.kind == RT_TEX_CHECKER) return texture_is_costly(d, t.a, depth + 1) || texture_is_costly(d, t.b, depth + 1);
    return false;
}
bool texture_needs_uv(const RtSceneDesc &d, uint32_t id, int depth) {
    if (id >= d.n_textures || depth > 16) return false;
    const RtTexture &t = d.textures[id];
    if (t.kind == RT_TEX_IMAGE) return true;
    if (t.kind == RT_TEX_CHECKER) return texture_needs_uv(d, t.a, depth + 1) || texture_needs_uv(d, t.b, depth + 1);
    return false;
}

}  // namespace

RtStatus compile_scene(const RtSceneDesc &d, CompiledScene &out, std::string &err, const CompileOptions &opts) {
    if (d.abi_version != RTB200_ABI_VERSION) {
        err = "abi version mismatch";
        return RT_ERR_BAD_ARGUMENT;
    }
    if (!d.nodes || d.n_nodes == 0) {
        err = "no object in the scene";
        return RT_ERR_EMPTY_SCENE;
    }
    if ((d.n_child_index && !d.child_index) || (d.n_materials && !d.materials) || (d.n_textures && !d.textures) ||
        (d.n_perlin && !d.perlin) || (d.n_images && !d.images) || (d.n_texel_bytes && !d.texels)) {
        err = "null table pointer with a non-zero count";
        return RT_ERR_BAD_ARGUMENT;
    }
    // ---- textures / materials ----
    for (uint64_t i = 0; i < d.n_textures; ++i) {
        const RtTexture &t = d.textures[i];
        DTexture dt;
        std::memset(&dt, 0, sizeof(dt));
        dt.kind = t.kind;
        dt.a = t.a;
        dt.b = t.b;
        for (int a = 0; a < 3; ++a) dt.color[a] = t.color[a];
        dt.scale = t.scale;
        switch (t.kind) {
            case RT_TEX_CONSTANT: break;
            case RT_TEX_CHECKER:
                // children must precede or follow; only the range is checked (cycles would loop: bound the walk on device)
                if (t.a >= d.n_textures || t.b >= d.n_textures) {
                    err = "checker texture child out of range";
                    return RT_ERR_BAD_ARGUMENT;
                }
                break;
            case RT_TEX_NOISE:
                if (t.a >= d.n_perlin) {
                    err = "noise texture perlin index out of range";
                    return RT_ERR_BAD_ARGUMENT;
                }
                break;
            case RT_TEX_IMAGE:
                if (t.a >= d.n_images) {
                    err = "image texture index out of range";
                    return RT_ERR_BAD_ARGUMENT;
                }
                break;
            default:
                err = "unknown texture kind";
                return RT_ERR_BAD_ARGUMENT;
        }
        out.textures.push_back(dt);
    }
    for (uint64_t i = 0; i < d.n_images; ++i) {
        const RtImage &im = d.images[i];
        if (im.width == 0 || im.height == 0 || im.offset + (uint64_t)im.width * im.height * 3 > d.n_texel_bytes) {
            err = "image outside the texel pool";
            return RT_ERR_BAD_ARGUMENT;
        }
        out.images.push_back(DImage{im.width, im.height, im.offset});
    }
    out.perlin.resize(d.n_perlin);
    for (uint64_t i = 0; i < d.n_perlin; ++i) {
        static_assert(sizeof(DPerlin) == sizeof(RtPerlin), "perlin layout");
        std::memcpy(&out.perlin[i], &d.perlin[i], sizeof(DPerlin));
        for (int k = 0; k < 256; ++k)
            if (d.perlin[i].perm_x[k] > 255 || d.perlin[i].perm_y[k] > 255 || d.perlin[i].perm_z[k] > 255) {
                err = "perlin permutation entry out of range";
                return RT_ERR_BAD_ARGUMENT;
            }
    }
    if (d.n_texel_bytes) out.texels.assign(d.texels, d.texels + d.n_texel_bytes);
    for (uint64_t i = 0; i < d.n_materials; ++i) {
        const RtMaterial &m = d.materials[i];
        DMaterial dm;
        std::memset(&dm, 0, sizeof(dm));
        dm.kind = m.kind;
        dm.texture = m.texture;
        for (int a = 0; a < 3; ++a) dm.albedo[a] = m.albedo[a];
        dm.fuzz = m.fuzz;
        dm.ir = m.ir;
        for (int a = 0; a < 10; ++a) dm.pbr[a] = m.pbr[a];
        if (m.kind > RT_MAT_PBR) {
            err = "unknown material kind";
            return RT_ERR_UNSUPPORTED;
        }
        bool textured = m.kind == RT_MAT_LAMBERTIAN || m.kind == RT_MAT_DIFFUSE_LIGHT || m.kind == RT_MAT_ISOTROPIC || m.kind == RT_MAT_PBR;
        if (textured && m.texture >= d.n_textures) {
            err = "texture index out of range";
            return RT_ERR_BAD_ARGUMENT;
        }
        dm.needs_uv = textured && texture_needs_uv(d, m.texture, 0) ? 1u : 0u;
        dm.costly = textured && texture_is_costly(d, m.texture, 0) ? 1u : 0u;
        out.materials.push_back(dm);
    }
    for (int a = 0; a < 3; ++a) out.background[a] = d.background[a];

    // ---- world ----
    auto T0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { if (getenv("RTB200_COMPILE_TIMING")) { auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[compile] %s %.3f s\n", what, std::chrono::duration<double>(t - T0).count()); T0 = t; } };
    Walker w(d, out, err);
    std::vector<GroupBuild> world_groups;
    std::vector<PendingMedium> pending;
    w.use_groups(&world_groups);
    w.media = &pending;
    w.cache_valid = false;
    if (!w.walk(d.world, 0)) return w.status;
    lap("walk");
    if (!finalize_groups(out, world_groups, err, opts)) return RT_ERR_BAD_ARGUMENT;
    lap("finalize (BVH build)");
    out.n_world_groups = (uint32_t)out.groups.size();
    if (out.n_world_groups == 0 && pending.empty()) {
        err = "no object in the scene";
        return RT_ERR_EMPTY_SCENE;
    }
    // ---- media: each boundary is its own sub-scene ----
    for (const PendingMedium &pm : pending) {
        const RtNode &n = d.nodes[pm.node];
        std::vector<GroupBuild> bgroups;
        w.use_groups(&bgroups);
        w.media = nullptr;
        w.stack = pm.stack;
        w.cache_valid = false;
        if (!w.walk(n.child, 0)) return w.status;
        DMedium dm;
        std::memset(&dm, 0, sizeof(dm));
        dm.first_group = (uint32_t)out.groups.size();
        if (!finalize_groups(out, bgroups, err, opts)) return RT_ERR_BAD_ARGUMENT;
        dm.n_groups = (uint32_t)out.groups.size() - dm.first_group;
        dm.chain = w.intern_chain(pm.stack);
        dm.material = n.material;
        dm.node = (int32_t)pm.node;
        dm.rank = pm.rank;
        dm.density = n.v[0];
        dm.convex_prim = 0xFFFFFFFFu;
        dm.pad0 = dm.pad1 = dm.pad2 = 0;
        if (dm.n_groups == 1 && !std::getenv("RTB200_NO_CONVEX_MEDIA")) {  // (the switch is for the A/B of r2-n)
            const DGroup &bg = out.groups[dm.first_group];
            // only under a transform: that is what the second query would repeat (measured, r2-n: Cornell smoke's
            // rotated boxes +11 %, the untransformed spheres of the Next Week final scene -2 %)
            if (bg.n_prims == 1 && bg.bvh_root < 0 && !(bg.flags & GROUP_CULL) && (bg.flags & GROUP_XFORM) &&
                (out.prims[bg.first_prim].kind == PRIM_BOX || out.prims[bg.first_prim].kind == PRIM_SPHERE))
                dm.convex_prim = bg.first_prim;
        }
        out.media.push_back(dm);
    }
    w.stack.clear();
    for (const DPrim &pr : out.prims)
        if (pr.kind == PRIM_MSPHERE && (pr.d[6] != 0.0 || pr.d[7] != 1.0)) out.shutter_limited = true;
    // ---- chains / ops ----
    for (const std::vector<DOp> &ops : w.chain_ops) {
        DChain c{(uint32_t)out.ops.size(), (uint32_t)ops.size()};
        out.ops.insert(out.ops.end(), ops.begin(), ops.end());
        out.chains.push_back(c);
    }
    // ---- lights (pdf.rs PDF::Hittable over the light list) ----
    if (d.lights >= d.n_nodes || d.nodes[d.lights].kind != RT_NODE_LIST) {
        err = "lights must be a LIST node";
        return RT_ERR_BAD_ARGUMENT;
    }
    const RtNode &ln = d.nodes[d.lights];
    if ((uint64_t)ln.child + ln.count > d.n_child_index) {
        err = "light list out of bounds";
        return RT_ERR_BAD_ARGUMENT;
    }
    if (ln.count > 2048) {
        err = "more than 2048 lights";  // the light index comes from 11 spare Philox bits
        return RT_ERR_UNSUPPORTED;
    }
    for (uint32_t i = 0; i < ln.count; ++i) {
        uint32_t id = d.child_index[ln.child + i];
        uint32_t guard = 0;
        // FlipNormal forwards pdf_value/random (hit.rs:126-132); every other wrapper keeps the
        // trait defaults (hit.rs:29-30): pdf 0 and direction (1,0,0).
        while (id < d.n_nodes && d.nodes[id].kind == RT_NODE_FLIP && guard++ < d.n_nodes) id = d.nodes[id].child;
        if (id >= d.n_nodes) {
            err = "light node out of range";
            return RT_ERR_BAD_ARGUMENT;
        }
        const RtNode &n = d.nodes[id];
        DLight l;
        std::memset(&l, 0, sizeof(l));
        // A HittableList forwards pdf_value / random to its members (hit.rs:90-96), so a list nested in the light
        // list is sampled through a second `choose` with random numbers of its own; the slot-addressed draw of
        // DESIGN.md §3 has one light index per scatter.  Refused rather than sampled differently from the reference.
        if (n.kind == RT_NODE_LIST) {
            err = "a HittableList nested in the light list is not supported (flatten it into the light list)";
            return RT_ERR_UNSUPPORTED;
        }
        if (n.kind == RT_NODE_RECT) {
            if (n.axis > 2) {
                err = "bad rect plane in the light list";
                return RT_ERR_BAD_ARGUMENT;
            }
            l.kind = LIGHT_RECT;
            l.axis = n.axis;
            for (int k = 0; k < 5; ++k) l.d[k] = n.v[k];
        } else if (n.kind == RT_NODE_SPHERE) {
            l.kind = LIGHT_SPHERE;
            for (int k = 0; k < 4; ++k) l.d[k] = n.v[k];
        } else {
            l.kind = LIGHT_DEFAULT;
        }
        out.lights.push_back(l);
    }
    // The traversal keeps its pending nodes on a fixed stack (tables.h: kStackSize entries; a visit pushes at most
    // one child, so a tree of depth d needs fewer than d) and would drop a subtree - geometry - rather than
    // overflow; a tree that could need more is refused here.  (Median splits below depth 32 keep every tree of up to
    // 2^28 primitives far from that: the 394k-triangle mesh is 22 levels deep.)
    if (out.max_bvh_depth >= (uint32_t)kStackSize) {
        err = "a BVH of this scene is " + std::to_string(out.max_bvh_depth) + " levels deep, the traversal stack of the kernels holds " +
              std::to_string(kStackSize) + " entries";
        return RT_ERR_UNSUPPORTED;
    }
    return RT_OK;
}

}  // namespace rtb200dev
