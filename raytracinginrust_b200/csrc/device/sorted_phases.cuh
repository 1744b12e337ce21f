// sorted_phases.cuh — the per-lane phases of render_sorted_kernel (sorted.inl), apart from the kernel so that the
// CPU test tier can run them lane by lane over a simulated block (tests/native) and compare the image with the plain
// sample loop: what has to be right here is the bookkeeping - every field of a path, its work item and the item's
// sum must survive the trip through shared memory.  Included inside the variant namespace, after trace.cuh.
//
// One segment of one lane:   A  next sample / work item (render_kernel's refill) + search  -> hit class
//                            -- block-wide counting sort by class: the lane's path gets position dst --
//                            B  file the path at dst        -- barrier --
//                            C  pick up the path filed at the lane's own index, resolve the hit, shade
constexpr int kSortedClasses = WF_N_CLASSES + 1;  // + lanes without a path
constexpr uint32_t kSortedIdle = WF_N_CLASSES;
constexpr int kSortedDoubles = 14, kSortedWords = 10;

template <int BLOCK>
struct SortedShared {
    double d[kSortedDoubles][BLOCK];  // ray o, d, time | beta | closest t | the item's sum
    uint32_t w[kSortedWords][BLOCK];  // Philox keys, depth, segments, winner, item, flags
};

struct SortedLane {
    PathState ps;
    Best win;
    double closest;
    V3 sum;                                        // of the work item the path belongs to, samples in sample order
    uint32_t out_pixel, s, s_end, chunk;           // the item: row * W + i (rows top-down), next sample, end, chunk
    bool alive, have_item, done;                   // done: the work counter ran out when this state asked
    unsigned long long n_paths, n_rays, n_bad;     // stay with the lane (only their totals matter)
};

RT_DEV void sorted_lane_init(SortedLane &L, uint32_t seed) {
    L.ps.ray.o = L.ps.ray.d = L.ps.beta = L.ps.radiance = mk(0.0, 0.0, 0.0);
    L.ps.ray.time = 0.0;
    L.ps.rng = Rng{seed, 0u, 0u, 0u};
    L.ps.depth_left = 0u;
    L.ps.segments = 0u;
    L.win = Best{RT_INF, kNoPrim, 0u, 0};
    L.closest = 0.0;
    L.sum = mk(0.0, 0.0, 0.0);
    L.out_pixel = L.s = L.s_end = L.chunk = 0u;
    L.alive = L.have_item = L.done = false;
    L.n_paths = L.n_rays = L.n_bad = 0ull;
}

// A: returns the class the lane's path is sorted by
template <bool MEDIA>
RT_DEV uint32_t sorted_generate_search(const DScene &sc, const RtCamera &cam, const RenderParams &P, double *planes,
                                       unsigned long long *counters, SortedLane &L) {
    if (!L.alive && !L.done) {
        if (!L.have_item || L.s == L.s_end) {
            if (L.have_item) {
                double *dst = planes + 3ull * ((uint64_t)L.chunk * P.width * P.height + L.out_pixel);
                dst[0] = L.sum.x;
                dst[1] = L.sum.y;
                dst[2] = L.sum.z;
                L.have_item = false;
            }
            for (;;) {  // next (chunk, pixel) item; skip the padding of partial tiles
                unsigned long long item = atomicAdd(&counters[kCounterWork], 1ull);
                if (item >= P.n_items) break;
                uint32_t i, row;
                L.chunk = (uint32_t)(item / P.items_per_chunk);
                uint64_t lin = item - (uint64_t)L.chunk * P.items_per_chunk;
                if (!item_pixel(P.tiles_x, P.width, P.height, lin, i, row)) continue;
                L.s = P.sample_begin + L.chunk * P.chunk_size;
                L.s_end = min(L.s + P.chunk_size, P.sample_end);
                L.out_pixel = row * P.width + i;
                L.sum = mk(0.0, 0.0, 0.0);
                L.have_item = true;
                break;
            }
            L.done = !L.have_item;
        }
        if (L.have_item) {
            const uint32_t row = L.out_pixel / P.width, i = L.out_pixel - row * P.width;
            path_begin(L.ps, cam, P.width, P.height, i, P.height - 1u - row, L.s, P.seed, P.max_depth);  // row 0 is j = H-1 (main.rs:772)
            ++L.s;
            ++L.n_paths;
            L.alive = true;
        }
    }
    L.win = Best{RT_INF, kNoPrim, 0u, 0};
    L.closest = 0.0;
    if (!L.alive) return kSortedIdle;
    L.ps.segments += 1;
    world_search<MEDIA>(sc, L.ps.ray, L.ps.rng, L.win, L.closest);  // world.hit without the hit record (main.rs:48)
    return hit_class(sc, L.win.prim);
}

// B
template <int BLOCK>
RT_DEV void sorted_file(SortedShared<BLOCK> &sh, unsigned dst, const SortedLane &L) {
    const PathState &ps = L.ps;
    sh.d[0][dst] = ps.ray.o.x; sh.d[1][dst] = ps.ray.o.y; sh.d[2][dst] = ps.ray.o.z;
    sh.d[3][dst] = ps.ray.d.x; sh.d[4][dst] = ps.ray.d.y; sh.d[5][dst] = ps.ray.d.z;
    sh.d[6][dst] = ps.ray.time;
    sh.d[7][dst] = ps.beta.x; sh.d[8][dst] = ps.beta.y; sh.d[9][dst] = ps.beta.z;
    sh.d[10][dst] = L.closest;
    sh.d[11][dst] = L.sum.x; sh.d[12][dst] = L.sum.y; sh.d[13][dst] = L.sum.z;
    sh.w[0][dst] = ps.rng.pixel; sh.w[1][dst] = ps.rng.sample;
    sh.w[2][dst] = ps.depth_left; sh.w[3][dst] = ps.segments;
    sh.w[4][dst] = L.win.prim; sh.w[5][dst] = (uint32_t)L.win.face;
    sh.w[6][dst] = L.out_pixel; sh.w[7][dst] = L.s; sh.w[8][dst] = L.s_end;
    sh.w[9][dst] = L.chunk | (L.alive ? 0x80000000u : 0u) | (L.have_item ? 0x40000000u : 0u) | (L.done ? 0x20000000u : 0u);
}

// C, first half
template <int BLOCK>
RT_DEV void sorted_pickup(const SortedShared<BLOCK> &sh, unsigned k, const RenderParams &P, SortedLane &L) {
    PathState &ps = L.ps;
    ps.ray.o = mk(sh.d[0][k], sh.d[1][k], sh.d[2][k]);
    ps.ray.d = mk(sh.d[3][k], sh.d[4][k], sh.d[5][k]);
    ps.ray.time = sh.d[6][k];
    ps.beta = mk(sh.d[7][k], sh.d[8][k], sh.d[9][k]);
    L.closest = sh.d[10][k];
    L.sum = mk(sh.d[11][k], sh.d[12][k], sh.d[13][k]);
    ps.rng.pixel = sh.w[0][k]; ps.rng.sample = sh.w[1][k];
    ps.depth_left = sh.w[2][k]; ps.segments = sh.w[3][k];
    ps.rng.bounce = P.max_depth - ps.depth_left;  // path_begin: 0, path_shade: += 1 as depth_left -= 1
    L.win.prim = sh.w[4][k]; L.win.face = (int)sh.w[5][k];
    L.win.t = L.closest;
    L.win.rank = 0u;
    L.out_pixel = sh.w[6][k]; L.s = sh.w[7][k]; L.s_end = sh.w[8][k];
    const uint32_t f = sh.w[9][k];
    L.chunk = f & 0x1FFFFFFFu;
    L.alive = (f & 0x80000000u) != 0u;
    L.have_item = (f & 0x40000000u) != 0u;
    L.done = (f & 0x20000000u) != 0u;
}

// C, second half: everything ray_color does after world.hit returned (main.rs:62-119)
RT_DEV void sorted_shade(const DScene &sc, const RenderParams &P, SortedLane &L) {
    if (!L.alive) return;
    HitRec rec;
    const bool hit = L.win.prim != kNoPrim;
    if (hit) {
        if (L.win.prim & kMediumFlag) resolve_medium(sc, L.ps.ray, L.win, L.closest, rec);
        else resolve_hit<false>(sc, L.ps.ray, L.win, L.closest, rec);
    }
    L.alive = path_shade(sc, L.ps, hit, rec, P.integrator, P.flags);
    if (!L.alive) {
        L.n_rays += L.ps.segments;
        if (!(isfinite(L.ps.radiance.x) && isfinite(L.ps.radiance.y) && isfinite(L.ps.radiance.z))) ++L.n_bad;  // §Q10: counted only
        L.sum = L.sum + L.ps.radiance;  // vec.rs:253-260 Sum, in sample order
    }
}
