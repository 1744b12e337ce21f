// wavefront.inl — the sample loop of src/main.rs:772-834 as a wavefront pipeline (sm_100a).
//
// The recursion of ray_color (src/main.rs:41-120) is cut at its one recursive call: a path is a
// slot of an HBM-resident pool (wavefront.h), and a round moves every live path forward by one
// segment with three stage kernels:
//
//   shade     (main.rs:62-119)  reads the slot's ray + what extend found, resolves the hit record,
//                               evaluates the material, writes the next ray or ends the path
//   generate  (main.rs:811-820) gives every slot whose path ended its next sample (or its next
//                               (chunk, pixel) work item) and writes the camera ray
//   extend    (main.rs:48)      world.hit for every live slot: persistent warps that pull slots
//                               from a cursor and REPLACE a finished ray by the next slot while the
//                               other lanes are still traversing (idle lanes are compacted onto the
//                               reserved range with __ballot_sync/__popc)
//   control                     one thread: round bookkeeping, loop condition
//
// All arithmetic is the device code of trace.cuh, shared with the megakernel; a slot runs the
// samples of its item in sample order into an f64 sum of its own, so the image is bit-identical
// to the megakernel's and to itself run after run.
// (included by pipelines.cu inside the variant namespace)

#ifndef RT_WF_INNER_THRESHOLD
#define RT_WF_INNER_THRESHOLD 12
#endif
#ifndef RT_WF_EXTEND_MIN_BLOCKS
#define RT_WF_EXTEND_MIN_BLOCKS 4
#endif
#ifndef RT_WF_SHADE_MIN_BLOCKS
#define RT_WF_SHADE_MIN_BLOCKS 6
#endif

constexpr unsigned kFull = 0xFFFFFFFFu;

// Shade pass 1 works on tiles of RT_WF_SHADE_TILE x 128 slots that it sorts by the class of what the ray hit, so
// that a warp shades one kind of material (32 neighbouring slots hold four or five kinds).  Extend knows the
// primitive and leaves the class in bits 8-10 of the slot's state word; the sort is a counting sort in shared
// memory (no global atomics, no extra pass over the pool).  A slot's result does not depend on the order.
#ifndef RT_WF_SHADE_TILE
#define RT_WF_SHADE_TILE 8  // measured on the final scene: 2 -> 149.0 ms, 4 -> 143.2, 8 -> 140.5 (unsorted pass: 162.2)
#endif
constexpr int kShadeTile = RT_WF_SHADE_TILE * kWfBlock;
static_assert(kShadeTile <= 4096, "local slot index and class share 16 bits");
// (the classes and hit_class live in trace.cuh: the sorted megakernel shares them)
// Bits 11-31 of the state word: where the slot's next ray starts, as the index of the primitive it leaves
// (primitives are stored in BVH leaf order, so close indices are close in space); camera rays get the last key.
// The simple extend stage sorts its tiles by it (RT_WF_EXT_TILE).  An octant-major key (direction octant, then 32
// origin bins) was measured too: 134.0 ms against 131.9 ms for the origin alone on the final scene.
constexpr uint32_t kOriginKeyMax = 0x1FFFFFu;
__device__ __forceinline__ uint32_t origin_key(uint32_t prim) {
    if (prim == kNoPrim) return kOriginKeyMax;                  // (the path ends in shade)
    if (prim & kMediumFlag) return kOriginKeyMax - 1u;          // scattered inside a medium
    return prim < kOriginKeyMax - 2u ? prim : kOriginKeyMax - 2u;
}
__device__ __forceinline__ uint32_t live_state(const DScene &sc, uint32_t prim) {
    return WF_LIVE | (hit_class(sc, prim) << 8) | (origin_key(prim) << 11);
}


// 128-bit views of a slot record
__device__ __forceinline__ const double2 *slot_d2(const WfPool &pool, uint32_t slot) { return reinterpret_cast<const double2 *>(pool.slots + slot); }
__device__ __forceinline__ double2 *slot_d2w(const WfPool &pool, uint32_t slot) { return reinterpret_cast<double2 *>(pool.slots + slot); }
__device__ __forceinline__ uint4 ld_u4(const double2 *p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ void st_u4(double2 *p, uint4 v) { *reinterpret_cast<uint4 *>(p) = v; }
__device__ __forceinline__ uint4 pack_time_state(double time, uint32_t state, uint32_t depth_left) {
    return make_uint4((uint32_t)__double2loint(time), (uint32_t)__double2hiint(time), state, depth_left);
}
__device__ __forceinline__ uint4 pack_bz_keys(double bz, uint32_t rng_pixel, uint32_t sample) {
    return make_uint4((uint32_t)__double2loint(bz), (uint32_t)__double2hiint(bz), rng_pixel, sample);
}
__device__ __forceinline__ double unpack_lo_double(uint4 v) { return __hiloint2double((int)v.y, (int)v.x); }

// Block-wide sum of a per-thread count, then ONE fire-and-forget atomic per block.
__device__ __forceinline__ void block_count_add(unsigned *dst, unsigned mine, unsigned long long *dst64 = nullptr) {
    __shared__ unsigned s_total;
    if (threadIdx.x == 0) s_total = 0u;
    __syncthreads();
    const unsigned w = __reduce_add_sync(kFull, mine);
    if ((threadIdx.x & 31u) == 0u && w) atomicAdd(&s_total, w);
    __syncthreads();
    if (threadIdx.x == 0 && s_total) {
        atomicAdd(dst, s_total);
        if (dst64) atomicAdd(dst64, (unsigned long long)s_total);
    }
}

// ---------------------------------------------------------------------------
// init: every slot is empty and asks generate for an item
// ---------------------------------------------------------------------------
__global__ void wf_init_kernel(const __grid_constant__ WfPool pool) {
    unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned stride = gridDim.x * blockDim.x;
    if (k == 0) {
        WfCtl c{};
        c.status_live = 1;
        *pool.ctl = c;
    }
    for (; k < pool.n_slots; k += stride) {
        double2 *u = slot_d2w(pool, k);
        st_u4(u + 3, pack_time_state(0.0, WF_REGEN, 0u));
        pool.state[k] = WF_REGEN;
        st_u4(u + 5, pack_bz_keys(0.0, 0u, 0u));
        st_u4(u + 7, make_uint4(0u, 0u, kWfNoItem, 0u));  // no item: sample + 1 >= s_end
        pool.sum[k] = make_double4(0.0, 0.0, 0.0, 0.0);
    }
}

// ---------------------------------------------------------------------------
// shade: everything ray_color does after world.hit returned (main.rs:62-119)
// ---------------------------------------------------------------------------
// The stage's cost is the latency of the slot record, which only parallelism hides, and - the code being
// 108 KB of SASS that a warp walks once per slot - instruction issue and fetch: every distinct material in
// a warp is another stretch of code issued for a few lanes.  So pass 1 SORTS before it shades: a block takes
// a tile of kShadeTile slots, sorts them by hit class (miss, medium, one class per material kind) with a
// counting sort in shared memory, and shades them in that order - a warp then holds one kind of material
// (measured on the Next Week final scene: the pass went from 11.8 to ~20 active lanes, the render from
// 162.2 to 140.5 ms; images unchanged, a slot's result does not depend on who shades it when).
//
// Two passes.  A few materials cost an order of magnitude more than the rest (the marble sphere:
// 7 octaves of f64 Perlin noise; image textures: atan2 + acos for the uv); they are a class of
// their own that pass 1 does not shade but COMPACTS into a queue (__ballot_sync/__popc, one
// atomic per warp that has any); pass 2 walks that queue, so those paths fill whole warps from all
// over the pool, not just from one tile.
__device__ __forceinline__ void shade_slot(const DScene &sc, const RenderParams &P, const WfPool &pool, uint32_t slot,
                                             const uint4 u3, const uint4 u6, bool &alive, bool &bad) {
    const double2 *u = slot_d2(pool, slot);
    const double2 u0 = u[0], u1 = u[1], u2 = u[2], u4 = u[4];
    const uint4 u5 = ld_u4(u + 5);
    PathState ps;
    ps.ray.o = mk(u0.x, u0.y, u1.x);
    ps.ray.d = mk(u1.y, u2.x, u2.y);
    ps.ray.time = unpack_lo_double(u3);
    ps.beta = mk(u4.x, u4.y, unpack_lo_double(u5));
    ps.radiance = mk(0.0, 0.0, 0.0);  // non-zero only at the segment that ends the path
    ps.rng = Rng{P.seed, u5.z, u5.w, P.max_depth - u3.w};
    ps.depth_left = u3.w;
    ps.segments = 0;
    const uint32_t prim = u6.x;
    const double t = __hiloint2double((int)u6.w, (int)u6.z);
    HitRec rec;
    const bool hit = prim != kNoPrim;
    if (hit) {
        Best win{t, prim, 0u, (int)u6.y};
        if (prim & kMediumFlag) resolve_medium(sc, ps.ray, win, t, rec);
        else resolve_hit<false>(sc, ps.ray, win, t, rec);
    }
    alive = path_shade(sc, ps, hit, rec, P.integrator, P.flags);
    double2 *w = slot_d2w(pool, slot);
    if (alive) {
        w[0] = make_double2(ps.ray.o.x, ps.ray.o.y);
        w[1] = make_double2(ps.ray.o.z, ps.ray.d.x);
        w[2] = make_double2(ps.ray.d.y, ps.ray.d.z);
        st_u4(w + 3, make_uint4(u3.x, u3.y, WF_LIVE, ps.depth_left));  // time is inherited (main.rs:95, mat.rs:219,270,368)
        w[4] = make_double2(ps.beta.x, ps.beta.y);
        st_u4(w + 5, pack_bz_keys(ps.beta.z, u5.z, u5.w));
    } else {
        pool.state[slot] = WF_REGEN;
        // vec.rs:253-260 Sum, in sample order (the slot runs its samples one after the other)
        const V3 L = ps.radiance;
        bad = !(isfinite(L.x) && isfinite(L.y) && isfinite(L.z));  // §Q10: counted, not guarded
        if (L.x != 0.0 || L.y != 0.0 || L.z != 0.0) {              // x + 0 == x
            double4 s = pool.sum[slot];
            s.x += L.x;
            s.y += L.y;
            s.z += L.z;
            pool.sum[slot] = s;
        }
    }
}

__global__ void __launch_bounds__(kWfBlock, RT_WF_SHADE_MIN_BLOCKS)
wf_shade_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RenderParams P,
                const __grid_constant__ WfPool pool, unsigned long long *__restrict__ counters) {
    __shared__ unsigned s_bin[WF_N_CLASSES];          // per class: count, then first position in s_order
    __shared__ unsigned s_live;                       // live slots of the tile
    __shared__ unsigned short s_order[kShadeTile];    // class << 12 | index in the tile, sorted by class
    const unsigned cur = pool.ctl->round & 1u;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t tile0 = blockIdx.x * (uint32_t)kShadeTile;
    if (threadIdx.x < WF_N_CLASSES) s_bin[threadIdx.x] = 0u;
    __syncthreads();
    // the state array is read first: 4 bytes per slot, so a round over a pool that is nearly
    // empty (the tail of a render) moves 16 MB, not the 512 MB of the slot records
    uint32_t cls[RT_WF_SHADE_TILE], pos[RT_WF_SHADE_TILE];
#pragma unroll
    for (int i = 0; i < RT_WF_SHADE_TILE; ++i) {
        const uint32_t slot = tile0 + (uint32_t)i * kWfBlock + threadIdx.x;
        uint32_t c = WF_CLS_NONE;
        if (slot < pool.n_slots) {
            const uint32_t st = pool.state[slot];
            if ((st & 0xFFu) == WF_LIVE) c = (st >> 8) & 7u;
        }
        cls[i] = c;
        // counting sort, step 1: position inside the class (one shared-memory atomic per class and warp)
        const unsigned peers = __match_any_sync(kFull, c);
        const unsigned leader = (unsigned)__ffs((int)peers) - 1u;
        unsigned base = 0u;
        if (c != WF_CLS_NONE && lane == leader) base = atomicAdd(&s_bin[c], (unsigned)__popc(peers));
        base = __shfl_sync(kFull, base, (int)leader);
        pos[i] = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // step 2: where each class starts
        unsigned acc = 0u;
        for (uint32_t c = 0; c < WF_N_CLASSES; ++c) {
            const unsigned n = s_bin[c];
            s_bin[c] = acc;
            acc += n;
        }
        s_live = acc;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RT_WF_SHADE_TILE; ++i)
        if (cls[i] != WF_CLS_NONE) s_order[s_bin[cls[i]] + pos[i]] = (unsigned short)((cls[i] << 12) | ((uint32_t)i * kWfBlock + threadIdx.x));
    __syncthreads();
    const unsigned n_live = s_live;
    unsigned alive_n = 0u, bad_n = 0u;
    for (unsigned k0 = 0; k0 < n_live; k0 += kWfBlock) {  // the same trip count for the whole block
        const unsigned k = k0 + threadIdx.x;
        bool alive = false, bad = false, defer = false;
        uint32_t slot = 0;
        if (k < n_live) {
            const uint32_t e = s_order[k];
            slot = tile0 + (e & 0xFFFu);
            defer = feat(F_TEX) && (e >> 12) == WF_CLS_COSTLY;
            if (!defer) {
                const double2 *u = slot_d2(pool, slot);
                shade_slot(sc, P, pool, slot, ld_u4(u + 3), ld_u4(u + 6), alive, bad);
            }
        }
        if (feat(F_TEX)) {  // compact the slots that hit a costly material for pass 2
            const unsigned m = __ballot_sync(kFull, defer);
            if (m != 0u) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&pool.ctl->defer_n, (unsigned)__popc(m));
                base = __shfl_sync(kFull, base, 0);
                if (defer) pool.defer_q[base + __popc(m & ((1u << lane) - 1u))] = slot;
            }
        }
        alive_n += alive ? 1u : 0u;
        bad_n += bad ? 1u : 0u;  // §Q10: counted, not guarded
    }
    block_count_add(&pool.ctl->live[cur], alive_n);
    if (__syncthreads_or(bad_n != 0u)) {
        const unsigned nb = __reduce_add_sync(kFull, bad_n);
        if (lane == 0u && nb) atomicAdd(&counters[kCounterNonFinite], (unsigned long long)nb);
    }
}

// pass 2: the compacted queue of slots that hit a costly material
__global__ void __launch_bounds__(kWfBlock, RT_WF_SHADE_MIN_BLOCKS)
wf_shade_deferred_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RenderParams P,
                         const __grid_constant__ WfPool pool, unsigned long long *__restrict__ counters) {
    const unsigned cur = pool.ctl->round & 1u;
    const unsigned n = pool.ctl->defer_n;
    const unsigned stride = gridDim.x * blockDim.x;
    bool bad_any = false;
    unsigned alive_n = 0;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        const uint32_t slot = pool.defer_q[k];
        const double2 *u = slot_d2(pool, slot);
        bool alive = false, bad = false;
        shade_slot(sc, P, pool, slot, ld_u4(u + 3), ld_u4(u + 6), alive, bad);
        alive_n += alive ? 1u : 0u;
        bad_any |= bad;
        if (bad) atomicAdd(&counters[kCounterNonFinite], 1ull);
    }
    block_count_add(&pool.ctl->live[cur], alive_n);
}

// ---------------------------------------------------------------------------
// generate: the sample closure of main.rs:811-820 for every slot whose path ended
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kWfBlock, 4)
wf_generate_kernel(const __grid_constant__ RtCamera cam, const __grid_constant__ RenderParams P,
                   const __grid_constant__ WfPool pool, double *__restrict__ planes,
                   unsigned long long *__restrict__ counters) {
    __shared__ unsigned s_need;
    __shared__ unsigned long long s_first;
    const unsigned cur = pool.ctl->round & 1u;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t n_pixels = (uint64_t)P.width * P.height;
    uint32_t rng_pixel = 0, out_pixel = 0, sample = 0, s_end = 0, chunk = kWfNoItem;
    bool regen = false, have = false, need_item = false;
    if (threadIdx.x == 0) s_need = 0u;
    if (slot < pool.n_slots) {
        regen = pool.state[slot] == WF_REGEN;
        if (regen) {
            const double2 *u = slot_d2(pool, slot);
            const uint4 u5 = ld_u4(u + 5), u7 = ld_u4(u + 7);
            rng_pixel = u5.z;
            sample = u5.w + 1u;
            out_pixel = u7.x;
            s_end = u7.y;
            chunk = u7.z;
            have = chunk != kWfNoItem && sample < s_end;
            need_item = !have;
            if (need_item && chunk != kWfNoItem) {  // the item is complete: its sum goes to its plane slot
                const double4 s = pool.sum[slot];
                double *dst = planes + 3 * ((uint64_t)chunk * n_pixels + out_pixel);
                dst[0] = s.x;
                dst[1] = s.y;
                dst[2] = s.z;
                pool.sum[slot] = make_double4(0.0, 0.0, 0.0, 0.0);
                chunk = kWfNoItem;
            }
        }
    }
    // Next (chunk, pixel) items.  First try: one atomic for the whole block - threads that ask
    // together get consecutive items, i.e. neighbouring pixels of 8x4 tiles.  Items that fall into
    // the padding of partial tiles are skipped and asked for again, warp by warp.
    __syncthreads();
    unsigned my_rank = 0;
    {
        const unsigned m = __ballot_sync(kFull, need_item);
        unsigned wbase = 0;
        if (lane == 0 && m) wbase = atomicAdd(&s_need, (unsigned)__popc(m));
        wbase = __shfl_sync(kFull, wbase, 0);
        my_rank = wbase + __popc(m & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_need) s_first = atomicAdd(&counters[kCounterWork], (unsigned long long)s_need);
    __syncthreads();
    unsigned long long item = s_first + my_rank;
    for (bool first_try = true;; first_try = false) {
        if (!first_try) {
            const unsigned m = __ballot_sync(kFull, need_item);
            if (m == 0u) break;
            unsigned long long f = 0;
            if (lane == 0) f = atomicAdd(&counters[kCounterWork], (unsigned long long)__popc(m));
            f = __shfl_sync(kFull, f, 0);
            item = f + __popc(m & ((1u << lane) - 1u));
        }
        if (need_item) {
            if (item >= P.n_items) {
                need_item = false;  // no work left: the slot retires
            } else {
                uint64_t lin;
                const uint32_t c = item_split(P, item, lin);
                uint32_t i, row;
                if (item_pixel(P, lin, i, row)) {
                    chunk = c;
                    sample = P.sample_begin + c * P.chunk_size;
                    s_end = min(sample + P.chunk_size, P.sample_end);
                    out_pixel = row * P.width + i;
                    rng_pixel = (P.height - 1u - row) * P.width + i;  // row 0 of the image is j = H-1 (main.rs:772)
                    have = true;
                    need_item = false;
                }
            }
        }
    }
    if (regen) {
        double2 *w = slot_d2w(pool, slot);
        if (have) {
            const uint32_t row = out_pixel / P.width, i = out_pixel - row * P.width;
            Rng rng{P.seed, rng_pixel, sample, 0};
            const Ray r = camera_ray(cam, P.width, P.height, i, P.height - 1u - row, rng);
            w[0] = make_double2(r.o.x, r.o.y);
            w[1] = make_double2(r.o.z, r.d.x);
            w[2] = make_double2(r.d.y, r.d.z);
            st_u4(w + 3, pack_time_state(r.time, WF_LIVE, P.max_depth));
            pool.state[slot] = WF_LIVE | (kOriginKeyMax << 11);  // a camera ray
            w[4] = make_double2(1.0, 1.0);
            st_u4(w + 5, pack_bz_keys(1.0, rng_pixel, sample));
            st_u4(w + 7, make_uint4(out_pixel, s_end, chunk, 0u));
        } else {
            pool.state[slot] = WF_EMPTY;
            st_u4(w + 7, make_uint4(0u, 0u, kWfNoItem, 0u));
        }
    }
    block_count_add(&pool.ctl->live[cur], have ? 1u : 0u, &counters[kCounterPaths]);
}

// ---------------------------------------------------------------------------
// extend: world.hit(ray, 0.00001, inf) (main.rs:48) for every live slot
// ---------------------------------------------------------------------------
// The control flow of world_hit / trace_groups / trace_group (trace.cuh) unrolled into a per-lane
// state machine, so that a lane whose ray is finished takes the next slot instead of waiting for
// the slowest ray of its warp:
//   A  refill      idle lanes take the next slots of the warp's reserved range (one global atomic
//                  per kWfReserve slots; idle lanes are ranked with __ballot_sync/__popc)
//   B  transition  lanes without a BVH node run world_hit's bookkeeping until they need one:
//                  next group of the query (cull, transform; a group that is a single leaf is
//                  scanned right here), query complete (exact_t; medium.rs:32-58), next query,
//                  result
//   C  traverse    while-while over BVH nodes and leaves, with votes: stop descending when few
//                  lanes still hold an inner node, leave when a quarter of the rays wait for B
#ifndef RT_WF_RESERVE
#define RT_WF_RESERVE 64
#endif
constexpr unsigned kWfReserve = RT_WF_RESERVE;  // slots a warp reserves per global atomic

template <bool MEDIA>
__global__ void __launch_bounds__(kWfBlock, RT_WF_EXTEND_MIN_BLOCKS)
wf_extend_kernel(const __grid_constant__ DScene sc, const __grid_constant__ WfPool pool, uint32_t seed, uint32_t max_depth,
                 uint32_t leave_threshold) {
    const int kDone = (int)0x80000000;
    const unsigned n = pool.n_slots;
    unsigned *cursor = &pool.ctl->ext_cursor;
    const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
    const uint32_t nq = 1u + (MEDIA ? 2u * sc.n_media : 0u);

    unsigned wbase = 0, wend = 0;  // the warp's reserved range of slots
    bool exhausted = false;
    bool has_ray = false;
    uint32_t slot = 0;
    Ray ray;
    V3 inv;
    Rng rng{seed, 0u, 0u, 0u};
    uint32_t q = 0, gi = 0, g_end = 0;
    double t_min = kTMin, closest = RT_INF, t1 = 0.0;
    bool have_t1 = false;
    Best b{RT_INF, kNoPrim, 0u, 0}, win{RT_INF, kNoPrim, 0u, 0};
    SRay r;
    FRay f;
    float t_min_f = 0.f, t_max_f = 0.f;
    int node = kDone, sp = 0;
    int stack[kStackSize];
    ray.o = ray.d = inv = r.o = r.d = r.inv = mk(0.0, 0.0, 0.0);
    ray.time = r.time = 0.0;
    f = FRay{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};

    for (;;) {
        // ---- A: refill ----------------------------------------------------------------------
        const unsigned want = __ballot_sync(kFull, !has_ray);
        // leave_threshold == 0: batch mode - new rays only when the whole warp is idle, so that all
        // its rays walk the query / group sequence of world_hit in step
        if (want != 0u && !exhausted && (leave_threshold != 0u || want == kFull)) {
            if (wbase == wend) {
                unsigned got = 0;
                if (lane == 0) got = atomicAdd(cursor, kWfReserve);
                got = __shfl_sync(kFull, got, 0);
                wbase = min(got, n);
                wend = min(got + kWfReserve, n);
                exhausted = wbase == wend;
            }
            const unsigned avail = wend - wbase;
            const unsigned rank = __popc(want & lt);
            if (!has_ray && rank < avail) {
                const uint32_t cand = wbase + rank;
                if ((pool.state[cand] & 0xFFu) == WF_LIVE) {  // not LIVE only in the tail of a render, when items have run out
                    slot = cand;
                    const double2 *u = slot_d2(pool, cand);
                    const uint4 u3 = ld_u4(u + 3);
                    const double2 r0 = u[0], r1 = u[1], r2 = u[2];
                    ray.o = mk(r0.x, r0.y, r1.x);
                    ray.d = mk(r1.y, r2.x, r2.y);
                    ray.time = unpack_lo_double(u3);
                    if (MEDIA) {
                        const uint4 u5 = ld_u4(u + 5);
                        rng.pixel = u5.z;
                        rng.sample = u5.w;
                        rng.bounce = max_depth - u3.w;
                    }
                    inv = mk(rcp_fast(ray.d.x), rcp_fast(ray.d.y), rcp_fast(ray.d.z));
                    q = 0;
                    gi = 0;
                    g_end = sc.n_world_groups;
                    t_min = kTMin;
                    b = Best{RT_INF, kNoPrim, 0u, 0};
                    win = b;
                    closest = RT_INF;
                    have_t1 = false;
                    node = kDone;
                    sp = 0;
                    has_ray = true;
                }
            }
            wbase += min(avail, (unsigned)__popc(want));
        }
        const unsigned n_has = __popc(__ballot_sync(kFull, has_ray));
        if (n_has == 0u) {
            if (exhausted) break;
            continue;
        }
        // leave the traversal loop once fewer than leave_threshold/32 of the rays in flight still traverse
        // (a launch parameter: 1 = never, the warp stays in lockstep; see wf_launch_round)
        const unsigned leave_below = (n_has * leave_threshold) >> 5;
        const unsigned inner_below = (n_has * (unsigned)RT_WF_INNER_THRESHOLD) >> 5;

        // ---- B: transition ------------------------------------------------------------------
        if (has_ray && node == kDone) {
            for (;;) {
                if (gi < g_end) {  // trace_groups: next group of this query
                    const DGroup &g = sc.groups[gi++];
                    double e;
                    if ((g.flags & GROUP_CULL) && !slab(ray.o, inv, g.bmin, g.bmax, t_min, b.t, e)) continue;
                    r.o = ray.o;
                    r.d = ray.d;
                    r.time = ray.time;
                    r.inv = inv;
                    if (g.flags & GROUP_XFORM) {
                        const double *m = g.m;
                        const V3 o = ray.o, d = ray.d;
                        r.o = mk(fma(m[0], o.x, fma(m[1], o.y, fma(m[2], o.z, g.t[0]))), fma(m[3], o.x, fma(m[4], o.y, fma(m[5], o.z, g.t[1]))),
                                 fma(m[6], o.x, fma(m[7], o.y, fma(m[8], o.z, g.t[2]))));
                        if (g.flags & GROUP_ROTATED) {
                            r.d = mk(fma(m[0], d.x, fma(m[1], d.y, m[2] * d.z)), fma(m[3], d.x, fma(m[4], d.y, m[5] * d.z)),
                                     fma(m[6], d.x, fma(m[7], d.y, m[8] * d.z)));
                            r.inv = mk(rcp_fast(r.d.x), rcp_fast(r.d.y), rcp_fast(r.d.z));
                        }
                    }
                    const int root = g.bvh_root;
                    if (root < 0) {  // the whole group is one leaf: scan it here, no trip through C
                        const uint32_t code = ~(uint32_t)root;
                        const uint32_t first = code >> 3, count = (code & 7u) + 1u;
                        for (uint32_t i = 0; i < count; ++i) s_prim(sc, first + i, r, t_min, b);
                        continue;
                    }
                    node = root;
                    sp = 0;
                    f = make_fray(r);
                    t_min_f = __double2float_rd(t_min);
                    t_max_f = __double2float_ru(b.t);
                    break;
                }
                // world_hit: the query is complete
                const bool found = b.prim != kNoPrim;
                const bool second = MEDIA && q > 0u && ((q - 1u) & 1u) != 0u;
                if (MEDIA && q > 0u && !second) have_t1 = found;
                if (found) {
                    const double te = exact_t(sc, ray, b, t_min);
                    if (q == 0u) {
                        win = b;
                        closest = te;
                    } else if (!second) {
                        t1 = te;
                    } else {  // medium.rs:32-58
                        const uint32_t mi = (q - 1u) >> 1;
                        const DMedium &m = sc.media[mi];
                        double h1 = t1, h2 = te;
                        if (h1 < kTMin) h1 = kTMin;
                        if (h2 > closest) h2 = closest;
                        if (h1 < h2) {
                            V3 o = ray.o, d = ray.d;
                            DChain c = sc.chains[m.chain];
                            chain_ray(sc, c.first_op, c.n_ops, o, d);
                            const double len = length(d);
                            const double distance_inside_boundary = (h2 - h1) * len;
                            const Draw dr = draw(rng, SLOT_MEDIUM, (uint32_t)m.node);
                            const double hit_distance = -(1.0 / m.density) * log(dr.a);
                            if (hit_distance < distance_inside_boundary) {
                                closest = h1 + hit_distance / len;
                                win.prim = kMediumFlag | mi;
                                win.rank = m.rank;
                                win.face = 0;
                            }
                        }
                    }
                }
                ++q;
                if (MEDIA && q < nq && ((q - 1u) & 1u) != 0u && !have_t1) ++q;  // no first boundary hit: skip the second query
                if (q >= nq) {
                    st_u4(slot_d2w(pool, slot) + 6, make_uint4(win.prim, (uint32_t)win.face, (uint32_t)__double2loint(closest), (uint32_t)__double2hiint(closest)));
                    pool.state[slot] = live_state(sc, win.prim);  // what shade (and extend) sort their tiles by
                    has_ray = false;
                    break;
                }
                if (MEDIA) {  // open a boundary query of medium (q-1)/2 (medium.rs:29-30)
                    const uint32_t mi = (q - 1u) >> 1;
                    const bool sec = ((q - 1u) & 1u) != 0u;
                    const DMedium &m = sc.media[mi];
                    gi = m.first_group;
                    g_end = gi + m.n_groups;
                    t_min = sec ? t1 + 0.0001 : -DBL_MAX;
                    b = Best{DBL_MAX, kNoPrim, 0u, 0};
                }
            }
        }

        // ---- C: traverse --------------------------------------------------------------------
        while (node != kDone) {
            while (node >= 0) {
                const float4 *np = reinterpret_cast<const float4 *>(sc.nodes + node);
                const float4 q0 = __ldg(np), q1 = __ldg(np + 1), q2 = __ldg(np + 2);
                const int4 ch = __ldg(reinterpret_cast<const int4 *>(np + 3));
                float e0, e1;
                const bool h0 = slab2f(f, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, t_min_f, t_max_f, e0);
                const bool h1 = slab2f(f, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, t_min_f, t_max_f, e1);
                if (h0 && h1) {
                    const bool swap = e1 < e0;
                    const int near_c = swap ? ch.y : ch.x, far_c = swap ? ch.x : ch.y;
                    if (sp < kStackSize) stack[sp++] = far_c;
                    node = near_c;
                } else if (h0) {
                    node = ch.x;
                } else if (h1) {
                    node = ch.y;
                } else {
                    node = sp ? stack[--sp] : kDone;
                }
                // lanes that hold a leaf wait here for the others: stop descending once few are left
                if ((unsigned)__popc(__ballot_sync(__activemask(), node >= 0)) < inner_below) break;
            }
            if (node < 0 && node != kDone) {
                const uint32_t code = ~(uint32_t)node;
                const uint32_t first = code >> 3, count = (code & 7u) + 1u;
                const double before = b.t;
                for (uint32_t i = 0; i < count; ++i) s_prim(sc, first + i, r, t_min, b);
                if (b.t != before) t_max_f = __double2float_ru(b.t);
                node = sp ? stack[--sp] : kDone;
            }
            if ((unsigned)__popc(__ballot_sync(__activemask(), node != kDone)) < leave_below) break;
        }
    }
}

// extend, simple form: one slot per thread, world_hit's search as straight-line code.  Used where
// every ray walks the same sequence of queries and groups (no triangle BVH): the warp stays in
// step through the bookkeeping, and the state machine above has nothing to replace.
#ifndef RT_WF_SIMPLE_MIN_BLOCKS
#define RT_WF_SIMPLE_MIN_BLOCKS 8
#endif
// RT_WF_EXT_TILE = N > 0: a block takes a tile of N x 128 slots and traces them sorted by where their rays start
// (origin_key, 256 bins, counting sort in shared memory), so that a warp's rays begin in the same part of the BVH.
#ifndef RT_WF_EXT_TILE
#define RT_WF_EXT_TILE 4  // measured on the final scene: off 140.6 ms, 2 -> 132.8, 4 -> 132.0, 8 -> 137.2
#endif
constexpr int kExtTile = (RT_WF_EXT_TILE > 0 ? RT_WF_EXT_TILE : 1) * kWfBlock;

template <bool MEDIA>
__device__ __forceinline__ void extend_slot(const DScene &sc, const WfPool &pool, uint32_t slot, uint32_t seed, uint32_t max_depth) {
    const double2 *u = slot_d2(pool, slot);
    const double2 r0 = u[0], r1 = u[1], r2 = u[2];
    const uint4 u3 = ld_u4(u + 3);
    Ray ray;
    ray.o = mk(r0.x, r0.y, r1.x);
    ray.d = mk(r1.y, r2.x, r2.y);
    ray.time = unpack_lo_double(u3);
    Rng rng{seed, 0u, 0u, 0u};
    if (MEDIA) {
        const uint4 u5 = ld_u4(u + 5);
        rng.pixel = u5.z;
        rng.sample = u5.w;
        rng.bounce = max_depth - u3.w;
    }
    Best win;
    double closest;
    world_search<MEDIA>(sc, ray, rng, win, closest);
    st_u4(slot_d2w(pool, slot) + 6, make_uint4(win.prim, (uint32_t)win.face, (uint32_t)__double2loint(closest), (uint32_t)__double2hiint(closest)));
    pool.state[slot] = live_state(sc, win.prim);  // what shade (and extend) sort their tiles by
}

template <bool MEDIA>
__global__ void __launch_bounds__(kWfBlock, RT_WF_SIMPLE_MIN_BLOCKS)
wf_extend_simple_kernel(const __grid_constant__ DScene sc, const __grid_constant__ WfPool pool, uint32_t seed, uint32_t max_depth) {
#if RT_WF_EXT_TILE > 0
    __shared__ unsigned s_bin[256];
    __shared__ unsigned s_warp[kWfBlock / 32];
    __shared__ unsigned s_live;
    __shared__ unsigned short s_order[kExtTile];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile0 = blockIdx.x * (uint32_t)kExtTile;
    // keys span [0, n_prims) plus the three special ones at the top: 256 bins over that range
    const uint32_t top = sc.n_prims + 3u;
    const int bits = 32 - __clz((int)(top | 1u));
    const int shift = bits > 8 ? bits - 8 : 0;
    s_bin[threadIdx.x] = 0u;
    s_bin[threadIdx.x + kWfBlock] = 0u;
    __syncthreads();
    uint32_t key[RT_WF_EXT_TILE], pos[RT_WF_EXT_TILE];
#pragma unroll
    for (int i = 0; i < RT_WF_EXT_TILE; ++i) {
        const uint32_t slot = tile0 + (uint32_t)i * kWfBlock + threadIdx.x;
        key[i] = 0xFFFFFFFFu;
        pos[i] = 0u;
        if (slot < pool.n_slots) {
            const uint32_t st = pool.state[slot];
            if ((st & 0xFFu) == WF_LIVE) {
                uint32_t k = st >> 11;
                k = k >= kOriginKeyMax - 2u ? sc.n_prims + (k - (kOriginKeyMax - 2u)) : k;  // the special keys follow the primitives
                key[i] = min(k >> shift, 255u);
                pos[i] = atomicAdd(&s_bin[key[i]], 1u);
            }
        }
    }
    __syncthreads();
    {  // exclusive scan of the 256 bins: two bins per thread, warp scan, four warp totals
        const unsigned a = s_bin[2u * threadIdx.x], b2 = s_bin[2u * threadIdx.x + 1u];
        unsigned inc = a + b2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned o = __shfl_up_sync(kFull, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        if (lane == 31u) s_warp[warp] = inc;
        __syncthreads();
        unsigned base = 0u, all = 0u;
#pragma unroll
        for (unsigned w = 0; w < kWfBlock / 32; ++w) {
            const unsigned t = s_warp[w];
            if (w < warp) base += t;
            all += t;
        }
        const unsigned excl = base + inc - (a + b2);
        s_bin[2u * threadIdx.x] = excl;
        s_bin[2u * threadIdx.x + 1u] = excl + a;
        if (threadIdx.x == 0) s_live = all;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RT_WF_EXT_TILE; ++i)
        if (key[i] != 0xFFFFFFFFu) s_order[s_bin[key[i]] + pos[i]] = (unsigned short)((uint32_t)i * kWfBlock + threadIdx.x);
    __syncthreads();
    const unsigned n_live = s_live;
    for (unsigned k = threadIdx.x; k < n_live; k += kWfBlock) extend_slot<MEDIA>(sc, pool, tile0 + s_order[k], seed, max_depth);
#else
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= pool.n_slots || (pool.state[slot] & 0xFFu) != WF_LIVE) return;
    extend_slot<MEDIA>(sc, pool, slot, seed, max_depth);
#endif
}

// ---------------------------------------------------------------------------
// control: end of a round
// ---------------------------------------------------------------------------
// cond: the handle of the CUDA-graph while-node whose body this round is (0: the host drives the loop)
__global__ void wf_control_kernel(const __grid_constant__ WfPool pool, unsigned long long *__restrict__ counters,
                                  cudaGraphConditionalHandle cond) {
    WfCtl *c = pool.ctl;
    const unsigned cur = c->round & 1u;
    const unsigned live = c->live[cur];
    counters[kCounterRays] += live;  // one segment per live slot (main.rs:48)
    c->status_live = live;
    c->live[cur ^ 1u] = 0u;
    c->ext_cursor = 0u;
    c->defer_n = 0u;
    c->round += 1u;
    c->rounds_done += 1u;
    if (cond) cudaGraphSetConditional(cond, live != 0u ? 1u : 0u);  // another round while any path is alive
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static cudaError_t wf_launch_init(const WfPool &pool, cudaStream_t stream) {
    unsigned blocks = (pool.n_slots + 255u) / 256u;
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    if (blocks < 1u) blocks = 1u;
    wf_init_kernel<<<blocks, 256, 0, stream>>>(pool);
    return cudaGetLastError();
}

static cudaError_t wf_launch_round(const DScene &sc, const RtCamera &cam, const RenderParams &P, const WfPool &pool,
                            double *planes, unsigned long long *counters, bool media, int sms, uint32_t leave_threshold,
                            unsigned long long cond_handle, cudaStream_t stream) {
    static int ext_per_sm[2] = {0, 0};
    if (ext_per_sm[0] == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ext_per_sm[0], wf_extend_kernel<false>, kWfBlock, 0);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ext_per_sm[1], wf_extend_kernel<true>, kWfBlock, 0);
        if (e != cudaSuccess) {
            ext_per_sm[0] = 0;
            return e;
        }
    }
    // shade and generate: one slot per thread.  extend: one resident wave of persistent warps, but
    // never more warps than there are reservations to hand out.
    const unsigned per_slot = (pool.n_slots + kWfBlock - 1) / kWfBlock;
    const unsigned per_reserve = (pool.n_slots + kWfReserve - 1) / kWfReserve;
    const int occ = ext_per_sm[media ? 1 : 0];
    unsigned ext_grid = (unsigned)(sms * (occ > 0 ? occ : 1));
    const unsigned ext_want = (per_reserve + (kWfBlock / 32) - 1) / (kWfBlock / 32);
    if (ext_grid > ext_want) ext_grid = ext_want ? ext_want : 1u;
    const unsigned per_tile = (pool.n_slots + kShadeTile - 1) / kShadeTile;
    const unsigned tiles = per_tile ? per_tile : 1u;
    wf_shade_kernel<<<tiles, kWfBlock, 0, stream>>>(sc, P, pool, counters);
    if (feat(F_TEX)) {  // pass 2 of shade; a grid-stride loop over a queue whose length only the device knows
        unsigned g = per_slot / 16u;
        wf_shade_deferred_kernel<<<g ? g : 1u, kWfBlock, 0, stream>>>(sc, P, pool, counters);
    }
    wf_generate_kernel<<<per_slot ? per_slot : 1u, kWfBlock, 0, stream>>>(cam, P, pool, planes, counters);
    if (leave_threshold == 33u) {  // the simple form: one slot per thread
        const unsigned per_ext = (pool.n_slots + kExtTile - 1) / kExtTile;
        if (media) wf_extend_simple_kernel<true><<<per_ext ? per_ext : 1u, kWfBlock, 0, stream>>>(sc, pool, P.seed, P.max_depth);
        else wf_extend_simple_kernel<false><<<per_ext ? per_ext : 1u, kWfBlock, 0, stream>>>(sc, pool, P.seed, P.max_depth);
    } else if (media) {
        wf_extend_kernel<true><<<ext_grid, kWfBlock, 0, stream>>>(sc, pool, P.seed, P.max_depth, leave_threshold);
    } else {
        wf_extend_kernel<false><<<ext_grid, kWfBlock, 0, stream>>>(sc, pool, P.seed, P.max_depth, leave_threshold);
    }
    wf_control_kernel<<<1, 1, 0, stream>>>(pool, counters, (cudaGraphConditionalHandle)cond_handle);
    return cudaGetLastError();
}

