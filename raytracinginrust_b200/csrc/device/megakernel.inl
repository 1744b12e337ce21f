// megakernel.inl — render_kernel: the nested pixel / sample loop of src/main.rs:772-834 as ONE
// persistent kernel (included by pipelines.cu inside the variant namespace).
// Design (DESIGN.md "Kernels"): persistent threads, one path per lane, per-lane regeneration.
// A work item is (sample chunk, pixel); a lane pulls items from a global counter, runs the
// chunk's samples one after the other in sample order, and writes the chunk's f64 sum to its
// own slot of a [chunk][pixel] plane — no atomics on pixel data, so the image is
// bit-reproducible run to run.  reduce_planes_kernel (kernels.cu) then adds the planes in chunk
// order into the fp32 image.

// Several register budgets of the same kernel (a launch bound is a compile-time property); which
// one a scene runs is chosen in rt_scene_create (see with_render_kernel below).
template <int MIN_BLOCKS, bool MEDIA>
__global__ void __launch_bounds__(kRenderBlock, MIN_BLOCKS)
render_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RtCamera cam,
              const __grid_constant__ RenderParams P, double *__restrict__ planes,
              unsigned long long *__restrict__ counters) {
    unsigned long long n_paths = 0, n_rays = 0, n_bad = 0;
    PathState ps;
    bool alive = false, have_item = false;
    uint32_t i = 0, row = 0, s = 0, s_end = 0;
    uint64_t slot = 0;
    V3 sum = mk(0.0, 0.0, 0.0);
    for (;;) {
        if (!alive) {
            if (!have_item || s == s_end) {
                if (have_item) {
                    double *dst = planes + 3 * slot;
                    dst[0] = sum.x;
                    dst[1] = sum.y;
                    dst[2] = sum.z;
                    have_item = false;
                }
                // next (chunk, pixel) item; skip the padding of partial tiles
                for (;;) {
                    unsigned long long item = atomicAdd(&counters[kCounterWork], 1ull);
                    if (item >= P.n_items) break;
                    uint32_t chunk = (uint32_t)(item / P.items_per_chunk);
                    uint64_t lin = item - (uint64_t)chunk * P.items_per_chunk;
                    if (!item_pixel(P.tiles_x, P.width, P.height, lin, i, row)) continue;
                    s = P.sample_begin + chunk * P.chunk_size;
                    s_end = min(s + P.chunk_size, P.sample_end);
                    slot = (uint64_t)chunk * P.width * P.height + (uint64_t)row * P.width + i;
                    sum = mk(0.0, 0.0, 0.0);
                    have_item = true;
                    break;
                }
                if (!have_item) break;
            }
            // row 0 of the image is j = H-1 (main.rs:772)
            path_begin(ps, cam, P.width, P.height, i, P.height - 1u - row, s, P.seed, P.max_depth);
            ++s;
            ++n_paths;
            alive = true;
        }
        alive = path_step<MEDIA>(sc, ps, P.integrator, P.flags);
        if (!alive) {
            n_rays += ps.segments;
            // no NaN guard, like the reference (§Q10); only counted
            if (!(isfinite(ps.radiance.x) && isfinite(ps.radiance.y) && isfinite(ps.radiance.z))) ++n_bad;
            sum = sum + ps.radiance;  // vec.rs:253-260 Sum, in sample order
        }
    }
    atomicAdd(&counters[kCounterPaths], n_paths);
    atomicAdd(&counters[kCounterRays], n_rays);
    atomicAdd(&counters[kCounterNonFinite], n_bad);
}

// render_deferred_kernel — the same loop for scenes whose world mixes flat groups (walls: a handful of rects scanned
// linearly) with BVH groups (meshes), e.g. BASELINE configs[4].  In render_kernel a warp walks a mesh BVH whenever ANY
// lane's ray passes the mesh's bounds, with the other lanes idle: measured 7.5 of 32 lanes through 68 % of the
// kernel's instructions (profiles/r2_b_render_kernel_mesh4spp.txt).  Here a segment's search is cut in two (the winner
// does not depend on the order groups are visited in, trace.cuh: trace_groups_sel): every lane first scans the flat
// groups; a lane whose ray still has a BVH to walk WAITS with its path in registers while the others shade, start
// new samples and scan again - until P.defer_threshold lanes of the warp wait (or nobody else can make progress).
// Then the warp walks the BVHs once, for all of them.  Same work items, same per-path arithmetic, same planes:
// the image is bit-identical to render_kernel's (test_deferred_traversal_equals_plain_megakernel).
//
// RT_SMEM_TOP = N > 0 (A/B build, BASELINE's "upper BVH levels staged in shared memory"): the first N nodes of the
// largest BVH of the world - its top levels, the nodes are stored breadth-first - are copied to shared memory by
// each block and read from there.

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kRenderBlock, MIN_BLOCKS)
render_deferred_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RtCamera cam,
                       const __grid_constant__ RenderParams P, double *__restrict__ planes,
                       unsigned long long *__restrict__ counters) {
    const unsigned kAll = 0xFFFFFFFFu;
    const float4 *s_top = nullptr;
    int top_first = 0;
#if RT_SMEM_TOP > 0
    __shared__ float4 s_top_nodes[4 * RT_SMEM_TOP];
    {
        uint32_t most = 0;
        for (uint32_t gi = 0; gi < sc.n_world_groups; ++gi) {
            const DGroup &g = sc.groups[gi];
            if (g.bvh_root >= 0 && g.n_prims > most) {
                most = g.n_prims;
                top_first = g.bvh_root;
            }
        }
        // a BVH over n primitives with at most two per leaf has at least n/2 - 1 inner nodes
        const uint32_t n_copy = min((uint32_t)RT_SMEM_TOP, most / 2u > 1u ? most / 2u - 1u : 0u);
        const float4 *src = reinterpret_cast<const float4 *>(sc.nodes + top_first);
        for (uint32_t k = threadIdx.x; k < 4u * n_copy; k += blockDim.x) s_top_nodes[k] = __ldg(src + k);
        if (n_copy < (uint32_t)RT_SMEM_TOP) top_first = 0x7fffffff;  // too small a tree: nothing staged
        __syncthreads();
        s_top = s_top_nodes;
    }
#endif
    unsigned long long n_paths = 0, n_rays = 0, n_bad = 0;
    PathState ps;
    bool alive = false, have_item = false, done = false, waiting = false;
    uint32_t i = 0, row = 0, s = 0, s_end = 0;
    uint64_t slot = 0;
    V3 sum = mk(0.0, 0.0, 0.0);
    Best win{RT_INF, kNoPrim, 0, 0};
    for (;;) {
        bool resolve = false, ended = false;
        if (!done && !waiting) {
            if (!alive) {
                if (!have_item || s == s_end) {
                    if (have_item) {
                        double *dst = planes + 3 * slot;
                        dst[0] = sum.x;
                        dst[1] = sum.y;
                        dst[2] = sum.z;
                        have_item = false;
                    }
                    for (;;) {  // next (chunk, pixel) item; skip the padding of partial tiles
                        unsigned long long item = atomicAdd(&counters[kCounterWork], 1ull);
                        if (item >= P.n_items) break;
                        uint32_t chunk = (uint32_t)(item / P.items_per_chunk);
                        uint64_t lin = item - (uint64_t)chunk * P.items_per_chunk;
                        if (!item_pixel(P.tiles_x, P.width, P.height, lin, i, row)) continue;
                        s = P.sample_begin + chunk * P.chunk_size;
                        s_end = min(s + P.chunk_size, P.sample_end);
                        slot = (uint64_t)chunk * P.width * P.height + (uint64_t)row * P.width + i;
                        sum = mk(0.0, 0.0, 0.0);
                        have_item = true;
                        break;
                    }
                    if (!have_item) done = true;
                }
                if (!done) {
                    path_begin(ps, cam, P.width, P.height, i, P.height - 1u - row, s, P.seed, P.max_depth);
                    ++s;
                    ++n_paths;
                    alive = true;
                }
            }
            if (!done) {  // path_step, first half: main.rs:42-48 up to the flat part of world.hit
                ps.radiance = mk(0.0, 0.0, 0.0);
                if (ps.depth_left == 0) {
                    alive = false;
                    ended = true;
                } else {
                    ps.segments += 1;
                    const V3 inv = mk(rcp_fast(ps.ray.d.x), rcp_fast(ps.ray.d.y), rcp_fast(ps.ray.d.z));
                    win = Best{RT_INF, kNoPrim, 0, 0};
                    waiting = trace_groups_sel<GROUPS_FLAT>(sc, 0, sc.n_world_groups, ps.ray, inv, kTMin, win);
                    resolve = !waiting;
                }
            }
        }
        const unsigned wait_mask = __ballot_sync(kAll, waiting);
        const unsigned free_mask = __ballot_sync(kAll, !done && !waiting);
        if ((wait_mask | free_mask) == 0u) break;
        if (waiting && ((unsigned)__popc(wait_mask) >= P.defer_threshold || free_mask == 0u)) {
            deferred_bvh_search(sc, ps.ray, win, s_top, top_first);
            waiting = false;
            resolve = true;
        }
        if (resolve) {  // path_step, second half: the hit record and main.rs:62-119
            HitRec rec;
            const bool hit = win.prim != kNoPrim;
            if (hit) {
                V3 o, d;
                object_ray(sc, sc.prims[win.prim].chain, ps.ray, o, d);
                resolve_hit_obj<false>(sc, ps.ray, win, exact_t_obj(sc, win, o, d, ps.ray.time, kTMin), o, d, rec);
            }
            alive = path_shade(sc, ps, hit, rec, P.integrator, P.flags);
            ended = !alive;
        }
        if (ended) {
            n_rays += ps.segments;
            if (!(isfinite(ps.radiance.x) && isfinite(ps.radiance.y) && isfinite(ps.radiance.z))) ++n_bad;
            sum = sum + ps.radiance;  // vec.rs:253-260 Sum, in sample order
        }
    }
    atomicAdd(&counters[kCounterPaths], n_paths);
    atomicAdd(&counters[kCounterRays], n_rays);
    atomicAdd(&counters[kCounterNonFinite], n_bad);
}

// variant bits 0-1: the register budget, as resident blocks per SM - 0: 6 blocks (80 registers), 1: 8 (64),
// 2: 12 (40).  Measured per scene class (profiles/r1_e_launch_bounds.md): flat scenes peak at 6, media
// and triangle-BVH scenes at 8, sphere-BVH scenes (cheap leaves, latency-bound) at 12.
// variant bit 2: the scene has media (the kernel carries the boundary-query loop of medium.rs)
// variant bit 3: deferred BVH traversal (render_deferred_kernel; scenes without media that mix flat and BVH groups)
// f(kernel, threads per block)
template <class F>
static cudaError_t with_render_kernel(int variant, F f) {
    if constexpr (feat(F_BVH)) {
        if (variant & 8) switch (variant & 3) {  // budgets of this kernel: 6, 8, 5 or 4 blocks per SM (80 / 64 / 96 / 128 registers)
            case 0: return f(render_deferred_kernel<6>, kRenderBlock);
            case 1: return f(render_deferred_kernel<8>, kRenderBlock);
            case 2: return f(render_deferred_kernel<5>, kRenderBlock);
            default: return f(render_deferred_kernel<4>, kRenderBlock);
        }
    }
    switch (variant & 7) {
        case 0: return f(render_kernel<6, false>, kRenderBlock);
        case 1: return f(render_kernel<8, false>, kRenderBlock);
        case 2: case 3: return f(render_kernel<12, false>, kRenderBlock);
        case 4: return f(render_kernel<6, true>, kRenderBlock);
        case 5: return f(render_kernel<8, true>, kRenderBlock);
        default: return f(render_kernel<12, true>, kRenderBlock);
    }
}
static cudaError_t render_grid_size(int device, int variant, int *blocks_out) {
    int sms = 0, per_sm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = with_render_kernel(variant, [&](auto k, int threads) { return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0); });
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    *blocks_out = sms * per_sm;  // persistent: exactly one resident wave
    return cudaSuccess;
}

static cudaError_t launch_render(const DScene &sc, const RtCamera &cam, const RenderParams &P, int variant, int blocks,
                          double *planes, unsigned long long *counters, cudaStream_t stream) {
    return with_render_kernel(variant, [&](auto k, int threads) {
        k<<<blocks, threads, 0, stream>>>(sc, cam, P, planes, counters);
        return cudaGetLastError();
    });
}
