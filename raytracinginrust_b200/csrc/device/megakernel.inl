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

// render_sliced_kernel — the same loop with the tree walk of world.hit cut into SLICES of RT_SLICE_STEPS node / leaf
// visits.  In render_kernel a warp's iteration lasts as long as its longest tree walk, and walk lengths are
// heavy-tailed (most rays that enter a mesh's bounds leave the tree after a few nodes, a few graze the surface for
// hundreds): the node loop runs at 4-5 of 32 lanes (profiles/r2_d_render_kernel_mesh4spp.txt).  Here a lane whose
// walk is not finished when the slice ends keeps its traversal state (node, stack, ray constants) and goes on in
// the next iteration, while the lanes that are done shade, start their next segment and join the next slice with
// fresh walks: short walks no longer wait for long ones, and a slice holds old and new walks together.
// The per-path arithmetic is render_kernel's, so the image is bit-identical
// (test_sliced_traversal_equals_plain_megakernel).
#ifndef RT_SLICE_STEPS
#define RT_SLICE_STEPS 32
#endif
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(kRenderBlock, MIN_BLOCKS)
render_sliced_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RtCamera cam,
                     const __grid_constant__ RenderParams P, double *__restrict__ planes,
                     unsigned long long *__restrict__ counters) {
    const int kDone = (int)0x80000000;
    unsigned long long n_paths = 0, n_rays = 0, n_bad = 0;
    PathState ps;
    bool alive = false, have_item = false, done = false, searching = false;
    uint32_t i = 0, row = 0, s = 0, s_end = 0;
    uint64_t slot = 0;
    V3 sum = mk(0.0, 0.0, 0.0);
    // the search of the current segment (world.hit, main.rs:48): next group to open, the walk in progress
    uint32_t gi = 0;
    int node = kDone, sp = 0;
    int stack[kStackSize];
    SRay r;
    FRay f;
    float t_max_f = 0.f;
    Best win{RT_INF, kNoPrim, 0, 0};
    const float t_min_f = __double2float_rd(kTMin);
    r.o = r.d = r.inv = mk(0.0, 0.0, 0.0);
    r.time = 0.0;
    f = FRay{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const unsigned kAll = 0xFFFFFFFFu;
    for (;;) {
        // The phases below are re-aligned with __syncwarp: left to itself the compiler keeps lanes that took
        // different branches apart through the whole body (measured: the flat scan at 10 of 32 lanes, the slice at
        // 3), and a slice only pays if the lanes that walk a tree walk it together.  So every lane stays in the loop
        // until the whole warp is done.
        bool ended = false;
        __syncwarp(kAll);
        if (!searching && !done) {
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
            if (!done) {  // path_step, first half (main.rs:42-48): a new segment's search starts
                ps.radiance = mk(0.0, 0.0, 0.0);
                if (ps.depth_left == 0) {
                    alive = false;
                    ended = true;
                } else {
                    ps.segments += 1;
                    win = Best{RT_INF, kNoPrim, 0, 0};
                    gi = 0;
                    node = kDone;
                    searching = true;
                }
            }
        }
        __syncwarp(kAll);
        if (searching && node == kDone) {  // open groups until one has a tree to walk (trace_groups / trace_one_group)
            const V3 inv = mk(rcp_fast(ps.ray.d.x), rcp_fast(ps.ray.d.y), rcp_fast(ps.ray.d.z));
            while (gi < sc.n_world_groups) {
                const DGroup &g = sc.groups[gi++];
                double e;
                if ((g.flags & GROUP_CULL) && !slab(ps.ray.o, inv, g.bmin, g.bmax, kTMin, win.t, e)) continue;
                r.o = ps.ray.o;
                r.d = ps.ray.d;
                r.time = ps.ray.time;
                r.inv = inv;
                if (g.flags & GROUP_XFORM) {
                    const double *m = g.m;
                    const V3 o = ps.ray.o, d = ps.ray.d;
                    r.o = mk(fma(m[0], o.x, fma(m[1], o.y, fma(m[2], o.z, g.t[0]))), fma(m[3], o.x, fma(m[4], o.y, fma(m[5], o.z, g.t[1]))),
                             fma(m[6], o.x, fma(m[7], o.y, fma(m[8], o.z, g.t[2]))));
                    if (g.flags & GROUP_ROTATED) {
                        r.d = mk(fma(m[0], d.x, fma(m[1], d.y, m[2] * d.z)), fma(m[3], d.x, fma(m[4], d.y, m[5] * d.z)),
                                 fma(m[6], d.x, fma(m[7], d.y, m[8] * d.z)));
                        r.inv = mk(rcp_fast(r.d.x), rcp_fast(r.d.y), rcp_fast(r.d.z));
                    }
                }
                if (g.bvh_root < 0) {  // the whole group is one leaf
                    const uint32_t code = ~(uint32_t)g.bvh_root;
                    const uint32_t first = code >> 3, count = (code & 7u) + 1u;
                    for (uint32_t k = 0; k < count; ++k) s_prim(sc, first + k, r, kTMin, win);
                    continue;
                }
                node = g.bvh_root;
                sp = 0;
                f = make_fray(r);
                t_max_f = __double2float_ru(win.t);
                break;
            }
        }
        __syncwarp(kAll);
        if (searching) {
            int budget = RT_SLICE_STEPS;  // one slice of the walk (trace_group's while-while, resumable)
            while (node != kDone && budget > 0) {
                while (node >= 0 && budget > 0) {
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
                    --budget;
                }
                if (node < 0 && node != kDone) {
                    const uint32_t code = ~(uint32_t)node;
                    const uint32_t first = code >> 3, count = (code & 7u) + 1u;
                    const double before = win.t;
                    for (uint32_t k = 0; k < count; ++k) s_prim(sc, first + k, r, kTMin, win);
                    if (win.t != before) t_max_f = __double2float_ru(win.t);
                    node = sp ? stack[--sp] : kDone;
                    --budget;
                }
            }
        }
        __syncwarp(kAll);
        if (searching && node == kDone && gi >= sc.n_world_groups) {  // the search is complete: the record and main.rs:62-119
            searching = false;
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
        if (__all_sync(kAll, done && !searching)) break;
    }
    atomicAdd(&counters[kCounterPaths], n_paths);
    atomicAdd(&counters[kCounterRays], n_rays);
    atomicAdd(&counters[kCounterNonFinite], n_bad);
}

// variant bits 0-1: the register budget, as resident blocks per SM - 0: 6 blocks (80 registers), 1: 8 (64),
// 2: 12 (40).  Measured per scene class (profiles/r1_e_launch_bounds.md): flat scenes peak at 6, media
// and triangle-BVH scenes at 8, sphere-BVH scenes (cheap leaves, latency-bound) at 12.  A 72-register build (7 blocks: the smallest budget at
// which the BVH node loop keeps its ray constants in registers) was measured in r2-g and changed nothing
// (profiles/r2_g_register_budgets.md).
// variant bit 2: the scene has media (the kernel carries the boundary-query loop of medium.rs)
// variant bit 3: sliced tree walk (render_sliced_kernel; scenes with a BVH and without media)
// f(kernel, threads per block)
template <class F>
static cudaError_t with_render_kernel(int variant, F f) {
    if constexpr (feat(F_BVH)) {
        if (variant & 8) switch (variant & 3) {  // budgets of this kernel: 6, 8, 5 or 4 blocks per SM (80 / 64 / 96 / 128 registers)
            case 0: return f(render_sliced_kernel<6>, kRenderBlock);
            case 1: return f(render_sliced_kernel<8>, kRenderBlock);
            case 2: return f(render_sliced_kernel<5>, kRenderBlock);
            default: return f(render_sliced_kernel<4>, kRenderBlock);
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
