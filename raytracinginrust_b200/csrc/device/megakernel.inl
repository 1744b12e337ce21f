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
                    uint64_t lin;
                    const uint32_t chunk = item_split(P, item, lin);
                    if (!item_pixel(P, lin, i, row)) continue;
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

// variant bits 0-1: the register budget, as resident blocks per SM - 0: 6 blocks (80 registers), 1: 8 (64),
// 2: 12 (40).  Measured per scene class (profiles/r1_e_launch_bounds.md): flat scenes peak at 6, media
// and triangle-BVH scenes at 8, sphere-BVH scenes (cheap leaves, latency-bound) at 12.  A 72-register build (7 blocks: the smallest budget at
// which the BVH node loop keeps its ray constants in registers) was measured in r2-g and changed nothing
// (profiles/r2_g_register_budgets.md).
// variant bit 2: the scene has media (the kernel carries the boundary-query loop of medium.rs)
// f(kernel, threads per block)
template <class F>
static cudaError_t with_render_kernel(int variant, F f) {
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
