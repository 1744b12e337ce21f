// sorted.inl — render_sorted_kernel: the megakernel with its lanes re-sorted by hit class once per segment
// (included by pipelines.cu inside the variant namespace).
//
// EXPERIMENT (RTB200_PIPELINE=sorted; not a default of any scene until measured on the B200).
// Why: in render_kernel the search runs with 31 of 32 lanes, everything after it with 9-14
// (profiles/r1_e_render_kernel_lines_cornell.txt): a warp holds misses, light hits, Lambertian and metal hits and
// lanes that start their next sample, each a stretch of code of its own - ~70 % of the kernel's issue slots.  The
// wavefront pipeline removes that by sorting, but pays a 128-byte record per path and stage in HBM (0.6x on the
// Cornell box).  Here the path state never leaves the SM: after the search every lane files its path in shared
// memory at the position a block-wide counting sort by hit class gives it, picks up the path filed at its own
// index, and resolves + shades that one; the lanes whose paths ended are neighbours, so whole warps start the
// next samples together.  A path (with its work item and the item's f64 sum) wanders between lanes; what it
// computes does not depend on the lane, and every item still adds its samples in sample order, so the image is
// bit-identical to render_kernel's.  The per-lane phases are in sorted_phases.cuh (and run on the CPU in the test
// tier); this file is the block-level part: the sort and the barriers.
#include "sorted_phases.cuh"

template <int BLOCK, int MIN_BLOCKS, bool MEDIA>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS)
render_sorted_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RtCamera cam,
                     const __grid_constant__ RenderParams P, double *__restrict__ planes,
                     unsigned long long *__restrict__ counters) {
    __shared__ SortedShared<BLOCK> sh;
    __shared__ unsigned s_bin[2][kSortedClasses + 1];
    const unsigned tid = threadIdx.x, lane = tid & 31u;
    SortedLane L;
    sorted_lane_init(L, P.seed);
    if (tid < (unsigned)kSortedClasses + 1u) s_bin[0][tid] = s_bin[1][tid] = 0u;
    __syncthreads();
    for (unsigned iter = 0;; ++iter) {
        const uint32_t cls = sorted_generate_search<MEDIA>(sc, cam, P, planes, counters, L);
        // block-wide counting sort by class: one shared atomic per class and warp
        unsigned *bin = s_bin[iter & 1u];
        const unsigned peers = __match_any_sync(kFull, cls);
        const unsigned leader = (unsigned)__ffs((int)peers) - 1u;
        unsigned base = 0u;
        if (lane == leader) base = atomicAdd(&bin[cls], (unsigned)__popc(peers));
        base = __shfl_sync(kFull, base, (int)leader);
        unsigned dst = base + (unsigned)__popc(peers & ((1u << lane) - 1u));
        __syncthreads();
        for (uint32_t c = 0; c < cls; ++c) dst += bin[c];
        const bool all_idle = bin[kSortedIdle] == (unsigned)BLOCK;  // block-uniform: every lane reads the same count
        // the other set of bins is idle between these two barriers (read before the second barrier of the previous
        // segment, counted into after the second barrier of this one): clear it for the next segment
        if (tid < (unsigned)kSortedClasses + 1u) s_bin[(iter & 1u) ^ 1u][tid] = 0u;
        sorted_file(sh, dst, L);
        __syncthreads();
        if (all_idle) break;
        sorted_pickup(sh, tid, P, L);
        sorted_shade(sc, P, L);
    }
    atomicAdd(&counters[kCounterPaths], L.n_paths);
    atomicAdd(&counters[kCounterRays], L.n_rays);
    atomicAdd(&counters[kCounterNonFinite], L.n_bad);
}
