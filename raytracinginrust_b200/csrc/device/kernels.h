// kernels.h — launch parameters and launcher declarations (host-visible).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/rtb200.h"
#include "tables.h"

namespace rtb200dev {

#ifndef RT_RENDER_BLOCK
#define RT_RENDER_BLOCK 128
#endif
constexpr int kRenderBlock = RT_RENDER_BLOCK;

enum Counter : int { kCounterWork = 0, kCounterPaths = 1, kCounterRays = 2, kCounterNonFinite = 3, kNumCounters = 4 };

struct RenderParams {
    uint32_t width, height;
    uint32_t max_depth;
    uint32_t seed;
    uint32_t integrator;
    uint32_t flags;
    uint32_t sample_begin, sample_end;  // this launch renders samples [begin, end)
    uint32_t chunk_size;                // samples per work item
    uint32_t n_chunks;
    uint32_t tiles_x, tiles_y;          // 8x4 pixel tiles
    uint64_t items_per_chunk;           // tiles_x * tiles_y * 32 (includes padding of partial tiles)
    uint64_t n_items;                   // n_chunks * items_per_chunk
    double inv_items_per_chunk, inv_tiles_x;  // reciprocals for the item -> (chunk, tile, pixel) arithmetic (trace.cuh: item_split)
};

cudaError_t measure_fp64_peak(int device, double *tflops);
cudaError_t launch_reduce_planes(const double *planes, float *out, uint64_t n_values, uint32_t n_chunks,
                                 cudaStream_t stream);
cudaError_t launch_first_hit(const DScene &sc, const RtRay *rays, uint64_t n, RtHit *hits, cudaStream_t stream);
cudaError_t launch_path_radiance(const DScene &sc, const RtCamera &cam, const RenderParams &P, const uint32_t *px,
                                 const uint32_t *py, const uint32_t *sample, uint64_t n, double *rgb,
                                 uint32_t *segments, cudaStream_t stream);
cudaError_t launch_camera_rays(const RtCamera &cam, const RenderParams &P, const uint32_t *px, const uint32_t *py,
                               const uint32_t *sample, uint64_t n, RtRay *rays, cudaStream_t stream);


// ---- gpu_bvh.cu: a linear BVH built on the device ----
cudaError_t build_bvh_on_device(DPrim *prims, uint32_t n, const float *boxes_dev, DBvhNode *nodes, uint32_t node_base,
                                uint32_t first_prim, const double lo[3], const double hi[3], cudaStream_t st);

// ---- image.cu: multi-GPU combine, format_color, P3 text ----
constexpr uint32_t kMaxGroupDevices = 16;
struct PeerImages {
    const float *image[kMaxGroupDevices - 1];  // W*H*3 fp32 sums on the peers (mapped into the root's address space)
    uint32_t n;
};
cudaError_t launch_sum_peers(float *out, const PeerImages &peers, uint64_t n_values, int sms, cudaStream_t stream);
cudaError_t launch_format_rgb8(const float *sum, uint8_t *out, uint64_t n_values, double spp, int sms, cudaStream_t stream);
uint32_t ppm_block_count(uint64_t n_pixels);
// packed: n_pixels u32; block_len / block_off: ppm_block_count(n_pixels) entries; total: the body's size in bytes
cudaError_t launch_ppm_measure(const float *sum, uint32_t *packed, uint32_t *block_len, uint64_t *block_off, uint64_t *total,
                               uint64_t n_pixels, double spp, cudaStream_t stream);
cudaError_t launch_ppm_write(const uint32_t *packed, const uint64_t *block_off, char *body, uint64_t n_pixels, cudaStream_t stream);

}  // namespace rtb200dev
