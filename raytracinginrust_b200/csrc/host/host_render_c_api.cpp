// host_render_c_api.cpp — C entry points of render() for the Python harness (librtb200_host.so: the half of the host
// layer that calls the device library).
#include <cstring>

#include "host_c_api.h"
#include "scene_api.hpp"

using namespace rtb200;

extern "C" {

// The whole render(world, camera, width, height, spp, max_depth) -> pixels call on a
// catalogue scene: flatten, compile + upload, render, read back.  This is the end-to-end
// path bench.py times with host buffers.
int rth_render(const RthScene *s, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
               const RtRenderOpts *opts, int device, float *out_rgb_sum, RtStats *stats) {
    try {
        RenderResult r = render(s->spec.world, s->spec.lights, s->spec.background, s->spec.camera, width, height, spp,
                                max_depth, *opts, device);
        std::memcpy(out_rgb_sum, r.rgb_sum.data(), r.rgb_sum.size() * sizeof(float));
        if (stats) *stats = r.stats;
        return RT_OK;
    } catch (const std::exception &e) {
        rth_set_error(e.what());
        return RT_ERR_INTERNAL;
    }
}

// render() over n_gpus GPUs (0 = all) with the P3 file produced on the GPU: what the CLI writes to stdout.
// out_ppm: host buffer of `capacity` bytes (32 + 12*W*H suffices); *length = file size.
int rth_render_ppm(const RthScene *s, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                   const RtRenderOpts *opts, uint32_t n_gpus, char *out_ppm, uint64_t capacity, uint64_t *length,
                   RtStats *stats) {
    try {
        RenderResult r = render_ppm(s->spec.world, s->spec.lights, s->spec.background, s->spec.camera, width, height, spp,
                                    max_depth, *opts, n_gpus);
        if (r.ppm.size() > capacity) {
            rth_set_error("output buffer too small");
            return RT_ERR_BAD_ARGUMENT;
        }
        std::memcpy(out_ppm, r.ppm.data(), r.ppm.size());
        *length = r.ppm.size();
        if (stats) *stats = r.stats;
        return RT_OK;
    } catch (const std::exception &e) {
        rth_set_error(e.what());
        return RT_ERR_INTERNAL;
    }
}

}  // extern "C"
