// render.cpp — render(world, camera, width, height, spp, max_depth) -> pixels: the one part of the
// host layer that calls the device library (include/rtb200.h).  Kept apart from the scene graph
// (scene_api.cpp, scenes.cpp -> librtb200_scenes.so) so that a process that only needs scene
// descriptions - the CPU reference arm of bench.py - never loads the CUDA library.
#include "scene_api.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace rtb200 {

// ---------------------------------------------------------------------------
// render(): flatten, compile+upload, run the device path, read the sums back
// ---------------------------------------------------------------------------
RenderResult render(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights,
                    Color background, const Camera &camera, uint32_t width, uint32_t height,
                    uint32_t spp, uint32_t max_depth, const RtRenderOpts &opts, int device) {
    FlatScene flat(world, lights, background);
    RtScene *scene = nullptr;
    if (rt_scene_create(&flat.desc, device, &scene) != RT_OK)
        throw std::runtime_error(std::string("rt_scene_create: ") + rt_last_error());
    RenderResult out;
    out.rgb_sum.resize((size_t)width * height * 3);
    RtStatus st = rt_render(scene, &camera.pod, width, height, spp, max_depth, &opts, out.rgb_sum.data(), &out.stats);
    std::string err = st == RT_OK ? "" : rt_last_error();
    rt_scene_destroy(scene);
    if (st != RT_OK) throw std::runtime_error("rt_render: " + err);
    return out;
}

namespace {
RenderResult render_group(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights, Color background,
                          const Camera &camera, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                          const RtRenderOpts &opts, uint32_t n_gpus, bool want_sums, bool want_ppm) {
    using clock = std::chrono::steady_clock;
    const bool timing = std::getenv("RTB200_MULTI_TIMING") != nullptr;
    auto t_prev = clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        auto t = clock::now();
        std::fprintf(stderr, "[render_group] %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_prev).count());
        t_prev = t;
    };
    FlatScene flat(world, lights, background);
    lap("flatten");
    RtSceneGroup *group = nullptr;
    if (rt_scene_group_create(&flat.desc, nullptr, n_gpus, &group) != RT_OK)
        throw std::runtime_error(std::string("rt_scene_group_create: ") + rt_last_error());
    lap("rt_scene_group_create");
    RenderResult out;
    if (want_sums) out.rgb_sum.resize((size_t)width * height * 3);
    RtStatus st = rt_render_multi(group, &camera.pod, width, height, spp, max_depth, &opts,
                                  want_sums ? out.rgb_sum.data() : nullptr, &out.stats);
    lap("rt_render_multi");
    std::string err = st == RT_OK ? "" : std::string("rt_render_multi: ") + rt_last_error();
    if (st == RT_OK && want_ppm) {
        // the sample range the image holds (RtRenderOpts: 0 = all of spp)
        const uint64_t n_samples = opts.sample_count ? opts.sample_count : (spp > opts.sample_begin ? spp - opts.sample_begin : 0);
        out.ppm.resize(32 + 12 * (size_t)width * height);
        lap("ppm buffer");
        uint64_t len = 0;
        st = rt_encode_ppm(rt_scene_group_scene(group, 0), nullptr, width, height, n_samples, &out.ppm[0], out.ppm.size(), &len);
        if (st != RT_OK) err = std::string("rt_encode_ppm: ") + rt_last_error();
        out.ppm.resize(st == RT_OK ? (size_t)len : 0);
        lap("rt_encode_ppm");
    }
    rt_scene_group_destroy(group);
    lap("rt_scene_group_destroy");
    if (st != RT_OK) throw std::runtime_error(err);
    return out;
}
}  // namespace

RenderResult render_gpus(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights, Color background,
                         const Camera &camera, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                         const RtRenderOpts &opts, uint32_t n_gpus) {
    return render_group(world, lights, background, camera, width, height, spp, max_depth, opts, n_gpus, true, false);
}

RenderResult render_ppm(const HittablePtr &world, const std::shared_ptr<const HittableList> &lights, Color background,
                        const Camera &camera, uint32_t width, uint32_t height, uint32_t spp, uint32_t max_depth,
                        const RtRenderOpts &opts, uint32_t n_gpus) {
    return render_group(world, lights, background, camera, width, height, spp, max_depth, opts, n_gpus, false, true);
}

}  // namespace rtb200
