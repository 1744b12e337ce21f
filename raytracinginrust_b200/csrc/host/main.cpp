// main.cpp — command-line driver: the reference's main() (src/main.rs:577-836) with the
// compile-time constants turned into flags.  Usage mirrors `cargo run --release > image.ppm`:
//   rtb200_render --scene cornell --width 600 --height 600 --spp 1000 > image.ppm
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "scene_api.hpp"

using namespace rtb200;

int main(int argc, char **argv) {
    std::string scene = "cornell", assets = "assets";
    uint32_t width = 0, height = 0, spp = 0, depth = 0, seed = 1, cseed = 1, detail = 0;
    int device = 0;
    int gpus = -1;  // -1: one GPU (--device); 0: all GPUs of the box; N: the first N
    bool host_ppm = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char * {
            if (i + 1 >= argc) {
                std::fprintf(stderr, "missing value for %s\n", a.c_str());
                std::exit(2);
            }
            return argv[++i];
        };
        if (a == "--scene") scene = next();
        else if (a == "--assets") assets = next();
        else if (a == "--width") width = (uint32_t)std::atoi(next());
        else if (a == "--height") height = (uint32_t)std::atoi(next());
        else if (a == "--spp") spp = (uint32_t)std::atoi(next());
        else if (a == "--depth") depth = (uint32_t)std::atoi(next());
        else if (a == "--seed") seed = (uint32_t)std::atoi(next());
        else if (a == "--construction-seed") cseed = (uint32_t)std::atoi(next());
        else if (a == "--mesh-detail") detail = (uint32_t)std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--gpus") gpus = std::atoi(next());
        else if (a == "--host-ppm") host_ppm = true;
        else {
            std::fprintf(stderr,
                         "usage: %s [--scene random|cornell|cornell_smoke|final|mesh|light_room|two_spheres] [--width W] "
                         "[--height H] [--spp N] [--depth D] [--seed S] [--construction-seed S] [--assets DIR] "
                         "[--mesh-detail K] [--device I | --gpus N (0 = all)] [--host-ppm] > image.ppm\n",
                         argv[0]);
            return 2;
        }
    }
    try {
        SceneSpec spec = make_scene(scene, cseed, assets, detail);
        if (!width) width = spec.width;
        if (!height) height = spec.height;
        if (!spp) spp = spec.spp;
        if (!depth) depth = spec.max_depth;
        RtRenderOpts opts{};
        opts.seed = seed;
        opts.integrator = spec.integrator;
        RenderResult r;
        if (host_ppm) {  // the reference's way: fp32 sums to the host, format_color + one line per pixel there
            r = gpus < 0 ? render(spec.world, spec.lights, spec.background, spec.camera, width, height, spp, depth, opts, device)
                         : render_gpus(spec.world, spec.lights, spec.background, spec.camera, width, height, spp, depth, opts, (uint32_t)gpus);
            write_ppm(stdout, r.rgb_sum.data(), width, height, spp);
        } else {  // default: the P3 text is produced on the GPU
            if (gpus < 0 && device != 0) {
                std::fprintf(stderr, "--device needs --host-ppm (the multi-GPU entry takes the first N devices)\n");
                return 2;
            }
            r = render_ppm(spec.world, spec.lights, spec.background, spec.camera, width, height, spp, depth, opts,
                           gpus < 0 ? 1u : (uint32_t)gpus);
            std::fwrite(r.ppm.data(), 1, r.ppm.size(), stdout);
        }
        std::fprintf(stderr, "Done. %llu paths, %llu rays, %.1f ms on device (%.1f Mpaths/s, %.1f Mrays/s), %llu non-finite samples\n",
                     (unsigned long long)r.stats.paths, (unsigned long long)r.stats.rays, r.stats.render_ms,
                     r.stats.paths / r.stats.render_ms / 1e3, r.stats.rays / r.stats.render_ms / 1e3,
                     (unsigned long long)r.stats.nonfinite_samples);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
