// host_c_api.cpp — C entry points over the scene-graph half of the C++ host layer (librtb200_scenes.so), for the
// Python test and bench harness (ctypes).  The Rust host would call the classes' equivalents directly.
// Nothing here calls the device library; the render entry points are in host_render_c_api.cpp.
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>

#include "host_c_api.h"
#include "scene_api.hpp"

using namespace rtb200;

static thread_local std::string g_host_err;

extern "C" {

const char *rth_last_error(void) { return g_host_err.c_str(); }
void rth_set_error(const char *message) { g_host_err = message ? message : ""; }  // for host_render_c_api.cpp

// Build one of the catalogue scenes (scenes.cpp) and flatten it.
int rth_scene_build(const char *name, uint32_t construction_seed, const char *assets_dir, uint32_t mesh_detail,
                    RthScene **out) {
    if (!name || !assets_dir || !out) {
        g_host_err = "null argument";
        return RT_ERR_BAD_ARGUMENT;
    }
    try {
        std::unique_ptr<RthScene> s(new RthScene(make_scene(name, construction_seed, assets_dir, mesh_detail)));
        s->flat.reset(new FlatScene(s->spec.world, s->spec.lights, s->spec.background));
        *out = s.release();
        return RT_OK;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return RT_ERR_BAD_ARGUMENT;
    }
}
void rth_scene_free(RthScene *s) { delete s; }
const RtSceneDesc *rth_scene_desc(const RthScene *s) { return &s->flat->desc; }
const RtCamera *rth_scene_camera(const RthScene *s) { return &s->spec.camera.pod; }
// integrator, width, height, spp, max_depth of the BASELINE config the scene belongs to
void rth_scene_config(const RthScene *s, uint32_t out[5]) {
    out[0] = s->spec.integrator;
    out[1] = s->spec.width;
    out[2] = s->spec.height;
    out[3] = s->spec.spp;
    out[4] = s->spec.max_depth;
}

// Camera::new (src/camera.rs:19-49)
void rth_camera_new(const double lookfrom[3], const double lookat[3], const double vup[3], double vfov,
                    double aspect_ratio, double aperture, double focus_dist, double time0, double time1, RtCamera *out) {
    Camera c(Point3(lookfrom[0], lookfrom[1], lookfrom[2]), Point3(lookat[0], lookat[1], lookat[2]),
             Vec3(vup[0], vup[1], vup[2]), vfov, aspect_ratio, aperture, focus_dist, time0, time1);
    *out = c.pod;
}

// Vec3::format_color (src/vec.rs:125-131) over a whole image of fp32 sums -> RGB8
void rth_format_image(const float *rgb_sum, uint64_t n_pixels, uint64_t samples_per_pixel, uint8_t *out_rgb8) {
    for (uint64_t p = 0; p < n_pixels; ++p) {
        uint64_t c[3];
        format_color(rgb_sum + 3 * p, samples_per_pixel, c);
        out_rgb8[3 * p] = (uint8_t)c[0];
        out_rgb8[3 * p + 1] = (uint8_t)c[1];
        out_rgb8[3 * p + 2] = (uint8_t)c[2];
    }
}
int rth_write_ppm(const char *path, const float *rgb_sum, uint32_t width, uint32_t height, uint64_t samples_per_pixel) {
    FILE *f = std::fopen(path, "w");
    if (!f) {
        g_host_err = std::string("cannot open ") + path;
        return RT_ERR_BAD_ARGUMENT;
    }
    write_ppm(f, rgb_sum, width, height, samples_per_pixel);
    std::fclose(f);
    return RT_OK;
}

// OBJ reader check (mesh.rs:40-52): number of triangles of the first model
int rth_obj_triangle_count(const char *path, uint64_t *n_vertices, uint64_t *n_triangles) {
    try {
        std::vector<float> pos;
        std::vector<uint32_t> idx;
        read_obj_first_model(path, pos, idx);
        *n_vertices = pos.size() / 3;
        *n_triangles = idx.size() / 3;
        return RT_OK;
    } catch (const std::exception &e) {
        g_host_err = e.what();
        return RT_ERR_BAD_ARGUMENT;
    }
}

}  // extern "C"
