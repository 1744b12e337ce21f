// compile.h — host-side scene compiler: RtSceneDesc (a serialised scene graph)
// -> the flat tables of tables.h.  Runs on the CPU once per rt_scene_create.
#pragma once
#include <string>
#include <vector>

#include "../../../include/rtb200.h"
#include "tables.h"

namespace rtb200dev {

struct CompiledScene {
    std::vector<DPrim> prims;
    std::vector<DOp> ops;
    std::vector<DChain> chains;
    std::vector<DGroup> groups;
    std::vector<DBvhNode> nodes;
    std::vector<DMedium> media;
    std::vector<DLight> lights;
    std::vector<DMaterial> materials;
    std::vector<DTexture> textures;
    std::vector<DImage> images;
    std::vector<DPerlin> perlin;
    std::vector<uint8_t> texels;
    uint32_t n_world_groups = 0;
    double background[3] = {0, 0, 0};
    uint32_t max_bvh_depth = 0;
    // Some MovingSphere has (time0, time1) != (0, 1): its centre extrapolates (sphere.rs:144-146) and the bounds
    // were built for shutter times in [0, 1] (compile.cpp: prim_box) - a camera whose shutter leaves that range is
    // refused at render time instead of being culled wrongly.
    bool shutter_limited = false;
};

// Returns RT_OK or an error status with a message in `err`.
RtStatus compile_scene(const RtSceneDesc &desc, CompiledScene &out, std::string &err);

}  // namespace rtb200dev
