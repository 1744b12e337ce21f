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
    // trees left for the device to build (CompileOptions::gpu_bvh_min_prims): the group, its primitive range, its
    // slice [node_base, node_base + n_prims - 1) of `nodes` (zeroed here; the root is node_base), its bounds, and
    // where its primitives' fp32 boxes (6 floats each: lo, hi, rounded outward) start in pending_boxes
    struct PendingBvh {
        uint32_t group, first_prim, n_prims, node_base;
        uint64_t first_box;
        double lo[3], hi[3];
    };
    std::vector<PendingBvh> pending_bvh;
    std::vector<float> pending_boxes;
};

struct CompileOptions {
    // > 0: groups with at least this many primitives get no host-built tree; their slice of the node table is
    // left empty for gpu_bvh.cu (CompiledScene::pending_bvh).  0: every tree is built here (binned SAH).
    uint32_t gpu_bvh_min_prims = 0;
};

// Returns RT_OK or an error status with a message in `err`.
RtStatus compile_scene(const RtSceneDesc &desc, CompiledScene &out, std::string &err, const CompileOptions &opts = CompileOptions());

}  // namespace rtb200dev
