// wavefront.h — the HBM-resident path pool and ray queues of the wavefront pipeline
// (generate / extend / shade stages; DESIGN.md "Kernels").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace rtb200dev {

constexpr int kWfBlock = 128;
constexpr uint32_t kWfNoItem = 0xFFFFFFFFu;

enum WfState : uint32_t {
    WF_EMPTY = 0,  // no path: before the first generate, or retired because no work item is left
    WF_LIVE = 1,   // a ray waits for extend, or (after extend) its hit waits for shade
    WF_REGEN = 2,  // the path ended in shade: generate gives the slot its next sample
};

struct WfCtl {
    unsigned round;
    unsigned live[2];      // live paths after generate of round r (index r & 1): what extend r traces
    unsigned ext_cursor;   // dynamic-fetch cursor of the extend stage (slots handed out so far)
    unsigned status_live;  // live count of the last finished round (the host loop reads it back)
    unsigned rounds_done;
    unsigned defer_n;      // entries of the deferred-shade queue of this round
    unsigned pad1;
};

// One 128-byte record per slot = one L2 line, eight 128-bit units; a lane moves a unit with one
// 128-bit access.  Path state lives in its slot for the whole life of a (chunk, pixel) work item.
// The stages walk the pool in slot order - no global index queues: every slot is LIVE in steady state, so
// a queue would be the identity, and each same-address atomic a warp spends on queue positions
// costs ~4 ns of serialised L2 time (measured: 64k of them per round were the whole shade stage).
// Reordering happens inside a block's tile of slots, in shared memory (the shade pass's class sort).
// Sectors 0-1: the ray, state, depth (all extend reads); sector 2: throughput + Philox keys;
// sector 3: what extend found + the work item.
struct alignas(128) WfSlot {
    double ox, oy;        // u0  ray origin
    double oz, dx;        // u1  ray direction (never normalised, ray.rs)
    double dy, dz;        // u2
    double time;          // u3  ray time
    uint32_t state_unused;  //   (the state lives in WfPool::state)
    uint32_t depth_left;  //     bounce = max_depth - depth_left
    double bx, by;        // u4  throughput (beta)
    double bz;            // u5
    uint32_t rng_pixel;   //     Philox pixel key (j*W+i, j bottom-up)
    uint32_t sample;      //     sample index (Philox key)
    uint32_t prim;        // u6  extend result: index into prims, kMediumFlag|medium, or kNoPrim
    int32_t face;         //     BOX: which side
    double t;             //     closest_so_far in reference arithmetic
    uint32_t out_pixel;   // u7  row*W+i, row top-down
    uint32_t s_end;       //     the item's sample range ends here
    uint32_t chunk;       //     the item's chunk (kWfNoItem: the slot holds no item)
    uint32_t pad;
};
static_assert(sizeof(WfSlot) == 128, "one slot per 128-byte line");

struct WfPool {
    WfSlot *slots;
    double4 *sum;       // per slot: the item's radiance sum (x,y,z), samples added in sample order
    uint32_t *state;    // per slot: WfState in bits 0-7 (apart from the record so that sparse rounds stay cheap);
                        // bits 8-10: class of what extend found (wavefront.inl: hit_class), the shade pass's sort key
    uint32_t *defer_q;  // shade pass 2: slots whose hit material is costly (Perlin noise, image textures)
    WfCtl *ctl;
    uint32_t capacity;  // slots allocated
    uint32_t n_slots;   // slots used by the current render (<= capacity)
};

// (launchers: per pipeline variant, see variants.h)
constexpr int kWfLaunchesPerRound = 4;  // + 1 (shade pass 2) in the variants with textures

}  // namespace rtb200dev
