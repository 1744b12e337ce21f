// variants.h — the pipelines (megakernel + wavefront stages) are compiled once per scene feature
// set; each compilation of pipelines.cu exports one descriptor.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "tables.h"
#include "wavefront.h"

namespace rtb200dev {

// name -> features compiled in.  Ordered from most to least specific; the last covers everything.
#define RT_MASK_vflat (F_RECT | F_BOX | F_METAL | F_HEAD)                                        /* Cornell box, Cornell smoke */
#define RT_MASK_vmesh (F_RECT | F_TRI | F_BVH | F_METAL | F_DIELECTRIC | F_HEAD)                  /* triangle-mesh scenes        */
#define RT_MASK_vspheres (F_SPHERE | F_MSPHERE | F_BVH | F_TEX | F_METAL | F_DIELECTRIC | F_LEGACY) /* RTiOW random spheres       */
#define RT_MASK_vnextweek (F_ALL & ~(F_TRI | F_LEGACY | F_SPHERE_LIGHT | F_PBR))                          /* Next Week final scene       */
#define RT_MASK_vall F_ALL
#define RT_VARIANT_LIST(X) X(vflat) X(vmesh) X(vspheres) X(vnextweek) X(vall)

struct PipelineVariant {
    const char *name;
    uint32_t mask;
    // megakernel; variant bit 0: 64-register build, bit 1: the scene has media
    cudaError_t (*render_grid_size)(int device, int variant, int *blocks_out);
    cudaError_t (*launch_render)(const DScene &sc, const RtCamera &cam, const RenderParams &P, int variant, int blocks,
                                 double *planes, unsigned long long *counters, cudaStream_t stream);
    // wavefront
    cudaError_t (*wf_launch_init)(const WfPool &pool, cudaStream_t stream);
    cudaError_t (*wf_launch_round)(const DScene &sc, const RtCamera &cam, const RenderParams &P, const WfPool &pool,
                                   double *planes, unsigned long long *counters, bool media, int sms, uint32_t leave_threshold,
                                   unsigned long long cond_handle, cudaStream_t stream);
};

#define RT_DECLARE_VARIANT(ns) const PipelineVariant *rtb200_variant_##ns();
RT_VARIANT_LIST(RT_DECLARE_VARIANT)
#undef RT_DECLARE_VARIANT

// The most specific variant whose mask covers `needed`.
inline const PipelineVariant *find_variant(uint32_t needed) {
#define RT_TRY_VARIANT(ns)                                   \
    {                                                        \
        const PipelineVariant *v = rtb200_variant_##ns();    \
        if ((needed & ~v->mask) == 0u) return v;             \
    }
    RT_VARIANT_LIST(RT_TRY_VARIANT)
#undef RT_TRY_VARIANT
    return rtb200_variant_vall();
}
inline const PipelineVariant *find_variant_by_name(const char *name) {
#define RT_NAME_VARIANT(ns) \
    if (!strcmp(name, #ns)) return rtb200_variant_##ns();
    RT_VARIANT_LIST(RT_NAME_VARIANT)
#undef RT_NAME_VARIANT
    return nullptr;
}

}  // namespace rtb200dev
