// pipelines.cu — the two render pipelines (megakernel.inl, wavefront.inl) for ONE scene feature
// set.  The Makefile compiles this file once per variant of variants.h with -DRT_VARIANT_NS=<name>;
// the feature mask follows from the name, the device code of trace.cuh lands in an inline
// namespace of that name, and the one exported symbol is rtb200_variant_<name>().
#include <cuda_runtime.h>

#include <cstring>

#include "variants.h"

#ifndef RT_VARIANT_NS
#define RT_VARIANT_NS vall
#endif
#define RT_CAT2(a, b) a##b
#define RT_CAT(a, b) RT_CAT2(a, b)
#define RT_STR2(a) #a
#define RT_STR(a) RT_STR2(a)
#define RT_FEAT_MASK RT_CAT(RT_MASK_, RT_VARIANT_NS)

#include "trace.cuh"

namespace rtb200dev {
inline namespace RT_VARIANT_NS {

#include "wavefront.inl"
#include "megakernel.inl"

}  // namespace RT_VARIANT_NS

const PipelineVariant *RT_CAT(rtb200_variant_, RT_VARIANT_NS)() {
    static const PipelineVariant v = {RT_STR(RT_VARIANT_NS), (uint32_t)(RT_FEAT_MASK), &RT_VARIANT_NS::render_grid_size, &RT_VARIANT_NS::launch_render,
                                      &RT_VARIANT_NS::wf_launch_init, &RT_VARIANT_NS::wf_launch_round};
    return &v;
}

}  // namespace rtb200dev
