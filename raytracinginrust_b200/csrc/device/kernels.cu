// kernels.cu — the kernels that exist once: the plane reduction, the parity hooks and the FP64
// probe (sm_100a).  The render pipelines live in pipelines.cu, compiled once per scene feature set.
#include <cuda_runtime.h>

#include "kernels.h"
#include "trace.cuh"

namespace rtb200dev {

// out[p] = (float)(sum over chunks of planes[c][p], c ascending): overwrites; a fixed summation order
__global__ void reduce_planes_kernel(const double *__restrict__ planes, float *__restrict__ out, uint64_t n_values,
                                     uint32_t n_chunks) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < n_values; k += stride) {
        double acc = 0.0;
        for (uint32_t c = 0; c < n_chunks; ++c) acc += planes[(uint64_t)c * n_values + k];
        out[k] = (float)acc;
    }
}

__global__ void first_hit_kernel(const __grid_constant__ DScene sc, const RtRay *__restrict__ rays, uint64_t n,
                                 RtHit *__restrict__ hits) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Ray r;
    r.o = ld3(rays[k].origin);
    r.d = ld3(rays[k].direction);
    r.time = rays[k].time;
    Rng rng{0, 0, 0, 0};
    HitRec rec;
    RtHit h;
    if (!world_hit<false, true>(sc, r, rng, rec)) {
        h.node = -1;
        h.face = 0;
        h.material = -1;
        h.front_face = 0;
        h.t = 0.0;
        h.u = h.v = 0.0;
        for (int a = 0; a < 3; ++a) h.position[a] = h.normal[a] = 0.0;
    } else {
        h.node = rec.node;
        h.face = rec.face;
        h.material = (int32_t)rec.material;
        h.front_face = rec.front_face ? 1 : 0;
        h.t = rec.t;
        h.u = rec.u;
        h.v = rec.v;
        h.position[0] = rec.p.x; h.position[1] = rec.p.y; h.position[2] = rec.p.z;
        h.normal[0] = rec.normal.x; h.normal[1] = rec.normal.y; h.normal[2] = rec.normal.z;
    }
    hits[k] = h;
}

__global__ void path_radiance_kernel(const __grid_constant__ DScene sc, const __grid_constant__ RtCamera cam,
                                     const __grid_constant__ RenderParams P, const uint32_t *__restrict__ px,
                                     const uint32_t *__restrict__ py, const uint32_t *__restrict__ sample, uint64_t n,
                                     double *__restrict__ rgb, uint32_t *__restrict__ segments) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    PathState ps;
    path_begin(ps, cam, P.width, P.height, px[k], py[k], sample[k], P.seed, P.max_depth);
    while (path_step<true>(sc, ps, P.integrator, P.flags)) {
    }
    rgb[3 * k] = ps.radiance.x;
    rgb[3 * k + 1] = ps.radiance.y;
    rgb[3 * k + 2] = ps.radiance.z;
    if (segments) segments[k] = ps.segments;
}

__global__ void camera_rays_kernel(const __grid_constant__ RtCamera cam, const __grid_constant__ RenderParams P,
                                   const uint32_t *__restrict__ px, const uint32_t *__restrict__ py,
                                   const uint32_t *__restrict__ sample, uint64_t n, RtRay *__restrict__ rays) {
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Rng rng{P.seed, py[k] * P.width + px[k], sample[k], 0};
    Ray r = camera_ray(cam, P.width, P.height, px[k], py[k], rng);
    RtRay o;
    o.origin[0] = r.o.x; o.origin[1] = r.o.y; o.origin[2] = r.o.z;
    o.direction[0] = r.d.x; o.direction[1] = r.d.y; o.direction[2] = r.d.z;
    o.time = r.time;
    rays[k] = o;
}

// FP64 FMA peak probe: 8 independent dependent-chains per thread
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
cudaError_t measure_fp64_peak(int device, double *tflops) {
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    const int threads = 256, blocks = sms * 8, iters = 1 << 14;
    double *buf = nullptr;
    e = cudaMalloc((void **)&buf, sizeof(double) * threads * blocks);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6 && e == cudaSuccess; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(buf, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double tf = 2.0 * 8.0 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops = best;
    return e;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
cudaError_t launch_reduce_planes(const double *planes, float *out, uint64_t n_values, uint32_t n_chunks,
                                 cudaStream_t stream) {
    uint64_t want = (n_values + 255) / 256;
    int blocks = (int)(want < 148ull * 16 ? (want ? want : 1) : 148ull * 16);
    reduce_planes_kernel<<<blocks, 256, 0, stream>>>(planes, out, n_values, n_chunks);
    return cudaGetLastError();
}
cudaError_t launch_first_hit(const DScene &sc, const RtRay *rays, uint64_t n, RtHit *hits, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    first_hit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc, rays, n, hits);
    return cudaGetLastError();
}
cudaError_t launch_path_radiance(const DScene &sc, const RtCamera &cam, const RenderParams &P, const uint32_t *px,
                                 const uint32_t *py, const uint32_t *sample, uint64_t n, double *rgb,
                                 uint32_t *segments, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    path_radiance_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc, cam, P, px, py, sample, n, rgb, segments);
    return cudaGetLastError();
}
cudaError_t launch_camera_rays(const RtCamera &cam, const RenderParams &P, const uint32_t *px, const uint32_t *py,
                               const uint32_t *sample, uint64_t n, RtRay *rays, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    camera_rays_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(cam, P, px, py, sample, n, rays);
    return cudaGetLastError();
}

}  // namespace rtb200dev
