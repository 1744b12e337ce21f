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
// self-test: Vec3 / f64 with the shared reciprocal (trace.cuh) against the compiler's own division
// ---------------------------------------------------------------------------
// Operand classes, chosen by the pair's index: random bit patterns (every exponent, NaN, inf, denormals), operands
// of the magnitudes a render sees (exponents within +-40 of 1), exact zeros of both signs, significands of all ones or
// one bit, and quotients at the edges of the normal range.  A mismatch is a quotient whose 64 bits differ from
// `a / b` (two NaNs count as equal: their payloads are not defined by IEEE 754).
__global__ void division_selftest_kernel(unsigned long long n, uint32_t seed, unsigned long long *mismatches) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
        uint32_t c0 = (uint32_t)k, c1 = (uint32_t)(k >> 32), c2 = 0x5eedu, c3 = seed;
        philox4x32_10(c0, c1, c2, c3, 0x243F6A88u, 0x85A308D3u);
        uint32_t d0 = c0 ^ 0x9E3779B9u, d1 = c1, d2 = c2, d3 = c3 + 1u;
        philox4x32_10(d0, d1, d2, d3, 0x13198A2Eu, 0x03707344u);
        unsigned long long w[4] = {((unsigned long long)c0 << 32) | c1, ((unsigned long long)c2 << 32) | c3,
                                   ((unsigned long long)d0 << 32) | d1, ((unsigned long long)d2 << 32) | d3};
        const uint32_t cls = (uint32_t)(k % 7ull);
        double v[4];
        for (int j = 0; j < 4; ++j) {
            unsigned long long bits = w[j];
            if (cls == 1 || cls == 2) {  // exponent within +-40 of 1.0
                const unsigned long long ex = 1023ull - 40ull + (bits >> 52) % 81ull;
                bits = (bits & 0x800FFFFFFFFFFFFFull) | (ex << 52);
            } else if (cls == 3 && j < 3) {
                bits = (bits & 0x8000000000000000ull);  // +-0 numerators
            } else if (cls == 4) {
                bits = (bits & 0xFFF0000000000000ull) | ((bits & 1ull) ? 0x000FFFFFFFFFFFFFull : 1ull << (bits % 52ull));
                const unsigned long long ex = 1023ull - 30ull + ((bits >> 52) & 0x7FFull) % 61ull;
                bits = (bits & 0x800FFFFFFFFFFFFFull) | (ex << 52);
            } else if (cls == 5) {  // quotients near the ends of the normal range
                const unsigned long long ex = j < 3 ? ((bits >> 52) & 1ull ? 40ull : 2000ull) : ((bits >> 52) & 2ull ? 1060ull : 980ull);
                bits = (bits & 0x800FFFFFFFFFFFFFull) | (ex << 52);
            }
            v[j] = __longlong_as_double((long long)bits);
        }
        const V3 q = mk(v[0], v[1], v[2]) / v[3];
        const double got[3] = {q.x, q.y, q.z};
        for (int j = 0; j < 3; ++j) {
            const double want = __ddiv_rn(v[j], v[3]);
            const bool same = (want != want && got[j] != got[j]) || __double_as_longlong(want) == __double_as_longlong(got[j]);
            if (!same) ++bad;
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}
cudaError_t selftest_division(int device, unsigned long long n, uint32_t seed, unsigned long long *mismatches_host) {
    (void)device;
    unsigned long long *dev = nullptr;
    cudaError_t e = cudaMalloc((void **)&dev, sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    e = cudaMemset(dev, 0, sizeof(unsigned long long));
    if (e == cudaSuccess) {
        division_selftest_kernel<<<148 * 8, 256>>>(n, seed, dev);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(mismatches_host, dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(dev);
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
